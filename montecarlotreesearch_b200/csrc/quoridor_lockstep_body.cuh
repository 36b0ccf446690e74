// quoridor_lockstep_body.cuh — body of the production REF-policy rollout kernel (sm_100a).
//
// Replaces the per-job `state->rollout()` of RolloutJob::run (reference mcts/include/mcts.h:82-91)
// for Quoridor: Quoridor_state::rollout / pick_semirandom_move (reference examples/Quoridor/
// Quoridor.cpp:680-855), bit-exact under the draw conventions of include/mctsb.h.
//
// Shape.  One thread = one rollout, position in registers (qcore.cuh).  The 32 rollouts of a warp
// advance PLY-SYNCHRONOUSLY: every trip of the main loop plays exactly one ply of every live
// rollout, and each phase of the ply (shortest-path layers, trace, draws, candidate walls, pawn
// step) is one code location that all lanes reach together, so the warp does not fall apart into
// 32 serial programs (v1: 2.7 live lanes per instruction).  A warp whose 32 rollouts have all ended
// takes the next 32 ids from a global counter (one atomic per warp) — rollouts are keyed by id, not
// by placement, and rollouts that start together stay alike.  Warp votes (__any_sync /
// __ballot_sync) steer every data-dependent loop, so a loop runs as long as ANY lane needs it and
// nobody branches alone.
//
// Work saved, exactly (qcore.cuh "Exact pruning"): a wall can only change a shortest path if it
// removes an edge on one; goal-rooted BFS layers + one trace give that edge set per ply, and the
// shuffle convention lets a candidate's place in the shuffled pool be computed from its own draw.
//
// Code size matters: the kernel must stay inside the instruction cache, so every loop is kept
// rolled, there is ONE flood-fill loop, ONE Philox routine and ONE finish path.
#pragma once
#include "rollout_kernels.cuh"
#include "qcore.cuh"

// The body below is plain per-lane code plus a handful of warp primitives.  tests/hostemu compiles
// it for the host with a ONE-LANE warp (votes and shuffles are the identity) to check the per-lane
// logic against the oracle without a GPU; the product only ever runs the __global__ kernels.
#if defined(MCTSB_HOST_EMU)
#include <string.h>
#define LS_DEV inline
#define LS_NOINLINE
#if defined(MCTSB_EMU_WARP32)                            // tests/hostemu/simt.h: 32 fibers per warp, real votes and shuffles
#define LS_ANY(x) simt::any((x), __LINE__)
#define LS_BALLOT(x) simt::ballot((x), __LINE__)
#define LS_SHFL(v, l) simt::shfl((v), (l), __LINE__)
#define LS_SHFL_UP(v, o) simt::shfl_up((v), (o), __LINE__)
#define LS_REDUCE_ADD(v) simt::reduce_add((v), __LINE__)
#define LS_SYNCWARP() simt::syncwarp(__LINE__)
#define LS_WARP 32
#define LS_TID simt::tid()
#else
#define LS_ANY(x) (x)
#define LS_BALLOT(x) ((x) ? 1u : 0u)
#define LS_SHFL(v, l) (v)
#define LS_SHFL_UP(v, o) (v)
#define LS_REDUCE_ADD(v) (v)
#define LS_SYNCWARP()
#define LS_WARP 1
#define LS_TID 0
#endif
#define LS_ATOMIC_MAX(p, v) do { if (*(p) < (v)) *(p) = (v); } while (0)
#define LS_POPC(x) __builtin_popcount(x)
#define LS_FFS(x) __builtin_ffs((int)(x))
#define LS_HIBIT(x) (31 - __builtin_clz(x))
#define LS_UMULHI(a, b) mulhi_u32(a, b)
#define LS_LDG(p) (*(p))
#define LS_ATOMIC_ADD(p, v) ls_emu_add(p, v)
static inline unsigned long long ls_emu_add(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p += v; return o; }
#define LS_D2ULL(x) ((unsigned long long)(x))
static inline unsigned long long ls_emu_bits(double x) { unsigned long long b; memcpy(&b, &x, 8); return b; }
#define LS_DBITS(x) ls_emu_bits(x)
#define LS_UNROLL1
#define LS_UNROLL2
#else
#define LS_DEV __device__ __forceinline__
#define LS_NOINLINE static __device__ __noinline__
#define LS_ANY(x) __any_sync(0xffffffffu, (x))
#define LS_BALLOT(x) __ballot_sync(0xffffffffu, (x))
#define LS_SHFL(v, l) __shfl_sync(0xffffffffu, (v), (l))
#define LS_SHFL_UP(v, o) __shfl_up_sync(0xffffffffu, (v), (o))
#define LS_REDUCE_ADD(v) __reduce_add_sync(0xffffffffu, (v))
#define LS_SYNCWARP() __syncwarp()
#define LS_WARP 32
#define LS_ATOMIC_MAX(p, v) atomicMax((p), (v))
#define LS_POPC(x) __popc(x)
#define LS_FFS(x) __ffs((int)(x))
#define LS_HIBIT(x) (31 - __clz((int)(x)))
#define LS_UMULHI(a, b) __umulhi(a, b)
#define LS_LDG(p) __ldg(p)
#define LS_TID ((int)threadIdx.x)
#define LS_ATOMIC_ADD(p, v) atomicAdd(p, v)
#define LS_D2ULL(x) __double2ull_rz(x)
#define LS_DBITS(x) ((unsigned long long)__double_as_longlong(x))
#define LS_UNROLL1 _Pragma("unroll 1")
#define LS_UNROLL2 _Pragma("unroll 2")
#endif

namespace mctsb {

static constexpr int LS_THREADS = 128;
static constexpr int LS_SHARE = 64;                     // candidates one lane may queue per batch (a quarter per pending word)
static constexpr int LS_QCAP = 32 * LS_SHARE;           // queue entries per warp
// dynamic shared memory of a block, in words:
// [layers of search A, overlaid by the task queues during P5][best slots: one u64 per thread][ring: last three reach sets of search B]
static constexpr int LS_RING = 3;
static constexpr int LS_SMEM_WORDS = MCTSB_MAX_LAYERS * 3 * LS_THREADS + 2 * LS_THREADS + (LS_RING + 1) * 3 * LS_THREADS;   // ring + one scratch slot

// Shared memory is warp-major: every warp owns one contiguous region [layers of search A | best slots | ring of
// search B], and inside it a board word is a row of 32 lanes (no bank conflicts).  The task queue of the candidate
// phase lies over the warp's layer area, entry i at word i: the layers are dead once the trace is done and are
// rebuilt at the top of the next ply, and no other warp ever touches the region.
static constexpr int LS_WARP_LAYER_WORDS = MCTSB_MAX_LAYERS * 3 * 32;
static constexpr int LS_WARP_WORDS = LS_WARP_LAYER_WORDS + 2 * 32 + (LS_RING + 1) * 3 * 32;
static_assert(LS_WARP_WORDS * (LS_THREADS / 32) == LS_SMEM_WORDS, "shared memory layout");
struct WarpBuf {
    uint32_t *qbase; unsigned long long *best;
    LS_DEV uint32_t &queue(int i) const { return qbase[i]; }
};
static_assert(LS_QCAP <= LS_WARP_LAYER_WORDS, "the task queue must fit in the warp's layer area");

// boards of one lane: three rows of the warp's region per slot
struct SmemLayers {
    uint32_t *base;   // &region[lane]
    LS_DEV void put(int j, B81 v) { base[(3 * j + 0) * 32] = v.a; base[(3 * j + 1) * 32] = v.b; base[(3 * j + 2) * 32] = v.c; }
    LS_DEV B81 get(int j) const { return b81(base[(3 * j + 0) * 32], base[(3 * j + 1) * 32], base[(3 * j + 2) * 32]); }
};

LS_DEV uint32_t umin32(uint32_t a, uint32_t b) { return a < b ? a : b; }

LS_DEV mctsb_qstate ls_load_qstate(const void *base, uint32_t idx) {
    const uint4 *p = reinterpret_cast<const uint4 *>(base) + 2ull * idx;
    uint4 lo = LS_LDG(p), hi = LS_LDG(p + 1);           // a 32-byte record is two 16-byte vector loads
    mctsb_qstate s;
    s.h_anchor = (uint64_t)lo.x | ((uint64_t)lo.y << 32);
    s.v_anchor = (uint64_t)lo.z | ((uint64_t)lo.w << 32);
    s.wx = hi.x & 0xff; s.wy = (hi.x >> 8) & 0xff; s.bx = (hi.x >> 16) & 0xff; s.by = hi.x >> 24;
    s.wwalls = hi.y & 0xff; s.bwalls = (hi.y >> 8) & 0xff; s.turn = (hi.y >> 16) & 0xff; s.flags = hi.y >> 24;
    s.move_counter = hi.z; s.reserved = hi.w;
    return s;
}

LS_DEV void ls_store_qstate(void *base, uint64_t idx, const mctsb_qstate &s) {
    uint4 lo, hi;
    lo.x = (uint32_t)s.h_anchor; lo.y = (uint32_t)(s.h_anchor >> 32);
    lo.z = (uint32_t)s.v_anchor; lo.w = (uint32_t)(s.v_anchor >> 32);
    hi.x = s.wx | (s.wy << 8) | (s.bx << 16) | ((uint32_t)s.by << 24);
    hi.y = s.wwalls | (s.bwalls << 8) | (s.turn << 16) | ((uint32_t)s.flags << 24);
    hi.z = s.move_counter; hi.w = 0;
    uint4 *p = reinterpret_cast<uint4 *>(base) + 2ull * idx;
    p[0] = lo; p[1] = hi;
}

// the one Philox routine of the kernel (a call, not an inline expansion: ~90 instructions)
LS_NOINLINE Philox4 ls_philox(uint32_t blk, uint32_t id_lo, uint32_t id_hi, uint32_t k0, uint32_t k1) {
    return philox4x32_10(blk, id_lo, id_hi, 0u, k0, k1);
}

// one score evaluation routine (two fp64 divisions: ~250 instructions)
LS_NOINLINE double ls_eval(int wp, int bp, int w, int b, int turn) { return q_eval_from(wp, bp, w, b, turn); }

// draw source of a lane: Philox (id-keyed) or a pre-drawn stream (replay parity mode)
template <bool REPLAY> struct LaneDraws;
template <> struct LaneDraws<false> {
    uint32_t k0, k1, id_lo, id_hi, pos;
    LS_DEV void init(const RolloutParams &p, uint64_t r) {
        uint64_t id = p.first_rollout_id + r;
        k0 = (uint32_t)p.seed; k1 = (uint32_t)(p.seed >> 32); id_lo = (uint32_t)id; id_hi = (uint32_t)(id >> 32); pos = 0;
    }
    LS_DEV LaneDraws<false> from_lane(int src) const {          // another lane's stream (same seed, its rollout id)
        LaneDraws<false> o; o.k0 = k0; o.k1 = k1; o.id_lo = LS_SHFL(id_lo, src); o.id_hi = LS_SHFL(id_hi, src); o.pos = 0; return o;
    }
    LS_DEV uint32_t at(uint32_t k) const {
        Philox4 A = ls_philox(k >> 2, id_lo, id_hi, k0, k1);
        uint32_t sh = k & 3u;
        return sh == 0 ? A.x : sh == 1 ? A.y : sh == 2 ? A.z : A.w;
    }
    // draws k0 (one) and k1..k1+3 (four consecutive) in one rolled loop over the Philox blocks involved
    LS_DEV void gather(uint32_t ka, bool need_a, uint32_t kb, uint32_t &a, uint32_t *q) const {
        a = 0; q[0] = q[1] = q[2] = q[3] = 0;
        LS_UNROLL1
        for (int t = 0; t < 3; t++) {
            uint32_t blk = t == 0 ? (ka >> 2) : (kb >> 2) + (uint32_t)(t - 1);
            bool need = t == 0 ? need_a : (t == 1 || (kb & 3u) != 0);
            if (!LS_ANY(need)) continue;
            Philox4 A = ls_philox(blk, id_lo, id_hi, k0, k1);
            uint32_t w[4] = {A.x, A.y, A.z, A.w};
            if (t == 0) { uint32_t sh = ka & 3u; a = sh == 0 ? w[0] : sh == 1 ? w[1] : sh == 2 ? w[2] : w[3]; }
            else {
                int first = (int)(blk << 2) - (int)kb;       // index in q[] of this block's word 0: 0..-3 (t=1), 1..4 (t=2)
#pragma unroll
                for (int i = 0; i < 4; i++) {
#pragma unroll
                    for (int c = 0; c < 4; c++) if (first + c == i) q[i] = w[c];
                }
            }
        }
    }
};
template <> struct LaneDraws<true> {
    const uint32_t *data; uint32_t len, pos;
    LS_DEV void init(const RolloutParams &p, uint64_t r) {
        data = p.streams + (size_t)r * p.stream_stride; len = p.stream_len; pos = 0;
    }
    LS_DEV LaneDraws<true> from_lane(int src) const {
        LaneDraws<true> o;
        unsigned long long a = (unsigned long long)(size_t)data;
        a = ((unsigned long long)LS_SHFL((uint32_t)(a >> 32), src) << 32) | (unsigned long long)LS_SHFL((uint32_t)a, src);
        o.data = (const uint32_t *)(size_t)a; o.len = len; o.pos = 0; return o;
    }
    LS_DEV uint32_t at(uint32_t k) const { return data[k % len]; }
    LS_DEV void gather(uint32_t ka, bool need_a, uint32_t kb, uint32_t &a, uint32_t *q) const {
        a = need_a ? at(ka) : 0u;
#pragma unroll
        for (int i = 0; i < 4; i++) q[i] = at(kb + (uint32_t)i);
    }
};

// Flood fills of all lanes that have one pending, in one converged loop (a pawn-rooted search
// with early exit, as q_flood_dist).  `go` lanes search from `start` to `player`'s goal row.  A trip of the
// loop makes TWO expansions and tests once: reach sets only grow, so a goal row reached by the first
// expansion is still reached after the second, and a search that stopped growing shows as n2 == n1.
// Finished lanes keep stepping (their result is already recorded).
LS_DEV int warp_flood(bool go, B81 oS, B81 oE, int start, int player, QWork &wk) {
    const QFloodMasks fm = q_flood_masks(oS, oE);
    B81 r = b81_bit(go ? start : 0);
    const uint32_t ga = player ? MCTSB_GOAL_B_A : 0u, gc = player ? 0u : MCTSB_GOAL_W_C;
    bool busy = go && ((r.a & ga) | (r.c & gc)) == 0u;
    int d = (go && !busy) ? 0 : -1;                        // expansions to the goal row, -1 = walled off (or idle lane)
    int j = 0;                                             // warp-uniform: expansions made so far
    if (go) wk.flood_queries++;
    LS_UNROLL1
    while (LS_ANY(busy)) {
        const B81 n1 = q_flood_step(r, fm);
        const B81 n2 = q_flood_step(n1, fm);
        const bool hit1 = ((n1.a & ga) | (n1.c & gc)) != 0u, hit2 = ((n2.a & ga) | (n2.c & gc)) != 0u;
        const bool stuck = b81_eq(n2, n1);
        j += 2;
        if (busy && hit2) d = hit1 ? j - 1 : j;
        busy = busy && !hit2 && !stuck;
        r = n2;
    }
    if (d > 0) wk.flood_steps += (uint32_t)d;
    return d;
}

// The 12 pawn-step candidates (q_step_delta order) all lie within 20 cells of the mover's pawn, so a 64-bit
// window of a board that starts 20 cells before the pawn holds every target at a fixed bit: window position of
// candidate i = 20 + q_step_delta(i) = {11, 2, 29, 38, 19, 18, 21, 22, 10, 12, 28, 30}.
LS_DEV void b81_window(B81 v, int wi, int sh, uint32_t &lo, uint32_t &hi) {
    // bits [32 * wi + sh, +64) of the 160-bit string (0, a, b, c, 0)
    const uint32_t w0 = wi == 0 ? 0u : wi == 1 ? v.a : v.b;
    const uint32_t w1 = wi == 0 ? v.a : wi == 1 ? v.b : v.c;
    const uint32_t w2 = wi == 0 ? v.b : wi == 1 ? v.c : 0u;
    lo = sh ? fsr(w0, w1, sh) : w0; hi = sh ? fsr(w1, w2, sh) : w1;
}
// candidate mask (bit i = candidate i) -> window coordinates, and back
LS_DEV void steps_to_window(uint32_t m, uint32_t &lo, uint32_t &hi) {
    lo = ((m & 0x0d0u) << 15) | ((m & 0x001u) << 11) | ((m & 0x002u) << 1) | ((m & 0x004u) << 27) | ((m & 0x020u) << 13) |
         ((m & 0x100u) << 2) | ((m & 0x200u) << 3) | ((m & 0x400u) << 18) | ((m & 0x800u) << 19);
    hi = (m & 0x008u) << 3;
}
LS_DEV uint32_t window_to_steps(uint32_t lo, uint32_t hi) {
    return ((lo >> 15) & 0x0d0u) | ((lo >> 11) & 0x001u) | ((lo >> 1) & 0x002u) | ((lo >> 27) & 0x004u) | ((lo >> 13) & 0x020u) |
           ((lo >> 2) & 0x100u) | ((lo >> 3) & 0x200u) | ((lo >> 18) & 0x400u) | ((lo >> 19) & 0x800u) | ((hi >> 3) & 0x008u);
}

// ------------------------------------------------------------------------------------------------
template <bool REPLAY>
LS_DEV void lockstep_body(const RolloutParams &p, unsigned long long *work_counter, uint32_t *smem) {
    const int lane = LS_TID & 31, warp = LS_TID >> 5;
    SmemLayers ls, ring;
    WarpBuf wb;
    {   // the warp's region: [layers, overlaid by the task queue during P5][best slots: 32 x u64][ring of search B + scratch slot]
        uint32_t *region = smem + warp * LS_WARP_WORDS;
        ls.base = region + lane;
        wb.qbase = region;
        wb.best = reinterpret_cast<unsigned long long *>(region + LS_WARP_LAYER_WORDS);
        ring.base = region + LS_WARP_LAYER_WORDS + 2 * 32 + lane;
    }

    QState s; LaneDraws<REPLAY> d;
    {   // every lane owns a valid (if unused) position from the start
        mctsb_qstate z; z.h_anchor = z.v_anchor = 0; z.wx = 0; z.wy = 4; z.bx = 8; z.by = 4; z.wwalls = z.bwalls = 0;
        z.turn = 0; z.flags = 0; z.move_counter = 0; z.reserved = 0;
        q_unpack(s, z); d.init(p, 0);
    }
    bool have = false, exhausted = false;
    uint64_t r = 0; uint32_t leaf = 0; int plies = 0;
    unsigned long long acc_lo = 0, acc_hi = 0, acc_n = 0; uint32_t acc_leaf = 0xffffffffu;
    QWork wk = {0u, 0u, 0u};
    uint32_t done_rollouts = 0, done_plies = 0;

    LS_UNROLL1
    for (;;) {
        // ---- P0: a warp takes its next 32 rollouts TOGETHER, once all of its lanes are done (one atomic per
        //      warp).  Lanes that finish early wait for the others: nearly every playout runs its 50 plies, and
        //      rollouts that started together from one leaf stay alike — wall counts, path lengths, candidate
        //      counts — so the per-ply loops below, which run as long as the slowest lane, stay short and full.
        //      (Refilling lane by lane mixes ply 5 with ply 45 in one warp: 29.2 M instead of 32.3 M rollouts/s
        //      on the opening position, 70 M instead of 99 M on the mid-game corpus.)
        {
            unsigned want = LS_BALLOT(!have && !exhausted && (uint32_t)lane < p.lanes_per_batch);
            if (want && !LS_ANY(have)) {
                unsigned long long base = 0;
                int leader = LS_FFS(want) - 1;
                if (lane == leader) base = LS_ATOMIC_ADD(work_counter, (unsigned long long)LS_POPC(want));
                base = LS_SHFL(base, leader);
                if ((want >> lane) & 1u) {
                    r = base + (unsigned long long)LS_POPC(want & ((1u << lane) - 1u));
                    if (r >= p.n_rollouts) exhausted = true;
                    else {
                        leaf = REPLAY ? (uint32_t)r : (uint32_t)(r / p.rollouts_per_leaf);
                        q_unpack(s, ls_load_qstate(p.states, leaf));
                        d.init(p, r);
                        plies = 0; have = true;
                    }
                }
            }
        }
        if (!LS_ANY(have)) break;

        // ---- P1: who does what this ply -------------------------------------------------------------------
        // terminal test at the top of the ply (Quoridor.cpp:822-824) — only while plies remain: after the 50th
        // move the position is evaluated even if that move won (Quoridor.cpp:820,854)
        const int won = (have && plies < 50) ? q_winner(s) : 0;
        const bool capped = have && plies >= 50;
        const bool playing = have && !capped && !won;
        const int pl = s.turn, en = 1 - pl;
        int fin_kind = won ? (won == 1 ? 0 : 1) : -1;          // >= 0: this lane's rollout ends in this trip
        double fin_score = won == 1 ? 1.0 : 0.0;
        bool fin_eval = false;

        // ---- P2: shortest-path information, two goal-rooted searches per lane in one loop ---------------
        // side A = the enemy of the mover (layers stored when its cut masks are needed), side B = the mover
        // (records the distance of the pawn and of every legal step target).  Capped lanes: A = White, B = Black.
        uint64_t cheap_h = 0, cheap_v = 0;
        if (playing && s.nwalls[1] != 0) q_cheap_wall_masks(s, cheap_h, cheap_v);
        const int n_pool = popc64(cheap_h) + popc64(cheap_v);
        const bool want_cut = playing && n_pool > 0;
        const uint32_t steps = playing ? q_step_mask(s) : 0u;
        const int qa = capped ? 0 : en, qb = 1 - qa;
        const int cell_a = qa ? s.cell[1] : s.cell[0], cell_b = qa ? s.cell[0] : s.cell[1];
        const B81 goal_a = qa ? b81(MCTSB_GOAL_B_A, 0u, 0u) : b81(0u, 0u, MCTSB_GOAL_W_C);
        int sp_a = -1, sp_b = -1;
        int best_target = -1;                                  // get_best_step_move's choice (cell), -1 none
        bool overflow_a = false;
        {
            const B81 goal_b = qb ? b81(MCTSB_GOAL_B_A, 0u, 0u) : b81(0u, 0u, MCTSB_GOAL_W_C);
            const B81 pawn_a = b81_bit(cell_a), pawn_b = b81_bit(cell_b);
            B81 ra = goal_a, rb = goal_b;
            const QFloodMasks fm = q_flood_masks(s.openS, s.openE);
            // One loop for both searches, no lane-private branches.  da / db: 0xffffffff while the search is still
            // looking for its pawn, then the number of expansions it took (an unsigned minimum keeps the first
            // hit); lanes with nothing to search start at 0.  A lane whose search has ended keeps stepping: what it
            // writes afterwards lands in slots nobody reads.  Search A stores every layer in the lane's column;
            // search B stores its reach set in a ring of three while it is still looking (afterwards: in a fourth,
            // scratch slot) — the distances of the step targets are read off the last three below.  A pawn that is
            // walled off (never in a legal position, but callers may hand in anything) ends with the trip cap:
            // 81 expansions exhaust every board.
            const bool searching = playing || capped;
            uint32_t da = (searching && !b81_any(pawn_a & goal_a)) ? 0xffffffffu : 0u;
            uint32_t db = (searching && !b81_any(pawn_b & goal_b)) ? 0xffffffffu : 0u;
            if (searching) wk.flood_queries += 2;
            ring.put(0, goal_b);
            uint32_t j = 0; int slot = 0;                      // warp-uniform: expansions so far, ring slot of R_j
            LS_UNROLL1
            while (LS_ANY((int)(da | db) < 0) && j < 81u) {
                j++; slot = slot == LS_RING - 1 ? 0 : slot + 1;
                const B81 grown_a = q_flood_step(ra, fm);
                const B81 layer = andn(grown_a, ra);
                if (j < MCTSB_MAX_LAYERS) ls.put((int)j, layer);
                da = umin32(da, b81_any(layer & pawn_a) ? j : 0xffffffffu);
                ra = grown_a;
                const B81 grown_b = q_flood_step(rb, fm);
                ring.put((int)db < 0 ? slot : LS_RING, grown_b);
                db = umin32(db, b81_any(grown_b & pawn_b) ? j : 0xffffffffu);
                rb = grown_b;
            }
            if (searching) { sp_a = (int)da; sp_b = (int)db; }   // -1: walled off
            if (sp_a > 0) wk.flood_steps += (uint32_t)sp_a;
            if (sp_b > 0) wk.flood_steps += (uint32_t)sp_b;
            overflow_a = sp_a >= MCTSB_MAX_LAYERS;
            // get_best_step_move (Quoridor.cpp:476-497) measures every legal step target with its own BFS and keeps
            // the first strict minimum in reversed candidate order.  A target is one or two moves from the pawn, so
            // its distance is within two of the pawn's: the reach sets R(sp-2) < R(sp-1) < R(sp) < R(sp+1) of search
            // B sort the targets into five distance classes.  All twelve candidates are classified at once in a
            // 64-bit window around the pawn; the nearest non-empty class wins, the highest candidate index in it.
            if (LS_ANY(steps != 0u)) {
                const B81 z = b81(0u, 0u, 0u);
                const int spb = sp_b < 0 ? 0 : sp_b;
                const B81 rb0 = ring.get(spb % LS_RING);                      // cells within sp_b of the goal row
                const B81 rb1 = spb >= 1 ? ring.get((spb + 2) % LS_RING) : z;  // within sp_b - 1
                const B81 rb2 = spb >= 2 ? ring.get((spb + 1) % LS_RING) : z;  // within sp_b - 2
                const B81 rb_next = q_flood_step(rb0, fm);
                const int wpos = cell_b + 12, wi = wpos >> 5, sh = wpos & 31;      // window start = cell_b - 20, in (0, a, b, c, 0)
                uint32_t slo, shi, lo, hi;
                steps_to_window(steps, slo, shi);
                uint32_t wlo = slo, whi = shi;                                   // farthest class: every legal step
                b81_window(rb_next, wi, sh, lo, hi); if (((lo & slo) | (hi & shi)) != 0u) { wlo = lo & slo; whi = hi & shi; }
                b81_window(rb0, wi, sh, lo, hi);     if (((lo & slo) | (hi & shi)) != 0u) { wlo = lo & slo; whi = hi & shi; }
                b81_window(rb1, wi, sh, lo, hi);     if (((lo & slo) | (hi & shi)) != 0u) { wlo = lo & slo; whi = hi & shi; }
                b81_window(rb2, wi, sh, lo, hi);     if (((lo & slo) | (hi & shi)) != 0u) { wlo = lo & slo; whi = hi & shi; }
                const uint32_t cls = window_to_steps(wlo, whi);
                if (sp_b >= 0 && cls != 0u) best_target = cell_b + q_step_delta(LS_HIBIT(cls));
            }
        }
        // rollouts that used their 50 plies are scored (evaluate_position, Quoridor.cpp:629-664; A = White there)
        if (capped) { fin_kind = 2; fin_eval = true; }

        // ---- P2b: trace the enemy's shortest paths through the stored layers -> cut masks -----------------
        uint64_t cut_h = ~0ull, cut_v = ~0ull;
        {
            const bool tracing = want_cut && !overflow_a && sp_a >= 0;
            QTrace t; q_trace_init(t, b81_bit(cell_a));
            int k = tracing ? sp_a - 1 : -1;
            LS_UNROLL1
            while (LS_ANY(k >= 0)) {
                if (k >= 0) {
                    q_trace_step(t, k == 0 ? goal_a : ls.get(k), s.openS, s.openE);
                    wk.flood_steps++;
                    k--;
                }
            }
            if (tracing) q_trace_to_anchors(t, cut_h, cut_v);
        }

        // ---- P4: the ply's draws and the branch of pick_semirandom_move (Quoridor.cpp:693-721,748) --------
        const int sp_e = sp_a, sp_p = sp_b;                    // for playing lanes
        int mode = 0;                                          // 0 no wall search, 1 best, 2 guided, 3 first legal
        uint32_t base = 0; int j0 = 0;
        uint32_t u_step = 0, u_rand = 0;                       // the two draws of the step branch (consumed lazily)
        {
            if (playing) {
                // force_playwall, Quoridor.cpp:666-678; when false one draw is consumed and ignored (Quoridor.cpp:696)
                int ow = s.nwalls[pl], ew = s.nwalls[en];
                bool fp = (sp_e <= 1 && sp_p > sp_e) || (sp_e <= 2 && sp_p > sp_e + 1 && ow >= ew) || (sp_e <= 3 && sp_p > sp_e + 2 && ow > ew);
                if (!fp) d.pos++;
            }
            const bool wall_branch = playing && s.nwalls[1] != 0;   // Quoridor.cpp:696: entered iff Black has walls left
            base = d.pos;
            const uint32_t first_after = base + ((wall_branch && n_pool >= 2) ? (uint32_t)n_pool : 0u);
            uint32_t head = 0, q[4];
            d.gather(base, wall_branch && n_pool >= 2, first_after, head, q);    // every lane calls: it votes inside
            if (playing) {
                if (wall_branch) {
                    if (n_pool >= 2) j0 = (int)LS_UMULHI(head, (uint32_t)n_pool);
                    d.pos = first_after + 1u;
                    if (q[0] < MCTSB_THR_0_10) mode = 1;
                    else { d.pos++; mode = q[1] < MCTSB_THR_0_75 ? 2 : 3; }
                    const bool one = mode == 1;                // draws used so far after the keys: 1 (best) or 2
                    u_step = one ? q[1] : q[2]; u_rand = one ? q[2] : q[3];
                    if (n_pool == 0) mode = 0;                 // empty pool: every branch falls through (draws already consumed)
                } else { u_step = q[0]; u_rand = q[1]; }
            }
        }

        // ---- P5: candidate walls, evaluated by the whole warp -----------------------------------------------------
        // Best-wall lanes must measure EVERY candidate that may hurt the enemy (Quoridor.cpp:726-743), guided
        // lanes the first acceptable one in shuffled order (:753-769) — on average ten tries, with a long tail.
        // Left to their owners these searches keep a warp waiting on one or two lanes.  Instead every lane
        // publishes its candidates to a per-warp queue in shared memory, all 32 lanes take tasks from the queue
        // (the owner's boards come over with shuffles), and each result is folded into the owner's slot with one
        // 64-bit atomicMax: score = {gain (best) | 1 (guided, first legal), ~shuffle order}, so the maximum is
        // the reference's choice — best gain / first accepted, earliest in the shuffled pool on ties.
        int chosen = -1;                                       // chosen wall k, -1 none
        {
            // pending candidates, four words: [0] horizontal anchors 0..31, [1] horizontal 32..63, [2] vertical 0..31,
            // [3] vertical 32..63; candidate k = 2 * anchor + orientation
            uint32_t pend0 = 0, pend1 = 0, pend2 = 0, pend3 = 0;
            if (mode == 1 || mode == 2) {
                const uint64_t ph = cheap_h & cut_h, pv = cheap_v & cut_v;
                pend0 = (uint32_t)ph; pend1 = (uint32_t)(ph >> 32); pend2 = (uint32_t)pv; pend3 = (uint32_t)(pv >> 32);
            }
            // first legal (:773-788): the head of the shuffle goes alone; only if it is illegal the rest follows
            int head_k = mode == 3 ? q_pool_select(cheap_h, cheap_v, j0) : -1;
            bool head_only = mode == 3;
            // (A head with an open detour around one end — q_detour_safe_one — would need no search at all, but the test
            // costs the dense build more in registers than the 2 % of flood steps it saves: measured 37.2 -> 36.7 M/s.)
            const uint32_t pk_cells = (uint32_t)(uint8_t)s.cell[0] | ((uint32_t)(uint8_t)s.cell[1] << 8) |
                                      ((uint32_t)(uint8_t)s.turn << 16) | ((uint32_t)mode << 24);
            const uint32_t pk_sp = (uint32_t)(sp_e & 0xff) | ((uint32_t)(sp_p & 0xff) << 8) | ((uint32_t)n_pool << 16) | ((uint32_t)j0 << 24);
            unsigned long long *slot = wb.best + lane;
            *slot = 0ull;
            LS_SYNCWARP();                                     // every lane is done reading its layers: the queue may overwrite them
            LS_UNROLL1
            while (LS_ANY((pend0 | pend1 | pend2 | pend3) != 0u || head_k >= 0)) {
                // -- publish: every lane queues up to LS_SHARE candidates (a quarter from each pending word), offsets
                //    by an exclusive warp scan.  Order inside the queue is free: every queued candidate is measured.
                const bool is_head = head_k >= 0;
                int c0 = LS_POPC(pend0), c1 = LS_POPC(pend1), c2 = LS_POPC(pend2), c3 = LS_POPC(pend3);
                c0 = c0 < LS_SHARE / 4 ? c0 : LS_SHARE / 4; c1 = c1 < LS_SHARE / 4 ? c1 : LS_SHARE / 4;
                c2 = c2 < LS_SHARE / 4 ? c2 : LS_SHARE / 4; c3 = c3 < LS_SHARE / 4 ? c3 : LS_SHARE / 4;
                const int mine = is_head ? 1 : c0 + c1 + c2 + c3;
                int off = mine;
                LS_UNROLL1
                for (int o = 1; o < LS_WARP; o <<= 1) { int t = LS_SHFL_UP(off, o); if (lane >= o) off += t; }
                const int total = LS_SHFL(off, LS_WARP - 1);
                off -= mine;
                if (is_head) { wb.queue(off) = (uint32_t)lane | ((uint32_t)head_k << 5); c0 = c1 = c2 = c3 = 0; }
                {
                    uint32_t *qp = wb.qbase + off;
                    // one short bit-iterator loop per word: entry = owner lane | k << 5, k = kbase + 2 * bit
#define LS_PUBLISH_WORD(PEND, CNT, KBASE)                                                        \
                    {                                                                            \
                        uint32_t m = PEND; int c = CNT; const uint32_t e0 = (uint32_t)lane | ((uint32_t)(KBASE) << 5); \
                        LS_UNROLL1                                                               \
                        while (LS_ANY(c > 0)) {                                                  \
                            if (c > 0) { *qp++ = e0 | ((uint32_t)ctz32(m) << 6); m &= m - 1u; c--; } \
                        }                                                                        \
                        PEND = m;                                                                \
                    }
                    LS_PUBLISH_WORD(pend0, c0, 0)
                    LS_PUBLISH_WORD(pend1, c1, 64)
                    LS_PUBLISH_WORD(pend2, c2, 1)
                    LS_PUBLISH_WORD(pend3, c3, 65)
#undef LS_PUBLISH_WORD
                }
                const bool head_sent = head_k >= 0;
                head_k = -1;
                LS_SYNCWARP();
                // -- stage 1, 32 tasks per trip: the first query (the enemy's path with the wall; first legal:
                //    White's).  Only candidates that pass it need the second query; they are compacted to the
                //    front of the same queue (ballot + popc), so stage 2 runs with full warps again.
                int passed = 0;
                LS_UNROLL1
                for (int t0 = 0; t0 < total; t0 += LS_WARP) {
                    const int t = t0 + lane;
                    const bool valid = t < total;
                    const uint32_t ent = valid ? wb.queue(t) : 0u;
                    const int owner = (int)(ent & 31u), k = (int)((ent >> 5) & 127u);
                    B81 oS = b81(LS_SHFL(s.openS.a, owner), LS_SHFL(s.openS.b, owner), LS_SHFL(s.openS.c, owner));
                    B81 oE = b81(LS_SHFL(s.openE.a, owner), LS_SHFL(s.openE.b, owner), LS_SHFL(s.openE.c, owner));
                    const uint32_t oc = LS_SHFL(pk_cells, owner), osp = LS_SHFL(pk_sp, owner);
                    const int o_turn = (int)((oc >> 16) & 0xffu), o_mode = (int)(oc >> 24);
                    const int o_sp_e = (int)(int8_t)(osp & 0xffu);
                    b81_add_wall(oS, oE, k);                   // the candidate wall on the owner's boards
                    const int p1 = o_mode == 3 ? 0 : 1 - o_turn;
                    const int c1 = (int)((p1 ? (oc >> 8) : oc) & 0xffu);
                    const int d1 = warp_flood(valid, oS, oE, c1, p1, wk);
                    const bool pass = valid && d1 >= 0 && (o_mode == 3 || d1 - o_sp_e > 0);
                    const unsigned bal = LS_BALLOT(pass);
                    LS_SYNCWARP();
                    if (pass) wb.queue(passed + LS_POPC(bal & ((1u << lane) - 1u))) = (ent & 0xfffu) | ((uint32_t)d1 << 12);
                    passed += LS_POPC(bal);
                }
                LS_SYNCWARP();
                // -- stage 2: the mover's own path with the wall (first legal: Black's), the verdict, and the
                //    candidate's place in the shuffled pool for the ones that are acceptable
                LS_UNROLL1
                for (int t0 = 0; t0 < passed; t0 += LS_WARP) {
                    const int t = t0 + lane;
                    const bool valid = t < passed;
                    const uint32_t ent = valid ? wb.queue(t) : 0u;
                    const int owner = (int)(ent & 31u), k = (int)((ent >> 5) & 127u), d1 = (int)(ent >> 12);
                    B81 oS = b81(LS_SHFL(s.openS.a, owner), LS_SHFL(s.openS.b, owner), LS_SHFL(s.openS.c, owner));
                    B81 oE = b81(LS_SHFL(s.openE.a, owner), LS_SHFL(s.openE.b, owner), LS_SHFL(s.openE.c, owner));
                    const uint32_t oc = LS_SHFL(pk_cells, owner), osp = LS_SHFL(pk_sp, owner);
                    const uint64_t o_ch = LS_SHFL(cheap_h, owner), o_cv = LS_SHFL(cheap_v, owner);
                    const uint32_t o_base = LS_SHFL(base, owner);
                    const LaneDraws<REPLAY> o_draws = d.from_lane(owner);
                    const int o_turn = (int)((oc >> 16) & 0xffu), o_mode = (int)(oc >> 24);
                    const int o_sp_e = (int)(int8_t)(osp & 0xffu), o_sp_p = (int)(int8_t)((osp >> 8) & 0xffu);
                    const int o_n = (int)((osp >> 16) & 0xffu), o_j0 = (int)(osp >> 24);
                    b81_add_wall(oS, oE, k);
                    const int p2 = o_mode == 3 ? 1 : o_turn;
                    const int c2 = (int)((p2 ? (oc >> 8) : oc) & 0xffu);
                    const int d2 = warp_flood(valid, oS, oE, c2, p2, wk);
                    int gain = 0;                              // > 0: acceptable
                    if (valid && d2 >= 0) {
                        if (o_mode == 3) gain = 1;
                        else { int ee = d1 - o_sp_e, oe = d2 - o_sp_p; if (ee > oe) gain = o_mode == 1 ? ee - oe : 1; }
                    }
                    const bool accepted = gain > 0;
                    int rk = 0;
                    bool need_key = false;
                    if (accepted && o_n >= 2) { rk = q_pool_rank(o_ch, o_cv, k); need_key = rk != o_j0; }
                    uint32_t key = 0;
                    if (LS_ANY(need_key)) { if (need_key) key = o_draws.at(o_base + 1u + (uint32_t)(rk < o_j0 ? rk : rk - 1)); }
                    if (accepted) {
                        unsigned long long ord = (unsigned long long)k;
                        if (need_key) ord |= ((unsigned long long)key + 1ull) << 7;
                        unsigned long long score = ((unsigned long long)gain << 40) | (~ord & ((1ull << 40) - 1ull));
                        LS_ATOMIC_MAX(wb.best + owner, score);
                    }
                }
                LS_SYNCWARP();
                if (head_sent && head_only && *slot == 0ull) {   // the head was not legal: the others, in key order (rare)
                    uint64_t ph = cheap_h, pv = cheap_v;
                    q_mask_clear(ph, pv, q_pool_select(cheap_h, cheap_v, j0));
                    pend0 = (uint32_t)ph; pend1 = (uint32_t)(ph >> 32); pend2 = (uint32_t)pv; pend3 = (uint32_t)(pv >> 32);
                    head_only = false;
                }
            }
            const unsigned long long sc = *slot;
            if (mode != 0 && sc) chosen = (int)((~sc) & 127ull);
        }

        // ---- P6: pawn step when no wall was played (Quoridor.cpp:791-801) ---------------------------------------
        int target = -1;
        if (playing && chosen < 0) {
            bool best_step = s.nwalls[en] == 0;
            if (!best_step) { d.pos++; best_step = u_step < MCTSB_THR_0_80; } else { u_rand = u_step; }
            if (best_step) {
                target = best_target;                            // classified with the reach sets of search B (P2)
            } else {
                d.pos++;                                         // rand() (Quoridor.cpp:796)
                int ns = LS_POPC(steps);
                if (ns > 0) {
                    int idx = (int)((u_rand & 0x7fffffffu) % (uint32_t)ns);
                    uint32_t m = steps;
                    LS_UNROLL1
                    for (; idx > 0; idx--) m &= m - 1u;
                    target = cell_b + q_step_delta(ctz32(m));
                }
            }
            // no legal step / all targets walled off: the pick returned NULL — the reference prints, play_move
            // fails, the loop breaks and the position is evaluated (Quoridor.cpp:832-849,854)
            if (target < 0) { fin_kind = 3; fin_eval = true; }
        }

        // ---- P7: play, or finish ----------------------------------------------------------------------------------
        if (playing) {
            if (chosen >= 0) { q_play_wall(s, chosen); plies++; }
            else if (target >= 0) { q_play_step(s, target); plies++; }
        }
        if (LS_ANY(fin_eval)) {
            if (fin_eval) {
                int wp = qa == 0 ? sp_a : sp_b, bp = qa == 0 ? sp_b : sp_a;
                fin_score = ls_eval(wp, bp, s.nwalls[0], s.nwalls[1], s.turn);
            }
        }
        if (fin_kind >= 0) {
            if (p.outcomes) {
                uint4 v; unsigned long long sb = LS_DBITS(fin_score);
                v.x = (uint32_t)sb; v.y = (uint32_t)(sb >> 32);
                v.z = (uint32_t)(plies & 0xffff) | ((uint32_t)(fin_kind & 0xff) << 16); v.w = d.pos;
                reinterpret_cast<uint4 *>(p.outcomes)[r] = v;
            }
            if (p.finals) { mctsb_qstate f; q_pack(s, f); ls_store_qstate(p.finals, r, f); }
            if (p.sums) {
                if (acc_leaf != leaf) {
                    if (acc_n) {
                        LS_ATOMIC_ADD((unsigned long long *)&p.sums[acc_leaf].lo, acc_lo);
                        LS_ATOMIC_ADD((unsigned long long *)&p.sums[acc_leaf].hi, acc_hi);
                        LS_ATOMIC_ADD((unsigned long long *)&p.sums[acc_leaf].n, acc_n);
                    }
                    acc_lo = acc_hi = acc_n = 0; acc_leaf = leaf;
                }
                unsigned long long fx = LS_D2ULL(fin_score * 144115188075855872.0);
                acc_lo += fx & ((1ull << 29) - 1); acc_hi += fx >> 29; acc_n += 1;
            }
            done_rollouts++; done_plies += (uint32_t)plies;
            have = false;
        }
    }

    if (p.sums && acc_n) {
        LS_ATOMIC_ADD((unsigned long long *)&p.sums[acc_leaf].lo, acc_lo);
        LS_ATOMIC_ADD((unsigned long long *)&p.sums[acc_leaf].hi, acc_hi);
        LS_ATOMIC_ADD((unsigned long long *)&p.sums[acc_leaf].n, acc_n);
    }
    if (p.counters) {
        uint32_t a = LS_REDUCE_ADD(done_rollouts), b = LS_REDUCE_ADD(done_plies);
        uint32_t c = LS_REDUCE_ADD(wk.flood_steps), e = LS_REDUCE_ADD(wk.flood_queries);
        if (lane == 0 && a) {
            LS_ATOMIC_ADD(p.counters + 0, (unsigned long long)a); LS_ATOMIC_ADD(p.counters + 1, (unsigned long long)b);
            LS_ATOMIC_ADD(p.counters + 2, (unsigned long long)c); LS_ATOMIC_ADD(p.counters + 3, (unsigned long long)e);
        }
    }
}

}  // namespace mctsb
