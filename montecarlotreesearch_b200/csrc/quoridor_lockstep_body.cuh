// quoridor_lockstep.cu — the production REF-policy rollout kernel (sm_100a).
//
// Replaces the per-job `state->rollout()` of RolloutJob::run (reference mcts/include/mcts.h:82-91)
// for Quoridor: Quoridor_state::rollout / pick_semirandom_move (reference examples/Quoridor/
// Quoridor.cpp:680-855), bit-exact under the draw conventions of include/mctsb.h.
//
// Shape.  One thread = one rollout, position in registers (qcore.cuh).  The 32 rollouts of a warp
// advance PLY-SYNCHRONOUSLY: every trip of the main loop plays exactly one ply of every live
// rollout, and each phase of the ply (shortest-path layers, trace, draws, candidate walls, pawn
// steps) is one code location that all lanes reach together, so the warp does not fall apart into
// 32 serial programs (v1: 2.7 live lanes per instruction).  Lanes whose rollout ended pick the next
// rollout id from a global counter (warp-aggregated atomic) — rollouts are keyed by id, not by
// placement.  Warp votes (__any_sync / __ballot_sync) steer every data-dependent loop, so a loop
// runs as long as ANY lane needs it and nobody branches alone.
//
// Work saved, exactly (qcore.cuh "Exact pruning"): a wall can only change a shortest path if it
// removes an edge on one; goal-rooted BFS layers + one trace give that edge set per ply, and the
// shuffle convention lets a candidate's place in the shuffled pool be computed from its own draw.
#pragma once
#include "rollout_kernels.cuh"
#include "qcore.cuh"

// The body below is plain per-lane code plus a handful of warp primitives.  tests/hostemu compiles
// it for the host with a ONE-LANE warp (votes and shuffles are the identity) to check the per-lane
// logic against the oracle without a GPU; the product only ever runs the __global__ kernels.
#if defined(MCTSB_HOST_EMU)
#include <string.h>
#define LS_DEV inline
#define LS_ANY(x) (x)
#define LS_BALLOT(x) ((x) ? 1u : 0u)
#define LS_SHFL(v, l) (v)
#define LS_SHFL_XOR(v, o) (0)
#define LS_SHFL_UP(v, o) (v)
#define LS_SYNCWARP()
#define LS_WARP 1
#define LS_ATOMIC_MAX(p, v) do { if (*(p) < (v)) *(p) = (v); } while (0)
#define LS_POPC(x) __builtin_popcount(x)
#define LS_FFS(x) __builtin_ffs((int)(x))
#define LS_UMULHI(a, b) mulhi_u32(a, b)
#define LS_LDG(p) (*(p))
#define LS_TID 0
#define LS_ATOMIC_ADD(p, v) ls_emu_add(p, v)
static inline unsigned long long ls_emu_add(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p += v; return o; }
#define LS_D2ULL(x) ((unsigned long long)(x))
static inline unsigned long long ls_emu_bits(double x) { unsigned long long b; memcpy(&b, &x, 8); return b; }
#define LS_DBITS(x) ls_emu_bits(x)
#else
#define LS_DEV __device__ __forceinline__
#define LS_ANY(x) __any_sync(0xffffffffu, (x))
#define LS_BALLOT(x) __ballot_sync(0xffffffffu, (x))
#define LS_SHFL(v, l) __shfl_sync(0xffffffffu, (v), (l))
#define LS_SHFL_XOR(v, o) __shfl_xor_sync(0xffffffffu, (v), (o))
#define LS_SHFL_UP(v, o) __shfl_up_sync(0xffffffffu, (v), (o))
#define LS_SYNCWARP() __syncwarp()
#define LS_WARP 32
#define LS_ATOMIC_MAX(p, v) atomicMax((p), (v))
#define LS_POPC(x) __popc(x)
#define LS_FFS(x) __ffs((int)(x))
#define LS_UMULHI(a, b) __umulhi(a, b)
#define LS_LDG(p) __ldg(p)
#define LS_TID ((int)threadIdx.x)
#define LS_ATOMIC_ADD(p, v) atomicAdd(p, v)
#define LS_D2ULL(x) __double2ull_rz(x)
#define LS_DBITS(x) ((unsigned long long)__double_as_longlong(x))
#endif

namespace mctsb {

#if defined(MCTSB_HOST_EMU)
static unsigned long long g_dbg_iters[4][130], g_dbg_pend[4][130];   // [mode][count] histograms (emulation only)
#define LS_DBG(x) x
#else
#define LS_DBG(x)
#endif

static constexpr int LS_THREADS = 128;
static constexpr int LS_SHARE = 32;                     // candidates one lane may queue per batch
static constexpr int LS_QCAP = 32 * LS_SHARE;           // queue entries per warp

// per-warp shared scratch of the candidate phase
struct WarpBuf { uint16_t *queue; unsigned long long *best; };

// layers of one lane live in shared memory, word-interleaved across the block: no bank conflicts
struct SmemLayers {
    uint32_t *base;   // &smem[tid]
    LS_DEV void put(int j, B81 v) {
        base[(3 * j + 0) * LS_THREADS] = v.a; base[(3 * j + 1) * LS_THREADS] = v.b; base[(3 * j + 2) * LS_THREADS] = v.c;
    }
    LS_DEV B81 get(int j) const {
        return b81(base[(3 * j + 0) * LS_THREADS], base[(3 * j + 1) * LS_THREADS], base[(3 * j + 2) * LS_THREADS]);
    }
};

LS_DEV mctsb_qstate ls_load_qstate(const void *base, uint32_t idx) {
    const uint4 *p = reinterpret_cast<const uint4 *>(base) + 2ull * idx;
    uint4 lo = LS_LDG(p), hi = LS_LDG(p + 1);
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

// draw source of a lane: Philox (id-keyed) or a pre-drawn stream (replay parity mode)
template <bool REPLAY> struct LaneDraws;
template <> struct LaneDraws<false> {
    uint32_t k0, k1, id_lo, id_hi, pos;
    LS_DEV void init(const RolloutParams &p, uint64_t r) {
        uint64_t id = p.first_rollout_id + r;
        k0 = (uint32_t)p.seed; k1 = (uint32_t)(p.seed >> 32); id_lo = (uint32_t)id; id_hi = (uint32_t)(id >> 32); pos = 0;
    }
    // four consecutive draws starting at k (two Philox blocks unless k is block aligned)
    LS_DEV void four(uint32_t k, uint32_t *o) const {
        // must be called by all lanes of the warp (warp vote below)
        Philox4 A = philox4x32_10(k >> 2, id_lo, id_hi, 0u, k0, k1);
        uint32_t sh = k & 3u;
        if (LS_ANY(sh != 0)) {
            Philox4 B = philox4x32_10((k >> 2) + 1u, id_lo, id_hi, 0u, k0, k1);
            uint32_t w[8] = {A.x, A.y, A.z, A.w, B.x, B.y, B.z, B.w};
#pragma unroll
            for (int i = 0; i < 4; i++) {
                uint32_t a0 = w[i], a1 = w[i + 1], a2 = w[i + 2], a3 = w[i + 3];
                o[i] = sh == 0 ? a0 : sh == 1 ? a1 : sh == 2 ? a2 : a3;
            }
        } else { o[0] = A.x; o[1] = A.y; o[2] = A.z; o[3] = A.w; }
    }
    LS_DEV LaneDraws<false> from_lane(int src) const {          // another lane's stream (same seed, its rollout id)
        LaneDraws<false> o; o.k0 = k0; o.k1 = k1; o.id_lo = LS_SHFL(id_lo, src); o.id_hi = LS_SHFL(id_hi, src); o.pos = 0; return o;
    }
    LS_DEV uint32_t at(uint32_t k) const {
        Philox4 A = philox4x32_10(k >> 2, id_lo, id_hi, 0u, k0, k1);
        uint32_t sh = k & 3u;
        return sh == 0 ? A.x : sh == 1 ? A.y : sh == 2 ? A.z : A.w;
    }
};
template <> struct LaneDraws<true> {
    const uint32_t *data; uint32_t len, pos;
    LS_DEV void init(const RolloutParams &p, uint64_t r) {
        data = p.streams + (size_t)r * p.stream_stride; len = p.stream_len; pos = 0;
    }
    LS_DEV void four(uint32_t k, uint32_t *o) const {
#pragma unroll
        for (int i = 0; i < 4; i++) o[i] = data[(k + (uint32_t)i) % len];
    }
    LS_DEV LaneDraws<true> from_lane(int src) const {
        LaneDraws<true> o;
        unsigned long long a = (unsigned long long)(size_t)data;
        a = ((unsigned long long)LS_SHFL((uint32_t)(a >> 32), src) << 32) | (unsigned long long)LS_SHFL((uint32_t)a, src);
        o.data = (const uint32_t *)(size_t)a; o.len = len; o.pos = 0; return o;
    }
    LS_DEV uint32_t at(uint32_t k) const { return data[k % len]; }
};

// candidate k = 2*anchor + orientation: boards with that wall added
LS_DEV void boards_with_wall(const QState &s, int k, B81 &oS, B81 &oE) {
    int anchor = k >> 1, pos = 9 * (anchor >> 3) + (anchor & 7);
    oS = s.openS; oE = s.openE;
    if (k & 1) oE = andn(oE, b81_shifted(0x201u, pos));
    else       oS = andn(oS, b81_shifted(0x3u, pos));
}

// Flood fills of all lanes that have one pending, in one converged loop (a pawn-rooted search
// with early exit, as q_flood_dist).  `go` lanes search from `start` to `player`'s goal row.
LS_DEV int warp_flood(bool go, B81 oS, B81 oE, int start, int player, QWork &wk) {
    B81 r = b81_bit(go ? start : 0);
    const uint32_t ga = player ? MCTSB_GOAL_B_A : 0u, gc = player ? 0u : MCTSB_GOAL_W_C;
    int dist = -1, d = 0;
    bool busy = go;
    if (busy && ((r.a & ga) | (r.c & gc))) { dist = 0; busy = false; }
    if (go) wk.flood_queries++;
    while (LS_ANY(busy)) {
        B81 n = q_flood_step(r, oS, oE);
        d++;
        if (busy) {
            wk.flood_steps++;
            if ((n.a & ga) | (n.c & gc)) { dist = d; busy = false; }
            else if (b81_eq(n, r)) { busy = false; }
            r = n;
        }
    }
    return dist;
}

// ------------------------------------------------------------------------------------------------
template <bool REPLAY>
LS_DEV void lockstep_body(const RolloutParams &p, unsigned long long *work_counter, uint32_t *smem) {
    const int lane = LS_TID & 31, warp = LS_TID >> 5;
    SmemLayers ls; ls.base = smem + LS_TID;
    WarpBuf wb;
    {   // layout: [layers][best slots: 32 x u64 per warp][queues: LS_QCAP x u16 per warp]
        uint32_t *after = smem + MCTSB_MAX_LAYERS * 3 * LS_THREADS;
        wb.best = reinterpret_cast<unsigned long long *>(after) + 32 * warp;
        wb.queue = reinterpret_cast<uint16_t *>(after + 2 * LS_THREADS) + LS_QCAP * warp;
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
    unsigned long long done_rollouts = 0, done_plies = 0;

    auto flush_acc = [&]() {
        if (p.sums && acc_n) {
            LS_ATOMIC_ADD((unsigned long long *)&p.sums[acc_leaf].lo, acc_lo);
            LS_ATOMIC_ADD((unsigned long long *)&p.sums[acc_leaf].hi, acc_hi);
            LS_ATOMIC_ADD((unsigned long long *)&p.sums[acc_leaf].n, acc_n);
        }
        acc_lo = acc_hi = acc_n = 0;
    };
    auto finish = [&](double score, int kind) {
        if (p.outcomes) {
            uint4 v; unsigned long long sb = LS_DBITS(score);
            v.x = (uint32_t)sb; v.y = (uint32_t)(sb >> 32);
            v.z = (uint32_t)(plies & 0xffff) | ((uint32_t)(kind & 0xff) << 16); v.w = d.pos;
            reinterpret_cast<uint4 *>(p.outcomes)[r] = v;
        }
        if (p.finals) { mctsb_qstate f; q_pack(s, f); ls_store_qstate(p.finals, r, f); }
        if (p.sums) {
            if (acc_leaf != leaf) { flush_acc(); acc_leaf = leaf; }
            unsigned long long fx = LS_D2ULL(score * 144115188075855872.0);
            acc_lo += fx & ((1ull << 29) - 1); acc_hi += fx >> 29; acc_n += 1;
        }
        done_rollouts++; done_plies += (unsigned long long)plies;
        have = false;
    };

    for (;;) {
        // ---- P0: lanes without a rollout take the next id (one atomic per warp) --------------------
        {
            unsigned want = LS_BALLOT(!have && !exhausted);
            if (want) {
                unsigned long long base = 0;
                int leader = LS_FFS(want) - 1;
                if (lane == leader) base = LS_ATOMIC_ADD(work_counter, (unsigned long long)LS_POPC(want));
                base = LS_SHFL(base, leader);
                if (!have && !exhausted) {
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

        // ---- P1: terminal test at the top of the ply (Quoridor.cpp:822-824) ------------------------------
        {
            // only while plies remain: after the 50th move the position is evaluated even if that move won
            int w = (have && plies < 50) ? q_winner(s) : 0;
            if (w) finish(w == 1 ? 1.0 : 0.0, w == 1 ? 0 : 1);
        }
        const bool capped = have && plies >= 50;              // ply budget spent -> evaluate (Quoridor.cpp:854)
        const bool playing = have && !capped;
        const int pl = s.turn, en = 1 - pl;

        // ---- P2: shortest-path information, two goal-rooted searches per lane in one loop ---------------
        // side A = the enemy of the mover (layers stored when its cut masks are needed), side B = the mover
        // (records the distance of the pawn and of every legal step target).  Capped lanes: A = White, B = Black.
        uint64_t cheap_h = 0, cheap_v = 0;
        if (playing && s.nwalls[1] != 0) q_cheap_wall_masks(s, cheap_h, cheap_v);
        const int n_pool = popc64(cheap_h) + popc64(cheap_v);
        const bool want_cut = playing && n_pool > 0;
        const uint32_t steps = playing ? q_step_mask(s) : 0u;
        const int qa = capped ? 0 : en, qb = capped ? 1 : pl;
        int sp_a = -1, sp_b = -1;
        uint32_t td0 = 0, td1 = 0, td2 = 0;                    // step target distances, 8 bits per candidate slot
        bool overflow_a = false;
        {
            const B81 goal_a = qa ? b81(MCTSB_GOAL_B_A, 0u, 0u) : b81(0u, 0u, MCTSB_GOAL_W_C);
            const B81 goal_b = qb ? b81(MCTSB_GOAL_B_A, 0u, 0u) : b81(0u, 0u, MCTSB_GOAL_W_C);
            const B81 pawn_a = b81_bit(s.cell[qa]), pawn_b = b81_bit(s.cell[qb]);
            B81 tmask = b81(0u, 0u, 0u);                       // step targets still to be reached by search B
            if (playing) {
#pragma unroll
                for (int i = 0; i < 12; i++) if ((steps >> i) & 1u) tmask = tmask | b81_bit(s.cell[pl] + q_step_delta(i));
            }
            B81 ra = goal_a, rb = goal_b;
            bool busy_a = have, busy_b = have;
            if (busy_a) { wk.flood_queries++; if (b81_any(pawn_a & goal_a)) { sp_a = 0; busy_a = false; } }
            if (busy_b) {
                wk.flood_queries++;
                B81 hit = tmask & goal_b;                      // targets on the goal row have distance 0 (fields start at 0)
                tmask = andn(tmask, hit);
                if (b81_any(pawn_b & goal_b)) sp_b = 0;
                busy_b = sp_b < 0 || b81_any(tmask);
            }
            int j = 0;
            while (LS_ANY(busy_a || busy_b)) {
                j++;
                if (busy_a) {
                    B81 grown = q_flood_step(ra, s.openS, s.openE);
                    B81 layer = andn(grown, ra);
                    wk.flood_steps++;
                    if (want_cut && !overflow_a) { if (j < MCTSB_MAX_LAYERS) ls.put(j, layer); else overflow_a = true; }
                    ra = grown;
                    if (b81_any(layer & pawn_a)) { sp_a = j; busy_a = false; }
                    else if (!b81_any(layer)) busy_a = false;  // walled off: sp stays -1 (never in a legal position)
                }
                if (busy_b) {
                    B81 grown = q_flood_step(rb, s.openS, s.openE);
                    B81 layer = andn(grown, rb);
                    wk.flood_steps++;
                    rb = grown;
                    if (b81_any(layer & pawn_b)) sp_b = j;
                    B81 hit = layer & tmask;
                    if (b81_any(hit)) {
                        tmask = andn(tmask, hit);
#pragma unroll
                        for (int i = 0; i < 12; i++) {
                            if (((steps >> i) & 1u) && b81_test(hit, s.cell[pl] + q_step_delta(i))) {
                                uint32_t f = (uint32_t)j << (8 * (i & 3));
                                if (i < 4) td0 |= f; else if (i < 8) td1 |= f; else td2 |= f;
                            }
                        }
                    }
                    if (!b81_any(layer)) { busy_b = false; }   // exhausted
                    else busy_b = sp_b < 0 || b81_any(tmask);
                }
            }
            // unreached step targets (cannot happen while the pawn has a path) keep distance "unreachable"
            if (b81_any(tmask)) {
#pragma unroll
                for (int i = 0; i < 12; i++) {
                    if (((steps >> i) & 1u) && b81_test(tmask, s.cell[pl] + q_step_delta(i))) {
                        uint32_t f = 0xffu << (8 * (i & 3));
                        if (i < 4) td0 |= f; else if (i < 8) td1 |= f; else td2 |= f;
                    }
                }
            }
        }

        // ---- P3: rollouts that used their 50 plies are scored (evaluate_position, Quoridor.cpp:629-664) ----
        if (capped) finish(q_eval_from(sp_a, sp_b, s.nwalls[0], s.nwalls[1], s.turn), 2);

        // ---- P2b: trace the enemy's shortest paths through the stored layers -> cut masks -----------------
        uint64_t cut_h = ~0ull, cut_v = ~0ull;
        {
            const bool tracing = want_cut && !overflow_a && sp_a >= 0;
            const B81 goal_a = qa ? b81(MCTSB_GOAL_B_A, 0u, 0u) : b81(0u, 0u, MCTSB_GOAL_W_C);
            QTrace t; q_trace_init(t, b81_bit(s.cell[qa]));
            int k = tracing ? sp_a - 1 : -1;
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
        if (playing) {
            // force_playwall, Quoridor.cpp:666-678
            int ow = s.nwalls[pl], ew = s.nwalls[en];
            bool fp = (sp_e <= 1 && sp_p > sp_e) || (sp_e <= 2 && sp_p > sp_e + 1 && ow >= ew) || (sp_e <= 3 && sp_p > sp_e + 2 && ow > ew);
            if (!fp) d.pos++;                                  // the value of this draw is never used (Quoridor.cpp:696)
        }
        {
            const bool wall_branch = playing && s.nwalls[1] != 0;
            base = d.pos;
            uint32_t first_after = base + (n_pool >= 2 ? (uint32_t)n_pool : 0u);
            uint32_t q[4] = {0u, 0u, 0u, 0u};
            uint32_t head = 0;
            if (LS_ANY(wall_branch && n_pool >= 2)) { if (wall_branch && n_pool >= 2) head = d.at(base); }
            d.four(wall_branch ? first_after : base, q);       // executed by every lane: it votes inside
            if (playing) {
                if (wall_branch) {
                    if (n_pool >= 2) j0 = (int)LS_UMULHI(head, (uint32_t)n_pool);
                    d.pos = first_after + 1u;
                    if (q[0] < MCTSB_THR_0_10) mode = 1;
                    else { d.pos++; mode = q[1] < MCTSB_THR_0_75 ? 2 : 3; }
                    uint32_t off = d.pos - first_after;        // 1 or 2
                    u_step = off == 1 ? q[1] : q[2]; u_rand = off == 1 ? q[2] : q[3];
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
            uint64_t pend_h = 0, pend_v = 0;
            if (mode == 1 || mode == 2) { pend_h = cheap_h & cut_h; pend_v = cheap_v & cut_v; }
            int head_k = mode == 3 ? q_pool_select(cheap_h, cheap_v, j0) : -1;   // first legal: the head of the shuffle
            const uint32_t pk_cells = (uint32_t)(uint8_t)s.cell[0] | ((uint32_t)(uint8_t)s.cell[1] << 8) |
                                      ((uint32_t)(uint8_t)s.turn << 16) | ((uint32_t)mode << 24);
            const uint32_t pk_sp = (uint32_t)(sp_e & 0xff) | ((uint32_t)(sp_p & 0xff) << 8) | ((uint32_t)n_pool << 16) | ((uint32_t)j0 << 24);
            unsigned long long *slot = wb.best + lane;
            *slot = 0ull;
            LS_DBG(if (playing) g_dbg_pend[mode][popc64(pend_h) + popc64(pend_v)]++;)
            while (LS_ANY((pend_h | pend_v) != 0ull || head_k >= 0)) {
                // -- publish: every lane queues up to LS_SHARE candidates, offsets by an exclusive warp scan
                int cnt = head_k >= 0 ? 1 : popc64(pend_h) + popc64(pend_v);
                int mine = cnt < LS_SHARE ? cnt : LS_SHARE;
                int off = mine;
#pragma unroll
                for (int o = 1; o < LS_WARP; o <<= 1) { int t = LS_SHFL_UP(off, o); if (lane >= o) off += t; }
                const int total = LS_SHFL(off, LS_WARP - 1);
                off -= mine;
                for (int i = 0; LS_ANY(i < mine); i++) {
                    if (i < mine) {
                        int k = head_k;
                        if (k < 0) { k = q_mask_first(pend_h, pend_v); q_mask_clear(pend_h, pend_v, k); }
                        wb.queue[off + i] = (uint16_t)((lane << 8) | k);
                    }
                }
                head_k = -1;
                LS_SYNCWARP();
                // -- work: 32 tasks per trip
                for (int t0 = 0; t0 < total; t0 += LS_WARP) {
                    const int t = t0 + lane;
                    const bool valid = t < total;
                    const uint32_t ent = valid ? (uint32_t)wb.queue[t] : 0u;
                    const int owner = (int)(ent >> 8), k = (int)(ent & 0xffu);
                    B81 oS = b81(LS_SHFL(s.openS.a, owner), LS_SHFL(s.openS.b, owner), LS_SHFL(s.openS.c, owner));
                    B81 oE = b81(LS_SHFL(s.openE.a, owner), LS_SHFL(s.openE.b, owner), LS_SHFL(s.openE.c, owner));
                    const uint32_t oc = LS_SHFL(pk_cells, owner), osp = LS_SHFL(pk_sp, owner);
                    const uint64_t o_ch = LS_SHFL(cheap_h, owner), o_cv = LS_SHFL(cheap_v, owner);
                    const uint32_t o_base = LS_SHFL(base, owner);
                    const LaneDraws<REPLAY> o_draws = d.from_lane(owner);
                    const int o_turn = (int)((oc >> 16) & 0xffu), o_mode = (int)(oc >> 24);
                    const int o_sp_e = (int)(int8_t)(osp & 0xffu), o_sp_p = (int)(int8_t)((osp >> 8) & 0xffu);
                    const int o_n = (int)((osp >> 16) & 0xffu), o_j0 = (int)(osp >> 24);
                    {   // the candidate wall on the owner's boards
                        int anchor = k >> 1, pos = 9 * (anchor >> 3) + (anchor & 7);
                        if (k & 1) oE = andn(oE, b81_shifted(0x201u, pos)); else oS = andn(oS, b81_shifted(0x3u, pos));
                    }
                    const int p1 = o_mode == 3 ? 0 : 1 - o_turn, p2 = o_mode == 3 ? 1 : o_turn;
                    const int c1 = (int)((p1 ? (oc >> 8) : oc) & 0xffu), c2 = (int)((p2 ? (oc >> 8) : oc) & 0xffu);
                    const int d1 = warp_flood(valid, oS, oE, c1, p1, wk);
                    const bool second = valid && d1 >= 0 && (o_mode == 3 || d1 - o_sp_e > 0);
                    const int d2 = warp_flood(second, oS, oE, c2, p2, wk);
                    int gain = 0;                              // > 0: acceptable
                    if (second && d2 >= 0) {
                        if (o_mode == 3) gain = 1;
                        else { int ee = d1 - o_sp_e, oe = d2 - o_sp_p; if (ee > oe) gain = o_mode == 1 ? ee - oe : 1; }
                    }
                    if (gain > 0) {
                        unsigned long long ord = (unsigned long long)k;
                        if (o_mode != 3 && o_n >= 2) {
                            int rk = q_pool_rank(o_ch, o_cv, k);
                            if (rk != o_j0) {
                                uint32_t key = o_draws.at(o_base + 1u + (uint32_t)(rk < o_j0 ? rk : rk - 1));
                                ord |= ((unsigned long long)key + 1ull) << 7;
                            }
                        }
                        unsigned long long score = ((unsigned long long)gain << 40) | (~ord & ((1ull << 40) - 1ull));
                        LS_ATOMIC_MAX(wb.best + owner, score);
                    }
                }
                LS_SYNCWARP();
            }
            const unsigned long long sc = *slot;
            if (mode != 0 && sc) chosen = (int)((~sc) & 127ull);
            // first-legal lanes whose head was not legal walk on in shuffled order on their own (rare)
            if (mode == 3 && chosen < 0) { pend_h = cheap_h; pend_v = cheap_v; q_mask_clear(pend_h, pend_v, q_pool_select(cheap_h, cheap_v, j0)); }
            bool searching = mode == 3 && chosen < 0 && (pend_h | pend_v) != 0ull;
            while (LS_ANY(searching)) {
                int k = -1;
                if (searching) {
                    uint64_t bk = ~0ull; uint64_t th = pend_h, tv = pend_v;
                    for (int c = q_mask_first(th, tv); c >= 0; q_mask_clear(th, tv, c), c = q_mask_first(th, tv)) {
                        int rk = q_pool_rank(cheap_h, cheap_v, c);
                        uint64_t key = (uint64_t)c;
                        if (n_pool >= 2 && rk != j0) key |= ((uint64_t)d.at(base + 1u + (uint32_t)(rk < j0 ? rk : rk - 1)) + 1ull) << 7;
                        if (key < bk) { bk = key; k = c; }
                    }
                    q_mask_clear(pend_h, pend_v, k);
                }
                B81 oS, oE;
                boards_with_wall(s, k < 0 ? 0 : k, oS, oE);
                int d1 = warp_flood(searching, oS, oE, s.cell[0], 0, wk);
                int d2 = warp_flood(searching && d1 >= 0, oS, oE, s.cell[1], 1, wk);
                if (searching) {
                    if (d1 >= 0 && d2 >= 0) { chosen = k; searching = false; }
                    else if (!(pend_h | pend_v)) searching = false;
                }
            }
        }

        // ---- P6: pawn step when no wall was played (Quoridor.cpp:791-801) ---------------------------------------
        int target = -1;
        bool no_move = false;
        if (playing && chosen < 0) {
            bool best_step = s.nwalls[en] == 0;
            if (!best_step) { d.pos++; best_step = u_step < MCTSB_THR_0_80; } else { u_rand = u_step; }
            if (best_step) {
                // get_best_step_move: reversed candidate order, first strict minimum (Quoridor.cpp:476-497)
                int mn = 9999999;
#pragma unroll
                for (int i = 11; i >= 0; i--) {
                    if (!((steps >> i) & 1u)) continue;
                    uint32_t w = i < 4 ? td0 : i < 8 ? td1 : td2;
                    int dist = (int)((w >> (8 * (i & 3))) & 0xffu);
                    if (dist != 0xff && dist < mn) { mn = dist; target = s.cell[pl] + q_step_delta(i); }
                }
                if (target < 0) no_move = true;
            } else {
                d.pos++;                                         // rand() (Quoridor.cpp:796)
                int ns = LS_POPC(steps);
                if (ns == 0) no_move = true;
                else {
                    int idx = (int)((u_rand & 0x7fffffffu) % (uint32_t)ns);
#pragma unroll
                    for (int i = 0; i < 12; i++) if ((steps >> i) & 1u) { if (idx == 0) target = s.cell[pl] + q_step_delta(i); idx--; }
                }
            }
        }

        // ---- P7: play -----------------------------------------------------------------------------------------------
        if (playing) {
            if (chosen >= 0) { q_play_wall(s, chosen); plies++; }
            else if (target >= 0) { q_play_step(s, target); plies++; }
            else if (no_move) {
                // the pick returned NULL: the reference prints, play_move fails, the loop breaks and the position is
                // evaluated (Quoridor.cpp:832-849,854)
                finish(q_eval_from(s.turn == 0 ? sp_p : sp_e, s.turn == 0 ? sp_e : sp_p, s.nwalls[0], s.nwalls[1], s.turn), 3);
            }
        }
    }

    flush_acc();
    if (p.counters) {
        unsigned long long a = done_rollouts, b = done_plies, c = wk.flood_steps, e = wk.flood_queries;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += LS_SHFL_XOR(a, o); b += LS_SHFL_XOR(b, o);
            c += LS_SHFL_XOR(c, o); e += LS_SHFL_XOR(e, o);
        }
        if (lane == 0 && a) { LS_ATOMIC_ADD(p.counters + 0, a); LS_ATOMIC_ADD(p.counters + 1, b); LS_ATOMIC_ADD(p.counters + 2, c); LS_ATOMIC_ADD(p.counters + 3, e); }
    }
}


}  // namespace mctsb
