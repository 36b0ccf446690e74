// tree_body.cuh — the search tree resident in HBM: Selection, Expansion and Backpropagation of
// MCTS_tree::grow_tree (reference mcts/src/mcts.cpp:182-210) for Quoridor, as warp-level device code.
//
// Why.  The host tree (host/mcts.cpp) costs ~1.2 us per iteration on one CPU thread; in the middle of a game a
// flush of K = 4 096 leaves x 16 playouts keeps a B200 busy for 0.6 ms but the host needs 4.9 ms to select it.
// Here the tree never leaves the device: one launch selects a whole flush, the rollout kernel reads the leaf
// positions where they lie, and one launch back-propagates — the host only queues launches.
//
// Semantics = the host tree's, statistic for statistic (tests: same nodes, same order, same doubles):
//   select      mcts.cpp:161-171  walk down by select_best_child until a node is terminal or not fully expanded
//   UCT         mcts.cpp:104-130  win rate of the side to move + c*sqrt(ln N / n), first strict maximum; simulations
//                                 still in flight count as losses (virtual loss), as in host/mcts.cpp
//   expand      mcts.cpp:37-55    children appear in actions_to_try order (generate_all_moves, Quoridor.cpp:594-616)
//   backprop    mcts.cpp:83-90    score += w, n += R up to the root, in selection order
//
// How K sequential selections run in parallel and still give the sequential result: the decision taken at a
// node depends only on that node's children (statistics + virtual losses) and on how many earlier descents of
// the flush passed through the node.  So the unit of work is an ITEM = (node, the ordered list of descents that
// reach it).  One warp plays the item's descents in order (tr_select_steps: a step is one multiply-add per child,
// a warp maximum and a vote) and hands every child the ordered sub-list of descents it received: the items of the
// next level, all independent.  The flush is cut into chunks of descents staggered by one level (TreeCtx::chunk),
// so the root's chain of K picks and the chains of the levels below it overlap in time.
//
// Exactness.  The ranking pass is float with a proven error budget; whenever more than one child comes within its
// margin of the maximum, candidates with identical statistics are settled by index, the others by a second pass in
// double, and what that cannot separate by the reference's own expression (IEEE division and square root, no
// contraction) — first strict maximum, as the host tree decides.  ln N comes from a table the host fills with ITS libm
// (N is always a multiple of R), so those doubles are the host's bit for bit.  Scores are summed per node in
// selection order, sequentially, as the host does.
#pragma once
#include "quoridor_lockstep_body.cuh"      // the LS_* warp primitives (device / tests/hostemu) and qcore.cuh

#if defined(MCTSB_HOST_EMU)
#include <math.h>
#if defined(MCTSB_EMU_WARP32)
#define TR_REDUCE_MAX(v) simt::reduce_max((int)(v), __LINE__)
#define TR_MATCH_ANY(v) simt::match_any((uint32_t)(v), __LINE__)
#define simt_or_self_maxu(v, site) simt::reduce_maxu((v), (site))
#else
#define TR_REDUCE_MAX(v) ((int)(v))
#define TR_MATCH_ANY(v) 1u
#define simt_or_self_maxu(v, site) (v)
#endif
static inline uint32_t tr_emu_add32(uint32_t *p, uint32_t v) { uint32_t o = *p; *p += v; return o; }
#define TR_ATOMIC_ADD32(p, v) tr_emu_add32((p), (v))
#define TR_ATOMIC_OR32(p, v) do { *(p) |= (v); } while (0)
#define TR_ATOMIC_MAX32(p, v) do { if (*(p) < (v)) *(p) = (v); } while (0)
#define TR_HIBIT64(x) (63 - __builtin_clzll(x))
#define TR_REDUCE_MAXU(v) simt_or_self_maxu((uint32_t)(v), __LINE__)
#define TR_FRSQRT(x) (1.0f / sqrtf(x))
#define TR_FDIV(a, b) ((a) / (b))
#define TR_HILO2D(h, l) tr_emu_hilo2d((h), (l))
static inline double tr_emu_hilo2d(uint32_t h, uint32_t l) { uint64_t b = ((uint64_t)h << 32) | l; double x; memcpy(&x, &b, 8); return x; }
#define TR_DHI(x) tr_emu_dhi(x)
#define TR_DLO(x) tr_emu_dlo(x)
static inline uint32_t tr_emu_dhi(double x) { uint64_t b; memcpy(&b, &x, 8); return (uint32_t)(b >> 32); }
static inline uint32_t tr_emu_dlo(double x) { uint64_t b; memcpy(&b, &x, 8); return (uint32_t)b; }
#define TR_FSQRT(x) sqrtf(x)
#define TR_FMAF(a, b, c) ((a) * (b) + (c))
#define TR_LOGF(x) logf(x)
#define TR_DSQRT(x) sqrt(x)
#define TR_DLOG(x) log(x)
#define TR_F2I(x) tr_emu_f2i(x)
static inline int tr_emu_f2i(float x) { int i; memcpy(&i, &x, 4); return i; }
static inline float tr_emu_i2f(int i) { float x; memcpy(&x, &i, 4); return x; }
#define TR_I2F(x) tr_emu_i2f(x)
#define TR_FENCE()
#else
#define TR_REDUCE_MAX(v) __reduce_max_sync(0xffffffffu, (int)(v))
#define TR_MATCH_ANY(v) __match_any_sync(0xffffffffu, (uint32_t)(v))
#define TR_ATOMIC_ADD32(p, v) atomicAdd((p), (v))
#define TR_ATOMIC_OR32(p, v) atomicOr((p), (v))
#define TR_ATOMIC_MAX32(p, v) atomicMax((p), (v))
#define TR_HIBIT64(x) (63 - __clzll((long long)(x)))
#define TR_REDUCE_MAXU(v) __reduce_max_sync(0xffffffffu, (unsigned)(v))
#define TR_FRSQRT(x) rsqrtf(x)
#define TR_FDIV(a, b) __fdividef((a), (b))
#define TR_HILO2D(h, l) __hiloint2double((int)(h), (int)(l))
#define TR_DHI(x) ((uint32_t)__double2hiint(x))
#define TR_DLO(x) ((uint32_t)__double2loint(x))
#define TR_FSQRT(x) __fsqrt_rn(x)
#define TR_FMAF(a, b, c) __fmaf_rn((a), (b), (c))
#define TR_LOGF(x) logf(x)
#define TR_DSQRT(x) __dsqrt_rn(x)
#define TR_DLOG(x) log(x)
#define TR_F2I(x) __float_as_int(x)
#define TR_I2F(x) __int_as_float(x)
#define TR_FENCE() __threadfence()
#endif

namespace mctsb {

static constexpr float TR_RANK_MARGIN = 3e-6f;
static constexpr double TR_RANK_MARGIN2 = 1e-11;    // second (double) ranking pass: 1000 x its error bound
static constexpr int TR_SLOTS = 7;                  // children per lane: 7 * 32 = 224 >= MCTSB_Q_NUM_MOVES
static constexpr int TR_MAX_LEVELS = 128;           // deepest selection path of one flush
static constexpr int TR_MAX_CHUNKS = 64;            // chunks of one flush (see TreeCtx::chunk)
static constexpr uint16_t TR_UNLISTED = 0xffffu;    // n_moves of a node whose move list has not been asked for yet
enum { TR_FLAG_TERMINAL = 1 };
enum { TR_ERR_POOL = 1, TR_ERR_ITEMS = 2, TR_ERR_DEPTH = 4, TR_ERR_LOG = 8, TR_ERR_COUNT = 16 };   // COUNT: a child with >= 2^31 simulations
//   // TR_ERR_LOG is informational: ln N taken from the device's log()

// Node pool, structure of arrays over node ids: the children of a node are `n_moves` consecutive ids reserved when
// the node is first expanded, so a warp reads the statistics of all children of a node in coalesced runs.
struct TreePool {
    mctsb_qstate *state;              // 32 B packed position
    uint64_t *lh, *lv;                // legal wall anchors (valid once n_moves != TR_UNLISTED)
    double *score;                    // sum of rollout results, accumulated in selection order (mcts.cpp:85)
    unsigned long long *sims;         // number_of_simulations
    uint32_t *virt;                   // simulations in flight (virtual losses)
    uint32_t *size;                   // MCTS_node::size
    int32_t *parent, *first_child;
    uint16_t *n_children, *n_moves;
    uint8_t *move;                    // canonical id of the move that led here
    uint8_t *flags;
    uint32_t *last_item, *last_stamp; // the node's latest item (flat index in its flush slot) and the flush that made it: a node
                                      // that is in several chunks of a flush has its items chained in chunk order
};

// n0 = the node's simulations (finished + in flight) before these descents (set by the parent, unused at the root);
// next = flat index + 1 of the node's item in a later chunk of the same flush (0 = none), has_prev = it is not the first:
// Backpropagation walks a node's items in this order, one warp per chain
struct TreeItem { int32_t node; uint32_t start, count, level; uint32_t n0, next, has_prev, pad; };

struct TreeCtl {                      // lives in device memory
    uint32_t pool_top;                // next free node id
    uint32_t error;                   // TR_ERR_* bits
    int32_t root;
    uint32_t level_begin, level_end;  // item range of the level being processed (select kernel)
    uint32_t depth_max;               // deepest level a flush has reached so far (statistics)
    uint32_t exact_evals;             // selections that needed the exact expression (statistics)
    uint32_t steps, ties_same;        // selection steps, ties settled because all candidates were alike (statistics)
    unsigned long long level_ns[8];   // time spent per level of the selection kernel, summed over flushes (levels >= 7 in the last slot)
};

struct TreeCtx {                      // the pool and the per-flush arrays of ONE flush slot (two slots: a flush may be
    TreePool pool;                    // selected while the playouts of the one before are still running)
    TreeCtl *ctl;
    // A flush is selected in CHUNKS of `chunk` consecutive descents, staggered by one level: wave w of the selection
    // kernel processes level w - c of chunk c for every chunk at once (the root takes its descents chunk by chunk, and
    // while it plays chunk c its children already play chunk c - 1 ...).  A deep, narrow tree — the endgame, where most
    // descents follow one line for twenty levels — then costs about K + depth * chunk steps instead of K * depth.
    // The chunks are independent because a node's simulations in flight are kept by its PARENT (added when the
    // descents are handed down), so an item reads nothing that an item of the same wave writes.
    uint32_t *n_items;                // [TR_MAX_CHUNKS] items of every chunk of this slot's flush (all levels; written when the selection ends)
    uint32_t *level_items;            // [TR_MAX_CHUNKS][TR_MAX_LEVELS] items per chunk and level: a wave appends to the counters of the
                                      // NEXT level while everybody reads the ones of the level just closed — never the same word
    TreeItem *items; uint32_t items_per_chunk;   // chunk c owns items[c * items_per_chunk ...]
    uint32_t chunk;                   // descents per chunk
    uint32_t stamp;                   // number of this flush (never 0)
    uint32_t good_moves;              // 1 = the tree expands generate_good_moves (Quoridor.cpp:518-592) instead of generate_all_moves
    uint32_t *idx;                    // [TR_MAX_LEVELS][K]: idx[l*K + p] = descent number, lists of level l items
    uint8_t *choice;                  // [K] scratch: child picked by the step at list position p
    int32_t *leaf_of;                 // [K] node every descent ended at
    uint8_t *created;                 // [K] 1 = the descent created its leaf (Expansion), 0 = it hit a terminal node
    uint32_t K;                       // capacity of the per-flush arrays
    uint32_t pool_cap;
    uint32_t R;                       // rollouts per leaf
    double c;                         // UCT constant (mcts.h default 1.41)
    const double *logtab; uint32_t logtab_n;     // logtab[m] = log((double)(m * R)) from the host's libm
};

// ---- the reference's UCT expression, term by term (mcts.cpp:104-130 as host/mcts.cpp evaluates it) ----------
LS_NOINLINE double tr_uct_exact(double score, double visits, double virt, bool p1, double c, double log_parent) {
    const double virt_as_loss = p1 ? 0.0 : 1.0, sign = p1 ? 1.0 : -1.0, base = p1 ? 0.0 : 1.0;
    const double s = d_add(score, d_mul(virt_as_loss, virt));
    const double exploit = d_add(base, d_mul(sign, d_div(s, visits)));
    if (!(c > 0)) return exploit;
    return d_add(exploit, d_mul(c, TR_DSQRT(d_div(log_parent, visits))));
}

// Error budget of the ranking pass.  Its value of a child is a = ex + coef * is in float, ex = wins / n (wins = the
// chooser's wins among the finished simulations, n = finished + in flight), is = n^-1/2, coef = c sqrt(ln N); against
// the reference's double expression, relative to (1 + a):
//   ex   : wins, n rounded to float (2 x 6e-8) + reciprocal 1.2e-7 + product 6e-8          <= 3.0e-7
//   is   : n rounded (3e-8) + rsqrt.approx 2.4e-7                                         <= 2.7e-7  (relative)
//   coef : logf 1.2e-7, sqrt halves it + 6e-8, c rounded 6e-8, product 6e-8               <= 2.4e-7  (relative)
//   a    : 3.0e-7 + (2.7e-7 + 2.4e-7 + 6e-8) * coef*is/(1+a) + 6e-8                        <= 0.93e-6
// Two values can therefore swap places only when they are closer than 1.9e-6 * (1 + a); the candidate test keeps
// everything within TR_RANK_MARGIN = 3e-6 * (1 + amax) of the float maximum (the host tree, whose ranking terms are
// rounded from doubles, uses 1e-5: any margin above the error bound selects the same child).

// ln N of a step: N = n0 + j*R; when n0 is a multiple of R (always, in a tree grown with one R) the host's table
// has it at m0 + j
LS_DEV double tr_log_parent(const TreeCtx &t, unsigned long long n0, unsigned long long m0, bool on_grid, uint32_t j) {
    if (on_grid && m0 + j < t.logtab_n) return t.logtab[m0 + j];
    TR_ATOMIC_OR32(&t.ctl->error, (uint32_t)TR_ERR_LOG);
    return TR_DLOG((double)(n0 + (unsigned long long)j * t.R));
}

// exact sum of a leaf -> the double the host computes (mctsb_leaf_sum_to_double): (hi * 2^29 + lo) * 2^-57,
// rounded to nearest even once
LS_DEV double tr_leaf_sum_to_double(unsigned long long lo, unsigned long long hi) {
    // t = hi * 2^29 + lo as a 128-bit integer (th:tl); t < 2^116 for any R < 2^59
    const unsigned long long tl = (hi << 29) + lo, th = (hi >> 35) + (tl < lo ? 1ull : 0ull);
    if ((th | tl) == 0ull) return 0.0;
    const double inv = 1.0 / 144115188075855872.0;     // 2^-57
    const int msb = th ? 64 + TR_HIBIT64(th) : TR_HIBIT64(tl);
    if (msb <= 52) return d_mul((double)tl, inv);
    const int shift = msb - 52;                        // 1 .. 63
    unsigned long long mant = (tl >> shift) | (th ? th << (64 - shift) : 0ull);
    const unsigned long long rem = tl & ((1ull << shift) - 1ull), half = 1ull << (shift - 1);
    if (rem > half || (rem == half && (mant & 1ull))) mant++;
    // mant * 2^shift * 2^-57: powers of two, exact
    const unsigned long long pbits = (unsigned long long)(shift - 57 + 1023) << 52;
    double pw;
#if defined(__CUDA_ARCH__)
    pw = __longlong_as_double((long long)pbits);
#else
    memcpy(&pw, &pbits, 8);
#endif
    return d_mul((double)mant, pw);
}

// a node's position is written by the warp that expands its parent and read, one level later, by another SM of the
// same launch: plain loads (the read-only path of ls_load_qstate is for data that never changes under a kernel)
LS_DEV mctsb_qstate tr_load_qstate(const mctsb_qstate *base, uint32_t idx) {
    const uint4 *p = reinterpret_cast<const uint4 *>(base) + 2ull * idx;
    const uint4 lo = p[0], hi = p[1];
    mctsb_qstate s;
    s.h_anchor = (uint64_t)lo.x | ((uint64_t)lo.y << 32);
    s.v_anchor = (uint64_t)lo.z | ((uint64_t)lo.w << 32);
    s.wx = hi.x & 0xff; s.wy = (hi.x >> 8) & 0xff; s.bx = (hi.x >> 16) & 0xff; s.by = hi.x >> 24;
    s.wwalls = hi.y & 0xff; s.bwalls = (hi.y >> 8) & 0xff; s.turn = (hi.y >> 16) & 0xff; s.flags = hi.y >> 24;
    s.move_counter = hi.z; s.reserved = hi.w;
    return s;
}

// ---- move list of a node, on demand (the node constructor's actions_to_try of mcts.cpp:19): lane l tests the
// walls anchored at l and l+32 in both orientations, votes assemble the two 64-bit masks -----------------------
LS_DEV void tr_list_moves(const QState &s, int lane, uint64_t &lh, uint64_t &lv) {
    uint64_t ch, cv, safe_h, safe_v;
    q_cheap_wall_masks(s, ch, cv);
    q_detour_safe_masks(s, safe_h, safe_v);
    uint32_t bal[4];
    LS_UNROLL1
    for (int j = 0; j < 4; j++) {
        const int a = 32 * (j >> 1) + lane, vertical = j & 1;
        bool ok = ((vertical ? cv : ch) >> a) & 1ull;
        if (ok && !(((vertical ? safe_v : safe_h) >> a) & 1ull)) ok = q_wall_keeps_paths(s, 2 * a + vertical, nullptr);
        bal[j] = LS_BALLOT(ok);
    }
    lh = (uint64_t)bal[0] | ((uint64_t)bal[2] << 32);
    lv = (uint64_t)bal[1] | ((uint64_t)bal[3] << 32);
}

// generate_good_moves by the warp (as quoridor_goodmoves_kernel): lane l measures the walls anchored at l and l + 32 in
// both orientations, lane 0 makes the reference's sequential pass; the ordered list lands in W.list, its length in W.list_n.
// The list is not stored with the node (144 bytes each): a node that is expanded again builds it again.
template <class WS>
LS_DEV int tr_good_moves(QState &s, int lane, WS &W) {
    uint64_t ch, cv;
    q_cheap_wall_masks(s, ch, cv);
    const int p = s.turn, e = 1 - p;
    LS_UNROLL1
    for (int j = 0; j < 4; j++) {
        const int a = 32 * (j >> 1) + lane, k = 2 * a + (j & 1);
        int pe = -1, po = -1;
        if ((((j & 1) ? cv : ch) >> a) & 1ull) { pe = q_sp_with_wall(s, e, k, nullptr); po = q_sp_with_wall(s, p, k, nullptr); }
        W.pe[k] = (int8_t)pe; W.po[k] = (int8_t)po;
    }
    LS_SYNCWARP();
    if (lane == 0) {
        const bool walls = (ch | cv) != 0;
        const int our_path = walls ? q_sp(s, p, nullptr) : 0, enemy_path = walls ? q_sp(s, e, nullptr) : 0;
        W.list_n = q_good_moves_select(s, q_step_mask(s), ch, cv, our_path, enemy_path, W.pe, W.po, W.list);
    }
    LS_SYNCWARP();
    return W.list_n;
}

// ---- Selection + Expansion of one item (all 32 lanes of a warp; W = the warp's shared memory) ------------------
struct TreeWarpSmem {
    double sc[32 * TR_SLOTS], vis[32 * TR_SLOTS];       // per child: score, simulations finished + in flight
    uint32_t vq[32 * TR_SLOTS], cnt[32 * TR_SLOTS];     // simulations in flight, descents handed to the child by this item
    uint32_t real[32 * TR_SLOTS];                       // finished simulations (constant while the item is processed)
    uint32_t vis0[32 * TR_SLOTS];                       // simulations finished + in flight when the item began
    int8_t pe[128], po[128];                            // good-moves trees: both players' path lengths with every candidate wall
    uint8_t list[MCTSB_Q_GOOD_MAX]; int32_t list_n;     // ... and the node's ordered move list
    uint32_t cursor[32 * TR_SLOTS];                     // hand-down: next free place of the child's list
};
#if LS_WARP == 32
// ---- the selection steps of one item: descents ne .. count-1 pick a child each, in order.  NS = slots in use ------
// One step = one float multiply-add per child, a warp maximum, a vote.  Everything the step-to-step chain needs of a
// child lives in registers of its owner lane (child i: lane i % 32, slot i / 32), indexed statically:
//   ex, is       exploitation term as the side choosing here sees it, and n^-1/2 (a missing child: -1 and 0 — its
//                value is -1 and it never wins).  Simulations in flight count as losses for the side choosing, so
//                the numerator of ex is a CONSTANT of the flush: the chooser's wins among the finished simulations
//                (score for the first player, finished - score for the second); only n moves.
//   rcpn, rsqn   1/(n+R) and (n+R)^-1/2, what a pick swaps in: the special-function unit delivered them after the
//                pick before, nobody waits for them
//   nf2          float(n + 2R), converted after the pick before: the operand of the next pair of reciprocals
//   nint         n as an integer (finished + in flight; < 2^31, checked)
// so a pick costs its owner two multiplies, two reciprocal issues, one conversion issue and an add — no memory, no
// double, nothing to wait for.  The exact statistics of shared memory (W.vis, W.vq) are brought up to date from the
// registers only when a tie needs them, and W.cnt once at the end.
struct TreeSteps { uint32_t start, ne, count; unsigned long long n0; int nc; bool p1; };

LS_DEV float tr_rcp(float x) {
#if defined(__CUDA_ARCH__)
    float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#else
    return 1.0f / x;
#endif
}
LS_DEV float tr_rsq(float x) {
#if defined(__CUDA_ARCH__)
    float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#else
    return 1.0f / sqrtf(x);
#endif
}
// double reciprocal and reciprocal square root to a few ulp, without the IEEE slow paths (second ranking pass)
LS_DEV double tr_drcp(double x) {
#if defined(__CUDA_ARCH__)
    double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = __fma_rn(r, __fma_rn(-x, r, 1.0), r);
    r = __fma_rn(r, __fma_rn(-x, r, 1.0), r);
    return r;
#else
    return 1.0 / x;
#endif
}
LS_DEV double tr_drsq(double x) {
#if defined(__CUDA_ARCH__)
    double r; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double hx = 0.5 * x;
    r = __fma_rn(r, __fma_rn(-hx * r, r, 0.5), r);
    r = __fma_rn(r, __fma_rn(-hx * r, r, 0.5), r);
    return r;
#else
    return 1.0 / sqrt(x);
#endif
}

template <int NS>
LS_DEV void tr_select_steps(const TreeCtx &t, TreeWarpSmem &W, const int lane, const TreeSteps st) {
    const uint32_t R = t.R;
    const bool p1 = st.p1;
    const int nc = st.nc;
    float ex[NS], is[NS], rcpn[NS], rsqn[NS], nf2[NS], wf[NS];
    uint32_t nint[NS];
    bool too_big = false;
#pragma unroll
    for (int q = 0; q < NS; q++) {
        const int i = lane + 32 * q;
        ex[q] = -1.0f; is[q] = 0.0f; rcpn[q] = 0.0f; rsqn[q] = 0.0f; nf2[q] = 1.0f; wf[q] = 0.0f; nint[q] = 0u;
        if (i < nc) {
            const double vis = W.vis[i], sc = W.sc[i];
            too_big = too_big || !(vis < 2147483648.0 - 3.0 * (double)R * (double)st.count);
            nint[q] = (uint32_t)vis;
            const uint32_t real = nint[q] - W.vq[i];                 // finished simulations
            W.real[i] = real; W.vis0[i] = nint[q];
            wf[q] = (float)(p1 ? sc : (double)real - sc);            // the chooser's wins among them
            const float nf = (float)nint[q];
            ex[q] = wf[q] * tr_rcp(nf); is[q] = tr_rsq(nf);
            const float nf1 = (float)(nint[q] + R);
            rcpn[q] = tr_rcp(nf1); rsqn[q] = tr_rsq(nf1);
            nf2[q] = (float)(nint[q] + 2u * R);
        }
    }
    if (LS_BALLOT(too_big) != 0u && lane == 0) TR_ATOMIC_OR32(&t.ctl->error, (uint32_t)TR_ERR_COUNT);
    const float cf = (float)t.c;
    float coef_lane = 0.0f;
    double lp_lane = 0.0, cs_lane = 0.0;
    uint32_t my_choice = 0;
    const unsigned long long m0 = st.n0 / R;
    const bool on_grid = m0 * R == st.n0;
    // the steps in blocks of 32: c*sqrt(ln N) of a block's steps is computed by its 32 lanes at once, and a step fetches the
    // coefficient of the NEXT one while its own multiply-adds are in flight
    LS_UNROLL1
    for (uint32_t jb = st.ne; jb < st.count; jb += 32u) {
        {
            const float np = (float)(st.n0 + (unsigned long long)(jb + (uint32_t)lane) * R);
            coef_lane = t.c > 0 ? cf * TR_FSQRT(TR_LOGF(np)) : 0.0f;
            lp_lane = t.c > 0 ? tr_log_parent(t, st.n0, m0, on_grid, jb + (uint32_t)lane) : 0.0;    // the exact ln N, should a tie need it
            cs_lane = t.c > 0 ? t.c * TR_DSQRT(lp_lane) : 0.0;
        }
        const uint32_t nrel = st.count - jb < 32u ? st.count - jb : 32u;
        float coef_next = LS_SHFL(coef_lane, 0);
        LS_UNROLL1
        for (uint32_t rel = 0; rel < nrel; rel++) {
        const uint32_t j = jb + rel;
        int pick;                                          // child index
        if (nc == 1) pick = 0;
        else {
            const float coef = coef_next;
            float a[NS];
#pragma unroll
            for (int q = 0; q < NS; q++) a[q] = TR_FMAF(coef, is[q], ex[q]);
            coef_next = LS_SHFL(coef_lane, (int)((rel + 1u) & 31u));
            float amax = a[0];
#pragma unroll
            for (int q = 1; q < NS; q++) amax = fmaxf(amax, a[q]);
            // every child has simulations (finished or in flight), so the values are finite and >= 0 (or -1 = no
            // child): integer order of the bit patterns = float order
            amax = TR_I2F(TR_REDUCE_MAX(TR_F2I(amax)));
            const float thr = amax - TR_RANK_MARGIN * (1.0f + amax);
            uint32_t cm = 0;                               // this lane's candidate slots
#pragma unroll
            for (int q = 0; q < NS; q++) cm |= a[q] >= thr ? 1u << q : 0u;
            // the candidate with the lowest index (slots of a lane ascend with the index)
            const int mine = cm ? lane + 32 * (LS_FFS(cm) - 1) : (1 << 20);
            const int lowest = -TR_REDUCE_MAX(-mine);
            const bool alone = (cm & (cm - 1u)) == 0u && (cm == 0u || mine == lowest);
            if (LS_BALLOT(!alone) == 0u) pick = lowest;    // one candidate: it is the maximum
            else {
                // Several children within the margin of the maximum.  A candidate's exact statistics come from its owner's
                // registers (n) and from the two constants of the flush in shared memory (score, finished simulations) —
                // nothing to write, nothing to wait for.
                bool settled = false;
                pick = lowest;
                {
                    // this lane's first candidate (the one with the lowest index), and whether its other candidates are like it
                    const int q1 = cm ? LS_FFS(cm) - 1 : 0;
                    uint32_t n1 = 0u;
#pragma unroll
                    for (int q = 0; q < NS; q++) n1 = q == q1 ? nint[q] : n1;
                    const int i1 = cm ? mine : 0;
                    const double sc1 = W.sc[i1]; const uint32_t real1 = W.real[i1];
                    bool alike = true;
#pragma unroll
                    for (int q = 0; q < NS; q++)
                        if (((cm >> q) & 1u) && q != q1) { const int i = lane + 32 * q; alike = alike && nint[q] == n1 && W.sc[i] == sc1 && W.real[i] == real1; }
                    // children with the same statistics have the same value: when all candidates are alike (leaves made
                    // by this flush, all of them R losses in flight) the first one wins without any arithmetic
                    const int ol = lowest & 31;
                    const double s0 = LS_SHFL(sc1, ol); const uint32_t n0c = LS_SHFL(n1, ol), r0 = LS_SHFL(real1, ol);
                    if (LS_BALLOT(cm != 0u && !(alike && sc1 == s0 && n1 == n0c && real1 == r0)) == 0u) {
                        settled = true;
                        if (lane == 0) TR_ATOMIC_ADD32(&t.ctl->ties_same, 1u);
                    } else if (LS_BALLOT((cm & (cm - 1u)) != 0u) == 0u) {           // at most one candidate per lane (nearly always)
                        // second ranking pass, in double: wins / n + c sqrt(ln N) n^-1/2 with reciprocals good to a few ulp
                        // is within 1e-14 (1 + v) of the reference's expression; a child that leads by TR_RANK_MARGIN2
                        // is the maximum
                        const double cs = LS_SHFL(cs_lane, (int)rel);
                        const double n = (double)(cm ? n1 : 1u), wins = p1 ? sc1 : (double)real1 - sc1;
                        const double v = cm ? wins * tr_drcp(n) + cs * tr_drsq(n) : -1.0;
                        // maximum over the warp: the values are >= 0, so the order of the doubles is the order of their bit patterns
                        const uint32_t hi = v >= 0.0 ? TR_DHI(v) : 0u, mhi = TR_REDUCE_MAXU(hi);
                        const uint32_t lo = v >= 0.0 && hi == mhi ? TR_DLO(v) : 0u, mlo = TR_REDUCE_MAXU(lo);
                        const double vmax = TR_HILO2D(mhi, mlo), thr2 = vmax - TR_RANK_MARGIN2 * (1.0 + vmax);
                        const uint32_t close = LS_BALLOT(cm != 0u && v >= thr2);
                        if ((close & (close - 1u)) == 0u) { settled = true; pick = LS_SHFL(mine, LS_FFS(close) - 1); }
                    }
                }
                if (!settled) {
                    // the general way (a lane with several candidates, or values that the double pass cannot separate):
                    // the candidates' statistics go to shared memory and the reference's expression decides — first strict
                    // maximum: maximum of the high words, then of the low words among those, then the lowest index
#pragma unroll
                    for (int q = 0; q < NS; q++)
                        if ((cm >> q) & 1u) { const int i = lane + 32 * q; W.vis[i] = (double)nint[q]; W.vq[i] = nint[q] - W.real[i]; }
                    LS_SYNCWARP();
                    const double lp = LS_SHFL(lp_lane, (int)rel);
                    double ev = -1.0; int bi = 1 << 20;
                    LS_UNROLL1
                    for (uint32_t mq = cm; mq; mq &= mq - 1u) {
                        const int i = lane + 32 * (LS_FFS(mq) - 1);
                        const double v = tr_uct_exact(W.sc[i], W.vis[i], (double)W.vq[i], p1, t.c, lp);
                        if (v > ev) { ev = v; bi = i; }
                    }
                    const bool has = ev >= 0.0;
                    const uint32_t hi = has ? TR_DHI(ev) : 0u, ehi = TR_REDUCE_MAXU(hi);
                    const bool top = has && hi == ehi;
                    const uint32_t lo = top ? TR_DLO(ev) : 0u, elo = TR_REDUCE_MAXU(lo);
                    const int win = -TR_REDUCE_MAX(top && lo == elo ? -bi : -(1 << 20));
                    pick = win < nc ? win : 0;
                    if (lane == 0) TR_ATOMIC_ADD32(&t.ctl->exact_evals, 1u);
                }
            }
        }
        // the picked child carries R more simulations in flight: its owner lane swaps the prepared terms in and starts the
        // reciprocals of the pick after this one
        {
            const bool own = (pick & 31) == lane;
            const int qs = pick >> 5;
#pragma unroll
            for (int q = 0; q < NS; q++)
                if (own && qs == q) {                      // straight-line, predicated: branches cost more than the six instructions
                    ex[q] = wf[q] * rcpn[q]; is[q] = rsqn[q]; rcpn[q] = tr_rcp(nf2[q]); rsqn[q] = tr_rsq(nf2[q]);
                    nint[q] += R; nf2[q] = (float)(nint[q] + 2u * R);
                }
        }
        if ((int)rel == lane) my_choice = (uint32_t)pick;
        (void)j;
        }
        if ((uint32_t)lane < nrel) t.choice[st.start + jb + lane] = (uint8_t)my_choice;      // 32 choices at a time go to memory
    }
    // how many descents every child received (the hand-down reads it)
#pragma unroll
    for (int q = 0; q < NS; q++) {
        const int i = lane + 32 * q;
        if (i < nc) W.cnt[i] = (nint[q] - W.vis0[i]) / R;
    }
}

LS_DEV void tree_select_item(const TreeCtx &t, const TreeItem it, const int lane, TreeWarpSmem &W, const uint32_t next_base) {
    // next_base = where the items of the next level of this item's chunk begin (= end of the current level's)
    const TreePool &P = t.pool;
    uint32_t *cursor = W.cursor;
    if (lane == 0 && it.level + 1 > t.ctl->depth_max) TR_ATOMIC_MAX32(&t.ctl->depth_max, it.level + 1);
    const int X = it.node;
    const uint32_t K = t.K, R = t.R;
    const uint32_t *my_idx = t.idx + (size_t)it.level * K + it.start;       // the descents that reach X, in order
    // --- terminal node: every descent ends here and re-scores it (mcts.cpp:38-40) ---
    if (P.flags[X] & TR_FLAG_TERMINAL) {
        for (uint32_t j = lane; j < it.count; j += 32) { const uint32_t d = my_idx[j]; t.leaf_of[d] = X; t.created[d] = 0; }
        if (lane == 0 && it.level == 0u) P.virt[X] += it.count * R;        // below the root the parent has already counted them
        return;
    }
    QState s;
    q_unpack(s, tr_load_qstate(P.state, (uint32_t)X));
    const bool p1 = s.turn == 0;
    // everything the lanes read of X's header is read here, before any lane writes to it
    int n_moves = P.n_moves[X], first = P.first_child[X], nc = P.n_children[X];
    uint64_t lh = P.lh[X], lv = P.lv[X];
    // N of the first step (every step adds R): the root reads its own counters; below it the parent wrote the number into
    // the item when it handed the descents down — P.virt[X] may already hold the next chunk's descents
    const unsigned long long n0 = it.level == 0u ? P.sims[X] + P.virt[X] : (unsigned long long)it.n0;
    LS_SYNCWARP();
    // --- move list on first use ---
    const uint32_t steps = q_step_mask(s);
    bool have_list = false;
    if (n_moves == TR_UNLISTED) {
        if (t.good_moves) { n_moves = tr_good_moves(s, lane, W); have_list = true; lh = lv = 0ull; }
        else { tr_list_moves(s, lane, lh, lv); n_moves = popc32(steps) + popc64(lh) + popc64(lv); }
        if (lane == 0) { P.lh[X] = lh; P.lv[X] = lv; P.n_moves[X] = (uint16_t)n_moves; }
    }
    if (n_moves == 0) {                                   // no legal move and not terminal: plays out where it stands
        for (uint32_t j = lane; j < it.count; j += 32) { const uint32_t d = my_idx[j]; t.leaf_of[d] = X; t.created[d] = 0; }
        if (lane == 0 && it.level == 0u) P.virt[X] += it.count * R;        // below the root the parent has already counted them
        return;
    }
    // --- the children's ids: n_moves consecutive ones, reserved at the first expansion ---
    if (first < 0) {
        uint32_t f = 0;
        if (lane == 0) f = TR_ATOMIC_ADD32(&t.ctl->pool_top, (uint32_t)n_moves);
        f = LS_SHFL(f, 0);
        if ((unsigned long long)f + (unsigned)n_moves > t.pool_cap) {            // pool full: the descents stop here
            if (lane == 0) TR_ATOMIC_OR32(&t.ctl->error, (uint32_t)TR_ERR_POOL);
            for (uint32_t j = lane; j < it.count; j += 32) { const uint32_t d = my_idx[j]; t.leaf_of[d] = X; t.created[d] = 0; }
            if (lane == 0 && it.level == 0u) P.virt[X] += it.count * R;        // below the root the parent has already counted them
            return;
        }
        first = (int)f;
        if (lane == 0) P.first_child[X] = first;
    }
    // --- children's statistics into shared memory: child i belongs to lane i % 32, slot i / 32 ---
    const double Rd = (double)R;
#pragma unroll
    for (int q = 0; q < TR_SLOTS; q++) {
        const int i = lane + 32 * q;
        W.cnt[i] = 0u;
        if (i < nc) {
            const uint32_t vq = P.virt[first + i];
            W.sc[i] = P.score[first + i]; W.vis[i] = (double)(P.sims[first + i] + vq); W.vq[i] = vq;
        }
    }
    // --- Expansion: while untried moves remain every descent that reaches X makes the next child ---
    const uint32_t untried = (uint32_t)(n_moves - nc);
    const uint32_t ne = it.count < untried ? it.count : untried;
    if (t.good_moves && ne > 0u && !have_list) tr_good_moves(s, lane, W);     // the ordered list, for the children made now
    for (uint32_t e = lane; e < ne; e += 32) {
        const int i = nc + (int)e, cid = first + i;
        int mv;                                             // encoded as q_nth_move returns it: cell, or 128 + wall
        if (t.good_moves) { const int id = W.list[i]; mv = id < 81 ? id : id < 145 ? 128 + 2 * (id - 81) : 128 + 2 * (id - 145) + 1; }
        else mv = q_nth_move(s, steps, lh, lv, i);
        QState ns = s;
        if (mv < 81) q_play_step(ns, mv); else q_play_wall(ns, mv - 128);
        mctsb_qstate packed; q_pack(ns, packed);
        ls_store_qstate(P.state, (uint64_t)cid, packed);
        P.score[cid] = 0.0; P.sims[cid] = 0ull; P.virt[cid] = R; P.size[cid] = 0u;
        P.parent[cid] = X; P.first_child[cid] = -1; P.n_children[cid] = 0; P.n_moves[cid] = TR_UNLISTED;
        P.move[cid] = (uint8_t)q_move_to_id(mv); P.flags[cid] = q_winner(ns) ? TR_FLAG_TERMINAL : 0; P.last_stamp[cid] = 0u;
        const uint32_t d = my_idx[e];
        t.leaf_of[d] = cid; t.created[d] = 1;
    }
#pragma unroll
    for (int q = 0; q < TR_SLOTS; q++) {
        const int i = lane + 32 * q;
        if (i >= nc && i < nc + (int)ne) { W.sc[i] = 0.0; W.vq[i] = R; W.vis[i] = Rd; }    // a new leaf: R simulations in flight, counted as losses
    }
    nc += (int)ne;
    LS_SYNCWARP();
    // --- Selection: the remaining descents pick a child each, in order (loop specialised by the number of slots in use) ---
    if (ne < it.count) {
        const TreeSteps st = {it.start, ne, it.count, n0, nc, p1};
        switch ((nc + 31) >> 5) {
            case 1: tr_select_steps<1>(t, W, lane, st); break;
            case 2: tr_select_steps<2>(t, W, lane, st); break;
            case 3: tr_select_steps<3>(t, W, lane, st); break;
            case 4: tr_select_steps<4>(t, W, lane, st); break;
            case 5: tr_select_steps<5>(t, W, lane, st); break;
            case 6: tr_select_steps<6>(t, W, lane, st); break;
            default: tr_select_steps<TR_SLOTS>(t, W, lane, st); break;
        }
    }
    if (lane == 0) { P.n_children[X] = (uint16_t)nc; if (it.level == 0u) P.virt[X] += it.count * R; TR_ATOMIC_ADD32(&t.ctl->steps, it.count - ne); }
    const uint32_t n_down = it.count - ne;
    if (n_down == 0u) return;
    if (it.level + 1 >= (uint32_t)TR_MAX_LEVELS) {         // deeper than the level buffers: the descents stop at X
        if (lane == 0) TR_ATOMIC_OR32(&t.ctl->error, (uint32_t)TR_ERR_DEPTH);
        for (uint32_t j = ne + lane; j < it.count; j += 32) { const uint32_t d = my_idx[j]; t.leaf_of[d] = X; t.created[d] = 0; }
        return;
    }
    // --- hand the descents down: child i receives the ordered sub-list of the steps that picked it ---
    uint32_t run = it.start + ne, n_new = 0;
    uint32_t off[TR_SLOTS];
    uint32_t cnt[TR_SLOTS];
    LS_SYNCWARP();
#pragma unroll
    for (int q = 0; q < TR_SLOTS; q++) {                   // exclusive prefix of cnt in child order
        cnt[q] = W.cnt[lane + 32 * q];
        if (cnt[q] != 0u) P.virt[first + lane + 32 * q] += cnt[q] * R;      // the children's simulations in flight are kept here, by the parent
        uint32_t v = cnt[q];
        for (int o = 1; o < 32; o <<= 1) { const uint32_t u = LS_SHFL_UP(v, o); if (lane >= o) v += u; }
        off[q] = run + v - cnt[q];
        run += LS_SHFL(v, 31);
        cursor[lane + 32 * q] = off[q];
        n_new += LS_POPC(LS_BALLOT(cnt[q] != 0u));
    }
    LS_SYNCWARP();
    uint32_t *next_idx = t.idx + (size_t)(it.level + 1) * K;
    for (uint32_t jb = ne; jb < it.count; jb += 32) {
        const uint32_t j = jb + lane;
        const bool on = j < it.count;
        const uint32_t ch = on ? t.choice[it.start + j] : 0xffu;
        const uint32_t peers = TR_MATCH_ANY(ch);
        if (on) {
            const uint32_t rank = LS_POPC(peers & ((1u << lane) - 1u));
            next_idx[cursor[ch] + rank] = my_idx[j];
        }
        LS_SYNCWARP();
        if (on && (peers >> lane) <= 1u) cursor[ch] += LS_POPC(peers);      // the highest lane of each group moves its cursor
        LS_SYNCWARP();
    }
    // --- the next level's items ---
    const uint32_t chunk_id = it.start / t.chunk;           // an item's list lies inside its chunk's range of descents
    TreeItem *chunk_items = t.items + (size_t)chunk_id * t.items_per_chunk;
    uint32_t ibase = 0;
    if (lane == 0) ibase = next_base + TR_ATOMIC_ADD32(t.level_items + chunk_id * (uint32_t)TR_MAX_LEVELS + it.level + 1u, n_new);
    ibase = LS_SHFL(ibase, 0);
    if (ibase + n_new > t.items_per_chunk) {                // cannot happen with items_per_chunk >= chunk * TR_MAX_LEVELS; stop here if it does
        if (lane == 0) TR_ATOMIC_OR32(&t.ctl->error, (uint32_t)TR_ERR_ITEMS);
        return;
    }
#pragma unroll
    for (int q = 0; q < TR_SLOTS; q++) {
        const uint32_t has = LS_BALLOT(cnt[q] != 0u);
        if (cnt[q] != 0u) {
            TreeItem ni; ni.node = first + lane + 32 * q; ni.start = off[q]; ni.count = cnt[q]; ni.level = it.level + 1; ni.n0 = W.vis0[lane + 32 * q]; ni.pad = 0u;
            // chain it behind the child's item of an earlier chunk of this flush, if there is one (made in an earlier wave)
            const uint32_t slot_i = ibase + LS_POPC(has & ((1u << lane) - 1u)), flat = chunk_id * t.items_per_chunk + slot_i;
            ni.next = 0u; ni.has_prev = 0u;
            if (P.last_stamp[ni.node] == t.stamp) { t.items[P.last_item[ni.node]].next = flat + 1u; ni.has_prev = 1u; }
            P.last_item[ni.node] = flat; P.last_stamp[ni.node] = t.stamp;
            chunk_items[slot_i] = ni;
        }
        ibase += LS_POPC(has);
    }
}

// ---- Backpropagation of one item: the results of the descents that passed through X are added in selection
// order (mcts.cpp:83-90); w[d] = result of descent d (sum of its R playouts) --------------------------------------
LS_DEV void tree_backprop_item(const TreeCtx &t, const TreeItem it, const int lane, const double *w) {
    const TreePool &P = t.pool;
    const int X = it.node;
    const uint32_t *my_idx = t.idx + (size_t)it.level * t.K + it.start;
    double sum = P.score[X];
    for (uint32_t jb = 0; jb < it.count; jb += 32) {
        const uint32_t j = jb + lane;
        const double wv = j < it.count ? w[my_idx[j]] : 0.0;
        const uint32_t n = it.count - jb < 32u ? it.count - jb : 32u;
        for (uint32_t k = 0; k < n; k++) sum = d_add(sum, LS_SHFL(wv, (int)k));
    }
    if (lane == 0) {
        P.score[X] = sum;
        P.sims[X] += (unsigned long long)it.count * t.R;
        P.virt[X] -= it.count * t.R;
        const bool own = (P.flags[X] & TR_FLAG_TERMINAL) || P.n_moves[X] == 0;   // the events of such an item are X's own leaf events
        if (!own) P.size[X] += it.count;                     // every event came up from a child (mcts.cpp:88)
    }
}
// one warp walks the items of one node, chunk after chunk
LS_DEV void tree_backprop_chain(const TreeCtx &t, uint32_t flat, const int lane, const double *w) {
    for (;;) {
        const TreeItem it = t.items[flat];
        tree_backprop_item(t, it, lane, w);
        LS_SYNCWARP();
        if (it.next == 0u) break;
        flat = it.next - 1u;
    }
}
#endif   // LS_WARP == 32

// Backpropagation, first half: the leaf a descent CREATED takes its first result (0 + w) before any later event
LS_DEV void tree_backprop_created(const TreeCtx &t, uint32_t d, const double *w) {
    if (!t.created[d]) return;
    const int leaf = t.leaf_of[d];
    t.pool.score[leaf] = d_add(0.0, w[d]);
    t.pool.sims[leaf] = t.R;
    t.pool.virt[leaf] -= t.R;
}

}  // namespace mctsb
