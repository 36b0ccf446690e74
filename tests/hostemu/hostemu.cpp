// hostemu.cpp — TEST-ONLY host compile of the device rule/rollout core (csrc/qcore.cuh, ttt_core.cuh)
// so that the kernel logic can be checked bit-for-bit against the oracle on a machine without a
// GPU.  It is built by tests/hostemu/build.py into tests/hostemu/_hostemu.so, is never linked
// into the product library and is not reachable through the C ABI.
#if defined(MCTSB_EMU_WARP32)
#include "simt.h"
#endif
#include "../../montecarlotreesearch_b200/csrc/qcore.cuh"
#include "../../montecarlotreesearch_b200/csrc/ttt_core.cuh"
#include "../../montecarlotreesearch_b200/csrc/quoridor_lockstep_body.cuh"
#include "../../montecarlotreesearch_b200/csrc/quoridor_uni_lockstep_body.cuh"
#include <vector>
#include <string.h>
using namespace mctsb;

extern "C" {

int emu_q_primitives(const mctsb_qstate *p, int32_t *info, uint64_t *legal4, double *eval, uint64_t *cheap2,
                     int32_t *steps12) {
    QState s; q_unpack(s, *p);
    info[0] = q_winner(s);
    info[1] = q_sp(s, 0, nullptr);
    info[2] = q_sp(s, 1, nullptr);
    info[3] = q_force_playwall(s, nullptr);
    *eval = q_eval(s, nullptr);
    q_legal_moves(s, legal4);
    q_cheap_wall_masks(s, cheap2[0], cheap2[1]);
    uint32_t m = q_step_mask(s);
    for (int i = 0; i < 12; i++) steps12[i] = ((m >> i) & 1u) ? s.cell[s.turn] + q_step_delta(i) : -1;
    return 0;
}

int emu_q_sp_with_wall(const mctsb_qstate *p, int black, int k) { QState s; q_unpack(s, *p); return q_sp_with_wall(s, black, k, nullptr); }
int emu_q_sp_from(const mctsb_qstate *p, int black, int cell) { QState s; q_unpack(s, *p); return q_flood_dist(s.openS, s.openE, cell, black, nullptr); }
int emu_q_best_step(const mctsb_qstate *p) { QState s; q_unpack(s, *p); return q_best_step(s, q_step_mask(s), nullptr); }
int emu_q_play(const mctsb_qstate *p, int id, mctsb_qstate *out) {
    QState s; q_unpack(s, *p);
    bool ok = q_legal_move(s, id);
    if (ok) q_play_id(s, id);
    q_pack(s, *out);
    return ok;
}

// Exactness of the pruning masks: every cheap-legal wall OUTSIDE q_cut_masks(player) must leave sp(player)
// unchanged.  Returns the number of violations; out2 = {candidates, candidates inside the mask}.
int emu_q_cut_check(const mctsb_qstate *p, int player, int64_t *out2) {
    QState s; q_unpack(s, *p);
    uint64_t ch, cv; q_cheap_wall_masks(s, ch, cv);
    LocalLayers ls;
    QCut cut = q_cut_masks(s, player, ls, nullptr);
    int bad = 0;
    for (int k = q_mask_first(ch, cv); k >= 0; q_mask_clear(ch, cv, k), k = q_mask_first(ch, cv)) {
        out2[0]++;
        if (q_mask_has(cut.h, cut.v, k)) { out2[1]++; continue; }
        if (q_sp_with_wall(s, player, k, nullptr) != cut.sp) bad++;
    }
    return bad;
}

// The search-free legality rule: every cheap-legal wall inside q_detour_safe_masks must keep both paths (and the
// one-wall form must agree with the mask form).  Returns the number of violations; out3 = {cheap-legal walls, walls
// the rule settles, cheap-legal walls that are in fact illegal}.
int emu_q_detour_check(const mctsb_qstate *p, int64_t *out3) {
    QState s; q_unpack(s, *p);
    uint64_t ch, cv, sh, sv; q_cheap_wall_masks(s, ch, cv); q_detour_safe_masks(s, sh, sv);
    int bad = 0;
    for (int k = 0; k < 128; k++) {
        const bool safe = q_mask_has(sh, sv, k);
        if (safe != q_detour_safe_one(s.openS, s.openE, k)) bad++;
        if (!q_mask_has(ch, cv, k)) continue;
        out3[0]++;
        const bool legal = q_wall_keeps_paths(s, k, nullptr);
        if (!legal) out3[2]++;
        if (safe) { out3[1]++; if (!legal) bad++; }
    }
    return bad;
}

int emu_q_rollout_streams(const mctsb_qstate *states, int policy, int n, const uint32_t *streams, uint32_t len,
                          uint32_t stride, mctsb_outcome *out, mctsb_qstate *finals, uint64_t *work3) {
    LocalLayers pool;
    for (int i = 0; i < n; i++) {
        QState s; q_unpack(s, states[i]);
        StreamDraws d; d.init(streams + (size_t)i * stride, len);
        int plies = 0, kind = 0; QWork wk = {0, 0, 0};
        double r = policy == MCTSB_POLICY_UNI ? q_rollout_uni(s, d, pool, plies, kind, &wk) : q_rollout_ref(s, d, pool, plies, kind, &wk);
        memset(out + i, 0, sizeof *out);
        out[i].score = r; out[i].plies = (uint16_t)plies; out[i].kind = (uint8_t)kind; out[i].draws = d.consumed();
        if (finals) q_pack(s, finals[i]);
        if (work3) { work3[0] += wk.flood_steps; work3[1] += wk.flood_queries; work3[2] += plies; }
    }
    return 0;
}

int emu_q_rollout_philox(const mctsb_qstate *state, int policy, uint64_t seed, uint64_t first_id, int n,
                         mctsb_outcome *out, mctsb_qstate *finals, uint64_t *work3) {
    LocalLayers pool;
    for (int i = 0; i < n; i++) {
        QState s; q_unpack(s, *state);
        PhiloxDraws d; d.init(seed, first_id + i);
        int plies = 0, kind = 0; QWork wk = {0, 0, 0};
        double r = policy == MCTSB_POLICY_UNI ? q_rollout_uni(s, d, pool, plies, kind, &wk) : q_rollout_ref(s, d, pool, plies, kind, &wk);
        memset(out + i, 0, sizeof *out);
        out[i].score = r; out[i].plies = (uint16_t)plies; out[i].kind = (uint8_t)kind; out[i].draws = d.consumed();
        if (finals) q_pack(s, finals[i]);
        if (work3) { work3[0] += wk.flood_steps; work3[1] += wk.flood_queries; work3[2] += plies; }
    }
    return 0;
}

int emu_t_rollout_philox(uint32_t x, uint32_t o, uint64_t seed, uint64_t first_id, int n, mctsb_outcome *out) {
    for (int i = 0; i < n; i++) {
        PhiloxDraws d; d.init(seed, first_id + i);
        int plies = 0, kind = 0;
        double r = t_rollout(x, o, d, plies, kind);
        memset(out + i, 0, sizeof *out);
        out[i].score = r; out[i].plies = (uint16_t)plies; out[i].kind = (uint8_t)kind; out[i].draws = d.consumed();
    }
    return 0;
}

int emu_t_rollout_streams(const mctsb_tstate *st, int n, const uint32_t *streams, uint32_t len, uint32_t stride, mctsb_outcome *out) {
    for (int i = 0; i < n; i++) {
        StreamDraws d; d.init(streams + (size_t)i * stride, len);
        int plies = 0, kind = 0;
        double r = t_rollout(st[i].x_mask, st[i].o_mask, d, plies, kind);
        memset(out + i, 0, sizeof *out);
        out[i].score = r; out[i].plies = (uint16_t)plies; out[i].kind = (uint8_t)kind; out[i].draws = d.consumed();
    }
    return 0;
}

int emu_t_winner(uint32_t x, uint32_t o) { return t_winner(x & 0x1ff, o & 0x1ff); }


// The production lockstep kernel body on the host: with a one-lane warp (votes and shuffles are the identity),
// or — built with -DMCTSB_EMU_WARP32 — on `blocks` emulated thread blocks of LS_THREADS fibers each
// (tests/hostemu/simt.h: real 32-lane votes, shuffles, shared-memory queues and the global work counter).
#if defined(MCTSB_EMU_WARP32)
struct EmuLaunch { RolloutParams p; unsigned long long *work; std::vector<std::vector<uint32_t>> *smem; bool replay; };
static void emu_thread(void *a) {
    EmuLaunch *L = (EmuLaunch *)a;
    uint32_t *sm = (*L->smem)[simt::block()].data();
    if (L->replay) lockstep_body<true>(L->p, L->work, sm); else lockstep_body<false>(L->p, L->work, sm);
}
#endif

int emu_lockstep(const mctsb_qstate *states, int n_leaves, uint32_t R, uint64_t seed, uint64_t first_id,
                 const uint32_t *streams, uint32_t stream_len, uint32_t stream_stride,
                 mctsb_outcome *out, mctsb_qstate *finals, mctsb_leaf_sum *sums, uint64_t *counters4, int blocks, int lanes) {
    RolloutParams p; memset(&p, 0, sizeof p);
    p.states = states; p.n_leaves = (uint32_t)n_leaves; p.rollouts_per_leaf = R; p.n_rollouts = (uint64_t)n_leaves * R;
    p.seed = seed; p.first_rollout_id = first_id; p.streams = streams; p.stream_len = stream_len; p.stream_stride = stream_stride;
    p.sums = sums; p.outcomes = out; p.finals = finals; p.lanes_per_batch = lanes > 0 ? (uint32_t)lanes : 32u;
    unsigned long long ctr[4] = {0, 0, 0, 0};
    p.counters = ctr;
    if (sums) memset(sums, 0, sizeof(mctsb_leaf_sum) * (size_t)n_leaves);
    unsigned long long work = 0;
#if defined(MCTSB_EMU_WARP32)
    if (blocks < 1) blocks = 1;
    std::vector<std::vector<uint32_t>> smem((size_t)blocks, std::vector<uint32_t>(LS_SMEM_WORDS + 64, 0xdeadbeefu));
    EmuLaunch L; L.p = p; L.work = &work; L.smem = &smem; L.replay = streams != nullptr;
    simt::launch(blocks, LS_THREADS, emu_thread, &L);
#else
    (void)blocks;
    std::vector<uint32_t> smem(LS_SMEM_WORDS + 64);
    if (streams) lockstep_body<true>(p, &work, smem.data());
    else lockstep_body<false>(p, &work, smem.data());
#endif
    if (counters4) for (int i = 0; i < 4; i++) counters4[i] = ctr[i];
    return 0;
}

// The UNI path on the host: phase A = uni_lockstep_body (one-lane warp, or emulated 32-lane warps), then phase B as
// quoridor_uni_walk_kernel does it — q_walk_uni over the hand-over records.  handed_over = rollouts that reached phase B.
#if defined(MCTSB_EMU_WARP32)
struct EmuUniLaunch { RolloutParams p; UniWalk *walk; unsigned long long *work; std::vector<std::vector<uint32_t>> *smem; bool replay; };
static void emu_uni_thread(void *a) {
    EmuUniLaunch *L = (EmuUniLaunch *)a;
    uint32_t *sm = (*L->smem)[simt::block()].data();
    if (L->replay) uni_lockstep_body<true>(L->p, L->walk, L->work, sm); else uni_lockstep_body<false>(L->p, L->walk, L->work, sm);
}
#endif

int emu_uni_lockstep(const mctsb_qstate *states, int n_leaves, uint32_t R, uint64_t seed, uint64_t first_id,
                     const uint32_t *streams, uint32_t stream_len, uint32_t stream_stride,
                     mctsb_outcome *out, mctsb_qstate *finals, mctsb_leaf_sum *sums, uint64_t *counters4, int blocks, int lanes,
                     uint64_t *handed_over) {
    RolloutParams p; memset(&p, 0, sizeof p);
    p.states = states; p.n_leaves = (uint32_t)n_leaves; p.rollouts_per_leaf = R; p.n_rollouts = (uint64_t)n_leaves * R;
    p.seed = seed; p.first_rollout_id = first_id; p.streams = streams; p.stream_len = stream_len; p.stream_stride = stream_stride;
    p.sums = sums; p.outcomes = out; p.finals = finals; p.lanes_per_batch = lanes > 0 ? (uint32_t)lanes : 32u;
    unsigned long long ctr[4] = {0, 0, 0, 0};
    p.counters = ctr;
    if (sums) memset(sums, 0, sizeof(mctsb_leaf_sum) * (size_t)n_leaves);
    std::vector<UniWalk> walk((size_t)p.n_rollouts);
    memset(walk.data(), 0xff, walk.size() * sizeof(UniWalk));
    p.walk = walk.data();
    unsigned long long work = 0;
#if defined(MCTSB_EMU_WARP32)
    if (blocks < 1) blocks = 1;
    std::vector<std::vector<uint32_t>> smem((size_t)blocks, std::vector<uint32_t>(LS_SMEM_WORDS + 64, 0xdeadbeefu));
    EmuUniLaunch L; L.p = p; L.walk = walk.data(); L.work = &work; L.smem = &smem; L.replay = streams != nullptr;
    simt::launch(blocks, LS_THREADS, emu_uni_thread, &L);
#else
    (void)blocks;
    std::vector<uint32_t> smem(LS_SMEM_WORDS + 64);
    if (streams) uni_lockstep_body<true>(p, walk.data(), &work, smem.data());
    else uni_lockstep_body<false>(p, walk.data(), &work, smem.data());
#endif
    uint64_t handed = 0;
    for (uint64_t r = 0; r < p.n_rollouts; r++) {
        const UniWalk &w = walk[(size_t)r];
        if (w.pending == 0u) continue;
        if (w.pending != 1u) return -1;                       // a rollout phase A never wrote
        handed++;
        QState s; q_unpack(s, w.state);
        int plies = (int)w.plies, kind = 0; double score; uint32_t draws;
        if (streams) { StreamDraws d; d.init(streams + (size_t)r * stream_stride, stream_len); d.pos = w.plies; score = q_walk_uni(s, d, plies, kind); draws = d.consumed(); }
        else { PhiloxDraws d; d.init(seed, first_id + r); d.pos = w.plies; score = q_walk_uni(s, d, plies, kind); draws = d.consumed(); }
        if (out) { memset(out + r, 0, sizeof *out); out[r].score = score; out[r].plies = (uint16_t)plies; out[r].kind = (uint8_t)kind; out[r].draws = draws; }
        if (finals) q_pack(s, finals[r]);
        if (sums) {
            size_t leaf = streams ? (size_t)r : (size_t)(r / R);
            unsigned long long fx = (unsigned long long)(score * 144115188075855872.0);
            sums[leaf].lo += fx & ((1ull << 29) - 1); sums[leaf].hi += fx >> 29; sums[leaf].n += 1;
        }
        ctr[0]++; ctr[1] += (unsigned long long)plies;
    }
    if (handed_over) *handed_over = handed;
    if (counters4) for (int i = 0; i < 4; i++) counters4[i] = ctr[i];
    return 0;
}

int emu_warp_width() { return LS_WARP; }

}  // extern "C"
