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
#include "../../montecarlotreesearch_b200/csrc/tree_body.cuh"
#include <vector>
#include <algorithm>
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

// ---- the device-resident tree (csrc/tree_body.cuh) on emulated 32-lane warps -----------------------------------
// Selection / Expansion / Backpropagation run exactly as the kernels of csrc/mctsb_tree.cu run them (level by
// level, one warp per item); the playouts of a flush are the caller's (the test feeds the same results to the
// host tree and to this one).
#if defined(MCTSB_EMU_WARP32)
struct EmuSlot {
    std::vector<TreeItem> items; std::vector<uint32_t> idx; std::vector<uint8_t> choice, created; std::vector<int32_t> leaf_of;
    uint32_t n_items[TR_MAX_CHUNKS], n_last;
    std::vector<uint32_t> level_items;
};
struct EmuTree {
    TreeCtx t[2]; TreeCtl ctl; EmuSlot slot[2];
    std::vector<mctsb_qstate> state; std::vector<uint64_t> lh, lv; std::vector<double> score; std::vector<unsigned long long> sims;
    std::vector<uint32_t> virt, size; std::vector<int32_t> parent, first_child; std::vector<uint16_t> n_children, n_moves;
    std::vector<uint8_t> move, flags;
    std::vector<uint32_t> last_item, last_stamp; uint32_t flush_no = 0;
    std::vector<double> logtab;
};
struct EmuTreeLaunch { EmuTree *T; int slot; std::vector<TreeItem> work; std::vector<uint32_t> next_base; const double *w; TreeWarpSmem *wsm; };
static void emu_tree_select_thread(void *a) {
    EmuTreeLaunch *L = (EmuTreeLaunch *)a;
    const int warp = simt::tid() >> 5, lane = simt::tid() & 31, nwarps = 4;
    for (size_t i = warp; i < L->work.size(); i += nwarps) tree_select_item(L->T->t[L->slot], L->work[i], lane, L->wsm[warp], L->next_base[i]);
}
static void emu_tree_backprop_thread(void *a) {
    EmuTreeLaunch *L = (EmuTreeLaunch *)a;
    const int warp = simt::tid() >> 5, lane = simt::tid() & 31, nwarps = 4;
    for (size_t i = warp; i < L->next_base.size(); i += nwarps) tree_backprop_chain(L->T->t[L->slot], L->next_base[i], lane, L->w);
}

void *emu_tree_create(const mctsb_qstate *root, uint32_t R, uint32_t K, uint32_t pool_cap, double c, uint32_t logtab_n, uint32_t chunk) {
    if (chunk < 32 || (K + chunk - 1) / chunk > (uint32_t)TR_MAX_CHUNKS) return nullptr;
    EmuTree *T = new EmuTree();
    T->state.resize(pool_cap); T->lh.resize(pool_cap); T->lv.resize(pool_cap); T->score.resize(pool_cap); T->sims.resize(pool_cap);
    T->virt.resize(pool_cap); T->size.resize(pool_cap); T->parent.resize(pool_cap); T->first_child.resize(pool_cap);
    T->n_children.resize(pool_cap); T->n_moves.resize(pool_cap); T->move.resize(pool_cap); T->flags.resize(pool_cap); T->last_item.assign(pool_cap, 0); T->last_stamp.assign(pool_cap, 0);
    T->logtab.resize(logtab_n);
    for (uint32_t m = 0; m < logtab_n; m++) T->logtab[m] = log((double)((unsigned long long)m * R));
    for (int k = 0; k < 2; k++) {
        EmuSlot &S = T->slot[k];
        const uint32_t nch = (K + chunk - 1) / chunk;
        S.items.resize((size_t)chunk * TR_MAX_LEVELS * nch); S.idx.resize((size_t)K * TR_MAX_LEVELS); S.choice.resize(K); S.created.resize(K); S.leaf_of.resize(K);
        memset(S.n_items, 0, sizeof S.n_items); S.n_last = 0; S.level_items.assign((size_t)TR_MAX_CHUNKS * TR_MAX_LEVELS, 0u);
        TreeCtx &t = T->t[k];
        t.pool.state = T->state.data(); t.pool.lh = T->lh.data(); t.pool.lv = T->lv.data(); t.pool.score = T->score.data(); t.pool.sims = T->sims.data();
        t.pool.virt = T->virt.data(); t.pool.size = T->size.data(); t.pool.parent = T->parent.data(); t.pool.first_child = T->first_child.data();
        t.pool.n_children = T->n_children.data(); t.pool.n_moves = T->n_moves.data(); t.pool.move = T->move.data(); t.pool.flags = T->flags.data();
        t.pool.last_item = T->last_item.data(); t.pool.last_stamp = T->last_stamp.data();
        t.ctl = &T->ctl; t.n_items = S.n_items; t.level_items = S.level_items.data(); t.items = S.items.data(); t.items_per_chunk = chunk * (uint32_t)TR_MAX_LEVELS; t.chunk = chunk; t.idx = S.idx.data(); t.choice = S.choice.data();
        t.leaf_of = S.leaf_of.data(); t.created = S.created.data(); t.K = K; t.pool_cap = pool_cap; t.R = R; t.c = c;
        t.logtab = T->logtab.data(); t.logtab_n = logtab_n; t.good_moves = 0u; t.stamp = 0u;
    }
    memset(&T->ctl, 0, sizeof T->ctl);
    QState s; q_unpack(s, *root);
    T->state[0] = *root; T->score[0] = 0; T->sims[0] = 0; T->virt[0] = 0; T->size[0] = 0; T->parent[0] = -1; T->first_child[0] = -1;
    T->n_children[0] = 0; T->n_moves[0] = TR_UNLISTED; T->move[0] = 0xff; T->flags[0] = q_winner(s) ? TR_FLAG_TERMINAL : 0; T->last_stamp[0] = 0;
    T->ctl.pool_top = 1; T->ctl.root = 0;
    return T;
}
void emu_tree_destroy(void *h) { delete (EmuTree *)h; }
void emu_tree_set_good_moves(void *h, int on) { EmuTree *T = (EmuTree *)h; T->t[0].good_moves = T->t[1].good_moves = on ? 1u : 0u; }

// Selection + Expansion of one flush of n descents into flush slot `slot`; leaf_states[d] = position of the leaf
// descent d ended at
int emu_tree_select(void *h, int slot, uint32_t n, mctsb_qstate *leaf_states) {
    EmuTree *T = (EmuTree *)h;
    if (n == 0 || n > T->t[0].K || slot < 0 || slot > 1) return -1;
    EmuSlot &S = T->slot[slot];
    for (uint32_t d = 0; d < n; d++) S.idx[d] = d;
    T->t[slot].stamp = ++T->flush_no;
    const TreeCtx &t = T->t[slot];
    const uint32_t nch = (n + t.chunk - 1) / t.chunk;
    std::vector<uint32_t> cbegin(nch, 0), cend(nch, 1);
    for (uint32_t c = 0; c < nch; c++) {
        TreeItem r; r.node = T->ctl.root; r.start = c * t.chunk; r.count = n - r.start < t.chunk ? n - r.start : t.chunk; r.level = 0; r.n0 = 0; r.pad = 0;
        r.has_prev = c > 0 ? 1u : 0u; r.next = c + 1 < nch ? (c + 1) * t.items_per_chunk + 1u : 0u;
        S.items[(size_t)c * t.items_per_chunk] = r; S.n_items[c] = 1;
    }
    std::fill(S.level_items.begin(), S.level_items.end(), 0u);
    std::vector<TreeWarpSmem> wsm(4);
    memset(wsm.data(), 0xa5, sizeof(TreeWarpSmem) * 4);
    // the waves of the selection kernel: level wave - c of every chunk c <= wave, all in one launch (their order inside
    // the launch is whatever the fibers make of it — the items of a wave must be independent)
    for (uint32_t wave = 0;; wave++) {
        const uint32_t started = wave + 1 < nch ? wave + 1 : nch;
        EmuTreeLaunch L{T, slot, {}, {}, nullptr, wsm.data()};
        for (uint32_t c = 0; c < started; c++)
            for (uint32_t i = cbegin[c]; i < cend[c]; i++) { L.work.push_back(S.items[(size_t)c * t.items_per_chunk + i]); L.next_base.push_back(cend[c]); }
        if (L.work.empty() && wave >= nch) break;
        // reverse order: the later chunks' shallow items run BEFORE the earlier chunks' deep ones
        std::reverse(L.work.begin(), L.work.end()); std::reverse(L.next_base.begin(), L.next_base.end());
        if (!L.work.empty()) simt::launch(1, 128, emu_tree_select_thread, &L);
        for (uint32_t c = 0; c < started; c++) { const uint32_t lvl = wave - c + 1; cbegin[c] = cend[c]; cend[c] += lvl < (uint32_t)TR_MAX_LEVELS ? S.level_items[(size_t)c * TR_MAX_LEVELS + lvl] : 0u; }
    }
    for (uint32_t c = 0; c < nch; c++) S.n_items[c] = cend[c];
    for (uint32_t d = 0; d < n; d++) leaf_states[d] = T->state[S.leaf_of[d]];
    S.n_last = n;
    return (int)T->ctl.error;
}

// Backpropagation of the flush selected last into `slot`: sums[d] = exact result of the R playouts of descent d
int emu_tree_backprop(void *h, int slot, const mctsb_leaf_sum *sums) {
    EmuTree *T = (EmuTree *)h;
    EmuSlot &S = T->slot[slot];
    std::vector<double> w(S.n_last);
    for (uint32_t d = 0; d < S.n_last; d++) w[d] = tr_leaf_sum_to_double(sums[d].lo, sums[d].hi);
    for (uint32_t d = 0; d < S.n_last; d++) tree_backprop_created(T->t[slot], d, w.data());
    const TreeCtx &t = T->t[slot];
    const uint32_t nch = (S.n_last + t.chunk - 1) / t.chunk;
    // one launch: the heads of the chains (a node's items, chunk after chunk), in reversed order for good measure
    EmuTreeLaunch L{T, slot, {}, {}, w.data(), nullptr};
    for (uint32_t c = 0; c < nch; c++)
        for (uint32_t i = 0; i < S.n_items[c]; i++)
            if (!S.items[(size_t)c * t.items_per_chunk + i].has_prev) L.next_base.push_back(c * t.items_per_chunk + i);
    std::reverse(L.next_base.begin(), L.next_base.end());
    simt::launch(1, 128, emu_tree_backprop_thread, &L);
    return (int)T->ctl.error;
}

double emu_leaf_sum_to_double(uint64_t lo, uint64_t hi) { return tr_leaf_sum_to_double(lo, hi); }

// pre-order walk from the root, children in creation order: one row per node
static void emu_tree_walk(const EmuTree *T, int node, int depth, uint64_t cap, uint64_t &n, int32_t *o_depth, int32_t *o_move, uint64_t *o_sims,
                          double *o_score, uint32_t *o_size, uint32_t *o_nchildren, uint32_t *o_virt) {
    if (n < cap) {
        o_depth[n] = depth; o_move[n] = T->move[node] == 0xff ? -1 : T->move[node]; o_sims[n] = T->sims[node]; o_score[n] = T->score[node];
        o_size[n] = T->size[node]; o_nchildren[n] = T->n_children[node]; o_virt[n] = T->virt[node];
    }
    n++;
    for (int i = 0; i < T->n_children[node]; i++)
        emu_tree_walk(T, T->first_child[node] + i, depth + 1, cap, n, o_depth, o_move, o_sims, o_score, o_size, o_nchildren, o_virt);
}
uint64_t emu_tree_dump(void *h, uint64_t cap, int32_t *o_depth, int32_t *o_move, uint64_t *o_sims, double *o_score, uint32_t *o_size,
                       uint32_t *o_nchildren, uint32_t *o_virt) {
    const EmuTree *T = (const EmuTree *)h;
    uint64_t n = 0;
    emu_tree_walk(T, T->ctl.root, 0, cap, n, o_depth, o_move, o_sims, o_score, o_size, o_nchildren, o_virt);
    return n;
}
int emu_tree_advance(void *h, int move_id) {            // the child that carries `move_id` becomes the root; -1 if it was never expanded
    EmuTree *T = (EmuTree *)h;
    const int r = T->ctl.root;
    for (int i = 0; i < T->n_children[r]; i++)
        if (T->move[T->first_child[r] + i] == move_id) { T->ctl.root = T->first_child[r] + i; T->parent[T->ctl.root] = -1; return 0; }
    return -1;
}
void emu_tree_stats(void *h, uint32_t *out4) {
    const EmuTree *T = (const EmuTree *)h;
    out4[0] = T->ctl.pool_top; out4[1] = T->slot[0].n_items[0] + T->slot[1].n_items[0]; out4[2] = T->ctl.error; out4[3] = T->ctl.exact_evals;
}
#endif

}  // extern "C"
