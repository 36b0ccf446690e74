// warpmodel.cpp — DEVELOPMENT TOOL (not product, not test): replays REF-policy rollouts on the host with the
// rule core of csrc/qcore.cuh, 32 rollouts per "warp" in ply lockstep, and prints how many warp-level loop
// trips the kernel's phases need under different schedules.  Used to rank kernel restructurings before
// spending GPU time.  g++ -O2 -std=c++17 -o /tmp/warpmodel tools/warpmodel.cpp
#include "../montecarlotreesearch_b200/csrc/qcore.cuh"
#include <vector>
#include <algorithm>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
using namespace mctsb;

struct Cand { int k; int len1, d1; bool pass; int len2, d2; bool accept; int gain; uint64_t key; int min4e, min4p; };
struct LanePly { int ecc_max; bool both_zero; bool mover_zero; int nw0, nw1;
    bool playing; int sp_e, sp_p; bool want_cut; int n_pool, mode; std::vector<Cand> c; bool wall_played; int plies;
    bool walls_left;   // any side has walls (pool could be non-empty later)
    int dirty;         // walls changed since the previous ply of this lane
};

static int flood_len(B81 oS, B81 oE, int start, int player, int &steps) {
    B81 r = b81_bit(start);
    const uint32_t ga = player ? MCTSB_GOAL_B_A : 0u, gc = player ? 0u : MCTSB_GOAL_W_C;
    steps = 0;
    if ((r.a & ga) | (r.c & gc)) return 0;
    const QFloodMasks fm = q_flood_masks(oS, oE);
    for (int d = 1;; d++) {
        B81 n = q_flood_step(r, fm); steps++;
        if ((n.a & ga) | (n.c & gc)) return d;
        if (b81_eq(n, r)) return -1;
        r = n;
    }
}
// goal-rooted distance field of `player` (no pawns): dist[cell], 99 = unreachable
static void goal_field(const QState &s, int player, int *dist) {
    const B81 goal = player ? b81(MCTSB_GOAL_B_A, 0u, 0u) : b81(0u, 0u, MCTSB_GOAL_W_C);
    const QFloodMasks fm = q_flood_masks(s.openS, s.openE);
    for (int c = 0; c < 81; c++) dist[c] = b81_test(goal, c) ? 0 : 99;
    B81 r = goal;
    for (int d = 1; d < 90; d++) {
        B81 n = q_flood_step(r, fm);
        if (b81_eq(n, r)) break;
        for (int c = 0; c < 81; c++) if (b81_test(n, c) && !b81_test(r, c)) dist[c] = d;
        r = n;
    }
}
static void wall_on(const QState &s, int k, B81 &oS, B81 &oE) {
    int anchor = k >> 1, pos = 9 * (anchor >> 3) + (anchor & 7);
    oS = s.openS; oE = s.openE;
    if (k & 1) oE = andn(oE, b81_shifted(0x201u, pos)); else oS = andn(oS, b81_shifted(0x3u, pos));
}
static int min4(const int *dist, int k) {
    int anchor = k >> 1, pos = 9 * (anchor >> 3) + (anchor & 7);
    return std::min(std::min(dist[pos], dist[pos + 1]), std::min(dist[pos + 9], dist[pos + 10]));
}

// one ply of pick_semirandom_move (same draw order as qcore.cuh q_pick_semirandom) with every flagged candidate measured
static int pick(QState &s, PhiloxDraws &d, LanePly &L) {
    LocalLayers ls;
    const int p = s.turn, e = 1 - p;
    uint64_t cheap_h = 0, cheap_v = 0;
    if (s.nwalls[1] != 0) q_cheap_wall_masks(s, cheap_h, cheap_v);
    const int n = popc64(cheap_h) + popc64(cheap_v);
    QCut cut_e; cut_e.h = cut_e.v = 0; cut_e.sp = 0;
    if (n > 0) { cut_e = q_cut_masks(s, e, ls, nullptr); s.sp[e] = (int8_t)cut_e.sp; }
    const int sp_p = q_sp(s, p, nullptr), sp_e = q_sp(s, e, nullptr);
    { int de[81], dp[81]; goal_field(s, 0, de); goal_field(s, 1, dp); int m = 0; for (int c = 0; c < 81; c++) { if (de[c] < 99) m = std::max(m, de[c]); if (dp[c] < 99) m = std::max(m, dp[c]); } L.ecc_max = m; }
    L.both_zero = s.nwalls[0] == 0 && s.nwalls[1] == 0; L.mover_zero = s.nwalls[p] == 0; L.nw0 = s.nwalls[0]; L.nw1 = s.nwalls[1];
    L.sp_e = sp_e; L.sp_p = sp_p; L.want_cut = n > 0; L.n_pool = n; L.mode = 0;
    if (!q_force_playwall(s, nullptr)) (void)d.next();
    int result = -2;
    if (s.nwalls[1] != 0) {
        const uint32_t base = d.consumed();
        int j0 = 0;
        if (n >= 2) { j0 = (int)mulhi_u32(d.at(base), (uint32_t)n); d.skip((uint32_t)n); }
        int mode;
        if (d.next() < MCTSB_THR_0_10) mode = 1; else if (d.next() < MCTSB_THR_0_75) mode = 2; else mode = 3;
        if (n == 0) mode = 0;
        L.mode = mode;
        if (mode == 1 || mode == 2) {
            int de[81], dp[81]; goal_field(s, e, de); goal_field(s, p, dp);
            uint64_t ch = cheap_h & cut_e.h, cv = cheap_v & cut_e.v;
            for (int k = q_mask_first(ch, cv); k >= 0; q_mask_clear(ch, cv, k), k = q_mask_first(ch, cv)) {
                Cand c; c.k = k; B81 oS, oE; wall_on(s, k, oS, oE);
                c.d1 = flood_len(oS, oE, s.cell[e], e, c.len1);
                c.pass = c.d1 >= 0 && c.d1 - sp_e > 0;
                c.len2 = 0; c.d2 = 0; c.accept = false; c.gain = 0;
                c.min4e = min4(de, k); c.min4p = min4(dp, k);
                if (c.pass) {
                    c.d2 = flood_len(oS, oE, s.cell[p], p, c.len2);
                    if (c.d2 >= 0) { int ee = c.d1 - sp_e, oe = c.d2 - sp_p; if (ee > oe) { c.gain = mode == 1 ? ee - oe : 1; c.accept = true; } }
                }
                c.key = q_shuffle_key(d, base, n, j0, q_pool_rank(cheap_h, cheap_v, k));
                L.c.push_back(c);
            }
            int best = -1, bg = 0; uint64_t bk = 0;
            for (auto &c : L.c) if (c.accept) {
                if (best < 0 || c.gain > bg || (c.gain == bg && c.key < bk)) { best = c.k; bg = c.gain; bk = c.key; }
            }
            if (best >= 0) result = 128 + best;
        } else if (mode == 3) {
            int k = q_pool_select(cheap_h, cheap_v, j0);
            Cand c; c.k = k; B81 oS, oE; wall_on(s, k, oS, oE);
            c.d1 = flood_len(oS, oE, s.cell[0], 0, c.len1); c.pass = c.d1 >= 0; c.len2 = 0; c.accept = false; c.gain = 0; c.key = 0; c.min4e = c.min4p = 0;
            if (c.pass) { c.d2 = flood_len(oS, oE, s.cell[1], 1, c.len2); c.accept = c.d2 >= 0; }
            L.c.push_back(c);
            if (c.accept) result = 128 + k;
            else {   // rare tail: use the library routine on a copy of the draw state
                uint64_t th = cheap_h, tv = cheap_v; q_mask_clear(th, tv, k);
                while (th | tv) {
                    int best = -1; uint64_t best_key = 0; uint64_t uh = th, uv = tv;
                    for (int cc = q_mask_first(uh, uv); cc >= 0; q_mask_clear(uh, uv, cc), cc = q_mask_first(uh, uv)) {
                        uint64_t key = q_shuffle_key(d, base, n, j0, q_pool_rank(cheap_h, cheap_v, cc));
                        if (best < 0 || key < best_key) { best = cc; best_key = key; }
                    }
                    if (q_wall_keeps_paths(s, best, nullptr)) { result = 128 + best; break; }
                    q_mask_clear(th, tv, best);
                }
            }
        }
    }
    if (result != -2) return result;
    uint32_t steps = q_step_mask(s);
    if (s.nwalls[e] == 0 || d.next() < MCTSB_THR_0_80) return q_best_step(s, steps, nullptr);
    uint32_t r = d.next() & 0x7fffffffu; int ns = popc32(steps);
    if (ns == 0) return -1;
    int idx = (int)(r % (uint32_t)ns);
    for (int i = 0; i < 12; i++) if ((steps >> i) & 1u) { if (idx == 0) return s.cell[s.turn] + q_step_delta(i); idx--; }
    return -1;
}

struct Acc { double v = 0; void add(double x) { v += x; } };

int main(int argc, char **argv) {
    int nwarps = argc > 1 ? atoi(argv[1]) : 64;
    uint64_t seed = 0x2026101900000001ull;
    mctsb_qstate open; memset(&open, 0, sizeof open); open.wx = 0; open.wy = 4; open.bx = 8; open.by = 4; open.wwalls = open.bwalls = 10;
    // accumulators, per ply-trip of a warp
    double trips = 0;
    double p2_cur = 0, p2_seq = 0, p2_sum_lane = 0, p2_lanes = 0, p2_ma = 0, p2_mb = 0, p2_mn = 0;
    double p2_cached_cur = 0, p2_cached_seq = 0; double xa = 0, xb = 0, xa_l = 0, xb_l = 0, xa_n = 0, xb_n = 0, xtrips_a = 0, xtrips_b = 0, xb_wall = 0, xb_nowall = 0; double st_norm = 0, st_build = 0, st_fastplies = 0, st_allfast = 0, st_lanes_fast = 0, st_lanes = 0; double cls[4] = {0,0,0,0}; double st2_norm = 0, st2_build = 0, st2_allfast = 0;   // searches skipped for lanes whose walls did not change (goal fields valid)
    double tr_cur = 0, tr_need = 0, tr_lane_sum = 0;
    double s1_trips = 0, s1_steps = 0, s1_tasks = 0, s1_lane_steps = 0;
    double s2_trips = 0, s2_steps = 0, s2_tasks = 0, s2_lane_steps = 0;
    double s1e_steps = 0, s2e_steps = 0;           // with early stop at min4
    double g_tasks1 = 0, g_tasks2 = 0, g_rounds = 0, g_s1_steps = 0, g_s2_steps = 0, g_s1_trips = 0, g_s2_trips = 0;   // guided sequential doubling
    double f_steps = 0, f_trips = 0;               // fused stage1+2 (both searches per candidate in one loop)
    double hist_mode[4] = {0, 0, 0, 0}, accept_n = 0, cand_n = 0, pass_n = 0;
    double ply_hist_p2[50] = {0}, ply_hist_s1[50] = {0}, ply_cnt[50] = {0}, ply_pool[50] = {0};
    for (int w = 0; w < nwarps; w++) {
        QState s[32]; PhiloxDraws d[32]; bool have[32]; int plies[32]; int dirty[32]; bool built[32] = {false}; bool valid2[32] = {false};
        for (int l = 0; l < 32; l++) { q_unpack(s[l], open); d[l].init(seed, (uint64_t)w * 32 + l); have[l] = true; plies[l] = 0; dirty[l] = 1; }
        for (int trip = 0;; trip++) {
            bool any = false; for (int l = 0; l < 32; l++) any |= have[l];
            if (!any) break;
            trips++;
            LanePly L[32];
            for (int l = 0; l < 32; l++) {
                L[l].playing = false; L[l].sp_e = L[l].sp_p = 0; L[l].want_cut = false; L[l].mode = 0; L[l].dirty = dirty[l];
                if (!have[l]) continue;
                if (plies[l] >= 50) {   // capped: evaluation needs both paths
                    L[l].sp_e = q_sp(s[l], 0, nullptr); L[l].sp_p = q_sp(s[l], 1, nullptr); L[l].playing = false; have[l] = false;
                    L[l].want_cut = false; L[l].mode = -1; continue;
                }
                if (q_winner(s[l])) { have[l] = false; continue; }
                L[l].playing = true;
                int mv = pick(s[l], d[l], L[l]);
                if (mv < 0) { have[l] = false; continue; }
                dirty[l] = mv >= 81; L[l].wall_played = mv >= 81;
                q_play_encoded(s[l], mv); plies[l]++;
            }
            // ---- P2
            int mx = 0, mxs = 0, cmx = 0, cmxs = 0, xma = 0, xmb = 0;
            for (int l = 0; l < 32; l++) if (L[l].playing || L[l].mode == -1) {
                int a = std::max(L[l].sp_e, 0), b = std::max(L[l].sp_p, 0);
                mx = std::max(mx, std::max(a, b)); mxs = std::max(mxs, a + b); xma = std::max(xma, a); xmb = std::max(xmb, b); p2_sum_lane += a + b; p2_lanes += 2;
                if (L[l].dirty) { cmx = std::max(cmx, std::max(a, b)); cmxs = std::max(cmxs, a + b); }
            }
            p2_ma += xma; p2_mb += xmb; p2_mn += std::min(xma, xmb); p2_cur += mx; p2_seq += mxs; p2_cached_cur += cmx; p2_cached_seq += cmxs;
            if (trip < 50) { ply_hist_p2[trip] += mx; ply_cnt[trip] += 1; }
            {   // static-phase policy S0: fields built at the first ply with both wall counts zero, reused afterwards
                int nm = 0, bm = 0; bool anyslow = false, anylane = false;
                int nm2 = 0, bm2 = 0; bool anyslow2 = false;
                for (int l = 0; l < 32; l++) if (L[l].playing) {
                    anylane = true; st_lanes++;
                    int a = std::max(L[l].sp_e, 0), b = std::max(L[l].sp_p, 0);
                    cls[L[l].both_zero ? 0 : L[l].mover_zero ? 1 : L[l].want_cut ? 2 : 3]++;
                    if (L[l].both_zero) { if (!built[l]) { bm = std::max(bm, L[l].ecc_max); built[l] = true; anyslow = true; } else st_lanes_fast++; }
                    else { nm = std::max(nm, std::max(a, b)); anyslow = true; }
                    // policy S2: dirty tracking, build whenever !want_cut && !valid; want_cut lanes with valid fields only search side A; a wall placement invalidates
                    if (!L[l].want_cut) { if (!valid2[l]) { bm2 = std::max(bm2, L[l].ecc_max); valid2[l] = true; anyslow2 = true; } }
                    else { nm2 = std::max(nm2, valid2[l] ? a : std::max(a, b)); anyslow2 = true; }
                }
                for (int l = 0; l < 32; l++) if (L[l].playing && dirty[l] && L[l].want_cut) valid2[l] = false;   // (dirty[] already holds THIS ply's move kind)
                st_norm += std::max(nm, 0); st_build += std::max(0, bm - nm); if (anylane && !anyslow) st_allfast++;
                st2_norm += nm2; st2_build += std::max(0, bm2 - nm2); if (anylane && !anyslow2) st2_allfast++;
            }
            {   // split searches with sp known incrementally: A only for mode 1/2 lanes, B only for stepping lanes
                int ma = 0, mb = 0; bool wallply = false;
                for (int l = 0; l < 32; l++) if (L[l].playing) {
                    if (L[l].mode == 1 || L[l].mode == 2) { ma = std::max(ma, L[l].sp_e); xa_l += L[l].sp_e; xa_n++; wallply = true; }
                    if (!L[l].wall_played) { mb = std::max(mb, L[l].sp_p + 1); xb_l += L[l].sp_p + 1; xb_n++; }
                }
                xa += ma; xb += mb; if (ma) xtrips_a++; if (mb) xtrips_b++; if (wallply) xb_wall += mb; else xb_nowall += mb;
            }
            // ---- trace
            int tmx = 0, tneed = 0;
            for (int l = 0; l < 32; l++) if (L[l].playing && L[l].want_cut) { tmx = std::max(tmx, L[l].sp_e); tr_lane_sum += L[l].sp_e; if (L[l].mode == 1 || L[l].mode == 2) tneed = std::max(tneed, L[l].sp_e); }
            tr_cur += tmx; tr_need += tneed;
            // ---- stage 1 / 2 current schedule: tasks in lane order, 32 per trip
            std::vector<Cand> q;
            for (int l = 0; l < 32; l++) { if (L[l].playing) { hist_mode[L[l].mode]++; if (trip < 50) ply_pool[trip] += L[l].c.size(); } for (auto &c : L[l].c) q.push_back(c); }
            s1_tasks += q.size();
            int st1 = 0;
            for (size_t t0 = 0; t0 < q.size(); t0 += 32) {
                int m = 0, me = 0, mf = 0;
                for (size_t t = t0; t < std::min(q.size(), t0 + 32); t++) {
                    m = std::max(m, q[t].len1); s1_lane_steps += q[t].len1;
                    int e1 = q[t].d1 >= 0 ? std::max(1, q[t].len1 - std::min(q[t].min4e, q[t].len1)) : q[t].len1; me = std::max(me, e1);
                    mf = std::max(mf, std::max(q[t].len1, q[t].len2));
                }
                s1_steps += m; s1_trips++; s1e_steps += me; st1 += m; f_steps += mf; f_trips++;
            }
            if (trip < 50) ply_hist_s1[trip] += st1;
            std::vector<Cand> q2; for (auto &c : q) { cand_n++; if (c.pass) { q2.push_back(c); pass_n++; } if (c.accept) accept_n++; }
            s2_tasks += q2.size();
            for (size_t t0 = 0; t0 < q2.size(); t0 += 32) {
                int m = 0, me = 0;
                for (size_t t = t0; t < std::min(q2.size(), t0 + 32); t++) {
                    m = std::max(m, q2[t].len2); s2_lane_steps += q2[t].len2;
                    int e2 = q2[t].d2 >= 0 ? std::max(1, q2[t].len2 - std::min(q2[t].min4p, q2[t].len2)) : q2[t].len2; me = std::max(me, e2);
                }
                s2_steps += m; s2_trips++; s2e_steps += me;
            }
            // ---- guided lanes sequential in key order with doubling batches (1,2,4,..), best lanes: everything in round 0
            {
                std::vector<std::vector<Cand>> lc(32);
                for (int l = 0; l < 32; l++) { lc[l] = L[l].c; if (L[l].mode == 2) std::sort(lc[l].begin(), lc[l].end(), [](const Cand &a, const Cand &b) { return a.key < b.key; }); }
                size_t pos[32] = {0}; bool done[32];
                for (int l = 0; l < 32; l++) done[l] = lc[l].empty();
                for (int round = 0;; round++) {
                    std::vector<Cand> rq;
                    int batch = 1 << std::min(round, 4);
                    for (int l = 0; l < 32; l++) if (!done[l]) {
                        size_t take = L[l].mode == 2 ? (size_t)batch : lc[l].size();
                        bool acc = false;
                        for (size_t i = 0; i < take && pos[l] < lc[l].size(); i++, pos[l]++) { rq.push_back(lc[l][pos[l]]); acc |= lc[l][pos[l]].accept; }
                        if (pos[l] >= lc[l].size() || (L[l].mode == 2 && acc)) done[l] = true;
                    }
                    if (rq.empty()) break;
                    g_rounds++; g_tasks1 += rq.size();
                    for (size_t t0 = 0; t0 < rq.size(); t0 += 32) { int m = 0; for (size_t t = t0; t < std::min(rq.size(), t0 + 32); t++) m = std::max(m, rq[t].len1); g_s1_steps += m; g_s1_trips++; }
                    std::vector<Cand> r2; for (auto &c : rq) if (c.pass) r2.push_back(c);
                    g_tasks2 += r2.size();
                    for (size_t t0 = 0; t0 < r2.size(); t0 += 32) { int m = 0; for (size_t t = t0; t < std::min(r2.size(), t0 + 32); t++) m = std::max(m, r2[t].len2); g_s2_steps += m; g_s2_trips++; }
                }
            }
        }
    }
    double R = nwarps * 32.0;
    printf("warps %d, ply-trips per warp %.2f\n", nwarps, trips / nwarps);
    printf("per ROLLOUT (warp-level trips / 32 lanes):\n");
    printf("P2 iterations: current(max of max) %.2f | sequential(max of sum) %.2f | mean lane sum %.2f  [kernel measured 32.27]\n", p2_cur / R, p2_seq / R, p2_sum_lane / R / 32 * 1);
    printf("P2 separate loops: A %.2f + B %.2f single-step trips; both-loop min(ma,mb) %.2f then single-loop rest %.2f\n", p2_ma / R, p2_mb / R, p2_mn / R, (p2_cur - p2_mn) / R);
    printf("P2 with goal-field caching (only lanes whose walls changed): max %.2f | seq %.2f\n", p2_cached_cur / R, p2_cached_seq / R);
    printf("static policy S0 (both zero): normal iters %.2f + extra build iters %.2f; all-fast ply-trips per warp %.2f; fast lane share %.3f\n", st_norm / R, st_build / R, st_allfast / nwarps, st_lanes_fast / st_lanes);
    printf("policy S2 (valid tracking): normal iters %.2f + build extra %.2f; all-fast ply-trips per warp %.2f\n", st2_norm / R, st2_build / R, st2_allfast / nwarps);
    printf("lane-ply classes: both-zero %.3f mover-zero %.3f want_cut %.3f other %.3f\n", cls[0]/st_lanes, cls[1]/st_lanes, cls[2]/st_lanes, cls[3]/st_lanes);
    printf("split P2: A iters %.2f (trips/warp %.1f, util %.2f) B iters %.2f (trips/warp %.1f, util %.2f; in wall plies %.2f, others %.2f)\n", xa / R, xtrips_a / nwarps, xa_l / (xa * 32), xb / R, xtrips_b / nwarps, xb_l / (xb * 32), xb_wall / R, xb_nowall / R);
    printf("trace iterations: current %.2f | only mode1/2 lanes %.2f  [kernel 18.09]\n", tr_cur / R, tr_need / R);
    printf("stage1: tasks %.1f trips %.2f steps %.2f (lane-useful %.1f => util %.2f) early-stop steps %.2f [kernel trips 8.05 steps 122.3]\n", s1_tasks / R, s1_trips / R, s1_steps / R, s1_lane_steps / R, s1_lane_steps / (s1_steps * 32), s1e_steps / R);
    printf("stage2: tasks %.1f trips %.2f steps %.2f (util %.2f) early-stop steps %.2f [kernel trips 5.53 steps 76.5]\n", s2_tasks / R, s2_trips / R, s2_steps / R, s2_lane_steps / (s2_steps * 32), s2e_steps / R);
    printf("fused s1+s2 double-steps: trips %.2f steps %.2f\n", f_trips / R, f_steps / R);
    printf("guided sequential doubling: rounds %.2f tasks1 %.1f tasks2 %.1f s1 trips %.2f steps %.2f | s2 trips %.2f steps %.2f\n", g_rounds / R, g_tasks1 / R, g_tasks2 / R, g_s1_trips / R, g_s1_steps / R, g_s2_trips / R, g_s2_steps / R);
    printf("modes (lane-plies): none %.0f best %.0f guided %.0f firstlegal %.0f; candidates %.0f pass %.3f accept %.3f\n", hist_mode[0], hist_mode[1], hist_mode[2], hist_mode[3], cand_n, pass_n / cand_n, accept_n / cand_n);
    printf("ply: P2 iters, stage1 steps, flagged cands per lane\n");
    for (int i = 0; i < 50; i += 1) printf("%2d: %5.1f %6.1f %5.2f\n", i, ply_hist_p2[i] / ply_cnt[i], ply_hist_s1[i] / ply_cnt[i], ply_pool[i] / ply_cnt[i] / 32);
    return 0;
}
