// quoridor_uni_lockstep_body.cuh — the UNI rollout policy (uniformly random legal move per ply over the
// reference's generate_all_moves order, Quoridor.cpp:594-616,331-392) on the lockstep machinery of the
// REF-policy kernel (quoridor_lockstep_body.cuh).
//
// A UNI playout has two very different halves.  While anybody still holds a wall, every ply needs the full
// legal-move set: up to 128 candidate walls, each legal only if both pawns keep a path (legal_wall with
// check_blocking, Quoridor.cpp:305-326).  125 of ~130 legal moves are walls, so the twenty walls are gone after
// twenty-odd plies — and from then on the game is a random walk of two pawns on a fixed maze, a few hundred plies
// of a dozen instructions each, with a length that varies by a factor of thirty between playouts.
//
//   phase A (this file): one thread = one rollout, the 32 rollouts of a warp advance ply-synchronously, exactly as
//     in the REF kernel.  Per ply: walls with an open three-step detour around one end are legal without a search
//     (q_detour_safe_masks); for the rest, goal-rooted layers + one path walked back for BOTH players give the walls
//     that lie over an edge of that path (q_path_*: every other wall leaves the path, hence the connection), the survivors go
//     to the warp's shared task queue, all 32 lanes flood-fill them, and a wall that closes a pawn in is struck
//     from its owner's legal set with one shared-memory atomicOr.  Then one draw picks the idx-th legal move.
//     A rollout whose two players have no wall left is handed over: its position and ply count go to a record
//     in global memory and the lane is free again.
//   phase B (quoridor_uni_walk_kernel, rollout_kernels.cu): one thread per handed-over rollout walks the
//     pawns to the end of the game — natural control flow, no shared memory; ragged lengths cost nothing there.
// Both phases draw from the same per-rollout stream (one draw per ply, position = ply), so the result is the
// oracle's, bit for bit, wherever the hand-over happens.
#pragma once
#include "quoridor_lockstep_body.cuh"

#if defined(MCTSB_HOST_EMU)
#define LS_ATOMIC_OR(p, v) do { *(p) |= (v); } while (0)
#else
#define LS_ATOMIC_OR(p, v) atomicOr((p), (v))
#endif

namespace mctsb {

// hand-over record of a rollout, 48 bytes (three 16-byte vector accesses)
struct UniWalk {
    mctsb_qstate state;
    uint32_t plies;        // plies played so far = draws consumed so far
    uint32_t pending;      // 1: phase B has to finish this rollout; 0: it ended in phase A (results already written)
    uint32_t pad[2];
};
static_assert(sizeof(UniWalk) == 48, "hand-over record layout");

LS_DEV void uni_store_walk(UniWalk *walk, uint64_t r, const mctsb_qstate &f, uint32_t plies, uint32_t pending) {
    uint4 *q = reinterpret_cast<uint4 *>(walk) + 3ull * r;
    uint4 lo, hi, tail;
    lo.x = (uint32_t)f.h_anchor; lo.y = (uint32_t)(f.h_anchor >> 32); lo.z = (uint32_t)f.v_anchor; lo.w = (uint32_t)(f.v_anchor >> 32);
    hi.x = f.wx | (f.wy << 8) | (f.bx << 16) | ((uint32_t)f.by << 24);
    hi.y = f.wwalls | (f.bwalls << 8) | (f.turn << 16) | ((uint32_t)f.flags << 24);
    hi.z = f.move_counter; hi.w = 0;
    tail.x = plies; tail.y = pending; tail.z = tail.w = 0u;
    q[0] = lo; q[1] = hi; q[2] = tail;
}

template <bool REPLAY>
LS_DEV void uni_lockstep_body(const RolloutParams &p, UniWalk *walk, unsigned long long *work_counter, uint32_t *smem) {
    const int lane = LS_TID & 31, warp = LS_TID >> 5;
    SmemLayers ls;
    WarpBuf wb;
    uint32_t *ill;                                             // this warp's struck-wall words: row w (0..3), column = owner lane
    {
        uint32_t *region = smem + warp * LS_WARP_WORDS;
        ls.base = region + lane;
        wb.qbase = region;
        wb.best = reinterpret_cast<unsigned long long *>(region + LS_WARP_LAYER_WORDS);
        ill = region + LS_WARP_LAYER_WORDS + 2 * 32;           // the REF kernel's ring area
    }

    QState s; LaneDraws<REPLAY> d;
    {
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
        // ---- P0: the warp takes its next 32 rollouts together (as the REF kernel does)
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

        // ---- P1: who does what this ply (q_rollout_uni, qcore.cuh: winner, ply cap, legal set, pick, play)
        const int won = have ? q_winner(s) : 0;
        const bool capped = have && !won && plies >= MCTSB_UNI_MAX_PLIES;
        const bool no_walls = s.nwalls[0] == 0 && s.nwalls[1] == 0;
        const bool handoff = have && !won && !capped && no_walls;
        const bool playing = have && !won && !capped && !no_walls;
        int fin_kind = won ? (won == 1 ? 0 : 1) : (capped ? 4 : -1);
        double fin_score = won == 1 ? 1.0 : (won == 2 ? 0.0 : 0.5);

        uint64_t cheap_h = 0, cheap_v = 0;
        if (playing) q_cheap_wall_masks(s, cheap_h, cheap_v);
        // walls with an open detour around one end cannot close anybody in (q_detour_safe_masks): no search for them
        uint64_t risky_h = 0, risky_v = 0;
        if (playing && (cheap_h | cheap_v) != 0ull) {
            uint64_t safe_h, safe_v;
            q_detour_safe_masks(s, safe_h, safe_v);
            risky_h = cheap_h & ~safe_h; risky_v = cheap_v & ~safe_v;
        }
        const bool need = (risky_h | risky_v) != 0ull;
        const uint32_t steps = playing ? q_step_mask(s) : 0u;

        // ---- P2: per player, goal-rooted layers until the pawn is reached, then ONE path walked back down the layers ->
        //      the walls over its edges.  Every other wall leaves that path, hence the pawn's connection, intact.
        uint64_t cut_h0 = ~0ull, cut_v0 = ~0ull, cut_h1 = ~0ull, cut_v1 = ~0ull;
        {
            const QFloodMasks fm = q_flood_masks(s.openS, s.openE);
            LS_UNROLL1
            for (int pl = 0; pl < 2; pl++) {
                const B81 goal = pl ? b81(MCTSB_GOAL_B_A, 0u, 0u) : b81(0u, 0u, MCTSB_GOAL_W_C);
                const B81 pawn = b81_bit(pl ? s.cell[1] : s.cell[0]);
                B81 ra = goal;
                uint32_t da = (need && !b81_any(pawn & goal)) ? 0xffffffffu : 0u;
                if (need) wk.flood_queries++;
                uint32_t j = 0;
                LS_UNROLL1
                while (LS_ANY((int)da < 0) && j < 81u) {
                    j++;
                    const B81 grown = q_flood_step(ra, fm);
                    const B81 layer = andn(grown, ra);
                    if (j < MCTSB_MAX_LAYERS) ls.put((int)j, layer);
                    da = umin32(da, b81_any(layer & pawn) ? j : 0xffffffffu);
                    ra = grown;
                }
                const int sp = need ? (int)da : -1;                // -1: walled off (never in a legal position)
                if (sp > 0) wk.flood_steps += (uint32_t)sp;
                const bool tracing = need && sp >= 0 && sp < MCTSB_MAX_LAYERS;
#if defined(MCTSB_UNI_FULL_TRACE)                            // A/B switch: the REF kernel's all-shortest-paths trace (fewer candidates, 4x the cost per layer)
                QTrace t; q_trace_init(t, pawn);
#else
                QPathWalk t; q_path_init(t, pawn);                  // one path down the layers: only walls over its edges can close this pawn in
#endif
                int k = tracing ? sp - 1 : -1;
                LS_UNROLL1
                while (LS_ANY(k >= 0)) {
                    if (k >= 0) {
#if defined(MCTSB_UNI_FULL_TRACE)
                        q_trace_step(t, k == 0 ? goal : ls.get(k), s.openS, s.openE);
#else
                        q_path_step(t, k == 0 ? goal : ls.get(k), s.openS, s.openE);
#endif
                        wk.flood_steps++;
                        k--;
                    }
                }
                uint64_t ch = ~0ull, cv = ~0ull;
#if defined(MCTSB_UNI_FULL_TRACE)
                if (tracing) q_trace_to_anchors(t, ch, cv);
#else
                if (tracing) q_path_to_anchors(t, ch, cv);
#endif
                if (pl == 0) { cut_h0 = ch; cut_v0 = cv; } else { cut_h1 = ch; cut_v1 = cv; }
            }
        }

        // ---- P5: the walls that may close somebody in, measured by the whole warp
        uint64_t legal_h = cheap_h, legal_v = cheap_v;
        {
            uint32_t pend0 = 0, pend1 = 0, pend2 = 0, pend3 = 0;
            if (need) {
                const uint64_t ph = risky_h & (cut_h0 | cut_h1), pv = risky_v & (cut_v0 | cut_v1);
                pend0 = (uint32_t)ph; pend1 = (uint32_t)(ph >> 32); pend2 = (uint32_t)pv; pend3 = (uint32_t)(pv >> 32);
            }
            // which player a pending wall has to be measured for: bit set = the wall may change that player's distance
            const uint32_t w00 = (uint32_t)cut_h0, w01 = (uint32_t)(cut_h0 >> 32), w02 = (uint32_t)cut_v0, w03 = (uint32_t)(cut_v0 >> 32);
            const uint32_t b00 = (uint32_t)cut_h1, b01 = (uint32_t)(cut_h1 >> 32), b02 = (uint32_t)cut_v1, b03 = (uint32_t)(cut_v1 >> 32);
            const uint32_t pk_cells = (uint32_t)(uint8_t)s.cell[0] | ((uint32_t)(uint8_t)s.cell[1] << 8);
            ill[0 * 32 + lane] = 0u; ill[1 * 32 + lane] = 0u; ill[2 * 32 + lane] = 0u; ill[3 * 32 + lane] = 0u;
            LS_SYNCWARP();                                     // every lane is done with its layers: the queue may overwrite them
            LS_UNROLL1
            while (LS_ANY((pend0 | pend1 | pend2 | pend3) != 0u)) {
                int c0 = LS_POPC(pend0), c1 = LS_POPC(pend1), c2 = LS_POPC(pend2), c3 = LS_POPC(pend3);
                c0 = c0 < LS_SHARE / 4 ? c0 : LS_SHARE / 4; c1 = c1 < LS_SHARE / 4 ? c1 : LS_SHARE / 4;
                c2 = c2 < LS_SHARE / 4 ? c2 : LS_SHARE / 4; c3 = c3 < LS_SHARE / 4 ? c3 : LS_SHARE / 4;
                const int mine = c0 + c1 + c2 + c3;
                int off = mine;
                LS_UNROLL1
                for (int o = 1; o < LS_WARP; o <<= 1) { int t = LS_SHFL_UP(off, o); if (lane >= o) off += t; }
                const int total = LS_SHFL(off, LS_WARP - 1);
                off -= mine;
                {
                    uint32_t *qp = wb.qbase + off;
                    // entry = owner lane | k << 5 | (measure for White) << 12 | (measure for Black) << 13
#define LS_PUBLISH_WORD(PEND, CNT, KBASE, WCUT, BCUT)                                            \
                    {                                                                            \
                        uint32_t m = PEND; int c = CNT; const uint32_t e0 = (uint32_t)lane | ((uint32_t)(KBASE) << 5); \
                        LS_UNROLL1                                                               \
                        while (LS_ANY(c > 0)) {                                                  \
                            if (c > 0) {                                                         \
                                const uint32_t bit = (uint32_t)ctz32(m);                         \
                                *qp++ = e0 | (bit << 6) | ((((WCUT) >> bit) & 1u) << 12) | ((((BCUT) >> bit) & 1u) << 13); \
                                m &= m - 1u; c--;                                                \
                            }                                                                    \
                        }                                                                        \
                        PEND = m;                                                                \
                    }
                    LS_PUBLISH_WORD(pend0, c0, 0, w00, b00)
                    LS_PUBLISH_WORD(pend1, c1, 64, w01, b01)
                    LS_PUBLISH_WORD(pend2, c2, 1, w02, b02)
                    LS_PUBLISH_WORD(pend3, c3, 65, w03, b03)
#undef LS_PUBLISH_WORD
                }
                LS_SYNCWARP();
                // -- stage 1: the first player concerned (White if the wall may touch White's paths, else Black).  A
                //    closed-in pawn strikes the wall; walls that also concern Black go on to stage 2, compacted.
                int passed = 0;
                LS_UNROLL1
                for (int t0 = 0; t0 < total; t0 += LS_WARP) {
                    const int t = t0 + lane;
                    const bool valid = t < total;
                    const uint32_t ent = valid ? wb.queue(t) : 0u;
                    const int owner = (int)(ent & 31u), k = (int)((ent >> 5) & 127u);
                    const bool for_w = (ent >> 12) & 1u, for_b = (ent >> 13) & 1u;
                    B81 oS = b81(LS_SHFL(s.openS.a, owner), LS_SHFL(s.openS.b, owner), LS_SHFL(s.openS.c, owner));
                    B81 oE = b81(LS_SHFL(s.openE.a, owner), LS_SHFL(s.openE.b, owner), LS_SHFL(s.openE.c, owner));
                    const uint32_t oc = LS_SHFL(pk_cells, owner);
                    b81_add_wall(oS, oE, k);
                    const int p1 = for_w ? 0 : 1;
                    const int c1 = (int)((p1 ? (oc >> 8) : oc) & 0xffu);
                    const int d1 = warp_flood(valid, oS, oE, c1, p1, wk);
                    if (valid && d1 < 0) LS_ATOMIC_OR(ill + (2 * (k & 1) + (k >> 6)) * 32 + owner, 1u << ((k >> 1) & 31));
                    const bool pass = valid && d1 >= 0 && for_w && for_b;
                    const unsigned bal = LS_BALLOT(pass);
                    LS_SYNCWARP();
                    if (pass) wb.queue(passed + LS_POPC(bal & ((1u << lane) - 1u))) = ent;
                    passed += LS_POPC(bal);
                }
                LS_SYNCWARP();
                // -- stage 2: Black's path with the wall
                LS_UNROLL1
                for (int t0 = 0; t0 < passed; t0 += LS_WARP) {
                    const int t = t0 + lane;
                    const bool valid = t < passed;
                    const uint32_t ent = valid ? wb.queue(t) : 0u;
                    const int owner = (int)(ent & 31u), k = (int)((ent >> 5) & 127u);
                    B81 oS = b81(LS_SHFL(s.openS.a, owner), LS_SHFL(s.openS.b, owner), LS_SHFL(s.openS.c, owner));
                    B81 oE = b81(LS_SHFL(s.openE.a, owner), LS_SHFL(s.openE.b, owner), LS_SHFL(s.openE.c, owner));
                    const uint32_t oc = LS_SHFL(pk_cells, owner);
                    b81_add_wall(oS, oE, k);
                    const int d2 = warp_flood(valid, oS, oE, (int)((oc >> 8) & 0xffu), 1, wk);
                    if (valid && d2 < 0) LS_ATOMIC_OR(ill + (2 * (k & 1) + (k >> 6)) * 32 + owner, 1u << ((k >> 1) & 31));
                }
                LS_SYNCWARP();
            }
            LS_SYNCWARP();
            const uint64_t ih = (uint64_t)ill[0 * 32 + lane] | ((uint64_t)ill[1 * 32 + lane] << 32);
            const uint64_t iv = (uint64_t)ill[2 * 32 + lane] | ((uint64_t)ill[3 * 32 + lane] << 32);
            legal_h &= ~ih; legal_v &= ~iv;
        }

        // ---- P6: one draw picks the idx-th legal move in generate_all_moves order: steps in reversed candidate
        //      order, then anchors row-major with the horizontal wall first (q_nth_move, qcore.cuh)
        {
            const int ns = LS_POPC(steps);
            const int n = ns + popc64(legal_h) + popc64(legal_v);
            if (playing && n == 0) { fin_kind = 3; fin_score = 0.5; }
            const bool drawing = playing && n > 0;
            uint32_t u = 0;
            if (LS_ANY(drawing)) u = d.at(d.pos);
            if (drawing) {
                d.pos++;
                int idx = (int)LS_UMULHI(u, (uint32_t)n);
                if (idx < ns) {
                    uint32_t m = steps;
                    LS_UNROLL1
                    for (; idx > 0; idx--) m &= ~(1u << LS_HIBIT(m));
                    q_play_step(s, (s.turn ? s.cell[1] : s.cell[0]) + q_step_delta(LS_HIBIT(m)));
                } else {
                    q_play_wall(s, q_pool_select(legal_h, legal_v, idx - ns));
                }
                plies++;
            }
        }

        // ---- P7: finish, or hand over to the walk kernel
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
        }
        if (fin_kind >= 0 || handoff) {
            mctsb_qstate f; q_pack(s, f);
            uni_store_walk(walk, r, f, (uint32_t)plies, handoff ? 1u : 0u);
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
        if (lane == 0 && (a | c)) {
            LS_ATOMIC_ADD(p.counters + 0, (unsigned long long)a); LS_ATOMIC_ADD(p.counters + 1, (unsigned long long)b);
            LS_ATOMIC_ADD(p.counters + 2, (unsigned long long)c); LS_ATOMIC_ADD(p.counters + 3, (unsigned long long)e);
        }
    }
}

}  // namespace mctsb
