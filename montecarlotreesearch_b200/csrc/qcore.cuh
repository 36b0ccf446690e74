// qcore.cuh — Quoridor rules and the reference rollout policy on register-resident bitboards.
//
// One rollout keeps its whole position in registers: two pawn cells, two wall counts, two 64-bit
// wall-anchor masks and two 81-bit "edge open" boards (3 x u32 each).  Shortest paths — the
// reference's per-move BFS with ten heap allocations (reference examples/Quoridor/Quoridor.cpp:
// 46-106) — become a bitboard flood fill: one frontier expansion of the whole 9x9 board is ~27
// integer instructions, and the BFS distance is the number of expansions until the goal row is hit.
//
// Everything here is `__host__ __device__` so that the host game class (host/Quoridor.cpp: move
// generation for the tree) and the kernels share one definition of the rules.  Rollouts themselves
// are only ever executed by the kernels (rollout_kernels.cu); there is no CPU rollout path in the
// product.
//
// Board geometry (identical to the reference, Quoridor.cpp:18,41-42): cell (x,y) = row x, column y,
// bit index 9*x+y.  White starts (0,4), goal row 8; Black starts (8,4), goal row 0.
#pragma once
#include <stdint.h>
#include "../../include/mctsb.h"
#include "philox.cuh"

namespace mctsb {

// ------------------------------------------------------------------------------------------------
// 81-bit board in three 32-bit words: a = cells 0..31, b = 32..63, c = 64..80.
// ------------------------------------------------------------------------------------------------
struct B81 { uint32_t a, b, c; };

MC_HD uint32_t fsl(uint32_t lo, uint32_t hi, int n) {   // high word of (hi:lo) << n, 0 < n < 32
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, n);
#else
    return (hi << n) | (lo >> (32 - n));
#endif
}
MC_HD uint32_t fsr(uint32_t lo, uint32_t hi, int n) {   // low word of (hi:lo) >> n, 0 < n < 32
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, n);
#else
    return (lo >> n) | (hi << (32 - n));
#endif
}
MC_HD int popc32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
MC_HD int hibit32(uint32_t v) {   // v != 0: index of the highest set bit
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int)v);
#else
    return 31 - __builtin_clz(v);
#endif
}
MC_HD int popc64(uint64_t v) { return popc32((uint32_t)v) + popc32((uint32_t)(v >> 32)); }
MC_HD int ctz32(uint32_t v) {   // v != 0
#if defined(__CUDA_ARCH__)
    return __ffs((int)v) - 1;
#else
    return __builtin_ctz(v);
#endif
}

MC_HD B81 b81(uint32_t a, uint32_t b, uint32_t c) { B81 r; r.a = a; r.b = b; r.c = c; return r; }
MC_HD B81 operator&(B81 x, B81 y) { return b81(x.a & y.a, x.b & y.b, x.c & y.c); }
MC_HD B81 operator|(B81 x, B81 y) { return b81(x.a | y.a, x.b | y.b, x.c | y.c); }
MC_HD B81 andn(B81 x, B81 y) { return b81(x.a & ~y.a, x.b & ~y.b, x.c & ~y.c); }
MC_HD bool b81_eq(B81 x, B81 y) { return ((x.a ^ y.a) | (x.b ^ y.b) | (x.c ^ y.c)) == 0; }
MC_HD bool b81_any(B81 x) { return (x.a | x.b | x.c) != 0; }
template <int N> MC_HD B81 shl(B81 v) { return b81(v.a << N, fsl(v.a, v.b, N), fsl(v.b, v.c, N)); }
template <int N> MC_HD B81 shr(B81 v) { return b81(fsr(v.a, v.b, N), fsr(v.b, v.c, N), v.c >> N); }

// `bits` (< 2^18) shifted to cell `pos` (0..80) as a board; bits running past cell 95 are dropped.
// Branch-free: the word index only steers selects (lanes of a warp call this with different cells).
MC_HD B81 b81_shifted(uint32_t bits, int pos) {
    const int w = pos >> 5, sh = pos & 31;
    const uint32_t lo = bits << sh, hi = (bits >> 1) >> (31 - sh);        // hi = bits >> (32 - sh), 0 when sh == 0
    const bool w0 = w == 0, w1 = w == 1;
    return b81(w0 ? lo : 0u, w0 ? hi : (w1 ? lo : 0u), w1 ? hi : (w0 ? 0u : lo));
}
MC_HD B81 b81_bit(int pos) {
    const int w = pos >> 5;
    const uint32_t bit = 1u << (pos & 31);
    return b81(w == 0 ? bit : 0u, w == 1 ? bit : 0u, w == 2 ? bit : 0u);
}
// the edge boards with candidate wall k = 2 * anchor + orientation added (get_shortest_path(player, wall),
// Quoridor.cpp:138-156): a horizontal wall closes the south edges of cells pos, pos+1, a vertical one the east
// edges of cells pos, pos+9.  One mask, applied to the board the orientation selects.
MC_HD void b81_add_wall(B81 &oS, B81 &oE, int k) {
    const int anchor = k >> 1, pos = 9 * (anchor >> 3) + (anchor & 7);
    const uint32_t vert = 0u - (uint32_t)(k & 1);                         // all ones for a vertical wall
    const B81 m = b81_shifted((k & 1) ? 0x201u : 0x3u, pos);
    oS = b81(oS.a & ~(m.a & ~vert), oS.b & ~(m.b & ~vert), oS.c & ~(m.c & ~vert));
    oE = b81(oE.a & ~(m.a & vert), oE.b & ~(m.b & vert), oE.c & ~(m.c & vert));
}
MC_HD bool b81_test(B81 v, int pos) {
    uint32_t w = pos < 32 ? v.a : pos < 64 ? v.b : v.c;
    return (w >> (pos & 31)) & 1u;
}

// cells 0..71 (rows 0..7) and all cells whose column is not 8
#define MCTSB_ROWS07 b81(0xffffffffu, 0xffffffffu, 0x000000ffu)
#define MCTSB_COLS07 b81(0xfbfdfeffu, 0xbfdfeff7u, 0x0000ff7fu)
#define MCTSB_GOAL_W_C 0x0001ff00u   // row 8 lives in word c, bits 8..16
#define MCTSB_GOAL_B_A 0x000001ffu   // row 0 lives in word a, bits 0..8

// ------------------------------------------------------------------------------------------------
// Working position.  Replaces Quoridor_state's members (Quoridor.h:32-48): walls[9][9] and
// wall_connections[8][8] are implied by the two anchor masks; the lazily cached BFS maps
// wdists/bdists shrink to two cached path lengths (only the scalar is ever consumed).
// ------------------------------------------------------------------------------------------------
struct QState {
    uint64_t hanc, vanc;   // wall anchors, bit 8*x+y
    B81 openS;             // cell c may step to c+9 (no 'h' segment under it, not the last row)
    B81 openE;             // cell c may step to c+1 (no 'v' segment right of it, not the last column)
    int8_t cell[2];        // pawn cells: [0] White, [1] Black
    int8_t nwalls[2];      // wwallsno, bwallsno
    int8_t turn;           // 0 White to move, 1 Black
    int8_t sp[2];          // cached shortest paths (get_shortest_path(player), Quoridor.cpp:111-131); -2 = stale
    uint32_t move_counter;
};

struct QWork {             // executed-work counters
    uint32_t flood_steps, flood_queries, plies;
};

MC_HD void q_add_wall_edges(QState &s, int anchor, bool horizontal) {
    int x = anchor >> 3, y = anchor & 7, pos = 9 * x + y;
    if (horizontal) s.openS = andn(s.openS, b81_shifted(0x3u, pos));      // cells (x,y),(x,y+1) lose their south edge
    else            s.openE = andn(s.openE, b81_shifted(0x201u, pos));    // cells (x,y),(x+1,y) lose their east edge
}

MC_HD bool q_unpack(QState &s, const mctsb_qstate &p) {
    s.hanc = p.h_anchor; s.vanc = p.v_anchor;
    s.openS = MCTSB_ROWS07; s.openE = MCTSB_COLS07;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int x = 0; x < 8; x++) {
        uint32_t hrow = (uint32_t)(p.h_anchor >> (8 * x)) & 0xffu;
        uint32_t vrow = (uint32_t)(p.v_anchor >> (8 * x)) & 0xffu;
        s.openS = andn(s.openS, b81_shifted(hrow | (hrow << 1), 9 * x));
        s.openE = andn(s.openE, b81_shifted(vrow | (vrow << 9), 9 * x));
    }
    s.cell[0] = (int8_t)(p.wx * 9 + p.wy); s.cell[1] = (int8_t)(p.bx * 9 + p.by);
    s.nwalls[0] = (int8_t)p.wwalls; s.nwalls[1] = (int8_t)p.bwalls;
    s.turn = p.turn ? 1 : 0;
    s.sp[0] = s.sp[1] = -2;
    s.move_counter = p.move_counter;
    return p.wx < 9 && p.wy < 9 && p.bx < 9 && p.by < 9 && p.wwalls <= 10 && p.bwalls <= 10 &&
           !(p.wx == p.bx && p.wy == p.by) && p.flags == 0;
}

MC_HD void q_pack(const QState &s, mctsb_qstate &p) {
    p.h_anchor = s.hanc; p.v_anchor = s.vanc;
    p.wx = (uint8_t)(s.cell[0] / 9); p.wy = (uint8_t)(s.cell[0] % 9);
    p.bx = (uint8_t)(s.cell[1] / 9); p.by = (uint8_t)(s.cell[1] % 9);
    p.wwalls = (uint8_t)s.nwalls[0]; p.bwalls = (uint8_t)s.nwalls[1];
    p.turn = (uint8_t)s.turn; p.flags = 0; p.move_counter = s.move_counter; p.reserved = 0;
}

// check_winner, Quoridor.cpp:40-44 (White tested first): 0 none, 1 White, 2 Black
MC_HD int q_winner(const QState &s) {
    if (s.cell[0] >= 72) return 1;
    if (s.cell[1] < 9) return 2;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Flood fill.  One call = one frontier expansion of all 81 cells in the four directions; replaces
// one BFS level of calculate_dists_from (Quoridor.cpp:80-104).
// ------------------------------------------------------------------------------------------------
#if defined(__CUDACC__)
// 2^9, 2^1, 2^23, 2^31 (and 1): kept in the constant bank so that ptxas cannot turn the multiplies below back into shifts
static __constant__ uint32_t mc_shift_mul[5] = {1u << 9, 1u << 1, 1u << 23, 1u << 31, 1u};
#endif

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void mc_mulwide(uint32_t a, uint32_t b, uint32_t &lo, uint32_t &hi) {
    asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
}
#endif

// The edge boards are prepared once per search: oS9 = oS << 9 and oE1 = oE << 1 mask the shifted board AFTER
// the shift ((r & m) << k == (r << k) & (m << k)), so the multiplies read r directly and every mask folds into
// one (shifted & mask) | accumulated LOP3.  On the device every 81-bit shift is a 32x32->64 multiply by a power
// of two (IMAD.WIDE on the FMA pipe: x * 2^N = {x << N, x >> (32-N)}); the straightforward version is 12 funnel
// shifts + 18 LOP3, all on the ALU pipe, which issues at half the scheduler's rate.
struct QFloodMasks { B81 oS, oE, oS9, oE1; };

MC_HD QFloodMasks q_flood_masks(B81 oS, B81 oE) {
    QFloodMasks m; m.oS = oS; m.oE = oE; m.oS9 = shl<9>(oS); m.oE1 = shl<1>(oE);
    return m;
}

#if defined(__CUDA_ARCH__)
// a + b for DISJOINT bit sets (= a | b), forced onto the FMA pipe as a * 1 + b with the 1 read from the constant bank
__device__ __forceinline__ uint32_t mc_join(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(mc_shift_mul[4]), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t mc_madlo(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// (x & mask) | acc as ONE LOP3 (ptxas otherwise peels the first AND and the last OR off the chain)
__device__ __forceinline__ uint32_t mc_andor(uint32_t x, uint32_t mask, uint32_t acc) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xf8;" : "=r"(d) : "r"(acc), "r"(x), "r"(mask));   // LUT 0xf8 = a | (b & c)
    return d;
}
// The word carries of a multi-word shift never overlap the bits shifted within the word, so they are joined
// by ADDITION on the FMA pipe (IMAD) instead of by OR on the half-rate ALU pipe: 18 IMAD + 12 LOP3, each LOP3
// being (shifted & mask) | accumulated.
__device__ __forceinline__ B81 q_flood_step(B81 r, const QFloodMasks &m) {
    const uint32_t L9 = mc_shift_mul[0], L1 = mc_shift_mul[1], R9 = mc_shift_mul[2], R1 = mc_shift_mul[3];
    uint32_t Sa_lo, Sa_hi, Sb_lo, Sb_hi, Ea_lo, Ea_hi, Eb_lo, Eb_hi;
    mc_mulwide(r.a, L9, Sa_lo, Sa_hi); mc_mulwide(r.b, L9, Sb_lo, Sb_hi);
    mc_mulwide(r.a, L1, Ea_lo, Ea_hi); mc_mulwide(r.b, L1, Eb_lo, Eb_hi);
    const uint32_t Sb = mc_join(Sb_lo, Sa_hi), Sc = mc_madlo(r.c, L9, Sb_hi);
    const uint32_t Eb = mc_join(Eb_lo, Ea_hi), Ec = mc_madlo(r.c, L1, Eb_hi);
    uint32_t Na_lo, Na_hi, Nb_lo, Nb_hi, Nc_lo, Nc_hi, Wa_lo, Wa_hi, Wb_lo, Wb_hi, Wc_lo, Wc_hi;
    mc_mulwide(r.a, R9, Na_lo, Na_hi); mc_mulwide(r.b, R9, Nb_lo, Nb_hi); mc_mulwide(r.c, R9, Nc_lo, Nc_hi);
    mc_mulwide(r.a, R1, Wa_lo, Wa_hi); mc_mulwide(r.b, R1, Wb_lo, Wb_hi); mc_mulwide(r.c, R1, Wc_lo, Wc_hi);
    (void)Na_lo; (void)Wa_lo;
    const uint32_t Na = mc_join(Na_hi, Nb_lo), Nb = mc_join(Nb_hi, Nc_lo);
    const uint32_t Wa = mc_join(Wa_hi, Wb_lo), Wb = mc_join(Wb_hi, Wc_lo);
    B81 n;
    n.a = mc_andor(Wa, m.oE.a, mc_andor(Na, m.oS.a, mc_andor(Ea_lo, m.oE1.a, mc_andor(Sa_lo, m.oS9.a, r.a))));
    n.b = mc_andor(Wb, m.oE.b, mc_andor(Nb, m.oS.b, mc_andor(Eb, m.oE1.b, mc_andor(Sb, m.oS9.b, r.b))));
    n.c = mc_andor(Wc_hi, m.oE.c, mc_andor(Nc_hi, m.oS.c, mc_andor(Ec, m.oE1.c, mc_andor(Sc, m.oS9.c, r.c))));
    return n;
}
#else
MC_HD B81 q_flood_step(B81 r, const QFloodMasks &m) {
    B81 south = shl<9>(r) & m.oS9, north = shr<9>(r) & m.oS, east = shl<1>(r) & m.oE1, west = shr<1>(r) & m.oE;
    return b81(r.a | south.a | north.a | east.a | west.a,
               r.b | south.b | north.b | east.b | west.b,
               r.c | south.c | north.c | east.c | west.c);
}
#endif

// Shortest path length from `start` to `player`'s goal row ignoring pawns, -1 if walled off:
// get_shortest_path / calculate_dists_from(stop_at_goal=true), Quoridor.cpp:46-106,161-169.
MC_HD int q_flood_dist(B81 oS, B81 oE, int start, int player, QWork *wk) {
    B81 r = b81_bit(start);
    const uint32_t ga = player ? MCTSB_GOAL_B_A : 0u, gc = player ? 0u : MCTSB_GOAL_W_C;
    if (wk) wk->flood_queries++;
    if ((r.a & ga) | (r.c & gc)) return 0;
    const QFloodMasks fm = q_flood_masks(oS, oE);
    for (int d = 1;; d++) {
        B81 n = q_flood_step(r, fm);
        if (wk) wk->flood_steps++;
        if ((n.a & ga) | (n.c & gc)) return d;
        if (b81_eq(n, r)) return -1;
        r = n;
    }
}

// get_shortest_path(player) with the wdists/bdists cache semantics (Quoridor.cpp:111-131)
MC_HD int q_sp(QState &s, int player, QWork *wk) {
    if (s.sp[player] == -2) s.sp[player] = (int8_t)q_flood_dist(s.openS, s.openE, s.cell[player], player, wk);
    return s.sp[player];
}

// get_shortest_path(player, extra wall), Quoridor.cpp:138-156: the wall is added to a register copy
// of the edge boards, nothing to undo.  k = 2*anchor + (0 horizontal | 1 vertical).
MC_HD int q_sp_with_wall(const QState &s, int player, int k, QWork *wk) {
    B81 oS = s.openS, oE = s.openE;
    b81_add_wall(oS, oE, k);
    return q_flood_dist(oS, oE, s.cell[player], player, wk);
}

// ------------------------------------------------------------------------------------------------
// Wall legality without the path test: legal_wall(..., check_blocking=false), Quoridor.cpp:291-303,
// for all 64 anchors of one orientation at once.  A horizontal wall at (x,y) needs segments (x,y)
// and (x,y+1) free and no vertical wall anchored at the same lattice point; in anchor terms: no
// horizontal anchor at (x,y-1),(x,y),(x,y+1) and no vertical anchor at (x,y).  Mirror for vertical.
// ------------------------------------------------------------------------------------------------
MC_HD void q_cheap_wall_masks(const QState &s, uint64_t &legal_h, uint64_t &legal_v) {
    if (s.nwalls[s.turn] <= 0) { legal_h = legal_v = 0; return; }
    const uint64_t col0 = 0x0101010101010101ull, col7 = 0x8080808080808080ull;
    uint64_t H = s.hanc, V = s.vanc;
    legal_h = ~(H | ((H << 1) & ~col0) | ((H >> 1) & ~col7) | V);
    legal_v = ~(V | (V << 8) | (V >> 8) | H);
}

// Pawn steps.  The 12 candidates in the order of get_legal_step_moves2 (Quoridor.cpp:461-472):
// (-1,0),(-2,0),(+1,0),(+2,0),(0,-1),(0,-2),(0,+1),(0,+2),(-1,-1),(-1,+1),(+1,-1),(+1,+1); target
// cell of candidate i = mover cell + q_step_delta(i).  Returns the 12-bit mask of legal ones
// (legal_step, Quoridor.cpp:200-287) for the side to move.
MC_HD int q_step_delta(int i) {
    // 9*dx+dy for the 12 candidates: {-9,-18,9,18,-1,-2,1,2,-10,-8,8,10}, biased by 18 and packed 6 bits each
    const uint64_t packed = (9ull) | (0ull << 6) | (27ull << 12) | (36ull << 18) | (17ull << 24) | (16ull << 30) |
                            (19ull << 36) | (20ull << 42) | (8ull << 48) | (10ull << 54);
    if (i >= 10) return i == 10 ? 8 : 10;
    return (int)((packed >> (6 * i)) & 63ull) - 18;
}

// 4-bit mask N,S,W,E of the directions a pawn on cell c can move in (board edge or wall blocks)
MC_HD uint32_t q_moves_from(const QState &s, int c) {
    uint32_t m = 0;
    if (c >= 9 && b81_test(s.openS, c - 9)) m |= 1u;   // north
    if (b81_test(s.openS, c)) m |= 2u;                  // south
    if (c >= 1 && b81_test(s.openE, c - 1)) m |= 4u;   // west (openE is 0 on column 8, so no wrap)
    if (b81_test(s.openE, c)) m |= 8u;                  // east
    return m;
}

MC_HD uint32_t q_step_mask(const QState &s) {
    int me = s.cell[s.turn], en = s.cell[1 - s.turn];
    uint32_t mm = q_moves_from(s, me), em = q_moves_from(s, en);
    bool aN = (mm & 1u) && en == me - 9, aS = (mm & 2u) && en == me + 9;
    bool aW = (mm & 4u) && en == me - 1, aE = (mm & 8u) && en == me + 1;
    bool eN = em & 1u, eS = em & 2u, eW = em & 4u, eE = em & 8u;
    uint32_t r = 0;
    if ((mm & 1u) && !aN) r |= 1u << 0;
    if (aN && eN)         r |= 1u << 1;
    if ((mm & 2u) && !aS) r |= 1u << 2;
    if (aS && eS)         r |= 1u << 3;
    if ((mm & 4u) && !aW) r |= 1u << 4;
    if (aW && eW)         r |= 1u << 5;
    if ((mm & 8u) && !aE) r |= 1u << 6;
    if (aE && eE)         r |= 1u << 7;
    // diagonal jumps: the enemy is adjacent, the square behind it is closed, the sidestep is open
    if ((aN && !eN && eW) || (aW && !eW && eN)) r |= 1u << 8;    // (-1,-1)
    if ((aN && !eN && eE) || (aE && !eE && eN)) r |= 1u << 9;    // (-1,+1)
    if ((aS && !eS && eW) || (aW && !eW && eS)) r |= 1u << 10;   // (+1,-1)
    if ((aS && !eS && eE) || (aE && !eE && eS)) r |= 1u << 11;   // (+1,+1)
    return r;
}

// Full wall legality (legal_wall with check_blocking=true, Quoridor.cpp:289-329): cheap test and
// both pawns keep a path.  The reference skips the path test for "isolated" walls; that shortcut
// never changes the answer (a wall touching nothing cannot close a region), so the flood fill may
// simply be run on every cheap-legal candidate.
MC_HD bool q_wall_keeps_paths(const QState &s, int k, QWork *wk) {
    return q_sp_with_wall(s, 0, k, wk) >= 0 && q_sp_with_wall(s, 1, k, wk) >= 0;
}

// get_best_step_move, Quoridor.cpp:476-497: candidates in REVERSED list order (push_front), first
// strict minimum of the shortest path from the target; -1 when no legal step exists.
MC_HD int q_best_step(const QState &s, uint32_t steps, QWork *wk) {
    int me = s.cell[s.turn], best = -1, mn = 9999999;
    for (int i = 11; i >= 0; i--) {
        if (!((steps >> i) & 1u)) continue;
        int t = me + q_step_delta(i);
        int path = q_flood_dist(s.openS, s.openE, t, s.turn, wk);
        if (path >= 0 && path < mn) { mn = path; best = t; }
    }
    return best;
}

// play_move, Quoridor.cpp:344-392, for a move already known to be legal.
MC_HD void q_play_step(QState &s, int target_cell) {
    s.cell[s.turn] = (int8_t)target_cell;
    s.sp[s.turn] = -2;
    s.turn ^= 1; s.move_counter++;
}
MC_HD void q_play_wall(QState &s, int k) {
    int anchor = k >> 1;
    if (k & 1) s.vanc |= 1ull << anchor; else s.hanc |= 1ull << anchor;
    q_add_wall_edges(s, anchor, !(k & 1));
    s.nwalls[s.turn]--;
    s.sp[0] = s.sp[1] = -2;
    s.turn ^= 1; s.move_counter++;
}

// ------------------------------------------------------------------------------------------------
// fp64 without contraction: the host reference is compiled without FMA, so every product and sum
// is rounded separately; on the device that has to be spelled out.
// ------------------------------------------------------------------------------------------------
MC_HD double d_mul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    volatile double r = a * b; return r;
#endif
}
MC_HD double d_add(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    volatile double r = a + b; return r;
#endif
}
MC_HD double d_div(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __ddiv_rn(a, b);
#else
    volatile double r = a / b; return r;
#endif
}

// evaluate_position, Quoridor.cpp:629-664, bug-faithful (MAX instead of MIN at :661).
MC_HD double q_eval_from(int wp, int bp, int w, int b, int turn) {
    bool white_ahead = wp + (turn != 0) <= bp - 1;
    bool black_ahead = bp + (turn != 1) <= wp - 1;
    if (b <= 0 && white_ahead) return 0.95;
    if (w <= 0 && black_ahead) return 1.0 - 0.95;
    if (b <= 1 && w >= 2 && white_ahead) return 0.95 - 0.1;
    if (w <= 1 && b >= 2 && black_ahead) return 1.0 - (0.95 - 0.1);
    double wallsdiff = 0.0;
    int mx = w > b ? w : b;
    if (mx > 0) {
        double q = d_div((double)((w - b) * (w - b)), (double)(mx * mx));
        wallsdiff = (w > b) ? q : -q;
    }
    double path_diff = (double)(bp - wp) + (turn == 0 ? 0.5 : -0.5);     // exact
    double distance = d_div(path_diff > 10.0 ? path_diff : 10.0, 10.0);
    return d_add(d_add(0.5, d_mul(0.2, wallsdiff)), d_mul(0.2, distance));
}
MC_HD double q_eval(QState &s, QWork *wk) {
    int wp = q_sp(s, 0, wk), bp = q_sp(s, 1, wk);
    return q_eval_from(wp, bp, s.nwalls[0], s.nwalls[1], s.turn);
}

// force_playwall, Quoridor.cpp:666-678
MC_HD bool q_force_playwall(QState &s, QWork *wk) {
    int p = s.turn, e = 1 - p;
    int ours = q_sp(s, p, wk), theirs = q_sp(s, e, wk), ow = s.nwalls[p], ew = s.nwalls[e];
    if (theirs <= 1 && ours > theirs) return true;
    if (theirs <= 2 && ours > theirs + 1 && ow >= ew) return true;
    if (theirs <= 3 && ours > theirs + 2 && ow > ew) return true;
    return false;
}

// U() < p for p in {0.1, 0.75, 0.8}: u32*2^-32 < p  <=>  u32 < ceil(p*2^32) (p taken as the double)
#define MCTSB_THR_0_10 429496730u
#define MCTSB_THR_0_75 3221225472u
#define MCTSB_THR_0_80 3435973837u

// ------------------------------------------------------------------------------------------------
// Exact pruning of shortest-path queries.
//
// The reference answers "how long is q's shortest path if wall m is added" with a fresh BFS for
// every candidate m (get_shortest_path(q, m), Quoridor.cpp:138-156; ~1 270 BFS per playout).  A
// wall only changes the distance if it removes an edge that lies on SOME shortest path of q, so one
// pass per player replaces most of those BFS by a bit test:
//   1. goal-rooted layers  L_0 = goal row, L_j = cells at distance j from the goal row, until the
//      pawn is reached (d = sp(q));
//   2. trace from the pawn down the layers: A_d = {pawn}, A_k = neighbours(A_{k+1}) in L_k; the edges
//      used are exactly the edges on shortest paths (stored at the upper / left cell of the edge);
//   3. while tracing, flag the wall anchors that may change the distance (see QTrace for the exact rule);
//      cut_h / cut_v bit a <=> the horizontal / vertical wall anchored at a must be measured.
// m not in cut  =>  sp(q, m) == sp(q) exactly; m in cut => run the flood fill as before.
// ------------------------------------------------------------------------------------------------
#define MCTSB_MAX_LAYERS 24          // deeper searches fall back to "every wall may cut" (still exact)

// Scratch for the layers: plain array on the host, strided shared memory in the kernels.
struct LocalLayers {
    B81 l[MCTSB_MAX_LAYERS];
    MC_HD void put(int j, B81 v) { l[j] = v; }
    MC_HD B81 get(int j) const { return l[j]; }
};

// 9-bit row x of a cell board
MC_HD uint32_t b81_row(B81 v, int x) {
    int pos = 9 * x, w = pos >> 5, sh = pos & 31;
    uint32_t lo = w == 0 ? v.a : w == 1 ? v.b : v.c, hi = w == 0 ? v.b : w == 1 ? v.c : 0u;
    uint32_t r = sh ? ((lo >> sh) | (hi << (32 - sh))) : lo;
    return r & 0x1ffu;
}

// Trace state.  Every step walks the shortest-path cells one layer down (a "transition") and decides
// which wall anchors may change the distance.  A wall removes two parallel neighbouring edges.  With
// T_j = the shortest-path cells of layer j (every one of them lies on a complete shortest path):
//   * an edge that is alone in its transition is a bottleneck: every wall over it changes the distance;
//   * a wall over ONE edge of a transition with several edges changes nothing;
//   * a wall over two edges of the same transition, or of transitions that are not neighbours, is kept
//     as "may change" (rare; measured by a flood fill);
//   * a wall over edge e of transition (j+1 -> j) and edge f of (j -> j-1) changes the distance iff every
//     cell v of T_j has in(v) = {e} or out(v) = {f}; at most one cell can have either, so this needs
//     |T_j| = 2.  When T_j has any other size the wall changes nothing — that alone removes two thirds
//     of the candidates the coarser rule "two multi edges" used to send to a flood fill.
// FS / FE collect the flagged anchors as cell boards (anchor (x,y) at cell 9x+y: the wall covers the
// south edges of cells c, c+1, resp. the east edges of cells c, c+9).
struct QTrace {
    B81 A;                 // T_j, the frontier
    B81 FS, FE;            // flagged horizontal / vertical anchors
    B81 SmOld, EmOld;      // multi edges of the transitions before the previous one
    B81 SmPrev, EmPrev;    // multi edges of the previous transition
};

MC_HD void q_trace_init(QTrace &t, B81 pawn) {
    t.A = pawn;
    t.FS = t.FE = t.SmOld = t.EmOld = t.SmPrev = t.EmPrev = b81(0u, 0u, 0u);
}

MC_HD void q_trace_step(QTrace &t, B81 L, B81 oS, B81 oE) {
    B81 sS = t.A & oS & shr<9>(L);            // u in A steps south into L: edge stored at u
    B81 tN = shr<9>(t.A) & oS & L;            // u in A steps north to v = u-9 in L: edge stored at v
    B81 sE = t.A & oE & shr<1>(L);
    B81 tW = shr<1>(t.A) & oE & L;
    const B81 es = sS | tN, ee = sE | tW;
    const int cnt = popc32(es.a) + popc32(es.b) + popc32(es.c) + popc32(ee.a) + popc32(ee.b) + popc32(ee.c);
    const bool single = cnt == 1;
    const bool mid2 = popc32(t.A.a) + popc32(t.A.b) + popc32(t.A.c) == 2;     // |T_j| of the layer between the two transitions
    // partners an edge of this transition pairs up with: same transition, far transitions, the previous
    // transition only when the layer in between has exactly two shortest-path cells — and EVERYTHING when
    // the edge is a bottleneck (both anchors over it are flagged)
    const uint32_t m2 = mid2 ? 0xffffffffu : 0u, m1 = single ? 0xffffffffu : 0u;
    const B81 xs = b81(es.a | t.SmOld.a | (t.SmPrev.a & m2) | m1, es.b | t.SmOld.b | (t.SmPrev.b & m2) | m1, es.c | t.SmOld.c | (t.SmPrev.c & m2) | m1);
    const B81 xe = b81(ee.a | t.EmOld.a | (t.EmPrev.a & m2) | m1, ee.b | t.EmOld.b | (t.EmPrev.b & m2) | m1, ee.c | t.EmOld.c | (t.EmPrev.c & m2) | m1);
    const B81 es1 = shr<1>(es), ee9 = shr<9>(ee);
    t.FS = t.FS | (es & shr<1>(xs)) | (xs & es1);
    t.FE = t.FE | (ee & shr<9>(xe)) | (xe & ee9);
    t.SmOld = t.SmOld | t.SmPrev; t.EmOld = t.EmOld | t.EmPrev;
    t.SmPrev = b81(es.a & ~m1, es.b & ~m1, es.c & ~m1); t.EmPrev = b81(ee.a & ~m1, ee.b & ~m1, ee.c & ~m1);
    t.A = shl<9>(sS) | tN | shl<1>(sE) | tW;
}

// Flag boards -> 64-bit anchor masks (bit 8x+y): rows 0..7, columns 0..7 of the boards.
MC_HD void q_trace_to_anchors(const QTrace &t, uint64_t &may_h, uint64_t &may_v) {
    uint64_t ch = 0, cv = 0;
#pragma unroll
    for (int x = 0; x < 8; x++) {
        ch |= (uint64_t)(b81_row(t.FS, x) & 0xffu) << (8 * x);
        cv |= (uint64_t)(b81_row(t.FE, x) & 0xffu) << (8 * x);
    }
    may_h = ch; may_v = cv;
}

// Cell board (bit 9x+y) -> anchor mask (bit 8x+y): rows 0..7, columns 0..7.
MC_HD uint64_t b81_to_anchors(B81 v) {
    uint64_t a = 0;
#pragma unroll
    for (int x = 0; x < 8; x++) a |= (uint64_t)(b81_row(v, x) & 0xffu) << (8 * x);
    return a;
}

// Walls that cannot close anybody in, found without a search.  A wall removes two parallel neighbouring edges, say
// A|A' and B|B'.  If a three-step detour around one END of the wall is open (A -> the cell beside it -> across ->
// A') and the two cells on either side of the wall are joined along it (A-B and A'-B' open), then A~A' and B~B'
// still hold, every path that used a removed edge has a replacement, and the connectivity of the whole board is
// what it was: both pawns keep their paths (legal_wall's blocking test, Quoridor.cpp:305-326, must say "legal").
// The same holds when BOTH ends have their detour (A~A' around one end, B~B' around the other), joined or not.
// On an open board this covers nearly every wall; the others go to the flood fill as before.
//   horizontal wall under cells c, c+1:  oE[c] & oE[c+9] & ( oE[c-1] & oS[c-1] & oE[c+8]  |  oE[c+1] & oS[c+2] & oE[c+10] )
//   vertical wall right of cells c, c+9: oS[c] & oS[c+1] & ( oS[c-9] & oE[c-9] & oS[c-8]  |  oS[c+9] & oE[c+18] & oS[c+10] )
// (openE is 0 on column 8 and openS on row 8, so terms that would wrap around a board edge vanish by themselves.)
MC_HD void q_detour_safe_masks(const QState &s, uint64_t &safe_h, uint64_t &safe_v) {
    const B81 oS = s.openS, oE = s.openE;
    const B81 e9 = shr<9>(oE), s1 = shr<1>(oS);
    const B81 h_left = shl<1>(oE & oS & shr<9>(oE));                     // bit c: oE[c-1] & oS[c-1] & oE[c+8]
    const B81 h_right = shr<1>(oE) & shr<2>(oS) & shr<10>(oE);           // bit c: oE[c+1] & oS[c+2] & oE[c+10]
    const B81 sh = (oE & e9 & (h_left | h_right)) | (h_left & h_right);   // one detour + joined sides, or a detour at either end
    const B81 v_top = shl<9>(oS & oE & s1);                              // bit c: oS[c-9] & oE[c-9] & oS[c-8]
    const B81 v_bottom = shr<9>(oS) & shr<18>(oE) & shr<10>(oS);         // bit c: oS[c+9] & oE[c+18] & oS[c+10]
    const B81 sv = (oS & s1 & (v_top | v_bottom)) | (v_top & v_bottom);
    safe_h = b81_to_anchors(sh); safe_v = b81_to_anchors(sv);
}

// the same test for ONE wall k = 2 * anchor + orientation (branch-free apart from the orientation select)
MC_HD bool q_detour_safe_one(B81 oS, B81 oE, int k) {
    const int a = k >> 1, c = 9 * (a >> 3) + (a & 7);
    if (k & 1) {
        const bool top = c >= 9 && b81_test(oS, c - 9) && b81_test(oE, c - 9) && b81_test(oS, c - 8);
        const bool bottom = b81_test(oS, c + 9) && b81_test(oE, c + 18) && b81_test(oS, c + 10);
        return (b81_test(oS, c) && b81_test(oS, c + 1) && (top || bottom)) || (top && bottom);
    }
    const bool left = c >= 1 && b81_test(oE, c - 1) && b81_test(oS, c - 1) && b81_test(oE, c + 8);
    const bool right = b81_test(oE, c + 1) && b81_test(oS, c + 2) && b81_test(oE, c + 10);
    return (b81_test(oE, c) && b81_test(oE, c + 9) && (left || right)) || (left && right);
}

// ------------------------------------------------------------------------------------------------
// Connectivity only (legality of a wall, not its effect on the distance): ONE path is enough.  A wall that covers no
// edge of some fixed pawn-to-goal path leaves that path, hence the pawn's connection, intact.  The walk picks one
// predecessor per layer (south, north, east, west in that order) — a single bit moves down the layers, ~40
// instructions per layer instead of the ~170 of q_trace_step's all-shortest-paths bookkeeping — and the walls over
// its sp edges (two anchors per edge) are the only ones that can close this pawn in.
// ------------------------------------------------------------------------------------------------
struct QPathWalk { B81 cur, es, ee; };       // the walker, the south-type and east-type edges it used (stored at the upper / left cell)

MC_HD void q_path_init(QPathWalk &w, B81 pawn) { w.cur = pawn; w.es = w.ee = b81(0u, 0u, 0u); }

MC_HD void q_path_step(QPathWalk &w, B81 L, B81 oS, B81 oE) {
    const B81 sS = w.cur & oS & shr<9>(L);           // step south: edge stored at cur
    const B81 tN = shr<9>(w.cur) & oS & L;           // step north to cur-9: edge stored there
    const B81 sE = w.cur & oE & shr<1>(L);
    const B81 tW = shr<1>(w.cur) & oE & L;
    const uint32_t hS = b81_any(sS) ? 0xffffffffu : 0u, hN = b81_any(tN) ? ~hS : 0u;
    const uint32_t vert = hS | hN, hE = b81_any(sE) ? ~vert : 0u, hW = ~(vert | hE);
    const B81 s9 = shl<9>(sS), e1 = shl<1>(sE);
    w.es = b81(w.es.a | (sS.a & hS) | (tN.a & hN), w.es.b | (sS.b & hS) | (tN.b & hN), w.es.c | (sS.c & hS) | (tN.c & hN));
    w.ee = b81(w.ee.a | (sE.a & hE) | (tW.a & hW), w.ee.b | (sE.b & hE) | (tW.b & hW), w.ee.c | (sE.c & hE) | (tW.c & hW));
    w.cur = b81((s9.a & hS) | (tN.a & hN) | (e1.a & hE) | (tW.a & hW), (s9.b & hS) | (tN.b & hN) | (e1.b & hE) | (tW.b & hW),
                (s9.c & hS) | (tN.c & hN) | (e1.c & hE) | (tW.c & hW));
}

// anchors whose wall covers an edge of the walked path: an edge stored at cell c lies under the horizontal (right of the
// vertical) walls anchored at c and at c-1 (c-9)
MC_HD void q_path_to_anchors(const QPathWalk &w, uint64_t &cut_h, uint64_t &cut_v) {
    cut_h = b81_to_anchors(w.es | shr<1>(w.es));
    cut_v = b81_to_anchors(w.ee | shr<9>(w.ee));
}

struct QCut { uint64_t h, v; int sp; };    // anchors whose wall may change sp(q) (must be measured), and sp(q)
    // anchors whose wall cuts a shortest-path edge, and sp(q)

template <class Layers>
MC_HD QCut q_cut_masks(const QState &s, int player, Layers &ls, QWork *wk) {
    QCut out;
    const B81 goal = player ? b81(MCTSB_GOAL_B_A, 0u, 0u) : b81(0u, 0u, MCTSB_GOAL_W_C);
    const B81 pawn = b81_bit(s.cell[player]);
    const B81 oS = s.openS, oE = s.openE;
    if (wk) wk->flood_queries++;
    // 1. layers from the goal row
    const QFloodMasks fm = q_flood_masks(oS, oE);
    B81 reached = goal;
    int d = 0;
    bool found = b81_any(pawn & goal), dead = false;
    while (!found && !dead) {
        B81 grown = q_flood_step(reached, fm);
        B81 layer = andn(grown, reached);
        if (wk) wk->flood_steps++;
        d++;
        if (!b81_any(layer) || d >= MCTSB_MAX_LAYERS) { dead = true; break; }
        ls.put(d, layer);
        reached = grown;
        found = b81_any(layer & pawn);
    }
    if (dead) {                                   // walled off (never in a legal position) or too deep to store
        out.h = out.v = ~0ull;
        out.sp = d >= MCTSB_MAX_LAYERS ? q_flood_dist(oS, oE, s.cell[player], player, wk) : -1;
        return out;
    }
    out.sp = d;
    // 2. trace the shortest paths back to the goal row
    QTrace t; q_trace_init(t, pawn);
    for (int k = d - 1; k >= 0; k--) {
        q_trace_step(t, k == 0 ? goal : ls.get(k), oS, oE);
        if (wk) wk->flood_steps++;
    }
    // 3. edges -> wall anchors
    q_trace_to_anchors(t, out.h, out.v);
    return out;
}

// the same with the one-path walk: anchors whose wall may close `player` in (every other wall leaves the walked path)
template <class Layers>
MC_HD QCut q_block_masks(const QState &s, int player, Layers &ls, QWork *wk) {
    QCut out;
    const B81 goal = player ? b81(MCTSB_GOAL_B_A, 0u, 0u) : b81(0u, 0u, MCTSB_GOAL_W_C);
    const B81 pawn = b81_bit(s.cell[player]);
    const B81 oS = s.openS, oE = s.openE;
    if (wk) wk->flood_queries++;
    const QFloodMasks fm = q_flood_masks(oS, oE);
    B81 reached = goal;
    int d = 0;
    bool found = b81_any(pawn & goal), dead = false;
    while (!found && !dead) {
        B81 grown = q_flood_step(reached, fm);
        B81 layer = andn(grown, reached);
        if (wk) wk->flood_steps++;
        d++;
        if (!b81_any(layer) || d >= MCTSB_MAX_LAYERS) { dead = true; break; }
        ls.put(d, layer);
        reached = grown;
        found = b81_any(layer & pawn);
    }
    if (dead) { out.h = out.v = ~0ull; out.sp = -1; return out; }      // walled off or too deep to store: every wall is measured
    out.sp = d;
    QPathWalk w; q_path_init(w, pawn);
    for (int k = d - 1; k >= 0; k--) {
        q_path_step(w, k == 0 ? goal : ls.get(k), oS, oE);
        if (wk) wk->flood_steps++;
    }
    q_path_to_anchors(w, out.h, out.v);
    return out;
}

// position of candidate k = 2*anchor + orientation in the un-shuffled pool (Quoridor.cpp:709-718:
// anchors row-major, horizontal before vertical) = number of cheap-legal candidates before it
MC_HD int q_pool_rank(uint64_t cheap_h, uint64_t cheap_v, int k) {
    int a = k >> 1;
    uint64_t below = a ? (~0ull >> (64 - a)) : 0ull;
    uint64_t below_h = (k & 1) ? (below | (1ull << a)) : below;
    return popc64(cheap_h & below_h) + popc64(cheap_v & below);
}

// candidate with pool rank r (inverse of q_pool_rank): binary search on the anchor index
MC_HD int q_pool_select(uint64_t cheap_h, uint64_t cheap_v, int r) {
    int a = 0;                                             // largest anchor with (#candidates before it) <= r
#pragma unroll
    for (int step = 32; step > 0; step >>= 1) {
        int t = a + step;                                  // 1..63
        uint64_t below = ~0ull >> (64 - t);
        if (popc64(cheap_h & below) + popc64(cheap_v & below) <= r) a = t;
    }
    uint64_t below = a ? (~0ull >> (64 - a)) : 0ull;
    int rem = r - (popc64(cheap_h & below) + popc64(cheap_v & below));
    bool has_h = (cheap_h >> a) & 1ull, has_v = (cheap_v >> a) & 1ull;
    if (rem == 0 && has_h) return 2 * a;
    if ((rem == 0 && has_v) || (rem == 1 && has_h && has_v)) return 2 * a + 1;
    return -1;
}

// Gain of wall k for the mover (enemy_enc - our_enc if > 0, else 0): the acceptance test shared by
// the best-wall and guided branches (Quoridor.cpp:728-733, 756-761).  `k` must be in cut(enemy) —
// every other wall has enemy_enc == 0 and is rejected without a query.  The mover's own path is
// re-measured only if the wall may touch it (own_cut_known && !in_own_cut => our_enc == 0).
MC_HD int q_wall_gain(const QState &s, int k, int sp_e, int sp_p, bool own_may_change, QWork *wk) {
    int p = s.turn, e = 1 - p;
    int ep = q_sp_with_wall(s, e, k, wk);
    int ee = ep - sp_e;
    if (!(ep >= 0 && ee > 0)) return 0;
    int oe = 0;
    if (own_may_change) {
        int op = q_sp_with_wall(s, p, k, wk);
        if (op < 0) return 0;
        oe = op - sp_p;
    }
    return ee > oe ? ee - oe : 0;
}

MC_HD bool q_mask_has(uint64_t h, uint64_t v, int k) { return (((k & 1) ? v : h) >> (k >> 1)) & 1ull; }
MC_HD void q_mask_clear(uint64_t &h, uint64_t &v, int k) { if (k & 1) v &= ~(1ull << (k >> 1)); else h &= ~(1ull << (k >> 1)); }
MC_HD int q_mask_first(uint64_t h, uint64_t v) {       // lowest candidate in pool order, -1 if empty
    int ah = h ? (((uint32_t)h) ? ctz32((uint32_t)h) : 32 + ctz32((uint32_t)(h >> 32))) : 64;
    int av = v ? (((uint32_t)v) ? ctz32((uint32_t)v) : 32 + ctz32((uint32_t)(v >> 32))) : 64;
    if (ah == 64 && av == 64) return -1;
    return ah <= av ? 2 * ah : 2 * av + 1;
}

// Encoded move: 0..80 pawn target cell; 128 + k wall (k = 2*anchor + orientation); -1 none.
MC_HD int q_move_to_id(int mv) {
    if (mv < 0) return -1;
    if (mv < 81) return mv;
    int k = mv - 128;
    return ((k & 1) ? 145 : 81) + (k >> 1);
}

// Place of pool element `r` (rank in the un-shuffled pool, Quoridor.cpp:709-718) in OUR shuffle
// (include/mctsb.h "Random draws"): element j0 = mulhi(draw(base), n) comes first, every other
// element is ordered by (its own draw, r).  Returned as one ascending 64-bit sort key.
template <class Draws>
MC_HD uint64_t q_shuffle_key(Draws &d, uint32_t base, int n, int j0, int r) {
    if (n < 2 || r == j0) return 0ull;
    uint32_t key = d.at(base + 1u + (uint32_t)(r < j0 ? r : r - 1));
    return (((uint64_t)key + 1ull) << 7) | (uint64_t)r;
}

// pick_semirandom_move, Quoridor.cpp:680-803, bug-faithful: the wall branch is
// entered iff Black still has walls, one draw is consumed iff force_playwall() is false.  The
// shuffled pool (Quoridor.cpp:707-720) is never built: only candidates that can pass the branch's
// test are ranked, each from its own draw.
template <class Draws, class Layers>
MC_HD int q_pick_semirandom(QState &s, Draws &d, Layers &ls, QWork *wk) {
    const int p = s.turn, e = 1 - p;
    uint64_t cheap_h = 0, cheap_v = 0;
    if (s.nwalls[1] != 0) q_cheap_wall_masks(s, cheap_h, cheap_v);
    const int n = popc64(cheap_h) + popc64(cheap_v);
    // enemy cut masks are needed by the best-wall and guided branches (77.5 % of the wall plies)
    QCut cut_e; cut_e.h = cut_e.v = 0; cut_e.sp = 0;
    if (n > 0) {
        cut_e = q_cut_masks(s, e, ls, wk);
        s.sp[e] = (int8_t)cut_e.sp;
    }
    const int sp_p = q_sp(s, p, wk), sp_e = q_sp(s, e, wk);
    if (!q_force_playwall(s, wk)) (void)d.next();
    if (s.nwalls[1] != 0) {
        const uint32_t base = d.consumed();
        int j0 = 0;
        if (n >= 2) { j0 = (int)mulhi_u32(d.at(base), (uint32_t)n); d.skip((uint32_t)n); }
        if (d.next() < MCTSB_THR_0_10) {                       // best wall (:721-746): max gain, first in shuffled order
            uint64_t ch = cheap_h & cut_e.h, cv = cheap_v & cut_e.v;
            int best = -1, best_gain = 0; uint64_t best_key = 0;
            if (ch | cv) {
                QCut cut_p = q_cut_masks(s, p, ls, wk);
                for (int k = q_mask_first(ch, cv); k >= 0; q_mask_clear(ch, cv, k), k = q_mask_first(ch, cv)) {
                    int g = q_wall_gain(s, k, sp_e, sp_p, q_mask_has(cut_p.h, cut_p.v, k), wk);
                    if (g <= 0 || g < best_gain) continue;
                    uint64_t key = q_shuffle_key(d, base, n, j0, q_pool_rank(cheap_h, cheap_v, k));
                    if (g > best_gain || key < best_key) { best = k; best_gain = g; best_key = key; }
                }
            }
            if (best >= 0) return 128 + best;
        } else if (d.next() < MCTSB_THR_0_75) {                // guided (:748-772): first in shuffled order with gain > 0
            uint64_t ch = cheap_h & cut_e.h, cv = cheap_v & cut_e.v;
            while (ch | cv) {
                int best = -1; uint64_t best_key = 0;                      // next candidate in shuffled order
                uint64_t th = ch, tv = cv;
                for (int k = q_mask_first(th, tv); k >= 0; q_mask_clear(th, tv, k), k = q_mask_first(th, tv)) {
                    uint64_t key = q_shuffle_key(d, base, n, j0, q_pool_rank(cheap_h, cheap_v, k));
                    if (best < 0 || key < best_key) { best = k; best_key = key; }
                }
                if (q_wall_gain(s, best, sp_e, sp_p, true, wk) > 0) return 128 + best;
                q_mask_clear(ch, cv, best);
            }
        } else if (n > 0) {                                    // first fully legal in shuffled order (:773-788)
            int k = q_pool_select(cheap_h, cheap_v, j0);       // the head of the shuffle: one draw
            if (q_detour_safe_one(s.openS, s.openE, k) || q_wall_keeps_paths(s, k, wk)) return 128 + k;
            uint64_t th = cheap_h, tv = cheap_v;
            q_mask_clear(th, tv, k);
            while (th | tv) {                                  // rare: walk on in key order
                int best = -1; uint64_t best_key = 0;
                uint64_t uh = th, uv = tv;
                for (int c = q_mask_first(uh, uv); c >= 0; q_mask_clear(uh, uv, c), c = q_mask_first(uh, uv)) {
                    uint64_t key = q_shuffle_key(d, base, n, j0, q_pool_rank(cheap_h, cheap_v, c));
                    if (best < 0 || key < best_key) { best = c; best_key = key; }
                }
                if (q_wall_keeps_paths(s, best, wk)) return 128 + best;
                q_mask_clear(th, tv, best);
            }
        }
    }
    uint32_t steps = q_step_mask(s);
    if (s.nwalls[e] == 0 || d.next() < MCTSB_THR_0_80) return q_best_step(s, steps, wk);   // :792-793
    uint32_t r = d.next() & 0x7fffffffu;                     // rand() (:796)
    int ns = popc32(steps);
    if (ns == 0) return -1;                                  // the reference divides by zero here
    int idx = (int)(r % (uint32_t)ns);
    for (int i = 0; i < 12; i++) if ((steps >> i) & 1u) { if (idx == 0) return s.cell[s.turn] + q_step_delta(i); idx--; }
    return -1;
}

MC_HD void q_play_encoded(QState &s, int mv) {
    if (mv < 81) q_play_step(s, mv); else q_play_wall(s, mv - 128);
}

// Quoridor_state::rollout, Quoridor.cpp:809-855: at most 50 plies of the policy above, terminal
// test at the top of each ply only, heuristic evaluation afterwards.  The evaluation inside the
// loop (:826-829) can never stop the playout and has no side effect besides filling the path
// cache, so it is not computed here.
template <class Draws, class Layers>
MC_HD double q_rollout_ref(QState &s, Draws &d, Layers &ls, int &plies, int &kind, QWork *wk) {
    plies = 0; kind = 2;
    for (int i = 0; i < 50; i++) {
        int w = q_winner(s);
        if (w) { kind = w == 1 ? 0 : 1; return w == 1 ? 1.0 : 0.0; }
        int mv = q_pick_semirandom(s, d, ls, wk);
        if (mv < 0) { kind = 3; break; }
        q_play_encoded(s, mv);
        plies++;
    }
    return q_eval(s, wk);
}

// ------------------------------------------------------------------------------------------------
// All legal moves of the side to move: the set generate_all_moves enumerates (Quoridor.cpp:594-616),
// as the 12-bit step mask plus one 64-bit anchor mask per wall orientation.  Thread-serial version
// (host class, UNI policy); the movegen kernel spreads the 128 wall candidates over a warp.
// ------------------------------------------------------------------------------------------------
// Exact pruning of the path test: a wall can only close a player in if it removes an edge of that
// player's shortest paths in a way that changes the distance — anchors outside q_cut_masks(player) leave
// sp(player) unchanged, hence finite.  Two layer searches + traces replace most of the ~200 flood fills
// of a mid-game position; the surviving candidates are flood-filled for the player(s) concerned only.
template <class Layers>
MC_HD void q_legal_parts(const QState &s, uint32_t &steps, uint64_t &legal_h, uint64_t &legal_v, Layers &ls, QWork *wk) {
    steps = q_step_mask(s);
    uint64_t lh, lv;
    q_cheap_wall_masks(s, lh, lv);
    legal_h = lh; legal_v = lv;
    if ((lh | lv) == 0ull) return;
    uint64_t safe_h, safe_v;
    q_detour_safe_masks(s, safe_h, safe_v);
    if (((lh & ~safe_h) | (lv & ~safe_v)) == 0ull) return;      // every candidate has its detour: nothing to search
    const QCut cw = q_block_masks(s, 0, ls, wk), cb = q_block_masks(s, 1, ls, wk);
    uint64_t th = lh & (cw.h | cb.h) & ~safe_h, tv = lv & (cw.v | cb.v) & ~safe_v;
    for (int k = q_mask_first(th, tv); k >= 0; q_mask_clear(th, tv, k), k = q_mask_first(th, tv)) {
        bool ok = true;
        if (q_mask_has(cw.h, cw.v, k)) ok = q_sp_with_wall(s, 0, k, wk) >= 0;
        if (ok && q_mask_has(cb.h, cb.v, k)) ok = q_sp_with_wall(s, 1, k, wk) >= 0;
        if (!ok) { if (k & 1) legal_v &= ~(1ull << (k >> 1)); else legal_h &= ~(1ull << (k >> 1)); }
    }
}

MC_HD void q_legal_parts(const QState &s, uint32_t &steps, uint64_t &legal_h, uint64_t &legal_v, QWork *wk) {
    LocalLayers ls;
    q_legal_parts(s, steps, legal_h, legal_v, ls, wk);
}

// the same set as a 209-bit mask over canonical move ids (include/mctsb.h)
MC_HD void q_parts_to_mask(const QState &s, uint32_t steps, uint64_t legal_h, uint64_t legal_v, uint64_t *mask4) {
    mask4[0] = mask4[1] = mask4[2] = mask4[3] = 0;
    for (int i = 0; i < 12; i++) if ((steps >> i) & 1u) { int id = s.cell[s.turn] + q_step_delta(i); mask4[id >> 6] |= 1ull << (id & 63); }
    // ids 81..144 = horizontal anchors, 145..208 = vertical anchors
    mask4[1] |= legal_h << 17; mask4[2] |= legal_h >> 47;
    mask4[2] |= legal_v << 17; mask4[3] |= legal_v >> 47;
}

MC_HD void q_legal_moves(const QState &s, uint64_t *mask4) {
    uint32_t steps; uint64_t lh, lv;
    q_legal_parts(s, steps, lh, lv, nullptr);
    q_parts_to_mask(s, steps, lh, lv, mask4);
}

// idx-th move in generate_all_moves order: steps in reversed candidate order (push_front,
// Quoridor.cpp:442-453), then anchors row-major with horizontal before vertical (:604-612).
// Returns the encoded move (cell, or 128 + k).
MC_HD int q_nth_move(const QState &s, uint32_t steps, uint64_t legal_h, uint64_t legal_v, int idx) {
    for (int i = 11; i >= 0; i--) if ((steps >> i) & 1u) { if (idx == 0) return s.cell[s.turn] + q_step_delta(i); idx--; }
    for (int a = 0; a < 64; a++) {
        if ((legal_h >> a) & 1ull) { if (idx == 0) return 128 + 2 * a; idx--; }
        if ((legal_v >> a) & 1ull) { if (idx == 0) return 128 + 2 * a + 1; idx--; }
    }
    return -1;
}

// UNI policy (no reference function; oracle = loop over the reference's public API): uniformly
// random legal move per ply, idx = mulhi(draw, |L|), ply cap MCTSB_UNI_MAX_PLIES scoring 0.5.
// `plies` comes in as the number of plies already played (the lockstep kernel hands a rollout over once nobody has a
// wall left; 0 for a whole playout) and goes out as the length of the playout.
template <class Draws, class Layers>
MC_HD double q_rollout_uni_from(QState &s, Draws &d, Layers &ls, int &plies, int &kind, QWork *wk) {
    kind = 4;
    for (;;) {
        int w = q_winner(s);
        if (w) { kind = w == 1 ? 0 : 1; return w == 1 ? 1.0 : 0.0; }
        if (plies >= MCTSB_UNI_MAX_PLIES) return 0.5;
        uint32_t steps; uint64_t lh, lv;
        q_legal_parts(s, steps, lh, lv, ls, wk);
        int n = popc32(steps) + popc64(lh) + popc64(lv);
        if (n == 0) { kind = 3; return 0.5; }
        int mv = q_nth_move(s, steps, lh, lv, (int)mulhi_u32(d.next(), (uint32_t)n));
        if (mv < 81) q_play_step(s, mv); else q_play_wall(s, mv - 128);
        plies++;
    }
}

template <class Draws, class Layers>
MC_HD double q_rollout_uni(QState &s, Draws &d, Layers &ls, int &plies, int &kind, QWork *wk) {
    plies = 0;
    return q_rollout_uni_from(s, d, ls, plies, kind, wk);
}

// The tail of a UNI playout once neither player has a wall left: the legal set is the step mask alone, so
// q_rollout_uni_from reduces to this loop (same draws, same order: reversed candidate order, Quoridor.cpp:442-453).
template <class Draws>
MC_HD double q_walk_uni(QState &s, Draws &d, int &plies, int &kind) {
    kind = 4;
    for (;;) {
        int w = q_winner(s);
        if (w) { kind = w == 1 ? 0 : 1; return w == 1 ? 1.0 : 0.0; }
        if (plies >= MCTSB_UNI_MAX_PLIES) return 0.5;
        uint32_t m = q_step_mask(s);
        int n = popc32(m);
        if (n == 0) { kind = 3; return 0.5; }
        int idx = (int)mulhi_u32(d.next(), (uint32_t)n);
        for (; idx > 0; idx--) m &= ~(1u << hibit32(m));
        q_play_step(s, s.cell[s.turn] + q_step_delta(hibit32(m)));
        plies++;
    }
}

// legal_move, Quoridor.cpp:331-342, for one canonical id
MC_HD bool q_legal_move(const QState &s, int id) {
    if (id < 0 || id >= MCTSB_Q_NUM_MOVES) return false;
    if (id < 81) {
        uint32_t steps = q_step_mask(s);
        for (int i = 0; i < 12; i++) if (((steps >> i) & 1u) && s.cell[s.turn] + q_step_delta(i) == id) return true;
        return false;
    }
    uint64_t lh, lv;
    q_cheap_wall_masks(s, lh, lv);
    int a = id < 145 ? id - 81 : id - 145, k = 2 * a + (id < 145 ? 0 : 1);
    if (!(((id < 145 ? lh : lv) >> a) & 1ull)) return false;
    return q_wall_keeps_paths(s, k, nullptr);
}

// generate_good_moves, Quoridor.cpp:518-592 — the alternative actions_to_try (compiled out of the shipped build by
// TEST_ALL_MOVES, Quoridor.cpp:7,618-627, but a public method).  The expensive part is "shortest path of both
// players with each cheap-legal wall added" (pe[k] enemy, po[k] ours; k = 2*anchor + orientation, -1 = closed in):
// independent flood fills the caller computes however it likes (one lane per wall on the device, a loop on the
// host).  What is left is this sequential pass in the reference's (row, column, h before v) order: the wall itself
// when it costs the enemy more than us and closes nobody in; and, when it would cost US at least
// MIN_ENC_FOR_STOPPING_ENEMY_WALLS = 3 more than the enemy and the enemy has walls left, the walls that pre-empt
// it (crossing wall on the same anchor, same orientation one square to either side), each at most once and only
// if fully legal.  out[] receives canonical move ids in queue order (at most MCTSB_Q_GOOD_MAX); returns the count.
MC_HD int q_good_moves_select(const QState &s, uint32_t steps, uint64_t cheap_h, uint64_t cheap_v, int our_path, int enemy_path,
                              const int8_t *pe, const int8_t *po, uint8_t *out) {
    int n = 0;
    for (int i = 11; i >= 0; i--) if ((steps >> i) & 1u) out[n++] = (uint8_t)(s.cell[s.turn] + q_step_delta(i));
    if (s.nwalls[s.turn] <= 0) return n;
    const bool enemy_has_walls = s.nwalls[1 - s.turn] > 0;
    uint64_t used[2] = {0, 0};                                   // already_used[i][j][k]
#define MCTSB_GM_CHEAP(k) ((((k) & 1 ? cheap_v : cheap_h) >> ((k) >> 1)) & 1ull)
#define MCTSB_GM_LEGAL(k) (MCTSB_GM_CHEAP(k) && pe[k] >= 0 && po[k] >= 0)
#define MCTSB_GM_TAKE(k) do { used[(k) & 1] |= 1ull << ((k) >> 1); out[n++] = (uint8_t)(((k) & 1 ? 145 : 81) + ((k) >> 1)); } while (0)
#define MCTSB_GM_USED(k) ((used[(k) & 1] >> ((k) >> 1)) & 1ull)
    for (int k = 0; k < 128; k++) {
        if (!MCTSB_GM_CHEAP(k)) continue;
        const int a = k >> 1, o = k & 1, row = a >> 3, col = a & 7;
        const int enemy_enc = pe[k] - enemy_path, our_enc = po[k] - our_path;
        if (!MCTSB_GM_USED(k) && enemy_enc > our_enc && pe[k] >= 0 && po[k] >= 0) MCTSB_GM_TAKE(k);
        if (enemy_has_walls && our_enc - enemy_enc >= 3) {
            const int kc = k ^ 1;                                // same anchor, crossing orientation
            if (!MCTSB_GM_USED(kc) && MCTSB_GM_LEGAL(kc)) MCTSB_GM_TAKE(kc);
            const int lo_ok = o == 0 ? col >= 1 : row >= 1, hi_ok = o == 0 ? col + 1 < 8 : row + 1 < 8;
            const int k_lo = o == 0 ? k - 2 : k - 16, k_hi = o == 0 ? k + 2 : k + 16;
            if (lo_ok && !MCTSB_GM_USED(k_lo) && MCTSB_GM_LEGAL(k_lo)) MCTSB_GM_TAKE(k_lo);
            if (hi_ok && !MCTSB_GM_USED(k_hi) && MCTSB_GM_LEGAL(k_hi)) MCTSB_GM_TAKE(k_hi);
        }
    }
#undef MCTSB_GM_CHEAP
#undef MCTSB_GM_LEGAL
#undef MCTSB_GM_TAKE
#undef MCTSB_GM_USED
    return n;
}

// the whole of generate_good_moves on one thread (host game class; the device spreads the flood fills over a warp)
MC_HD int q_good_moves(QState &s, uint8_t *out, QWork *wk) {
    uint64_t ch, cv;
    q_cheap_wall_masks(s, ch, cv);
    int8_t pe[128], po[128];
    const int p = s.turn, e = 1 - p;
    for (int k = 0; k < 128; k++) {
        pe[k] = po[k] = -1;
        if ((((k & 1) ? cv : ch) >> (k >> 1)) & 1ull) {
            pe[k] = (int8_t)q_sp_with_wall(s, e, k, wk);
            po[k] = (int8_t)q_sp_with_wall(s, p, k, wk);
        }
    }
    const int our_path = (ch | cv) ? q_sp(s, p, wk) : 0, enemy_path = (ch | cv) ? q_sp(s, e, wk) : 0;
    return q_good_moves_select(s, q_step_mask(s), ch, cv, our_path, enemy_path, pe, po, out);
}

MC_HD void q_play_id(QState &s, int id) {    // canonical id, already known legal
    if (id < 81) q_play_step(s, id);
    else q_play_wall(s, id < 145 ? 2 * (id - 81) : 2 * (id - 145) + 1);
}

}  // namespace mctsb
