// ttt_core.cuh — Tic-Tac-Toe on two 9-bit masks (host+device).  Replaces TicTacToe_state
// (reference examples/TicTacToe/TicTacToe.h:7-24): board[3][3] -> x_mask / o_mask, bit 3*row+col.
#pragma once
#include <stdint.h>
#include "philox.cuh"

namespace mctsb {

// player_won, TicTacToe.cpp:26-34: the eight lines as 9-bit masks
MC_HD bool t_line(uint32_t m) {
    return (m & 0x007) == 0x007 || (m & 0x038) == 0x038 || (m & 0x1c0) == 0x1c0 ||   // rows
           (m & 0x049) == 0x049 || (m & 0x092) == 0x092 || (m & 0x124) == 0x124 ||   // columns
           (m & 0x111) == 0x111 || (m & 0x054) == 0x054;                             // diagonals
}

// calculate_winner, TicTacToe.cpp:116-128: 'x' before 'o' before draw.  0 none, 1 x, 2 o, 3 draw
MC_HD int t_winner(uint32_t x, uint32_t o) {
    if (t_line(x)) return 1;
    if (t_line(o)) return 2;
    return ((x | o) & 0x1ff) == 0x1ff ? 3 : 0;
}

// TicTacToe_state::rollout, TicTacToe.cpp:73-107.  `available` holds the empty cells in DESCENDING
// order (push_front), r = rand() % size picks the r-th of them.  kind: 0 x won, 1 o won, 2 draw.
template <class Draws>
MC_HD double t_rollout(uint32_t x_mask, uint32_t o_mask, Draws &d, int &plies, int &kind) {
    uint32_t x = x_mask & 0x1ffu, o = o_mask & 0x1ffu;
    int turn = (o_mask & 0x8000u) ? 1 : 0;
    int w = t_winner(x, o);
    plies = 0;
    while (!w) {
        uint32_t free_cells = ~(x | o) & 0x1ffu;
        int na = 0;
        for (uint32_t f = free_cells; f; f &= f - 1) na++;
        uint32_t r = (d.next() & 0x7fffffffu) % (uint32_t)na;
        uint32_t f = free_cells;                       // drop the r highest free cells, take the next
        for (uint32_t i = 0; i < r; i++) { uint32_t top = 0x100u; while (!(f & top)) top >>= 1; f &= ~top; }
        uint32_t top = 0x100u; while (!(f & top)) top >>= 1;
        if (turn) o |= top; else x |= top;
        turn ^= 1; plies++;
        w = t_winner(x, o);
    }
    kind = w == 1 ? 0 : w == 2 ? 1 : 2;
    return w == 1 ? 1.0 : w == 3 ? 0.5 : 0.0;
}

}  // namespace mctsb
