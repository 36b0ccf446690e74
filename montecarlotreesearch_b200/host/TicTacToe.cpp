// TicTacToe.cpp — host game plugin on two 9-bit masks (rules: csrc/ttt_core.cuh).
#include "TicTacToe.h"
#include "../csrc/ttt_core.cuh"
#include "rollout_backend.h"
#include "mctsb.h"
#include <cstdio>
#include <cstring>

TicTacToe_state::TicTacToe_state() : x_mask(0), o_mask(0), turn('x'), winner(' ') {}

TicTacToe_state::TicTacToe_state(uint16_t x, uint16_t o, char t) : x_mask(x & 0x1ff), o_mask(o & 0x1ff), turn(t) {
    winner = calculate_winner();
}

char TicTacToe_state::calculate_winner() const {       // TicTacToe.cpp:116-128
    int w = mctsb::t_winner(x_mask, o_mask);
    return w == 1 ? 'x' : w == 2 ? 'o' : w == 3 ? 'd' : ' ';
}

bool TicTacToe_state::is_terminal() const { return winner != ' '; }
char TicTacToe_state::get_turn() const { return turn; }
char TicTacToe_state::get_winner() const { return winner; }

MCTS_state *TicTacToe_state::next_state(const MCTS_move *move) const {   // TicTacToe.cpp:48-61
    const TicTacToe_move *m = (const TicTacToe_move *) move;
    uint16_t bit = (uint16_t)(1u << (3 * m->x + m->y));
    if ((x_mask | o_mask) & bit) {
        cerr << "Warning: Illegal move (" << m->x << ", " << m->y << ")" << endl;
        return NULL;
    }
    TicTacToe_state *n = new TicTacToe_state(*this);
    if (m->player == 'x') n->x_mask |= bit; else n->o_mask |= bit;
    n->winner = n->calculate_winner();
    n->turn = (turn == 'x') ? 'o' : 'x';
    return n;
}

queue<MCTS_move *> *TicTacToe_state::actions_to_try() const {            // TicTacToe.cpp:63-71: ascending cells
    queue<MCTS_move *> *Q = new queue<MCTS_move *>();
    for (int c = 0; c < 9; c++)
        if (!((x_mask | o_mask) >> c & 1)) Q->push(new TicTacToe_move(c / 3, c % 3, turn));
    return Q;
}

int TicTacToe_state::gpu_game_id() const { return MCTSB_GAME_TICTACTOE; }

bool TicTacToe_state::gpu_pack(void *dst) const {
    mctsb_tstate p;
    p.x_mask = x_mask;
    p.o_mask = (uint16_t)(o_mask | (turn == 'o' ? 0x8000 : 0));
    memcpy(dst, &p, sizeof p);
    return true;
}

double TicTacToe_state::rollout() const {                                // TicTacToe.cpp:73-107, on the GPU
    mctsb_tstate p; gpu_pack(&p);
    double sum = 0.0;
    default_rollout_backend()->rollout_batch(MCTSB_GAME_TICTACTOE, 0, &p, 1, 1, next_rollout_ids(1), &sum);
    return sum;
}

void TicTacToe_state::print() const {
    char b[9];
    for (int c = 0; c < 9; c++) b[c] = (x_mask >> c & 1) ? 'x' : (o_mask >> c & 1) ? 'o' : ' ';
    printf(" %c | %c | %c\n---+---+---\n %c | %c | %c\n---+---+---\n %c | %c | %c\n", b[0], b[1], b[2], b[3], b[4], b[5], b[6], b[7], b[8]);
}

bool TicTacToe_move::operator==(const MCTS_move &other) const {
    const TicTacToe_move &o = (const TicTacToe_move &) other;
    return x == o.x && y == o.y && player == o.player;
}
