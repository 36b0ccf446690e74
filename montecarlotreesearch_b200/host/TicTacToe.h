// TicTacToe.h — host game plugin; public interface as the reference's examples/TicTacToe/TicTacToe.h.
#ifndef MCTS_TICTACTOE_H
#define MCTS_TICTACTOE_H

#include "state.h"
#include <cstdint>


class TicTacToe_state : public MCTS_state {
    uint16_t x_mask, o_mask;   // bit 3*row+col
    char turn, winner;         // 'x'/'o'; ' ' none, 'x', 'o', 'd' draw
    char calculate_winner() const;
public:
    TicTacToe_state();
    TicTacToe_state(uint16_t x_mask, uint16_t o_mask, char turn);
    char get_turn() const;
    char get_winner() const;
    bool is_terminal() const override;
    MCTS_state *next_state(const MCTS_move *move) const override;
    queue<MCTS_move *> *actions_to_try() const override;
    double rollout() const override;          // one playout ON THE GPU backend
    void print() const override;
    bool player1_turn() const override { return turn == 'x'; }
    int gpu_game_id() const override;
    bool gpu_pack(void *dst) const override;
};


struct TicTacToe_move : public MCTS_move {
    int x, y;
    char player;
    TicTacToe_move(int x, int y, char p) : x(x), y(y), player(p) {}
    bool operator==(const MCTS_move &other) const override;
};

#endif
