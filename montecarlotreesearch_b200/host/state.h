// state.h — the plugin interface a game implements; source-compatible with the reference's
// mcts/include/state.h:11-35 (same class names, virtuals and ownership rules), plus an OPTIONAL
// packing hook that lets MCTS_node::rollout() hand the position to the GPU rollout backend instead
// of scheduling RolloutJobs on the pthread JobScheduler (reference mcts/src/mcts.cpp:57-81).
#ifndef MCTS_STATE_H
#define MCTS_STATE_H

#include <cstdint>
#include <iostream>
#include <queue>
#include <stdexcept>
#include <vector>
#include <string>

using namespace std;   // the reference header does this and game code written against it relies on it


struct MCTS_move {
    virtual ~MCTS_move() = default;
    virtual bool operator==(const MCTS_move& other) const = 0;
    virtual string sprint() const { return "Not implemented"; }
};


/** rollout() must return the winning chance of player1 in [0, 1] (reference state.h:19-21). */
class MCTS_state {
public:
    virtual ~MCTS_state() = default;
    virtual queue<MCTS_move *> *actions_to_try() const = 0;           // heap queue, caller owns
    virtual MCTS_state *next_state(const MCTS_move *move) const = 0;  // heap state, caller owns
    virtual double rollout() const = 0;
    virtual bool is_terminal() const = 0;
    virtual void print() const { cout << "Printing not implemented" << endl; }
    virtual bool player1_turn() const = 0;

    // ---- B200 backend hook (not in the reference) -----------------------------------------------
    // A state the rollout backend knows how to play returns its game id (MCTSB_GAME_*) and writes
    // its packed wire record (include/mctsb.h: 32 bytes Quoridor, 4 bytes Tic-Tac-Toe) to dst.
    // States that return -1 are simulated through their own virtual rollout().
    virtual int gpu_game_id() const { return -1; }
    virtual int gpu_policy_id() const { return 0; }
    virtual bool gpu_pack(void *dst) const { (void)dst; return false; }
    // Optional compact form of actions_to_try(): the same moves in the same order as small integers of the game's own
    // choosing.  A node then keeps ~130 bytes instead of ~130 heap-allocated moves, and materialises a move
    // (move_of_id) only when it expands it.  Return false if the game does not offer it.
    virtual bool actions_to_try_ids(std::vector<uint8_t> &ids) const { (void)ids; return false; }
    virtual MCTS_move *move_of_id(int id) const { (void)id; return NULL; }
    // The legal-move mask of this position as the backend's move-generation kernel computed it (4 words,
    // canonical move ids): a state that keeps it answers its next actions_to_try() from the mask instead of
    // generating the moves on the CPU.  Same list, same order.
    virtual void gpu_set_legal(const uint64_t *mask4) { (void)mask4; }
    // next_state for a move that came out of this state's own actions_to_try(): legal by construction, so a
    // game may skip the legality test the reference repeats in play_move (Quoridor.cpp:345).
    virtual MCTS_state *next_state_of_listed_move(const MCTS_move *move) const { return next_state(move); }
};


#endif
