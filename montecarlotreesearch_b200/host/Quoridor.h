// Quoridor.h — host game plugin: Quoridor on bitboards.  Public interface as the reference's
// examples/Quoridor/Quoridor.h:15-88 (Quoridor_move, Quoridor_state and its public methods), written
// from scratch on top of the rules core shared with the kernels (csrc/qcore.cuh).  rollout() does
// NOT play on the CPU: it sends the position to the GPU rollout backend.
#ifndef MCTS_QUORIDOR_H
#define MCTS_QUORIDOR_H

#include "state.h"
#include "../csrc/qcore.cuh"
#include <forward_list>
#include <vector>


struct Quoridor_move : public MCTS_move {
    short int x, y;
    char player;       // 'W' / 'B'
    char type;         // 'h' / 'v' wall, anything else a pawn step
    Quoridor_move(short int x, short int y, char p, char t) : x(x), y(y), player(p), type(t) {}
    bool operator==(const MCTS_move &other) const override {
        const Quoridor_move &o = (const Quoridor_move &) other;
        return x == o.x && y == o.y && player == o.player && type == o.type;
    }
    string sprint() const override;    // "White moves to E2", same wording as Quoridor.h:24-28
    int canonical_id() const;          // include/mctsb.h move ids (0..208)
    static Quoridor_move *from_canonical_id(int id, char player);
};


class Quoridor_state : public MCTS_state {
    mutable mctsb::QState s;           // mutable: the cached shortest paths are filled lazily, as wdists/bdists are
    int rollout_policy;                // MCTSB_POLICY_REF (default) or MCTSB_POLICY_UNI
    bool good_moves_only;              // actions_to_try() = generate_good_moves() (the build without TEST_ALL_MOVES, Quoridor.cpp:7,618-627)
    uint64_t legal_mask[4]; bool have_legal_mask;   // the device's answer to generate_all_moves for this position, if it came
public:
    Quoridor_state();
    explicit Quoridor_state(const mctsb_qstate &packed, int rollout_policy = 0);
    char whose_turn() const { return s.turn ? 'B' : 'W'; }
    unsigned int get_number_of_turns() const { return s.move_counter; }
    char check_winner() const;                                   // Quoridor.cpp:40-44
    short int remaining_walls(char p) const { return (p == 'W') ? s.nwalls[0] : s.nwalls[1]; }
    bool legal_move(const Quoridor_move *move) const;            // Quoridor.cpp:331-342
    bool play_move(const Quoridor_move *move);                   // Quoridor.cpp:344-392
    int get_shortest_path(char player, const Quoridor_move *extra_wall_move = NULL, short int posx = -1, short int posy = -1) const;
    forward_list<MCTS_move *> get_legal_step_moves(char p) const;   // reversed candidate order (Quoridor.cpp:438-455)
    vector<MCTS_move *> get_legal_step_moves2(char p) const;        // forward candidate order  (Quoridor.cpp:457-474)
    Quoridor_move *get_best_step_move(char player) const;           // Quoridor.cpp:476-497
    queue<MCTS_move *> *generate_all_moves() const;                 // Quoridor.cpp:594-616
    int all_move_ids(uint8_t *ids) const;                           // the same list as canonical ids (room for 209)
    queue<MCTS_move *> *generate_good_moves() const;                // Quoridor.cpp:518-592
    // which of the two actions_to_try() returns; inherited by every next_state (the reference chooses at compile time)
    void set_good_moves_only(bool on) { good_moves_only = on; have_legal_mask = false; }
    bool get_good_moves_only() const { return good_moves_only; }
    double evaluate() const;                                        // evaluate_position, Quoridor.cpp:629-664
    bool force_playwall() const;                                    // Quoridor.cpp:666-678
    void pack(mctsb_qstate &out) const;
    void set_rollout_policy(int p) { rollout_policy = p; }
    /** Overrides **/
    bool is_terminal() const override;
    MCTS_state *next_state(const MCTS_move *move) const override;
    queue<MCTS_move *> *actions_to_try() const override;
    double rollout() const override;                                // one playout ON THE GPU backend
    void print() const override;
    bool player1_turn() const override { return s.turn == 0; }
    int gpu_game_id() const override { return MCTSB_GAME_QUORIDOR; }
    int gpu_policy_id() const override { return rollout_policy; }
    bool gpu_pack(void *dst) const override;
    void gpu_set_legal(const uint64_t *mask4) override;
    bool actions_to_try_ids(std::vector<uint8_t> &ids) const override;      // canonical move ids, actions_to_try() order
    MCTS_move *move_of_id(int id) const override;
    MCTS_state *next_state_of_listed_move(const MCTS_move *move) const override;
};

#endif
