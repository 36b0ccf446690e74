// mcts.h — host search tree.  Same public interface as the reference's mcts/include/mcts.h:26-79
// (MCTS_node / MCTS_tree / MCTS_agent: select, select_best_child, grow_tree, advance_tree, get_size,
// get_current_state, print_stats, genmove, feedback); written from scratch.  The tree (UCT
// selection, expansion, backpropagation) stays on the host; the Simulation phase — the reference's
// RolloutJob fan-out onto a pthread JobScheduler (mcts.h:82-91, mcts.cpp:57-81) — is replaced by
// batched flushes to the GPU rollout backend (rollout_backend.h -> include/mctsb.h).
#ifndef MCTS_H
#define MCTS_H

#include "state.h"
#include "rollout_backend.h"
#include <cstdint>
#include <iomanip>
#include <deque>
#include <queue>
#include <vector>

#define STARTING_NUMBER_OF_CHILDREN 32   // same preallocation hint as the reference (mcts.h:11)

using namespace std;

// How the Simulation phase is batched.  The reference plays NUMBER_OF_THREADS = 4 rollouts of one
// leaf per iteration (JobScheduler.h:10, mcts.cpp:62-64); here one iteration plays
// `rollouts_per_leaf` of them on the GPU, and `leaves_per_flush` > 1 selects that many distinct
// leaves (virtual loss keeps the selections apart) and flushes them in one kernel launch.
struct MCTS_batching {
    uint32_t rollouts_per_leaf = 4;
    uint32_t leaves_per_flush = 1;
    // leaves_per_flush > 1 only: select the next batch of leaves on the host while the previous flush is
    // still running on the GPU (its leaves keep their virtual loss until the results arrive)
    bool overlap = false;
    // with overlap: how many flushes may be queued on the GPU at once (1 or 2; capped by the backend).  Two
    // keep the GPU busy across the hand-over: a 65 536-playout flush occupies only 2 048 of the 2 960 resident
    // warps of a B200, the second one fills the rest and takes over when the first drains.
    uint32_t flushes_in_flight = 1;
    // Expansion prepared on the device: every flush also runs its leaves through the backend's move-generation
    // kernel, and a leaf keeps the legal-move mask that comes back — when it is selected again its move list is
    // read off the mask (same list, same order) instead of being generated on the CPU (mcts.cpp:19).
    bool gpu_expand = false;
};

class MCTS_tree;

// What UCT reads of a child, kept in one array per parent: selection walks 131 children per level at the
// root of a Quoridor tree, and with a child's statistics inside the child every one of them is a cache miss.
struct MCTS_child_stat {
    double score;
    double visits;          // simulations + virtual visits (exact: far below 2^53)
    double virtual_visits;
};

class MCTS_node {
    friend class MCTS_tree;
    bool terminal;
    unsigned int size;
    uint64_t number_of_simulations;        // the reference counts in an unsigned int (mcts.h:29): at 4 rollouts per
                                           // iteration that lasts; at 65 536 per flush it wraps within minutes
    double score;
    MCTS_state *state;
    const MCTS_move *move;
    vector<MCTS_node *> *children;
    vector<MCTS_child_stat> child_stats;   // child_stats[i] mirrors children->at(i)
    // what the ranking pass of select_best_child reads, 8 bytes per child in two dense float arrays (a Quoridor node
    // has ~131 children: 1 KB instead of the 5 KB of pointers + statistics): the exploitation term as the side
    // choosing here sees it, and n^-1/2; refreshed whenever the child's statistics change
    vector<float> rank_exploit, rank_inv_sqrt;
    bool p1_to_move;                       // state->player1_turn(), asked once
    uint32_t index_in_parent;
    MCTS_node *parent;
    // the moves not tried yet, generated on first use: as the reference's queue of moves, or — when the game offers
    // actions_to_try_ids — as one byte per move with a cursor (a move object is then only made for the move expanded)
    mutable queue<MCTS_move *> *untried_actions;
    mutable vector<uint8_t> untried_ids; mutable uint32_t next_untried_id; mutable bool listed, listed_as_ids;
    MCTS_tree *tree;                       // owner (batching parameters, backend, rollout ids); may be NULL
    uint64_t virtual_visits;               // pending simulations counted as losses while a flush is in flight
    void backpropagate(double w, uint64_t n);
    void list_actions() const;
    bool no_untried_action() const;
    void add_virtual(int64_t n);
    void mirror_to_parent();
    void rebuild_child_stats();
    MCTS_node *expand_no_rollout();        // Expansion without Simulation; NULL when nothing to expand
public:
    MCTS_node(MCTS_node *parent, MCTS_state *state, const MCTS_move *move, MCTS_tree *tree = NULL);
    ~MCTS_node();
    bool is_fully_expanded() const;
    bool is_terminal() const;
    const MCTS_move *get_move() const;
    unsigned int get_size() const;
    void expand();
    void rollout();
    MCTS_node *select_best_child(double c) const;
    MCTS_node *advance_tree(const MCTS_move *m);
    const MCTS_state *get_current_state() const;
    void print_stats() const;
    double calculate_winrate(bool player1turn) const;
    // read-only accessors (not in the reference; used by tests and the multi-GPU root merge)
    uint64_t get_number_of_simulations() const { return number_of_simulations; }
    double get_score() const { return score; }
    const vector<MCTS_node *> &get_children() const { return *children; }
};


class MCTS_tree {
    friend class MCTS_node;
    MCTS_node *root;
    RolloutBackend *backend;               // not owned; NULL = process default (created on first use)
    MCTS_batching batching;
    uint64_t flushes, rollouts_played;
    double seconds_total, seconds_backend;  // wall time inside grow_tree / of it, inside backend calls (submit, wait)
    RolloutBackend *get_backend();
    void apply_legal_masks(const vector<MCTS_node *> &leaves, size_t first, size_t count);
    void drain(deque<vector<MCTS_node *>> &inflight, vector<MCTS_node *> &leaves);
    void flush_leaves(vector<MCTS_node *> &leaves);
    bool submit_leaves(const vector<MCTS_node *> &leaves);      // asynchronous half of flush_leaves; false = not batchable
    void collect_leaves(vector<MCTS_node *> &leaves);           // wait + backpropagate
public:
    MCTS_tree(MCTS_state *starting_state);
    MCTS_tree(MCTS_state *starting_state, RolloutBackend *backend, MCTS_batching batching);
    ~MCTS_tree();
    MCTS_node *select(double c=1.41);
    MCTS_node *select_best_child();
    void grow_tree(int max_iter, double max_time_in_seconds);
    void advance_tree(const MCTS_move *move);
    unsigned int get_size() const;
    const MCTS_state *get_current_state() const;
    void print_stats() const;
    // extensions
    void set_batching(MCTS_batching b) { batching = b; }
    MCTS_batching get_batching() const { return batching; }
    const MCTS_node *get_root() const { return root; }
    uint64_t get_flushes() const { return flushes; }
    uint64_t get_rollouts_played() const { return rollouts_played; }
    // where grow_tree's wall time went: all of it, and the part spent inside backend calls (copying leaves in,
    // waiting for the GPU); the difference is the host tree — selection, expansion, backpropagation
    double get_seconds_total() const { return seconds_total; }
    double get_seconds_backend() const { return seconds_backend; }
    // Root-parallel merge: add other trees' root-child statistics, keyed by the
    // position of the child's move in the root's actions_to_try() order.
    void export_root_stats(vector<double> &visits_and_scores) const;     // 2 doubles per root action
    void import_root_stats(const vector<double> &visits_and_scores);     // REPLACES the children's statistics
    // export -> sum over the GPUs the backend is joined with (RolloutBackend::allreduce_sum: NCCL behind the C
    // ABI) -> import: afterwards select_best_child() applies the reference's final choice rule
    // (mcts.cpp:104-129 with c = 0, :268-270) to the same numbers on every rank.  No-op for a lone backend.
    void merge_root_stats();
};


class MCTS_agent {
    MCTS_tree *tree;
    int max_iter, max_seconds;
public:
    MCTS_agent(MCTS_state *starting_state, int max_iter = 100000, int max_seconds = 30);
    MCTS_agent(MCTS_state *starting_state, int max_iter, int max_seconds, RolloutBackend *backend, MCTS_batching batching);
    ~MCTS_agent();
    const MCTS_move *genmove(const MCTS_move *enemy_move);
    const MCTS_state *get_current_state() const;
    void feedback() const { tree->print_stats(); }
    MCTS_tree *get_tree() { return tree; }
};

// quiet mode: the reference prints progress lines from grow_tree/genmove (mcts.cpp:185-209,289-297)
void mcts_set_verbose(bool on);

#endif
