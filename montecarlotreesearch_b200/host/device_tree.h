// device_tree.h — MCTS_tree's interface (reference mcts/include/mcts.h:56-68: grow_tree, select_best_child,
// advance_tree, get_size, get_current_state) over the search tree that lives in HBM (include/mctsb.h, mctsb_tree_*).
// Same searches as MCTS_tree with MCTS_batching{rollouts_per_leaf, leaves_per_flush, overlap}: node for node and
// double for double (tests/test_device_tree.py), but Selection, Expansion and Backpropagation run as kernels next to
// the playouts and the host thread only queues launches.  Quoridor trees (generate_all_moves, or generate_good_moves when
// the starting state says so: Quoridor_state::set_good_moves_only).
#ifndef MCTS_DEVICE_TREE_H
#define MCTS_DEVICE_TREE_H

#include "mcts.h"
#include "Quoridor.h"

struct mctsb_tree;

class MCTS_device_tree {
    CudaRolloutBackend *backend;            // not owned
    mctsb_tree *tree;
    MCTS_batching batching;
    mutable Quoridor_state *current;        // cache for get_current_state()
    void check(int status, const char *what) const;
public:
    // max_nodes: node pool (94 bytes each; a node reserves ids for all its children when it is first expanded)
    MCTS_device_tree(const Quoridor_state &starting_state, CudaRolloutBackend *backend, MCTS_batching batching,
                     uint64_t max_nodes = 1ull << 24, double uct_c = 1.41);
    ~MCTS_device_tree();
    MCTS_device_tree(const MCTS_device_tree &) = delete;
    MCTS_device_tree &operator=(const MCTS_device_tree &) = delete;
    void grow_tree(int max_iter, double max_time_in_seconds);      // mcts.cpp:182-210; the time limit is checked between flushes
    Quoridor_move *select_best_child() const;                       // the move of MCTS_tree::select_best_child() (mcts.cpp:268-270); caller owns it; NULL = no child
    void advance_tree(const MCTS_move *move);                       // mcts.cpp:260-264
    unsigned int get_size() const;
    const MCTS_state *get_current_state() const;
    void print_stats() const;                                       // mcts.cpp:222-250 (does not re-sort anything: the children stay in creation order)
    uint64_t get_rollouts_played() const;
};

#endif
