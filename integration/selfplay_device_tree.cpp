// selfplay_device_tree.cpp — full-game Quoridor MCTS self-play as a plain C++ program with the tree itself on the GPU:
// MCTS_device_tree (montecarlotreesearch_b200/host/device_tree.h) has the interface of the reference's MCTS_tree
// (mcts/include/mcts.h:56-68), the loop below is the reference's own self-play shape (examples/Quoridor/main.cpp:
// 129-169: grow, pick, advance) — only the class name differs.  With `host` as the last argument the same loop runs
// on MCTS_tree (the tree on this host thread, rollouts on the GPU): both play the same game, move for move.
//
//   g++ -O2 -std=c++17 -pthread -Iinclude -Imontecarlotreesearch_b200/host integration/selfplay_device_tree.cpp \
//       -Lmontecarlotreesearch_b200 -lmcts_host -lmctsb -Wl,-rpath,$PWD/montecarlotreesearch_b200 -o selfplay_device_tree
//   ./selfplay_device_tree [iterations/move=16384] [leaves/flush=4096] [rollouts/leaf=16] [max moves=300] [device|host]
#include "device_tree.h"
#include "mctsb.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

template <class Tree> static int play(Tree &tree, int iters, int max_moves, std::vector<int> &moves, bool &terminal) {
    terminal = tree.get_current_state()->is_terminal();
    while (!terminal && (int)moves.size() < max_moves) {
        tree.grow_tree(iters, 1e9);
        int id = -1;
        if constexpr (std::is_same<Tree, MCTS_tree>::value) {
            MCTS_node *best = tree.select_best_child();
            if (best == NULL) break;
            id = ((const Quoridor_move *)best->get_move())->canonical_id();
            Quoridor_move *m = Quoridor_move::from_canonical_id(id, ((const Quoridor_state *)tree.get_current_state())->whose_turn());
            tree.advance_tree(m);
            delete m;
        } else {
            Quoridor_move *m = tree.select_best_child();
            if (m == NULL) break;
            id = m->canonical_id();
            tree.advance_tree(m);
            delete m;
        }
        moves.push_back(id);
        terminal = tree.get_current_state()->is_terminal();
    }
    return 0;
}

int main(int argc, char **argv) {
    int have = 0;
    if (mctsb_device_count(&have) != MCTSB_OK || have < 1) { fprintf(stderr, "no CUDA device: this program has no CPU path\n"); return 2; }
    const int iters = argc > 1 ? atoi(argv[1]) : 16384;
    MCTS_batching b;
    b.leaves_per_flush = argc > 2 ? (uint32_t)atoi(argv[2]) : 4096u;
    b.rollouts_per_leaf = argc > 3 ? (uint32_t)atoi(argv[3]) : 16u;
    const int max_moves = argc > 4 ? atoi(argv[4]) : 300;
    const bool on_host = argc > 5 && strcmp(argv[5], "host") == 0;
    mcts_set_verbose(false);
    try {
        CudaRolloutBackend gpu(0);
        std::vector<int> moves;
        bool terminal = false;
        uint64_t rollouts = 0;
        const auto t0 = std::chrono::steady_clock::now();
        if (on_host) {
            MCTS_tree tree(new Quoridor_state(), &gpu, b);
            play(tree, iters, max_moves, moves, terminal);
            rollouts = tree.get_rollouts_played();
        } else {
            MCTS_device_tree tree(Quoridor_state(), &gpu, b, 1ull << 26);
            play(tree, iters, max_moves, moves, terminal);
            rollouts = tree.get_rollouts_played();
        }
        const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::string list;
        for (size_t i = 0; i < moves.size(); i++) list += (i ? "," : "") + std::to_string(moves[i]);
        printf("{\"tree\": \"%s\", \"moves\": %zu, \"terminal\": %s, \"rollouts_total\": %llu, \"wall_s\": %.4f, \"rollouts_per_sec\": %.1f, "
               "\"leaves_per_flush\": %u, \"rollouts_per_leaf\": %u, \"move_ids\": [%s]}\n",
               on_host ? "host (MCTS_tree)" : "device (MCTS_device_tree)", moves.size(), terminal ? "true" : "false", (unsigned long long)rollouts, wall,
               rollouts / wall, b.leaves_per_flush, b.rollouts_per_leaf, list.c_str());
    } catch (const std::exception &e) { fprintf(stderr, "%s\n", e.what()); return 3; }
    return 0;
}
