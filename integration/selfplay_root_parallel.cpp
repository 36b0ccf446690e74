// selfplay_root_parallel.cpp — BASELINE.json configs[3] / configs[4] as a plain C++ program: no Python, no
// torch.distributed.  One process, one host thread and one MCTS_tree per GPU (root-parallel: same root, different
// Philox keys), every rollout on the B200 through the C ABI; after each search the trees' root-child statistics are
// summed over the GPUs (MCTS_tree::merge_root_stats -> mctsb_allreduce_root_stats, NCCL over NVLink) and every tree
// applies the reference's final choice rule (mcts.cpp:104-129 with c = 0, :268-270) to identical numbers.
// The loop is the reference's own self-play shape (examples/Quoridor/main.cpp:129-169: grow, pick, advance).
//
//   g++ -O2 -std=c++17 -pthread -Iinclude -Imontecarlotreesearch_b200/host integration/selfplay_root_parallel.cpp \
//       -Lmontecarlotreesearch_b200 -lmcts_host -lmctsb -Wl,-rpath,$PWD/montecarlotreesearch_b200 -o selfplay_cpp
//   ./selfplay_cpp [gpus=all] [iterations/move=2048] [leaves/flush=16] [rollouts/leaf=4096] [max moves=300]
#include "mcts.h"
#include "Quoridor.h"
#include "rollout_backend.h"
#include "mctsb.h"
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace {
struct Barrier {                               // every tree waits here before it reads the others' choice
    std::mutex m; std::condition_variable cv; int n, waiting = 0, generation = 0;
    explicit Barrier(int n) : n(n) {}
    void wait() {
        std::unique_lock<std::mutex> lk(m);
        int g = generation;
        if (++waiting == n) { waiting = 0; generation++; cv.notify_all(); }
        else cv.wait(lk, [&] { return g != generation; });
    }
};
}  // namespace

int main(int argc, char **argv) {
    int have = 0;
    if (mctsb_device_count(&have) != MCTSB_OK || have < 1) { fprintf(stderr, "no CUDA device: this program has no CPU path\n"); return 2; }
    const int gpus = argc > 1 && atoi(argv[1]) > 0 ? atoi(argv[1]) : have;
    const int iters = argc > 2 ? atoi(argv[2]) : 2048;
    MCTS_batching batching;
    batching.leaves_per_flush = argc > 3 ? (uint32_t)atoi(argv[3]) : 16u;
    batching.rollouts_per_leaf = argc > 4 ? (uint32_t)atoi(argv[4]) : 4096u;
    batching.gpu_expand = true;
    const int max_moves = argc > 5 ? atoi(argv[5]) : 300;
    if (gpus > have) { fprintf(stderr, "%d GPUs asked for, %d present\n", gpus, have); return 2; }
    mcts_set_verbose(false);

    std::vector<std::unique_ptr<CudaRolloutBackend>> backends;
    std::vector<CudaRolloutBackend *> raw;
    try {
        for (int g = 0; g < gpus; g++) { backends.emplace_back(new CudaRolloutBackend(g, 0x2026101900000001ull + (uint64_t)g)); raw.push_back(backends.back().get()); }
        if (gpus > 1) CudaRolloutBackend::comm_init_all(raw);
    } catch (const std::exception &e) { fprintf(stderr, "%s\n", e.what()); return 3; }

    std::vector<int> choice((size_t)gpus, -1);
    std::vector<int> moves;
    std::atomic<bool> failed(false);
    std::vector<uint64_t> rollouts((size_t)gpus, 0);
    std::vector<double> merge_seconds((size_t)gpus, 0.0);
    Barrier barrier(gpus);
    bool terminal = false;
    const auto t0 = std::chrono::steady_clock::now();
    auto worker = [&](int g) {
        try {
            MCTS_tree tree(new Quoridor_state(), raw[(size_t)g], batching);
            for (int mv = 0; mv < max_moves; mv++) {
                tree.grow_tree(iters, 1e9);
                const auto a = std::chrono::steady_clock::now();
                tree.merge_root_stats();                        // blocks until every GPU has contributed
                merge_seconds[(size_t)g] += std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count();
                MCTS_node *best = tree.select_best_child();
                choice[(size_t)g] = best ? ((const Quoridor_move *)best->get_move())->canonical_id() : -1;
                barrier.wait();
                for (int o = 0; o < gpus; o++) if (choice[(size_t)o] != choice[0]) failed = true;   // identical statistics, identical rule
                if (failed || choice[0] < 0) break;
                if (g == 0) moves.push_back(choice[0]);
                tree.advance_tree(best->get_move());
                const bool over = tree.get_current_state()->is_terminal();
                barrier.wait();
                if (over) { if (g == 0) terminal = true; break; }
            }
            rollouts[(size_t)g] = tree.get_rollouts_played();
        } catch (const std::exception &e) { fprintf(stderr, "gpu %d: %s\n", g, e.what()); failed = true; }
    };
    std::vector<std::thread> threads;
    for (int g = 0; g < gpus; g++) threads.emplace_back(worker, g);
    for (auto &t : threads) t.join();
    const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (failed) { fprintf(stderr, "the trees disagreed on a move or a backend call failed\n"); return 4; }
    uint64_t total = 0;
    for (uint64_t r : rollouts) total += r;
    printf("{\"program\": \"integration/selfplay_root_parallel.cpp\", \"n_gpus\": %d, \"moves\": %zu, \"terminal\": %s, "
           "\"iterations_per_move\": %d, \"leaves_per_flush\": %u, \"rollouts_per_leaf\": %u, \"rollouts_total\": %llu, "
           "\"rollouts_per_sec\": %.1f, \"wall_s\": %.3f, \"merge_ms_per_move\": %.3f, \"same_move_on_every_gpu\": true}\n",
           gpus, moves.size(), terminal ? "true" : "false", iters, batching.leaves_per_flush, batching.rollouts_per_leaf,
           (unsigned long long)total, (double)total / wall, wall, moves.empty() ? 0.0 : 1e3 * merge_seconds[0] / (double)moves.size());
    return 0;
}
