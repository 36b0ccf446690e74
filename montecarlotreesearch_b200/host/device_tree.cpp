// device_tree.cpp — see device_tree.h.  A thin shell: every method is one or two calls of the C ABI.
#include "device_tree.h"
#include "mctsb.h"
#include <algorithm>
#include <ctime>
#include <vector>

void MCTS_device_tree::check(int st, const char *what) const {
    if (st == MCTSB_OK) return;
    std::string msg = std::string(what) + ": " + mctsb_strerror(st);
    if (st == MCTSB_ERR_CUDA || st == MCTSB_ERR_TREE_FULL) msg += std::string(" / ") + mctsb_last_cuda_error(backend->raw());
    throw BackendError(st, msg);
}

MCTS_device_tree::MCTS_device_tree(const Quoridor_state &starting_state, CudaRolloutBackend *backend, MCTS_batching batching,
                                   uint64_t max_nodes, double uct_c)
    : backend(backend), tree(nullptr), batching(batching), current(nullptr) {
    if (backend == nullptr) throw std::invalid_argument("MCTS_device_tree needs a CUDA backend");
    mctsb_qstate root;
    starting_state.pack(root);
    const uint32_t K = batching.leaves_per_flush == 0 ? 1 : batching.leaves_per_flush;
    check(mctsb_tree_create(backend->raw(), &tree, &root, starting_state.gpu_policy_id(), batching.rollouts_per_leaf, K, max_nodes, uct_c), "mctsb_tree_create");
    check(mctsb_tree_set_overlap(tree, batching.overlap ? 1 : 0), "mctsb_tree_set_overlap");
    // the state's own choice of actions_to_try travels with it, as it does from node to node on the host
    check(mctsb_tree_set_good_moves(tree, starting_state.get_good_moves_only() ? 1 : 0), "mctsb_tree_set_good_moves");
}

MCTS_device_tree::~MCTS_device_tree() {
    delete current;
    mctsb_tree_destroy(tree);
}

void MCTS_device_tree::grow_tree(int max_iter, double max_time_in_seconds) {
    const uint32_t K = batching.leaves_per_flush == 0 ? 1 : batching.leaves_per_flush;
    // whole seconds, strict, like the reference (mcts.cpp:199-204); the clock is looked at every 64 flushes
    time_t start_t, now_t;
    time(&start_t);
    int done = 0;
    while (done < max_iter) {
        const int chunk = std::min<long long>((long long)max_iter - done, 64ll * K);
        const uint64_t first = next_rollout_ids(               // the process-wide allocator MCTS_tree uses: ids are never reused
            (uint64_t)chunk * batching.rollouts_per_leaf);
        check(mctsb_tree_grow(tree, (uint32_t)chunk, K, first), "mctsb_tree_grow");
        done += chunk;
        time(&now_t);
        if (difftime(now_t, start_t) > max_time_in_seconds) break;
    }
}

Quoridor_move *MCTS_device_tree::select_best_child() const {
    int32_t id = -1;
    check(mctsb_tree_best_move(tree, &id), "mctsb_tree_best_move");
    if (id < 0) return NULL;
    return Quoridor_move::from_canonical_id(id, ((const Quoridor_state *)get_current_state())->whose_turn());
}

void MCTS_device_tree::advance_tree(const MCTS_move *move) {
    check(mctsb_tree_advance(tree, ((const Quoridor_move *)move)->canonical_id()), "mctsb_tree_advance");
    delete current; current = nullptr;
}

unsigned int MCTS_device_tree::get_size() const {
    mctsb_tree_info info;
    check(mctsb_tree_get_info(tree, &info), "mctsb_tree_get_info");
    return info.root_size;
}

uint64_t MCTS_device_tree::get_rollouts_played() const {
    mctsb_tree_info info;
    check(mctsb_tree_get_info(tree, &info), "mctsb_tree_get_info");
    return info.rollouts;
}

const MCTS_state *MCTS_device_tree::get_current_state() const {
    if (current == nullptr) {
        mctsb_qstate s;
        check(mctsb_tree_root_state(tree, &s), "mctsb_tree_root_state");
        current = new Quoridor_state(s);
    }
    return current;
}

void MCTS_device_tree::print_stats() const {
    mctsb_tree_info info;
    check(mctsb_tree_get_info(tree, &info), "mctsb_tree_get_info");
    if (info.root_sims == 0) { cout << "Tree not expanded yet" << endl; return; }
    int n = 0;
    std::vector<int32_t> ids(MCTSB_Q_NUM_MOVES); std::vector<uint64_t> sims(MCTSB_Q_NUM_MOVES); std::vector<double> scores(MCTSB_Q_NUM_MOVES);
    check(mctsb_tree_root_children(tree, ids.data(), sims.data(), scores.data(), MCTSB_Q_NUM_MOVES, &n), "mctsb_tree_root_children");
    cout << "___ INFO _______________________" << endl
         << "Tree size: " << info.root_size << endl
         << "Number of simulations: " << info.root_sims << endl
         << "Branching factor at root: " << n << endl
         << "Chances of P1 winning: " << setprecision(4) << 100.0 * (info.root_score / info.root_sims) << "%" << endl;
    const Quoridor_state *s = (const Quoridor_state *)get_current_state();
    const bool p1 = s->player1_turn();
    std::vector<int> order(n);
    for (int i = 0; i < n; i++) order[i] = i;
    auto rate = [&](int i) { return p1 ? scores[i] / sims[i] : 1.0 - scores[i] / sims[i]; };
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return rate(a) > rate(b); });
    cout << "Best moves:" << endl;
    for (int k = 0; k < n && k < 10; k++) {
        Quoridor_move *m = Quoridor_move::from_canonical_id(ids[order[k]], s->whose_turn());
        cout << "  " << k + 1 << ". " << m->sprint() << "  -->  " << setprecision(4) << 100.0 * rate(order[k]) << "%" << endl;
        delete m;
    }
    cout << "________________________________" << endl;
}
