// rollout_backend.h — what MCTS_node::rollout() talks to in place of the reference's
// function-local `static JobScheduler scheduler` (reference mcts/src/mcts.cpp:60).
//
// RolloutBackend is the seam; CudaRolloutBackend (rollout_backend.cpp) is the only implementation
// the product ships: it forwards to the C ABI in include/mctsb.h and therefore to the sm_100a
// kernels.  There is no CPU implementation in this package.
#ifndef MCTS_ROLLOUT_BACKEND_H
#define MCTS_ROLLOUT_BACKEND_H

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

struct mctsb_backend;

class RolloutBackend {
public:
    virtual ~RolloutBackend() = default;
    // Play `rollouts_per_leaf` independent playouts of each of `n_leaves` packed states (all of one
    // game / policy).  score_sum[i] = sum of the results of leaf i — what the reference accumulates
    // in `score_sum` (mcts.cpp:68-75) before backpropagate(score_sum, N) (mcts.cpp:76).
    virtual void rollout_batch(int game, int policy, const void *packed_states, int n_leaves,
                               uint32_t rollouts_per_leaf, uint64_t first_rollout_id, double *score_sum) = 0;
    // asynchronous pair (submit copies the states before returning).  Up to max_in_flight() submissions may be
    // outstanding; wait() returns the results of the OLDEST one.
    virtual void submit(int game, int policy, const void *packed_states, int n_leaves,
                        uint32_t rollouts_per_leaf, uint64_t first_rollout_id) = 0;
    virtual void wait(double *score_sum) = 0;
    virtual int max_in_flight() const { return 1; }
};

class BackendError : public std::runtime_error {
public:
    int status;
    BackendError(int status, const std::string &what) : std::runtime_error(what), status(status) {}
};

class CudaRolloutBackend : public RolloutBackend {
    // one C-ABI handle = one CUDA stream with its staging buffers; a second one (created on first use) lets two
    // flushes be queued on the GPU at once — a 65 536-playout flush fills only 2 048 of the 2 960 resident warps
    static const int MAX_HANDLES = 2;
    mctsb_backend *handle;                 // handles[0]
    mctsb_backend *handles[MAX_HANDLES];
    int device; uint64_t seed;
    unsigned submitted, collected;         // FIFO positions
    mctsb_backend *slot(unsigned i);
public:
    explicit CudaRolloutBackend(int device = 0, uint64_t seed = 0x2026101900000001ull);   // throws BackendError without a GPU
    ~CudaRolloutBackend() override;
    CudaRolloutBackend(const CudaRolloutBackend &) = delete;
    CudaRolloutBackend &operator=(const CudaRolloutBackend &) = delete;
    void rollout_batch(int game, int policy, const void *packed_states, int n_leaves, uint32_t rollouts_per_leaf,
                       uint64_t first_rollout_id, double *score_sum) override;
    void submit(int game, int policy, const void *packed_states, int n_leaves, uint32_t rollouts_per_leaf,
                uint64_t first_rollout_id) override;
    void wait(double *score_sum) override;
    int max_in_flight() const override { return MAX_HANDLES; }
    mctsb_backend *raw() const { return handle; }
};

// Process-wide default used by MCTS_tree / MCTS_state::rollout() when none is given: created on
// first use on device 0 (the analogue of the reference's function-local static scheduler).
RolloutBackend *default_rollout_backend();
void set_default_rollout_backend(RolloutBackend *b);    // not owned
uint64_t next_rollout_ids(uint64_t count);              // process-wide Philox stream id allocator

#endif
