// rollout_backend.cpp — CudaRolloutBackend: the host tree's view of the C ABI (include/mctsb.h).
#include "rollout_backend.h"
#include "mctsb.h"
#include <atomic>
#include <memory>
#include <mutex>

static void check(mctsb_backend *h, int st, const char *what) {
    if (st == MCTSB_OK) return;
    std::string msg = std::string(what) + ": " + mctsb_strerror(st);
    if (st == MCTSB_ERR_CUDA && h) msg += std::string(" / ") + mctsb_last_cuda_error(h);
    throw BackendError(st, msg);
}

CudaRolloutBackend::CudaRolloutBackend(int device, uint64_t seed) : handle(nullptr) {
    check(nullptr, mctsb_create(&handle, device, seed), "mctsb_create");
}

CudaRolloutBackend::~CudaRolloutBackend() { mctsb_destroy(handle); }

void CudaRolloutBackend::rollout_batch(int game, int policy, const void *packed_states, int n_leaves,
                                       uint32_t rollouts_per_leaf, uint64_t first_rollout_id, double *score_sum) {
    check(handle, mctsb_rollout_batch(handle, game, policy, packed_states, n_leaves, rollouts_per_leaf, first_rollout_id,
                                      score_sum, nullptr), "mctsb_rollout_batch");
}

void CudaRolloutBackend::submit(int game, int policy, const void *packed_states, int n_leaves, uint32_t rollouts_per_leaf,
                                uint64_t first_rollout_id) {
    check(handle, mctsb_submit(handle, game, policy, packed_states, n_leaves, rollouts_per_leaf, first_rollout_id), "mctsb_submit");
}

void CudaRolloutBackend::wait(double *score_sum) { check(handle, mctsb_wait(handle, score_sum, nullptr), "mctsb_wait"); }

static RolloutBackend *g_default = nullptr;
static std::unique_ptr<CudaRolloutBackend> g_owned;
static std::mutex g_mu;
static std::atomic<uint64_t> g_next_id{0};

RolloutBackend *default_rollout_backend() {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_default) {
        g_owned.reset(new CudaRolloutBackend(0));     // throws without a GPU: no CPU fallback
        g_default = g_owned.get();
    }
    return g_default;
}

void set_default_rollout_backend(RolloutBackend *b) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_default = b;
}

uint64_t next_rollout_ids(uint64_t count) { return g_next_id.fetch_add(count); }
