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

CudaRolloutBackend::CudaRolloutBackend(int device, uint64_t seed)
    : handle(nullptr), device(device), seed(seed), submitted(0), collected(0) {
    for (int i = 0; i < MAX_HANDLES; i++) handles[i] = nullptr;
    check(nullptr, mctsb_create(&handle, device, seed), "mctsb_create");
    handles[0] = handle;
}

CudaRolloutBackend::~CudaRolloutBackend() {
    for (int i = 0; i < MAX_HANDLES; i++) if (handles[i]) mctsb_destroy(handles[i]);
}

mctsb_backend *CudaRolloutBackend::slot(unsigned i) {
    mctsb_backend *&h = handles[i % MAX_HANDLES];
    if (!h) check(nullptr, mctsb_create(&h, device, seed), "mctsb_create");   // same device and Philox key: results depend on ids only
    return h;
}

void CudaRolloutBackend::rollout_batch(int game, int policy, const void *packed_states, int n_leaves,
                                       uint32_t rollouts_per_leaf, uint64_t first_rollout_id, double *score_sum) {
    if (submitted != collected) throw BackendError(MCTSB_ERR_BUSY, "rollout_batch while submissions are in flight");
    check(handle, mctsb_rollout_batch(handle, game, policy, packed_states, n_leaves, rollouts_per_leaf, first_rollout_id,
                                      score_sum, nullptr), "mctsb_rollout_batch");
}

void CudaRolloutBackend::submit(int game, int policy, const void *packed_states, int n_leaves, uint32_t rollouts_per_leaf,
                                uint64_t first_rollout_id) {
    if (submitted - collected >= (unsigned)MAX_HANDLES) throw BackendError(MCTSB_ERR_BUSY, "too many submissions in flight");
    mctsb_backend *h = slot(submitted);
    check(h, mctsb_submit(h, game, policy, packed_states, n_leaves, rollouts_per_leaf, first_rollout_id), "mctsb_submit");
    submitted++;
}

void CudaRolloutBackend::wait(double *score_sum) {
    if (submitted == collected) throw BackendError(MCTSB_ERR_NOT_SUBMITTED, "wait without a submission");
    mctsb_backend *h = slot(collected);
    check(h, mctsb_wait(h, score_sum, nullptr), "mctsb_wait");
    collected++;
}

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
