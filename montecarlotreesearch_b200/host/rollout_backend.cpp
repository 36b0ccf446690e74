// rollout_backend.cpp — CudaRolloutBackend: the host tree's view of the C ABI (include/mctsb.h).
#include "rollout_backend.h"
#include "mctsb.h"
#include <atomic>
#include <memory>
#include <mutex>

static void check(mctsb_backend *h, int st, const char *what) {
    if (st == MCTSB_OK) return;
    std::string msg = std::string(what) + ": " + mctsb_strerror(st);
    if ((st == MCTSB_ERR_CUDA || st == MCTSB_ERR_COMM) && h) msg += std::string(" / ") + mctsb_last_cuda_error(h);
    throw BackendError(st, msg);
}

CudaRolloutBackend::CudaRolloutBackend(int device, uint64_t seed)
    : handle(nullptr), device(device), seed(seed), submitted(0), collected(0), expand(false), have_legal(false) {
    for (int i = 0; i < MAX_HANDLES; i++) { handles[i] = nullptr; slot_leaves[i] = 0; slot_legal[i] = false; }
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
    submit(game, policy, packed_states, n_leaves, rollouts_per_leaf, first_rollout_id);
    wait(score_sum);
}

void CudaRolloutBackend::submit(int game, int policy, const void *packed_states, int n_leaves, uint32_t rollouts_per_leaf,
                                uint64_t first_rollout_id) {
    if (submitted - collected >= (unsigned)MAX_HANDLES) throw BackendError(MCTSB_ERR_BUSY, "too many submissions in flight");
    mctsb_backend *h = slot(submitted);
    const bool want_legal = expand && game == MCTSB_GAME_QUORIDOR;
    check(h, mctsb_submit_x(h, game, policy, packed_states, n_leaves, rollouts_per_leaf, first_rollout_id,
                            want_legal ? MCTSB_SUBMIT_LEGAL : 0u), "mctsb_submit_x");
    slot_leaves[submitted % MAX_HANDLES] = n_leaves; slot_legal[submitted % MAX_HANDLES] = want_legal;
    submitted++;
}

void CudaRolloutBackend::wait(double *score_sum) {
    if (submitted == collected) throw BackendError(MCTSB_ERR_NOT_SUBMITTED, "wait without a submission");
    mctsb_backend *h = slot(collected);
    have_legal = slot_legal[collected % MAX_HANDLES];
    if (have_legal) legal.resize((size_t)slot_leaves[collected % MAX_HANDLES] * 4);
    check(h, mctsb_wait_x(h, score_sum, nullptr, have_legal ? legal.data() : nullptr), "mctsb_wait_x");
    collected++;
}

static RolloutBackend *g_default = nullptr;
std::vector<unsigned char> CudaRolloutBackend::comm_unique_id() {
    std::vector<unsigned char> id(MCTSB_COMM_ID_BYTES);
    check(nullptr, mctsb_comm_unique_id(id.data()), "mctsb_comm_unique_id");
    return id;
}

void CudaRolloutBackend::comm_init_rank(const void *id128, int world, int rank) {
    check(handle, mctsb_comm_init_rank(handle, id128, world, rank), "mctsb_comm_init_rank");
}

void CudaRolloutBackend::comm_init_all(const std::vector<CudaRolloutBackend *> &backends) {
    std::vector<mctsb_backend *> h;
    for (CudaRolloutBackend *b : backends) h.push_back(b->handle);
    check(h.empty() ? nullptr : h[0], mctsb_comm_init_all(h.data(), (int)h.size()), "mctsb_comm_init_all");
}

void CudaRolloutBackend::allreduce_sum_group(const std::vector<CudaRolloutBackend *> &backends, const std::vector<double *> &vecs, int n) {
    std::vector<mctsb_backend *> h;
    for (CudaRolloutBackend *b : backends) {
        if (b->submitted != b->collected) throw BackendError(MCTSB_ERR_BUSY, "allreduce while submissions are in flight");
        h.push_back(b->handle);
    }
    std::vector<double *> v(vecs);
    check(h.empty() ? nullptr : h[0], mctsb_allreduce_root_stats_group(h.data(), v.data(), (int)h.size(), n), "mctsb_allreduce_root_stats_group");
}

int CudaRolloutBackend::world() const { int w = 1; mctsb_comm_info(handle, &w, nullptr); return w; }
int CudaRolloutBackend::rank() const { int r = 0; mctsb_comm_info(handle, nullptr, &r); return r; }

void CudaRolloutBackend::allreduce_sum(double *v, int n) {
    if (world() == 1) return;
    if (submitted != collected) throw BackendError(MCTSB_ERR_BUSY, "allreduce while submissions are in flight");
    check(handle, mctsb_allreduce_root_stats(handle, v, n), "mctsb_allreduce_root_stats");
}

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
void set_next_rollout_id(uint64_t id) { g_next_id.store(id); }
