// mctsb_comm.cu — the one exchange step of the multi-GPU rollout path, behind the C ABI (include/mctsb.h).
//
// The path shards with no data-path collective: rollouts are keyed by id, every GPU plays its
// own id range.  What crosses NVLink is a few hundred bytes per step or per move:
//   * root-parallel trees: the root-child statistics, 2 doubles per root action (MCTS_tree::export_root_stats),
//     summed over ranks before every rank applies the reference's final choice rule (mcts.cpp:104-129 with
//     c = 0, :268-270) to identical numbers;
//   * sharded throughput runs: the exact per-leaf accumulators {lo, hi, n}, summed as 64-bit integers.
// NCCL is loaded at run time (dlopen of libnccl.so.2 — the copy the process already has when PyTorch is
// around, the system one otherwise), so libmctsb.so itself has no link-time dependency on it and a
// single-GPU user never touches it.  A failed NCCL call is reported as MCTSB_ERR_COMM with NCCL's own text in
// mctsb_last_cuda_error.
#include "../../include/mctsb.h"
#include "mctsb_backend_internal.cuh"
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <mutex>
#include <stdio.h>
#include <string.h>

namespace {

struct NcclApi {
    void *lib = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_lock;

bool load_nccl() {
    std::lock_guard<std::mutex> g(g_nccl_lock);
    if (g_nccl.lib) return true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return false;
    NcclApi a; a.lib = h;
#define SYM(name) a.name = (decltype(a.name))dlsym(h, "nccl" #name); if (!a.name) { dlclose(h); return false; }
    SYM(GetUniqueId) SYM(CommInitRank) SYM(CommInitAll) SYM(AllReduce) SYM(CommDestroy) SYM(GetErrorString) SYM(GroupStart) SYM(GroupEnd)
#undef SYM
    g_nccl = a;
    return true;
}

int fail_nccl(mctsb_backend *b, ncclResult_t r, const char *what) {
    if (b) snprintf(b->cuda_err, sizeof b->cuda_err, "%s: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "NCCL error");
    return MCTSB_ERR_COMM;
}
#define NK(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) return fail_nccl(b, r_, #call); } while (0)
#define CKC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { if (b) snprintf(b->cuda_err, sizeof b->cuda_err, "%s: %s", #call, cudaGetErrorString(e_)); return MCTSB_ERR_CUDA; } } while (0)

int ensure_comm_buffer(mctsb_backend *b, size_t bytes) {
    if (bytes <= b->cap_comm) return MCTSB_OK;
    if (b->d_comm) CKC(cudaFree(b->d_comm));
    b->d_comm = nullptr; b->cap_comm = 0;
    CKC(cudaMalloc(&b->d_comm, bytes));
    b->cap_comm = bytes;
    return MCTSB_OK;
}

// in-place sum over the ranks of the handle's communicator, through a device staging buffer on the handle's stream
int allreduce_bytes(mctsb_backend *b, void *host_buf, size_t count, ncclDataType_t type, size_t elem) {
    if (!b || !host_buf || count == 0) return MCTSB_ERR_BAD_ARG;
    if (!b->comm) return MCTSB_ERR_NOT_SUBMITTED;              // no communicator: mctsb_comm_init_* first
    if (b->pending) return MCTSB_ERR_BUSY;
    CKC(cudaSetDevice(b->device));
    int st = ensure_comm_buffer(b, count * elem);
    if (st != MCTSB_OK) return st;
    CKC(cudaMemcpyAsync(b->d_comm, host_buf, count * elem, cudaMemcpyHostToDevice, b->stream));
    NK(g_nccl.AllReduce(b->d_comm, b->d_comm, count, type, ncclSum, (ncclComm_t)b->comm, b->stream));
    CKC(cudaMemcpyAsync(host_buf, b->d_comm, count * elem, cudaMemcpyDeviceToHost, b->stream));
    CKC(cudaStreamSynchronize(b->stream));
    return MCTSB_OK;
}

}  // namespace

extern "C" {

int mctsb_comm_unique_id(void *id128) {
    if (!id128) return MCTSB_ERR_BAD_ARG;
    if (!load_nccl()) return MCTSB_ERR_COMM;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return MCTSB_ERR_COMM;
    static_assert(sizeof id == MCTSB_COMM_ID_BYTES, "NCCL unique id size");
    memcpy(id128, &id, sizeof id);
    return MCTSB_OK;
}

int mctsb_comm_init_rank(mctsb_backend *b, const void *id128, int world, int rank) {
    if (!b || !id128 || world < 1 || rank < 0 || rank >= world) return MCTSB_ERR_BAD_ARG;
    if (b->comm) return MCTSB_ERR_BUSY;
    if (!load_nccl()) { snprintf(b->cuda_err, sizeof b->cuda_err, "libnccl.so.2 not found"); return MCTSB_ERR_COMM; }
    CKC(cudaSetDevice(b->device));
    ncclUniqueId id; memcpy(&id, id128, sizeof id);
    ncclComm_t c = nullptr;
    NK(g_nccl.CommInitRank(&c, world, id, rank));
    b->comm = c; b->comm_world = world; b->comm_rank = rank;
    return MCTSB_OK;
}

int mctsb_comm_init_all(mctsb_backend **handles, int n) {
    if (!handles || n < 1 || n > 64) return MCTSB_ERR_BAD_ARG;
    for (int i = 0; i < n; i++) if (!handles[i] || handles[i]->comm) return MCTSB_ERR_BAD_ARG;
    mctsb_backend *b = handles[0];
    if (!load_nccl()) { snprintf(b->cuda_err, sizeof b->cuda_err, "libnccl.so.2 not found"); return MCTSB_ERR_COMM; }
    int devs[64]; ncclComm_t comms[64];
    for (int i = 0; i < n; i++) devs[i] = handles[i]->device;
    NK(g_nccl.CommInitAll(comms, n, devs));
    for (int i = 0; i < n; i++) { handles[i]->comm = comms[i]; handles[i]->comm_world = n; handles[i]->comm_rank = i; }
    return MCTSB_OK;
}

int mctsb_comm_destroy(mctsb_backend *b) {
    if (!b) return MCTSB_ERR_BAD_ARG;
    if (b->comm) { g_nccl.CommDestroy((ncclComm_t)b->comm); b->comm = nullptr; }
    if (b->d_comm) { cudaSetDevice(b->device); cudaFree(b->d_comm); b->d_comm = nullptr; b->cap_comm = 0; }
    return MCTSB_OK;
}

int mctsb_comm_info(const mctsb_backend *b, int *world, int *rank) {
    if (!b) return MCTSB_ERR_BAD_ARG;
    if (world) *world = b->comm ? b->comm_world : 1;
    if (rank) *rank = b->comm ? b->comm_rank : 0;
    return MCTSB_OK;
}

int mctsb_allreduce_root_stats(mctsb_backend *b, double *stats, int n) {
    if (n <= 0) return MCTSB_ERR_BAD_ARG;
    return allreduce_bytes(b, stats, (size_t)n, ncclDouble, sizeof(double));
}

int mctsb_allreduce_leaf_sums(mctsb_backend *b, mctsb_leaf_sum *sums, int n_leaves) {
    if (n_leaves <= 0) return MCTSB_ERR_BAD_ARG;
    static_assert(sizeof(mctsb_leaf_sum) == 3 * sizeof(uint64_t), "leaf sums are three 64-bit integers");
    return allreduce_bytes(b, sums, (size_t)n_leaves * 3, ncclUint64, sizeof(uint64_t));
}

// One process driving several handles from ONE thread (e.g. the C++ host with a tree per GPU): the reductions
// of all handles must be issued as one NCCL group, or the first one would wait for peers that are never called.
int mctsb_allreduce_root_stats_group(mctsb_backend **handles, double **stats, int n_handles, int n) {
    if (!handles || !stats || n_handles < 1 || n <= 0) return MCTSB_ERR_BAD_ARG;
    mctsb_backend *b = handles[0];
    for (int i = 0; i < n_handles; i++) {
        b = handles[i];
        if (!b || !stats[i]) return MCTSB_ERR_BAD_ARG;
        if (!b->comm) return MCTSB_ERR_NOT_SUBMITTED;
        if (b->pending) return MCTSB_ERR_BUSY;
        CKC(cudaSetDevice(b->device));
        int st = ensure_comm_buffer(b, (size_t)n * sizeof(double));
        if (st != MCTSB_OK) return st;
        CKC(cudaMemcpyAsync(b->d_comm, stats[i], (size_t)n * sizeof(double), cudaMemcpyHostToDevice, b->stream));
    }
    NK(g_nccl.GroupStart());
    for (int i = 0; i < n_handles; i++) {
        b = handles[i];
        ncclResult_t r = g_nccl.AllReduce(b->d_comm, b->d_comm, (size_t)n, ncclDouble, ncclSum, (ncclComm_t)b->comm, b->stream);
        if (r != ncclSuccess) { g_nccl.GroupEnd(); return fail_nccl(b, r, "ncclAllReduce"); }
    }
    NK(g_nccl.GroupEnd());
    for (int i = 0; i < n_handles; i++) {
        b = handles[i];
        CKC(cudaSetDevice(b->device));
        CKC(cudaMemcpyAsync(stats[i], b->d_comm, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
        CKC(cudaStreamSynchronize(b->stream));
    }
    return MCTSB_OK;
}

}  // extern "C"
