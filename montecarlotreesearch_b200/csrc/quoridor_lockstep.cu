// quoridor_lockstep.cu — the production REF-policy rollout kernel (sm_100a).
// (kernel entry points and launch; the body lives in quoridor_lockstep_body.cuh)
#include "quoridor_lockstep_body.cuh"
#include <mutex>

namespace mctsb {

// Two builds of each kernel: DENSE caps the registers for five resident blocks per SM (a few bytes spilled; the
// extra warps pay at full load), SPARSE leaves room for four without spills — faster whenever a launch cannot
// fill more than four blocks per SM anyway.
#ifndef LS_BLOCKS_DENSE
#define LS_BLOCKS_DENSE 5
#endif
#ifndef LS_BLOCKS_SPARSE
#define LS_BLOCKS_SPARSE 4
#endif

template <bool REPLAY, int BLOCKS_PER_SM>
__global__ void __launch_bounds__(LS_THREADS, BLOCKS_PER_SM) quoridor_lockstep_kernel(RolloutParams p, unsigned long long *work_counter) {
    extern __shared__ uint32_t smem[];
    lockstep_body<REPLAY>(p, work_counter, smem);
}

static constexpr size_t LS_SMEM = (size_t)LS_SMEM_WORDS * sizeof(uint32_t);

// Launch configuration, per device: SM count and resident blocks per SM of the kernels, and the opt-in to the
// kernel's dynamic shared memory.  Filled the first time a device launches (under a lock: handles on different
// devices, driven by different host threads, are independent).
struct LsDeviceConfig { bool ready; int sms; int dense[2], sparse[2]; };   // [replay]
static std::mutex ls_config_lock;
static LsDeviceConfig ls_config[MCTSB_MAX_DEVICES];

template <class K> static cudaError_t ls_prepare(K kernel, int &blocks_per_sm) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LS_SMEM);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kernel, LS_THREADS, LS_SMEM);
}

static cudaError_t ls_device_config(LsDeviceConfig &out) {
    int dev = 0;
    cudaError_t e;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if (dev < 0 || dev >= MCTSB_MAX_DEVICES) return cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> g(ls_config_lock);
    LsDeviceConfig &c = ls_config[dev];
    if (!c.ready) {
        if ((e = cudaDeviceGetAttribute(&c.sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
        if ((e = ls_prepare(quoridor_lockstep_kernel<false, LS_BLOCKS_DENSE>, c.dense[0])) != cudaSuccess) return e;
        if ((e = ls_prepare(quoridor_lockstep_kernel<true, LS_BLOCKS_DENSE>, c.dense[1])) != cudaSuccess) return e;
        if ((e = ls_prepare(quoridor_lockstep_kernel<false, LS_BLOCKS_SPARSE>, c.sparse[0])) != cudaSuccess) return e;
        if ((e = ls_prepare(quoridor_lockstep_kernel<true, LS_BLOCKS_SPARSE>, c.sparse[1])) != cudaSuccess) return e;
        c.ready = true;
    }
    out = c;
    return cudaSuccess;
}

// Grid shape.  Large launches: one resident wave of blocks, every warp takes 32 rollouts at a time from the work
// counter until none are left.  Launches of up to two waves are bound by the most loaded SM and by the time
// one warp needs for its batch, so they are spread evenly instead: a whole multiple of the SM count of blocks,
// and every warp takes `lanes` (< 32) rollouts per batch such that all warps play the same number of batches.
void ls_grid_shape(unsigned long long n, int sms, int bps, unsigned &blocks, unsigned &lanes) {
    const int warps_per_block = LS_THREADS / 32;
    const unsigned long long resident_warps = (unsigned long long)sms * bps * warps_per_block;
    const unsigned long long batches = (n + 31) / 32;
    if (batches > 2 * resident_warps) { blocks = (unsigned)(sms * bps); lanes = 32; return; }
    if (batches <= resident_warps) {                       // less than one wave: one batch per warp
        unsigned long long need = (batches + warps_per_block - 1) / warps_per_block;
        unsigned long long b = (need + sms - 1) / sms * sms;
        if (b > (unsigned long long)sms * bps) b = (unsigned long long)sms * bps;
        blocks = (unsigned)b;
        lanes = (unsigned)((n + b * warps_per_block - 1) / (b * warps_per_block));
    } else {                                               // between one and two waves: two equal batches per warp
        blocks = (unsigned)(sms * bps);
        lanes = (unsigned)((n + resident_warps * 2 - 1) / (resident_warps * 2));
    }
    if (lanes < 1) lanes = 1;
    if (lanes > 32) lanes = 32;
}

cudaError_t launch_quoridor_lockstep(const RolloutParams &params, bool replay, unsigned long long *work_counter, cudaStream_t st) {
    if (params.n_rollouts == 0) return cudaSuccess;
    cudaError_t e;
    LsDeviceConfig c;
    if ((e = ls_device_config(c)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(work_counter, 0, sizeof(unsigned long long), st)) != cudaSuccess) return e;
    const int r = replay ? 1 : 0;
    const int dense = c.dense[r] < 1 ? 1 : c.dense[r], sparse = c.sparse[r] < 1 ? 1 : c.sparse[r];
    RolloutParams p = params;
    unsigned blocks = 0, lanes = 32;
    const int sms = params.sm_limit > 0 && (int)params.sm_limit < c.sms ? (int)params.sm_limit : c.sms;
    ls_grid_shape(p.n_rollouts, sms, dense, blocks, lanes);
    p.lanes_per_batch = lanes;
    const bool use_sparse = blocks <= (unsigned)(sms * sparse) && !params.leave_room;
    if (use_sparse) {
        if (replay) quoridor_lockstep_kernel<true, LS_BLOCKS_SPARSE><<<blocks, LS_THREADS, LS_SMEM, st>>>(p, work_counter);
        else        quoridor_lockstep_kernel<false, LS_BLOCKS_SPARSE><<<blocks, LS_THREADS, LS_SMEM, st>>>(p, work_counter);
    } else {
        if (replay) quoridor_lockstep_kernel<true, LS_BLOCKS_DENSE><<<blocks, LS_THREADS, LS_SMEM, st>>>(p, work_counter);
        else        quoridor_lockstep_kernel<false, LS_BLOCKS_DENSE><<<blocks, LS_THREADS, LS_SMEM, st>>>(p, work_counter);
    }
    return cudaGetLastError();
}

}  // namespace mctsb
