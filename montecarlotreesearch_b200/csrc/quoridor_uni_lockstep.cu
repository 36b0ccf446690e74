// quoridor_uni_lockstep.cu — the production UNI-policy rollout path (sm_100a): kernel entry points and launch of
// phase A (the wall phase on the lockstep machinery, body in quoridor_uni_lockstep_body.cuh) followed by phase B
// (quoridor_uni_walk_kernel, rollout_kernels.cu) on the same stream.
#include "quoridor_uni_lockstep_body.cuh"
#include <mutex>

namespace mctsb {

#ifndef UNI_BLOCKS_PER_SM
#define UNI_BLOCKS_PER_SM 4
#endif

template <bool REPLAY>
__global__ void __launch_bounds__(LS_THREADS, UNI_BLOCKS_PER_SM) quoridor_uni_lockstep_kernel(RolloutParams p, unsigned long long *work_counter) {
    extern __shared__ uint32_t smem[];
    uni_lockstep_body<REPLAY>(p, reinterpret_cast<UniWalk *>(p.walk), work_counter, smem);
}

static constexpr size_t UNI_SMEM = (size_t)LS_SMEM_WORDS * sizeof(uint32_t);

void ls_grid_shape(unsigned long long n, int sms, int bps, unsigned &blocks, unsigned &lanes);   // quoridor_lockstep.cu

// per-device launch configuration, filled on the device's first launch (under a lock)
struct UniDeviceConfig { bool ready; int sms; int bps[2]; };
static std::mutex uni_config_lock;
static UniDeviceConfig uni_config[MCTSB_MAX_DEVICES];

template <class K> static cudaError_t uni_prepare(K kernel, int &blocks_per_sm) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UNI_SMEM);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kernel, LS_THREADS, UNI_SMEM);
}

static cudaError_t uni_device_config(UniDeviceConfig &out) {
    int dev = 0;
    cudaError_t e;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if (dev < 0 || dev >= MCTSB_MAX_DEVICES) return cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> g(uni_config_lock);
    UniDeviceConfig &c = uni_config[dev];
    if (!c.ready) {
        if ((e = cudaDeviceGetAttribute(&c.sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
        if ((e = uni_prepare(quoridor_uni_lockstep_kernel<false>, c.bps[0])) != cudaSuccess) return e;
        if ((e = uni_prepare(quoridor_uni_lockstep_kernel<true>, c.bps[1])) != cudaSuccess) return e;
        c.ready = true;
    }
    out = c;
    return cudaSuccess;
}

cudaError_t launch_quoridor_uni_lockstep(const RolloutParams &params, bool replay, unsigned long long *work_counter, cudaStream_t st) {
    if (params.n_rollouts == 0) return cudaSuccess;
    if (!params.walk) return cudaErrorInvalidValue;
    cudaError_t e;
    UniDeviceConfig c;
    if ((e = uni_device_config(c)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(work_counter, 0, sizeof(unsigned long long), st)) != cudaSuccess) return e;
    const int bps = c.bps[replay ? 1 : 0] < 1 ? 1 : c.bps[replay ? 1 : 0];
    RolloutParams p = params;
    unsigned blocks = 0, lanes = 32;
    ls_grid_shape(p.n_rollouts, c.sms, bps, blocks, lanes);
    p.lanes_per_batch = lanes;
    if (replay) quoridor_uni_lockstep_kernel<true><<<blocks, LS_THREADS, UNI_SMEM, st>>>(p, work_counter);
    else        quoridor_uni_lockstep_kernel<false><<<blocks, LS_THREADS, UNI_SMEM, st>>>(p, work_counter);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    return launch_quoridor_uni_walk(p, replay, work_counter + 1, c.sms, st);     // the walk's own counter: the slot after phase A's
}

}  // namespace mctsb
