// quoridor_lockstep.cu — the production REF-policy rollout kernel (sm_100a).
// (kernel entry points and launch; the body lives in quoridor_lockstep_body.cuh)
#include "quoridor_lockstep_body.cuh"

namespace mctsb {

#ifndef LS_BLOCKS_PER_SM
#define LS_BLOCKS_PER_SM 5
#endif

template <bool REPLAY>
__global__ void __launch_bounds__(LS_THREADS, LS_BLOCKS_PER_SM) quoridor_lockstep_kernel(RolloutParams p, unsigned long long *work_counter) {
    extern __shared__ uint32_t smem[];
    lockstep_body<REPLAY>(p, work_counter, smem);
}

static constexpr size_t LS_SMEM = (size_t)LS_SMEM_WORDS * sizeof(uint32_t);

cudaError_t launch_quoridor_lockstep(const RolloutParams &p, bool replay, unsigned long long *work_counter, cudaStream_t st) {
    if (p.n_rollouts == 0) return cudaSuccess;
    static int blocks_per_sm[2] = {0, 0}, sms = 0;
    cudaError_t e;
    if (!sms) {
        int dev = 0;
        if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
        if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(quoridor_lockstep_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LS_SMEM)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(quoridor_lockstep_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LS_SMEM)) != cudaSuccess) return e;
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm[0], quoridor_lockstep_kernel<false>, LS_THREADS, LS_SMEM)) != cudaSuccess) return e;
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm[1], quoridor_lockstep_kernel<true>, LS_THREADS, LS_SMEM)) != cudaSuccess) return e;
    }
    if ((e = cudaMemsetAsync(work_counter, 0, sizeof(unsigned long long), st)) != cudaSuccess) return e;
    int bps = blocks_per_sm[replay ? 1 : 0];
    if (bps < 1) bps = 1;
    unsigned long long want = (p.n_rollouts + LS_THREADS - 1) / LS_THREADS;
    unsigned blocks = (unsigned)(want < (unsigned long long)sms * bps ? want : (unsigned long long)sms * bps);   // persistent: one resident wave
    if (replay) quoridor_lockstep_kernel<true><<<blocks, LS_THREADS, LS_SMEM, st>>>(p, work_counter);
    else        quoridor_lockstep_kernel<false><<<blocks, LS_THREADS, LS_SMEM, st>>>(p, work_counter);
    return cudaGetLastError();
}

}  // namespace mctsb
