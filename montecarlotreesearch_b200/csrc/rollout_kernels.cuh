// rollout_kernels.cuh — sm_100a kernels of the rollout backend (declarations + launch params).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "../../include/mctsb.h"

namespace mctsb {

#define MCTSB_MAX_DEVICES 64           // per-device launch configuration slots

struct RolloutParams {
    const void *states;            // mctsb_qstate[n_leaves] or mctsb_tstate[n_leaves]
    uint32_t n_leaves;
    uint32_t rollouts_per_leaf;    // R; rollout r belongs to leaf r / R
    uint64_t n_rollouts;           // n_leaves * R
    uint64_t seed;
    uint64_t first_rollout_id;
    // replay mode (one rollout per state): pre-drawn u32 streams
    const uint32_t *streams;
    uint32_t stream_len, stream_stride;
    // outputs (any may be null)
    mctsb_leaf_sum *sums;          // [n_leaves], accumulated with integer atomics
    mctsb_outcome *outcomes;       // [n_rollouts]
    void *finals;                  // mctsb_qstate[n_rollouts]
    unsigned long long *counters;  // [4]: rollouts, plies, flood_steps, flood_queries
    uint32_t lanes_per_batch;      // lockstep kernel: rollouts a warp takes per batch (1..32; set by the launch)
    void *walk;                    // UNI lockstep: hand-over records between the two phases, 48 B per rollout (device)
    uint32_t sm_limit;             // REF lockstep: SMs the launch may count on (0 = all of the device): the launching stream belongs to a
                                   // partition of the GPU (device tree with one flush in flight), the persistent grid is sized for it
    uint32_t leave_room;           // REF lockstep: 1 = use the 96-register build even for a small grid, so that a warp of the
                                   // tree's selection kernel (128 registers) fits in every SM sub-partition next to it
};
#define MCTSB_UNI_WALK_BYTES 48

// REF policy, production kernel (quoridor_lockstep.cu): one thread = one rollout, the 32 rollouts of a
// warp advance ply-synchronously; `work_counter` is a device u64 the launch resets
cudaError_t launch_quoridor_lockstep(const RolloutParams &p, bool replay, unsigned long long *work_counter, cudaStream_t st);

// UNI policy, production path (quoridor_uni_lockstep.cu): phase A = the wall phase on the lockstep machinery, phase
// B = one thread per rollout walking the pawns once nobody has a wall left.  p.walk must hold n_rollouts records.
cudaError_t launch_quoridor_uni_lockstep(const RolloutParams &p, bool replay, unsigned long long *work_counter, cudaStream_t st);
cudaError_t launch_quoridor_uni_walk(const RolloutParams &p, bool replay, unsigned long long *work_counter, int sms, cudaStream_t st);

// one thread = one rollout, natural control flow (cross-check of both policies: MCTSB_SIMPLE_KERNEL=1)
cudaError_t launch_quoridor_rollouts(const RolloutParams &p, int policy, bool replay, cudaStream_t st);
cudaError_t launch_ttt_rollouts(const RolloutParams &p, bool replay, cudaStream_t st);

// one warp = one position; the 128 wall candidates are spread over the lanes
cudaError_t launch_quoridor_movegen(const mctsb_qstate *states, int n, uint64_t *legal4, mctsb_qinfo *info,
                                    double *eval, cudaStream_t st);
cudaError_t launch_quoridor_goodmoves(const mctsb_qstate *states, int n, uint8_t *moves, int32_t *counts, cudaStream_t st);
cudaError_t launch_quoridor_play(const mctsb_qstate *states, const int32_t *ids, int n, mctsb_qstate *out,
                                 uint8_t *ok, cudaStream_t st);

// integer-issue roofline denominator: independent LOP3/IADD3 chains on all SMs
cudaError_t launch_int_peak(uint32_t *sink, int blocks, int threads, int iters, cudaStream_t st);
cudaError_t launch_lop3_peak(uint32_t *sink, int blocks, int threads, int iters, cudaStream_t st);   // LOP3 only (ALU pipe)
static const int INT_PEAK_OPS_PER_ITER = 64;   // thread-level integer ops per loop trip

}  // namespace mctsb
