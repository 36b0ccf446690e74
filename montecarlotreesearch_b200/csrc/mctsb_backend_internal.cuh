// mctsb_backend_internal.cuh — the opaque handle of include/mctsb.h, shared by the translation units that
// implement the C ABI (mctsb_api.cu: rollouts, move generation; mctsb_comm.cu: the multi-GPU exchange).
#pragma once
#include "../../include/mctsb.h"
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

struct mctsb_backend {
    int device;
    uint64_t seed;
    cudaStream_t stream;
    cudaEvent_t ev0, ev1;
    // device buffers (grown on demand, never shrunk)
    void *d_states; size_t cap_states;               // bytes
    mctsb_leaf_sum *d_sums; size_t cap_sums;         // elements
    mctsb_outcome *d_outcomes; size_t cap_outcomes;  // elements
    void *d_finals; size_t cap_finals;               // bytes
    uint32_t *d_streams; size_t cap_streams;         // elements
    void *d_aux; size_t cap_aux;                     // bytes (movegen / play outputs)
    void *d_walk; size_t cap_walk;                   // bytes (UNI policy: hand-over records between the two kernels)
    unsigned long long *d_counters;                  // [8]: 0-3 rollouts, plies, flood steps, flood queries; 7 = work counter of the lockstep kernel
    uint32_t *d_sink;                                // int-peak microbenchmark
    void *d_flush; size_t flush_bytes;                // L2 flush scratch (bench hygiene)
    // pinned host staging
    void *h_in; size_t cap_h_in;
    void *h_out; size_t cap_h_out;
    // pending submit
    bool pending; int pend_game; uint32_t pend_leaves; bool pend_legal;
    void *h_legal; size_t cap_h_legal;               // pinned: legal-move masks of the pending submit
    // resident state
    int res_game; uint32_t res_leaves;
    uint64_t launches;
    bool use_simple_kernel;                          // MCTSB_SIMPLE_KERNEL=1: REF policy on the v1 kernel (A/B cross-check)
    // multi-GPU exchange (mctsb_comm.cu): NCCL communicator of this handle and its staging buffer
    void *comm; int comm_world, comm_rank;
    void *d_comm; size_t cap_comm;
    char cuda_err[256];
};


namespace mctsb { struct RolloutParams; }
// mctsb_api.cu: picks the kernel(s) of (game, policy) and launches them on the handle's stream
int mctsb_internal_launch_rollouts(mctsb_backend *b, int game, int policy, bool replay, const mctsb::RolloutParams &p);
