// philox.cuh — counter-based random draws for the rollout kernels (host+device).
//
// The reference draws from a process-wide std::default_random_engine, std::shuffle and rand()
// (reference examples/Quoridor/Quoridor.h:50, Quoridor.cpp:720,796; TicTacToe.cpp:86,93), which
// cannot be reproduced on a GPU and is racy even on the CPU.  Here draw k of rollout r is a pure
// function of (seed, r, k): Philox4x32-10 with counter {k>>2, lo32(r), hi32(r), 0} and key
// {lo32(seed), hi32(seed)}, word k&3 — independent of lane, block or GPU placement.  The draw ->
// value conventions (U() = u32*2^-32, rand() = u32 & 0x7fffffff, the keyed shuffle: head of the pool
// by mulhi, every other element ordered by its own draw) are documented once in include/mctsb.h /
// DESIGN.md.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MC_HD __host__ __device__ __forceinline__
#else
#define MC_HD inline
#endif

namespace mctsb {

MC_HD uint32_t mulhi_u32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

struct Philox4 { uint32_t x, y, z, w; };

MC_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t lo0 = 0xD2511F53u * c0, hi0 = mulhi_u32(0xD2511F53u, c0);
        uint32_t lo1 = 0xCD9E8D57u * c2, hi1 = mulhi_u32(0xCD9E8D57u, c2);
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    Philox4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

// Draw source 1: Philox stream keyed by (seed, rollout id).  `pos` is the index of the next draw;
// skipping draws (pos += n) costs nothing, which the lazy shuffle relies on.
struct PhiloxDraws {
    uint32_t k0, k1, id_lo, id_hi;
    uint32_t pos;
    uint32_t blk;      // index of the cached block, 0xffffffff = none
    Philox4 buf;
    MC_HD void init(uint64_t seed, uint64_t rollout_id) {
        k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32);
        id_lo = (uint32_t)rollout_id; id_hi = (uint32_t)(rollout_id >> 32);
        pos = 0; blk = 0xffffffffu;
    }
    MC_HD uint32_t at(uint32_t k) {
        uint32_t b = k >> 2;
        if (b != blk) { buf = philox4x32_10(b, id_lo, id_hi, 0u, k0, k1); blk = b; }
        uint32_t s = k & 3u;
        return s == 0 ? buf.x : s == 1 ? buf.y : s == 2 ? buf.z : buf.w;
    }
    MC_HD uint32_t next() { return at(pos++); }
    MC_HD void skip(uint32_t n) { pos += n; }
    MC_HD uint32_t consumed() const { return pos; }
};

// Draw source 2: caller-supplied pre-drawn stream, read modulo its length (replay parity mode).
struct StreamDraws {
    const uint32_t *data;
    uint32_t len;
    uint32_t pos;
    MC_HD void init(const uint32_t *d, uint32_t n) { data = d; len = n; pos = 0; }
    MC_HD uint32_t at(uint32_t k) const { return data[k % len]; }
    MC_HD uint32_t next() { return at(pos++); }
    MC_HD void skip(uint32_t n) { pos += n; }
    MC_HD uint32_t consumed() const { return pos; }
};

}  // namespace mctsb
