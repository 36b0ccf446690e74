/*
 * philox_ref.h — TEST INFRASTRUCTURE (oracle side).  Plain-C Philox4x32-10 and the draw
 * conventions shared by the shimmed reference (oracle/shim.h), the C restatement
 * (oracle/quoridor_oracle.c) and — re-implemented independently in CUDA — the product kernels
 * (montecarlotreesearch_b200/csrc/philox.cuh).  The reference itself has no counter-based RNG
 * (it uses std::default_random_engine / rand(), Quoridor.h:50, Quoridor.cpp:796,
 * TicTacToe.cpp:86,93); these conventions are ours and are defined exactly once, here:
 *
 *   draw k of rollout r under seed s  = philox4x32_10(ctr = {k>>2, lo32(r), hi32(r), 0},
 *                                                     key = {lo32(s), hi32(s)})[k & 3]
 *   U()      = u32 * 2^-32                       (replaces dist(gen), Quoridor.cpp:696,721,748,792)
 *   rand()   = u32 & 0x7fffffff                  (replaces rand(),    Quoridor.cpp:796, TicTacToe.cpp:93)
 *   shuffle  (replaces std::shuffle, Quoridor.cpp:720), for n >= 2 elements, n draws, none if n <= 1:
 *              j0 = mulhi32(u32, n)                 -> input element j0 goes first;
 *              every other input element i (ascending i) then gets key_i = next u32 and the rest
 *              of the output is those elements ordered by (key_i, i) ascending.
 *              A uniform shuffle whose first element costs one draw and in which the relative order
 *              of any subset depends only on that subset's own draws, so a GPU can rank a handful
 *              of candidates without materialising the permutation.
 */
#ifndef ORACLE_PHILOX_REF_H
#define ORACLE_PHILOX_REF_H
#include <stdint.h>

static inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline uint32_t philox_draw(uint64_t seed, uint64_t rollout_id, uint32_t k) {
    uint32_t ctr[4] = { k >> 2, (uint32_t)rollout_id, (uint32_t)(rollout_id >> 32), 0u };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t out[4];
    philox4x32_10(ctr, key, out);
    return out[k & 3];
}

static inline uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }

#endif
