// simt.h — TEST-ONLY 32-lane warp emulation for the host build of the lockstep kernel body.
//
// Every GPU thread of an emulated block is a fiber (own stack, cooperative switch).  A warp primitive
// (vote, ballot, shuffle, reduce, __syncwarp) deposits the lane's operand in a per-warp exchange buffer and
// yields to the scheduler; the scheduler resumes every live fiber once per pass, so when a lane comes back
// all 32 operands of its warp are there.  Two buffers alternate, so a fast lane's next deposit never
// overwrites values a slower lane still has to read.  The call site (__LINE__) travels with the operand:
// lanes of a warp that meet at different primitives — divergence around a full-mask collective, undefined
// behaviour on the GPU — abort the run.  Lanes are resumed in a different order every pass, so shared-memory
// hand-offs that lack a __syncwarp show up as wrong results.
//
// Never linked into the product; tests/hostemu/build.py compiles it into _hostemu32.so.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace simt {

static constexpr int WARP = 32;
static constexpr size_t STACK_BYTES = 256 * 1024;

extern "C" void simt_switch(void **from_sp, void *to_sp);
#if defined(__x86_64__)
asm(R"(
.text
.globl simt_switch
.type simt_switch,@function
simt_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size simt_switch,.-simt_switch
)");
#else
#error "simt.h: the fiber switch is written for x86-64"
#endif

struct WarpX { uint64_t val[2][WARP]; int site[2][WARP]; };

struct Grid {
    int n_threads = 0, threads_per_block = 0;
    std::vector<void *> sp; std::vector<char *> stack; std::vector<char> done; std::vector<uint32_t> gen;
    std::vector<WarpX> warps;
    void *sched_sp = nullptr;
    int cur = -1;
    void (*body)(void *) = nullptr; void *arg = nullptr;
    uint64_t rng = 0x9E3779B97F4A7C15ull;
};
static Grid *g_grid = nullptr;

inline int tid() { return g_grid->cur % g_grid->threads_per_block; }
inline int block() { return g_grid->cur / g_grid->threads_per_block; }

static void fiber_entry() {
    Grid *g = g_grid;
    g->body(g->arg);
    g->done[g->cur] = 1;
    simt_switch(&g->sp[g->cur], g->sched_sp);
    abort();                                   // a finished fiber is never resumed
}

// operand exchange: returns the 32 operands of the caller's warp
inline const uint64_t *exchange(uint64_t v, int site) {
    Grid *g = g_grid;
    const int me = g->cur, lane = me % WARP;
    WarpX &w = g->warps[me / WARP];
    const int b = (int)(g->gen[me]++ & 1u);
    w.val[b][lane] = v; w.site[b][lane] = site;
    simt_switch(&g->sp[me], g->sched_sp);
    for (int l = 0; l < WARP; l++)
        if (w.site[b][l] != site) { fprintf(stderr, "simt: warp %d diverged around a collective (line %d vs %d)\n", me / WARP, site, w.site[b][l]); abort(); }
    return w.val[b];
}

template <class T> inline uint64_t to_bits(T v) { uint64_t b = 0; memcpy(&b, &v, sizeof(T) < 8 ? sizeof(T) : 8); return b; }
template <class T> inline T from_bits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T) < 8 ? sizeof(T) : 8); return v; }

inline uint32_t ballot(bool p, int site) {
    const uint64_t *v = exchange(p ? 1u : 0u, site);
    uint32_t m = 0; for (int l = 0; l < WARP; l++) if (v[l]) m |= 1u << l;
    return m;
}
inline bool any(bool p, int site) { return ballot(p, site) != 0u; }
template <class T> inline T shfl(T x, int src, int site) { const uint64_t *v = exchange(to_bits(x), site); return from_bits<T>(v[src & 31]); }
template <class T> inline T shfl_up(T x, int delta, int site) {
    const int lane = g_grid->cur % WARP;
    const uint64_t *v = exchange(to_bits(x), site);
    return lane >= delta ? from_bits<T>(v[lane - delta]) : x;
}
inline uint32_t reduce_add(uint32_t x, int site) { const uint64_t *v = exchange(x, site); uint32_t s = 0; for (int l = 0; l < WARP; l++) s += (uint32_t)v[l]; return s; }
inline void syncwarp(int site) { (void)exchange(0, site); }
inline int reduce_max(int x, int site) { const uint64_t *v = exchange((uint64_t)(uint32_t)x, site); int m = (int)(uint32_t)v[0]; for (int l = 1; l < WARP; l++) { int y = (int)(uint32_t)v[l]; if (y > m) m = y; } return m; }
inline uint32_t reduce_maxu(uint32_t x, int site) { const uint64_t *v = exchange(x, site); uint32_t m = 0; for (int l = 0; l < WARP; l++) if ((uint32_t)v[l] > m) m = (uint32_t)v[l]; return m; }
inline uint32_t match_any(uint32_t x, int site) { const uint64_t *v = exchange(x, site); uint32_t m = 0; for (int l = 0; l < WARP; l++) if ((uint32_t)v[l] == x) m |= 1u << l; return m; }

// run body(arg) on n_blocks x threads_per_block fibers; returns when every fiber has returned
inline void launch(int n_blocks, int threads_per_block, void (*body)(void *), void *arg) {
    Grid g; g.threads_per_block = threads_per_block; g.n_threads = n_blocks * threads_per_block;
    g.sp.assign(g.n_threads, nullptr); g.stack.assign(g.n_threads, nullptr); g.done.assign(g.n_threads, 0); g.gen.assign(g.n_threads, 0);
    g.warps.resize(g.n_threads / WARP); g.body = body; g.arg = arg;
    for (int t = 0; t < g.n_threads; t++) {
        g.stack[t] = (char *)malloc(STACK_BYTES);
        uintptr_t top = ((uintptr_t)(g.stack[t] + STACK_BYTES)) & ~(uintptr_t)15;
        void **s = (void **)top;
        s[-1] = nullptr;                       // return address of fiber_entry (never used)
        s[-2] = (void *)&fiber_entry;          // popped by simt_switch's ret
        for (int i = 3; i <= 8; i++) s[-i] = nullptr;   // rbp rbx r12 r13 r14 r15
        g.sp[t] = (void *)(s - 8);
    }
    Grid *prev = g_grid; g_grid = &g;
    std::vector<int> order(g.n_threads);
    for (int t = 0; t < g.n_threads; t++) order[t] = t;
    for (;;) {
        int live = 0;
        // a new resume order every pass (xorshift), warp by warp
        for (int w0 = 0; w0 < g.n_threads; w0 += WARP)
            for (int i = WARP - 1; i > 0; i--) {
                g.rng ^= g.rng << 13; g.rng ^= g.rng >> 7; g.rng ^= g.rng << 17;
                int j = (int)(g.rng % (uint64_t)(i + 1));
                int t = order[w0 + i]; order[w0 + i] = order[w0 + j]; order[w0 + j] = t;
            }
        for (int i = 0; i < g.n_threads; i++) {
            int t = order[i];
            if (g.done[t]) continue;
            live++; g.cur = t;
            simt_switch(&g.sched_sp, g.sp[t]);
        }
        if (!live) break;
        for (int w0 = 0; w0 < g.n_threads; w0 += WARP) {          // lanes of a warp leave together
            int d = 0; for (int l = 0; l < WARP; l++) d += g.done[w0 + l];
            if (d != 0 && d != WARP) { fprintf(stderr, "simt: warp %d: %d lanes returned, the others wait at a collective\n", w0 / WARP, d); abort(); }
        }
    }
    g_grid = prev;
    for (int t = 0; t < g.n_threads; t++) free(g.stack[t]);
}

}  // namespace simt
