// mctsb_tree.cu — the search tree resident in HBM behind the C ABI (include/mctsb.h, mctsb_tree_*): kernels around
// the warp-level code of tree_body.cuh and the host side that queues them.  Replaces the loop of
// MCTS_tree::grow_tree (reference mcts/src/mcts.cpp:182-210) for Quoridor trees with batched Simulation.
//
// One flush = four launches on the backend's stream, no host synchronisation in between:
//   tree_select_kernel    cooperative, one CTA per SM: the flush's descents level by level (grid barrier between
//                         levels), then the leaf positions gathered into the rollout kernel's input
//   rollout kernel(s)     the production REF / UNI path (mctsb_internal_launch_rollouts)
//   tree_results_kernel   exact leaf sums -> doubles; leaves created by the flush take their first result
//   tree_backprop_kernel  one warp per item: results added along the paths in selection order
#include "tree_body.cuh"
#include "mctsb_backend_internal.cuh"
#include <cooperative_groups.h>
#include <cuda.h>                      // types of the driver's green contexts; the entry points are fetched at run time
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

namespace cg = cooperative_groups;
using namespace mctsb;

namespace {

constexpr int TREE_THREADS = 128;                  // results / backprop kernels: 4 warps per CTA
constexpr int TREE_WARPS = TREE_THREADS / 32;
// The selection kernel runs one warp per CTA: 32 threads x <= 128 registers and 7 KB of shared memory fit on an SM
// next to four resident blocks of the rollout kernel's 96-register build, so the next flush can be selected while the
// playouts of the current one are running (mctsb_tree_set_overlap), and all its CTAs stay resident for the grid barrier.
constexpr int SELECT_THREADS = 32;

__global__ void __launch_bounds__(SELECT_THREADS, 16) tree_select_kernel(TreeCtx t, uint32_t n, mctsb_qstate *leaf_states) {
    cg::grid_group grid = cg::this_grid();
    __shared__ TreeWarpSmem wsm[1];
    const int lane = threadIdx.x & 31, wib = 0;
    const uint32_t gtid = blockIdx.x * SELECT_THREADS + threadIdx.x, gthreads = gridDim.x * SELECT_THREADS;
    const uint32_t warp = blockIdx.x, nwarps = gridDim.x;
    // The flush in chunks of t.chunk descents, chunk c one level behind chunk c - 1: wave w = level w - c of every chunk c
    // that has started and still has items.  All CTAs keep the same table of item ranges (read after the grid barrier).
    __shared__ uint32_t cbegin[TR_MAX_CHUNKS], cend[TR_MAX_CHUNKS];
    const uint32_t nch = (n + t.chunk - 1) / t.chunk;
    for (uint32_t d = gtid; d < n; d += gthreads) t.idx[d] = d;               // level 0: every descent starts at the root
    if (gtid < nch) {
        const uint32_t c = gtid;
        TreeItem r; r.node = t.ctl->root; r.start = c * t.chunk; r.count = n - r.start < t.chunk ? n - r.start : t.chunk; r.level = 0; r.n0 = 0; r.pad = 0;
        r.has_prev = c > 0 ? 1u : 0u; r.next = c + 1 < nch ? (c + 1) * t.items_per_chunk + 1u : 0u;          // the root's items, chunk after chunk
        t.items[(size_t)c * t.items_per_chunk] = r; t.n_items[c] = 1;
    }
    for (uint32_t k = gtid; k < nch * (uint32_t)TR_MAX_LEVELS; k += gthreads) t.level_items[k] = 0u;
    if (lane < (int)TR_MAX_CHUNKS / 32 * 32) for (int c = lane; c < TR_MAX_CHUNKS; c += 32) { cbegin[c] = 0; cend[c] = c < (int)nch ? 1u : 0u; }
    grid.sync();
    unsigned long long t0 = 0;
    if (gtid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (uint32_t wave = 0;; wave++) {
        // this wave's work: items [cbegin[c], cend[c]) of every chunk c <= wave
        const uint32_t started = wave + 1 < nch ? wave + 1 : nch;
        uint32_t total = 0;
        for (uint32_t c = 0; c < started; c++) total += cend[c] - cbegin[c];
        if (total == 0 && wave >= nch) break;
        for (uint32_t g = warp; g < total; g += nwarps) {
            uint32_t c = 0, r = g;
            while (r >= cend[c] - cbegin[c]) { r -= cend[c] - cbegin[c]; c++; }
            tree_select_item(t, t.items[(size_t)c * t.items_per_chunk + cbegin[c] + r], lane, wsm[wib], cend[c]);
        }
        grid.sync();
        if (gtid == 0) {
            unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            t.ctl->level_ns[wave < 7 ? wave : 7] += t1 - t0; t0 = t1;
        }
        // chunk c closed level wave - c; its items of level wave - c + 1 were counted in a word nobody touches any more
        for (int c = lane; c < (int)started; c += 32) { const uint32_t lvl = wave - (uint32_t)c + 1u; cbegin[c] = cend[c]; cend[c] += lvl < (uint32_t)TR_MAX_LEVELS ? *(volatile uint32_t *)(t.level_items + c * TR_MAX_LEVELS + lvl) : 0u; }
        __syncwarp();
    }
    if (blockIdx.x == 0) for (int c = lane; c < (int)nch; c += 32) t.n_items[c] = cend[c];      // what Backpropagation walks
    for (uint32_t d = gtid; d < n; d += gthreads)                              // the rollout kernel's input, in selection order
        ls_store_qstate(leaf_states, d, tr_load_qstate(t.pool.state, (uint32_t)t.leaf_of[d]));
}

__global__ void __launch_bounds__(TREE_THREADS) tree_results_kernel(TreeCtx t, uint32_t n, const mctsb_leaf_sum *sums, double *w) {
    const uint32_t d = blockIdx.x * TREE_THREADS + threadIdx.x;
    if (d >= n) return;
    w[d] = tr_leaf_sum_to_double(sums[d].lo, sums[d].hi);
    tree_backprop_created(t, d, w);
}

// a node's results are added in selection order and a node may be in several chunks: its items are chained, the warp
// that finds the first one walks the chain
__global__ void __launch_bounds__(TREE_THREADS) tree_backprop_kernel(TreeCtx t, const double *w, uint32_t nch) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * TREE_THREADS + threadIdx.x) >> 5, nwarps = gridDim.x * TREE_WARPS;
    for (uint32_t c = 0; c < nch; c++) {
        const uint32_t n_items = t.n_items[c], base = c * t.items_per_chunk;
        for (uint32_t i = warp; i < n_items; i += nwarps)
            if (!t.items[base + i].has_prev) tree_backprop_chain(t, base + i, lane, w);
    }
}

// advance_tree drops every subtree but one (mcts.cpp:132-157: `delete child`).  Here the pool is a bump allocator, so
// the kept subtree is COPIED, level by level, into the second pool and the two pools change places: one thread per
// live node copies its fields, reserves the ids of its children's block in the new pool (a block stays n_moves
// consecutive ids) and queues its expanded children for the next level.
struct TreeGcEntry { uint32_t old_id, new_id; int32_t new_parent; };
__global__ void __launch_bounds__(TREE_THREADS) tree_gc_level_kernel(TreePool from, TreePool to, const TreeGcEntry *cur, uint32_t n_cur,
                                                                    TreeGcEntry *next, uint32_t next_cap, uint32_t *counters /* [0] to_top, [1] n_next, [2] overflow */) {
    const uint32_t i = blockIdx.x * TREE_THREADS + threadIdx.x;
    if (i >= n_cur) return;
    const TreeGcEntry e = cur[i];
    const uint32_t o = e.old_id, n = e.new_id;
    ls_store_qstate(to.state, n, tr_load_qstate(from.state, o));
    to.lh[n] = from.lh[o]; to.lv[n] = from.lv[o]; to.score[n] = from.score[o]; to.sims[n] = from.sims[o]; to.virt[n] = from.virt[o];
    to.size[n] = from.size[o]; to.parent[n] = e.new_parent; to.move[n] = from.move[o]; to.flags[n] = from.flags[o]; to.last_stamp[n] = 0u;
    const uint16_t nc = from.n_children[o], nm = from.n_moves[o];
    to.n_children[n] = nc; to.n_moves[n] = nm;
    const int32_t fc = from.first_child[o];
    int32_t nf = -1;
    if (fc >= 0 && nm != TR_UNLISTED && nm > 0) {
        nf = (int32_t)atomicAdd(&counters[0], (uint32_t)nm);
        if (nc > 0) {
            const uint32_t pos = atomicAdd(&counters[1], (uint32_t)nc);
            if (pos + nc <= next_cap) {
                for (uint32_t k = 0; k < nc; k++) { TreeGcEntry c; c.old_id = (uint32_t)fc + k; c.new_id = (uint32_t)nf + k; c.new_parent = (int32_t)n; next[pos + k] = c; }
            } else atomicOr(&counters[2], 1u);
        }
    }
    to.first_child[n] = nf;
}

int fail(mctsb_backend *b, cudaError_t e, const char *what) {
    if (b) snprintf(b->cuda_err, sizeof b->cuda_err, "%s: %s", what, cudaGetErrorString(e));
    return e == cudaErrorMemoryAllocation ? MCTSB_ERR_OOM : MCTSB_ERR_CUDA;
}
#define TCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(b, e_, #call); } while (0)

}  // namespace

struct mctsb_tree {
    mctsb_backend *b;
    int policy;
    uint32_t R, K; uint64_t pool_cap; double c;
    TreeCtx ctx[2];                     // device pointers; one set of per-flush arrays per flush slot
    TreePool spare;                     // the second pool: advance_tree copies the kept subtree into it and swaps
    TreeGcEntry *d_gc[2]; uint32_t gc_cap; uint32_t *d_gc_counters;
    uint64_t collections;
    mctsb_qstate *d_leaf_states[2]; mctsb_leaf_sum *d_sums[2]; double *d_w[2];
    double *d_logtab;
    void *allocs[96]; int n_allocs;
    int select_blocks;
    bool overlap;                       // select flush f+1 while the playouts of flush f run (host tree: MCTS_batching::overlap)
    cudaStream_t tree_stream;           // Selection / Backpropagation kernels in overlap mode (the playouts stay on the backend's stream)
    // overlap mode, when the driver can partition the GPU (green contexts): the tree's kernels own a few SMs and the
    // playouts the rest, so that the root's chain of picks — one warp's issue latency — does not share its SM with
    // sixteen playout warps.  Falls back to the two plain streams above where that is not available.
    bool part_tried, part_ok;
    CUgreenCtx part_tree_ctx, part_play_ctx;
    cudaStream_t part_tree_stream, part_play_stream;
    int part_tree_sms, part_play_sms, part_select_blocks;
    cudaEvent_t ev_selected[2], ev_played[2];
    uint64_t flushes, rollouts;
    float ms[3], ms_total;
    cudaEvent_t ev_begin, ev_end;       // GPU span of the last grow
};

namespace {

template <class T> int dev_alloc(mctsb_tree *t, T **p, size_t n) {
    mctsb_backend *b = t->b;
    TCK(cudaMalloc((void **)p, n * sizeof(T)));
    t->allocs[t->n_allocs++] = *p;
    return MCTSB_OK;
}
#define TTRY(call) do { int s_ = (call); if (s_ != MCTSB_OK) { mctsb_tree_destroy(t); return s_; } } while (0)

int read_ctl(mctsb_tree *t, TreeCtl *out) {
    mctsb_backend *b = t->b;
    TCK(cudaStreamSynchronize(b->stream));
    if (t->tree_stream) TCK(cudaStreamSynchronize(t->tree_stream));
    if (t->part_tree_stream) TCK(cudaStreamSynchronize(t->part_tree_stream));
    if (t->part_play_stream) TCK(cudaStreamSynchronize(t->part_play_stream));
    TCK(cudaMemcpy(out, t->ctx[0].ctl, sizeof *out, cudaMemcpyDeviceToHost));
    return MCTSB_OK;
}

// a fresh node written from the host (the first root, or a root that advance_tree has to start over)
int write_node(mctsb_tree *t, uint32_t id, const mctsb_qstate &st) {
    mctsb_backend *b = t->b;
    QState s;
    q_unpack(s, st);
    const TreePool &P = t->ctx[0].pool;
    const double zero = 0.0; const unsigned long long zero64 = 0; const uint32_t zero32 = 0; const int32_t minus1 = -1;
    const uint16_t zero16 = 0, unlisted = TR_UNLISTED; const uint8_t nomove = 0xff, flags = q_winner(s) ? TR_FLAG_TERMINAL : 0;
    TCK(cudaMemcpy(P.state + id, &st, sizeof st, cudaMemcpyHostToDevice));
    TCK(cudaMemcpy(P.score + id, &zero, 8, cudaMemcpyHostToDevice));
    TCK(cudaMemcpy(P.sims + id, &zero64, 8, cudaMemcpyHostToDevice));
    TCK(cudaMemcpy(P.virt + id, &zero32, 4, cudaMemcpyHostToDevice));
    TCK(cudaMemcpy(P.size + id, &zero32, 4, cudaMemcpyHostToDevice));
    TCK(cudaMemcpy(P.parent + id, &minus1, 4, cudaMemcpyHostToDevice));
    TCK(cudaMemcpy(P.first_child + id, &minus1, 4, cudaMemcpyHostToDevice));
    TCK(cudaMemcpy(P.n_children + id, &zero16, 2, cudaMemcpyHostToDevice));
    TCK(cudaMemcpy(P.n_moves + id, &unlisted, 2, cudaMemcpyHostToDevice));
    TCK(cudaMemcpy(P.move + id, &nomove, 1, cudaMemcpyHostToDevice));
    TCK(cudaMemcpy(P.flags + id, &flags, 1, cudaMemcpyHostToDevice));
    TCK(cudaMemcpy(P.last_stamp + id, &zero32, 4, cudaMemcpyHostToDevice));
    return MCTSB_OK;
}

bool valid_root(const mctsb_qstate &p) {
    if (p.wx > 8 || p.wy > 8 || p.bx > 8 || p.by > 8 || p.wwalls > 10 || p.bwalls > 10 || p.turn > 1 || p.flags != 0) return false;
    if (p.wx == p.bx && p.wy == p.by) return false;
    if (p.h_anchor & p.v_anchor) return false;
    if (p.h_anchor & ((p.h_anchor << 1) & ~0x0101010101010101ull)) return false;
    if (p.v_anchor & (p.v_anchor << 8)) return false;
    return true;
}

struct RootView { TreeCtl ctl; int first, nc; std::vector<uint8_t> move; std::vector<unsigned long long> sims; std::vector<double> score; mctsb_qstate state; };

int read_root(mctsb_tree *t, RootView &v) {
    mctsb_backend *b = t->b;
    int s = read_ctl(t, &v.ctl);
    if (s != MCTSB_OK) return s;
    const TreePool &P = t->ctx[0].pool;
    const int r = v.ctl.root;
    uint16_t nc = 0;
    TCK(cudaMemcpy(&v.first, P.first_child + r, 4, cudaMemcpyDeviceToHost));
    TCK(cudaMemcpy(&nc, P.n_children + r, 2, cudaMemcpyDeviceToHost));
    TCK(cudaMemcpy(&v.state, P.state + r, sizeof v.state, cudaMemcpyDeviceToHost));
    v.nc = nc;
    v.move.resize(nc); v.sims.resize(nc); v.score.resize(nc);
    if (nc) {
        TCK(cudaMemcpy(v.move.data(), P.move + v.first, nc, cudaMemcpyDeviceToHost));
        TCK(cudaMemcpy(v.sims.data(), P.sims + v.first, (size_t)nc * 8, cudaMemcpyDeviceToHost));
        TCK(cudaMemcpy(v.score.data(), P.score + v.first, (size_t)nc * 8, cudaMemcpyDeviceToHost));
    }
    return MCTSB_OK;
}

}  // namespace

namespace {

// The kept subtree (root = `root`) moves into the spare pool; afterwards the root is node 0 and pool_top is what the
// subtree needs.  Called between grows (no simulation in flight).
int collect(mctsb_tree *t, int32_t root, uint32_t *new_top) {
    mctsb_backend *b = t->b;
    cudaStream_t st = b->stream;
    const TreePool from = t->ctx[0].pool, to = t->spare;
    const TreeGcEntry first = {(uint32_t)root, 0u, -1};
    uint32_t counters[4] = {1u, 0u, 0u, 0u};                 // the root takes id 0
    TCK(cudaMemcpyAsync(t->d_gc[0], &first, sizeof first, cudaMemcpyHostToDevice, st));
    TCK(cudaMemcpyAsync(t->d_gc_counters, counters, sizeof counters, cudaMemcpyHostToDevice, st));
    uint32_t n_cur = 1;
    int cur = 0;
    while (n_cur > 0) {
        tree_gc_level_kernel<<<(n_cur + TREE_THREADS - 1) / TREE_THREADS, TREE_THREADS, 0, st>>>(from, to, t->d_gc[cur], n_cur, t->d_gc[cur ^ 1], t->gc_cap, t->d_gc_counters);
        TCK(cudaGetLastError());
        b->launches++;
        TCK(cudaMemcpyAsync(counters, t->d_gc_counters, sizeof counters, cudaMemcpyDeviceToHost, st));
        TCK(cudaStreamSynchronize(st));
        if (counters[2]) { snprintf(b->cuda_err, sizeof b->cuda_err, "device tree: a level of the kept subtree is wider than the collector's queue"); return MCTSB_ERR_TREE_FULL; }
        n_cur = counters[1];
        const uint32_t zero = 0;
        TCK(cudaMemcpyAsync(t->d_gc_counters + 1, &zero, 4, cudaMemcpyHostToDevice, st));
        cur ^= 1;
    }
    *new_top = counters[0];
    for (int k = 0; k < 2; k++) t->ctx[k].pool = to;
    t->spare = from;
    t->collections++;
    return MCTSB_OK;
}

// Split the device's SMs into a small group for the tree and the rest for the playouts (CUDA driver API, green
// contexts; entry points through cudaGetDriverEntryPoint so that the library does not link libcuda).  Any failure
// leaves part_ok false: the caller then runs both kinds of kernels on all SMs, as before.
void make_partition(mctsb_tree *t) {
    t->part_tried = true; t->part_ok = false;
    if (getenv("MCTSB_TREE_NO_PARTITION")) return;
    typedef CUresult (*fn_dev_get)(CUdevice *, int);
    typedef CUresult (*fn_get_res)(CUdevice, CUdevResource *, CUdevResourceType);
    typedef CUresult (*fn_split)(CUdevResource *, unsigned int *, const CUdevResource *, CUdevResource *, unsigned int, unsigned int);
    typedef CUresult (*fn_desc)(CUdevResourceDesc *, CUdevResource *, unsigned int);
    typedef CUresult (*fn_gctx)(CUgreenCtx *, CUdevResourceDesc, CUdevice, unsigned int);
    typedef CUresult (*fn_gstream)(CUstream *, CUgreenCtx, unsigned int, int);
    void *p[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    const char *names[6] = {"cuDeviceGet", "cuDeviceGetDevResource", "cuDevSmResourceSplitByCount", "cuDevResourceGenerateDesc", "cuGreenCtxCreate", "cuGreenCtxStreamCreate"};
    for (int i = 0; i < 6; i++) {
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint(names[i], &p[i], cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p[i]) { cudaGetLastError(); return; }
    }
    CUdevice dev;
    if (((fn_dev_get)p[0])(&dev, t->b->device) != CUDA_SUCCESS) return;
    CUdevResource all, tree_res, play_res;
    if (((fn_get_res)p[1])(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return;
    unsigned int groups = 1;
    const char *ev = getenv("MCTSB_TREE_SMS");
    // groups come in multiples of 8 SMs on sm_90+; measured over whole games at 4 096 leaves per flush: 8 SMs 25.6 M rollouts/s,
    // 16 SMs 26.7 M, 32 SMs 26.3 M (the waves below the root spread over the group's warps)
    const unsigned int want = ev && atoi(ev) > 0 ? (unsigned)atoi(ev) : 16u;
    if (((fn_split)p[2])(&tree_res, &groups, &all, &play_res, 0u, want) != CUDA_SUCCESS || groups != 1) return;
    if (tree_res.sm.smCount == 0 || play_res.sm.smCount < 64) return;
    CUdevResourceDesc dtree, dplay;
    if (((fn_desc)p[3])(&dtree, &tree_res, 1) != CUDA_SUCCESS || ((fn_desc)p[3])(&dplay, &play_res, 1) != CUDA_SUCCESS) return;
    if (((fn_gctx)p[4])(&t->part_tree_ctx, dtree, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return;
    if (((fn_gctx)p[4])(&t->part_play_ctx, dplay, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return;
    CUstream st = nullptr, sp = nullptr;
    if (((fn_gstream)p[5])(&st, t->part_tree_ctx, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS) return;
    if (((fn_gstream)p[5])(&sp, t->part_play_ctx, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS) return;
    t->part_tree_stream = (cudaStream_t)st; t->part_play_stream = (cudaStream_t)sp;
    t->part_tree_sms = (int)tree_res.sm.smCount; t->part_play_sms = (int)play_res.sm.smCount;
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tree_select_kernel, SELECT_THREADS, 0);
    t->part_select_blocks = t->part_tree_sms * (per_sm < 1 ? 1 : per_sm);
    { const char *eb = getenv("MCTSB_TREE_SELECT_BLOCKS"); if (eb && atoi(eb) > 0 && atoi(eb) < t->part_select_blocks) t->part_select_blocks = atoi(eb); }
    t->part_ok = true;
    if (getenv("MCTSB_TREE_DEBUG")) fprintf(stderr, "tree: SM partition %d (tree) + %d (playouts), selection grid %d\n", t->part_tree_sms, t->part_play_sms, t->part_select_blocks);
}

}  // namespace

extern "C" {

int mctsb_tree_create(mctsb_backend *b, mctsb_tree **out, const mctsb_qstate *root, int policy,
                      uint32_t rollouts_per_leaf, uint32_t max_leaves_per_flush, uint64_t max_nodes, double uct_c) {
    if (!b || !out || !root || rollouts_per_leaf == 0 || max_leaves_per_flush == 0) return MCTSB_ERR_BAD_ARG;
    *out = nullptr;
    if (policy != MCTSB_POLICY_REF && policy != MCTSB_POLICY_UNI) return MCTSB_ERR_BAD_ARG;
    if (max_nodes < 512 || max_nodes > 0x7fffff00ull) return MCTSB_ERR_BAD_ARG;
    // virtual losses are counted in 32 bits: two flushes' worth of playouts must fit
    if ((uint64_t)rollouts_per_leaf * max_leaves_per_flush >= (1ull << 30)) return MCTSB_ERR_BAD_ARG;
    if (!valid_root(*root)) return MCTSB_ERR_BAD_STATE;
    if (b->pending) return MCTSB_ERR_BUSY;
    TCK(cudaSetDevice(b->device));
    mctsb_tree *t = new (std::nothrow) mctsb_tree();
    if (!t) return MCTSB_ERR_OOM;
    memset(t, 0, sizeof *t);
    t->b = b; t->policy = policy; t->R = rollouts_per_leaf; t->K = max_leaves_per_flush; t->pool_cap = max_nodes; t->c = uct_c;
    TreeCtx &c = t->ctx[0];
    const size_t N = (size_t)max_nodes, K = t->K;
    TTRY(dev_alloc(t, &c.pool.state, N)); TTRY(dev_alloc(t, &c.pool.lh, N)); TTRY(dev_alloc(t, &c.pool.lv, N));
    TTRY(dev_alloc(t, &c.pool.score, N)); TTRY(dev_alloc(t, &c.pool.sims, N)); TTRY(dev_alloc(t, &c.pool.virt, N));
    TTRY(dev_alloc(t, &c.pool.size, N)); TTRY(dev_alloc(t, &c.pool.parent, N)); TTRY(dev_alloc(t, &c.pool.first_child, N));
    TTRY(dev_alloc(t, &c.pool.n_children, N)); TTRY(dev_alloc(t, &c.pool.n_moves, N)); TTRY(dev_alloc(t, &c.pool.move, N));
    TTRY(dev_alloc(t, &c.pool.flags, N)); TTRY(dev_alloc(t, &c.pool.last_item, N)); TTRY(dev_alloc(t, &c.pool.last_stamp, N));
    TTRY(dev_alloc(t, &c.ctl, 1));
    {
        TreePool &q = t->spare;
        TTRY(dev_alloc(t, &q.state, N)); TTRY(dev_alloc(t, &q.lh, N)); TTRY(dev_alloc(t, &q.lv, N));
        TTRY(dev_alloc(t, &q.score, N)); TTRY(dev_alloc(t, &q.sims, N)); TTRY(dev_alloc(t, &q.virt, N));
        TTRY(dev_alloc(t, &q.size, N)); TTRY(dev_alloc(t, &q.parent, N)); TTRY(dev_alloc(t, &q.first_child, N));
        TTRY(dev_alloc(t, &q.n_children, N)); TTRY(dev_alloc(t, &q.n_moves, N)); TTRY(dev_alloc(t, &q.move, N));
        TTRY(dev_alloc(t, &q.flags, N)); TTRY(dev_alloc(t, &q.last_item, N)); TTRY(dev_alloc(t, &q.last_stamp, N));
        t->gc_cap = (uint32_t)(N < (1u << 24) ? N : (1u << 24));       // widest level of a kept subtree (nodes that exist, not reserved ids)
        TTRY(dev_alloc(t, &t->d_gc[0], (size_t)t->gc_cap)); TTRY(dev_alloc(t, &t->d_gc[1], (size_t)t->gc_cap));
        TTRY(dev_alloc(t, &t->d_gc_counters, 4));
    }
    c.K = t->K; c.pool_cap = (uint32_t)max_nodes; c.R = t->R; c.c = uct_c;
    // chunks of 256 descents (measured best over whole games at 4 096 leaves per flush: 23.7 M rollouts/s against 22.5 M
    // at 512 and 19.0 M at 1 024; more for flushes beyond 64 x 256), a whole number of warps' worth each
    c.chunk = (uint32_t)(K <= 256 * (size_t)TR_MAX_CHUNKS ? 256 : ((K + TR_MAX_CHUNKS - 1) / TR_MAX_CHUNKS + 31) / 32 * 32);
    { const char *ev = getenv("MCTSB_TREE_CHUNK"); if (ev && atoi(ev) >= 32) { c.chunk = (uint32_t)atoi(ev) / 32 * 32; if ((K + c.chunk - 1) / c.chunk > TR_MAX_CHUNKS) c.chunk = (uint32_t)(((K + TR_MAX_CHUNKS - 1) / TR_MAX_CHUNKS + 31) / 32 * 32); } }
    c.items_per_chunk = c.chunk * (uint32_t)TR_MAX_LEVELS;
    t->ctx[1] = c;
    for (int k = 0; k < 2; k++) {                      // per-flush arrays, one set per slot
        TreeCtx &f = t->ctx[k];
        TTRY(dev_alloc(t, &f.n_items, TR_MAX_CHUNKS)); TTRY(dev_alloc(t, &f.level_items, (size_t)TR_MAX_CHUNKS * TR_MAX_LEVELS)); TTRY(dev_alloc(t, &f.items, (size_t)f.items_per_chunk * ((K + f.chunk - 1) / f.chunk)));
        TTRY(dev_alloc(t, &f.idx, K * TR_MAX_LEVELS)); TTRY(dev_alloc(t, &f.choice, K)); TTRY(dev_alloc(t, &f.leaf_of, K)); TTRY(dev_alloc(t, &f.created, K));
        TTRY(dev_alloc(t, &t->d_leaf_states[k], K)); TTRY(dev_alloc(t, &t->d_sums[k], K)); TTRY(dev_alloc(t, &t->d_w[k], K));
    }
    // ln N for every N = m * R a node of this pool can reach, from the host's libm: the exact UCT expression then
    // sees the doubles the host tree sees.  One entry per iteration the pool can hold.
    {
        const size_t n = (size_t)(max_nodes < (1u << 22) ? max_nodes : (1u << 22)) + 2 * K + 16;
        std::vector<double> tab(n);
        for (size_t m = 0; m < n; m++) tab[m] = log((double)((unsigned long long)m * t->R));
        TTRY(dev_alloc(t, &t->d_logtab, n));
        cudaError_t e = cudaMemcpy(t->d_logtab, tab.data(), n * sizeof(double), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { int s = fail(b, e, "log table upload"); mctsb_tree_destroy(t); return s; }
        for (int k = 0; k < 2; k++) { t->ctx[k].logtab = t->d_logtab; t->ctx[k].logtab_n = (uint32_t)n; }
    }
    TreeCtl ctl; memset(&ctl, 0, sizeof ctl);
    ctl.pool_top = 1; ctl.root = 0;
    cudaError_t e = cudaMemcpy(c.ctl, &ctl, sizeof ctl, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { int s = fail(b, e, "tree control block upload"); mctsb_tree_destroy(t); return s; }
    TTRY(write_node(t, 0, *root));
    for (int i = 0; i < 2; i++) {
        e = cudaEventCreateWithFlags(&t->ev_selected[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&t->ev_played[i], cudaEventDisableTiming);
        if (e != cudaSuccess) { int s = fail(b, e, "cudaEventCreate"); mctsb_tree_destroy(t); return s; }
    }
    e = cudaEventCreate(&t->ev_begin);
    if (e == cudaSuccess) e = cudaEventCreate(&t->ev_end);
    if (e != cudaSuccess) { int s = fail(b, e, "cudaEventCreate"); mctsb_tree_destroy(t); return s; }
    {
        int lo = 0, hi = 0;                              // the tree's kernels are short and on the critical path: highest priority
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        e = cudaStreamCreateWithPriority(&t->tree_stream, cudaStreamNonBlocking, hi);
        if (e != cudaSuccess) { int s = fail(b, e, "cudaStreamCreateWithPriority"); mctsb_tree_destroy(t); return s; }
    }
    // the selection kernel synchronises its CTAs between levels: all of them must be resident (cooperative launch)
    int dev_sms = 0, per_sm = 0, coop = 0;
    cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, b->device);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, b->device);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tree_select_kernel, SELECT_THREADS, 0);
    if (!coop || per_sm < 1 || dev_sms < 1) { snprintf(b->cuda_err, sizeof b->cuda_err, "cooperative launch not available"); mctsb_tree_destroy(t); return MCTSB_ERR_CUDA; }
    // one-warp CTAs, four per SM: the waves below the root hold hundreds of items in the middle of a game (measured over
    // whole games at 4 096 leaves per flush: 1 / 2 / 4 / 8 CTAs per SM = 1.84 / 1.73 / 1.68 / 1.69 ms of Selection per flush)
    t->select_blocks = dev_sms * (per_sm < 4 ? per_sm : 4);
    { const char *eb = getenv("MCTSB_TREE_BLOCKS_PER_SM"); if (eb && atoi(eb) > 0 && atoi(eb) <= per_sm) t->select_blocks = dev_sms * atoi(eb); }
    *out = t;
    return MCTSB_OK;
}

int mctsb_tree_destroy(mctsb_tree *t) {
    if (!t) return MCTSB_OK;
    cudaSetDevice(t->b->device);
    cudaStreamSynchronize(t->b->stream);
    for (int i = 0; i < t->n_allocs; i++) cudaFree(t->allocs[i]);
    if (t->tree_stream) { cudaStreamSynchronize(t->tree_stream); cudaStreamDestroy(t->tree_stream); }
    if (t->part_tree_stream) { cudaStreamSynchronize(t->part_tree_stream); cudaStreamDestroy(t->part_tree_stream); }
    if (t->part_play_stream) { cudaStreamSynchronize(t->part_play_stream); cudaStreamDestroy(t->part_play_stream); }
    if (t->part_tree_ctx || t->part_play_ctx) {
        typedef CUresult (*fn_gdestroy)(CUgreenCtx);
        void *p = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuGreenCtxDestroy", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && p) {
            if (t->part_tree_ctx) ((fn_gdestroy)p)(t->part_tree_ctx);
            if (t->part_play_ctx) ((fn_gdestroy)p)(t->part_play_ctx);
        }
    }
    if (t->ev_begin) cudaEventDestroy(t->ev_begin);
    if (t->ev_end) cudaEventDestroy(t->ev_end);
    for (int i = 0; i < 2; i++) { if (t->ev_selected[i]) cudaEventDestroy(t->ev_selected[i]); if (t->ev_played[i]) cudaEventDestroy(t->ev_played[i]); }
    delete t;
    return MCTSB_OK;
}

int mctsb_tree_set_good_moves(mctsb_tree *t, int on) {
    if (!t) return MCTSB_ERR_BAD_ARG;
    if (t->flushes > 0) return MCTSB_ERR_BUSY;            // the move lists of a grown tree are what they are
    for (int k = 0; k < 2; k++) t->ctx[k].good_moves = on ? 1u : 0u;
    return MCTSB_OK;
}

int mctsb_tree_set_overlap(mctsb_tree *t, int on) {
    if (!t) return MCTSB_ERR_BAD_ARG;
    t->overlap = on != 0;
    if (t->overlap && !t->part_tried) { cudaSetDevice(t->b->device); make_partition(t); }
    return MCTSB_OK;
}

namespace {

// the launches of one flush, split so that the two modes can order them differently
int launch_select(mctsb_tree *t, int slot, uint32_t n, cudaStream_t st) {
    mctsb_backend *b = t->b;
    t->ctx[slot].stamp = (uint32_t)(t->flushes + 1);      // flush numbers start at 1 (0 = "no item yet" in the pool)
    void *args[] = {(void *)&t->ctx[slot], (void *)&n, (void *)&t->d_leaf_states[slot]};
    const int blocks = t->overlap && t->part_ok ? t->part_select_blocks : t->select_blocks;
    TCK(cudaLaunchCooperativeKernel((const void *)tree_select_kernel, dim3(blocks), dim3(SELECT_THREADS), args, 0, st));
    b->launches++;
    return MCTSB_OK;
}

int launch_playouts(mctsb_tree *t, int slot, uint32_t n, uint64_t first_id) {      // on the backend's stream
    mctsb_backend *b = t->b;
    TCK(cudaMemsetAsync(t->d_sums[slot], 0, (size_t)n * sizeof(mctsb_leaf_sum), b->stream));
    RolloutParams p; memset(&p, 0, sizeof p);
    p.states = t->d_leaf_states[slot]; p.n_leaves = n; p.rollouts_per_leaf = t->R; p.n_rollouts = (uint64_t)n * t->R;
    p.seed = b->seed; p.first_rollout_id = first_id; p.sums = t->d_sums[slot]; p.counters = b->d_counters;
    if (t->overlap && t->part_ok) p.sm_limit = (uint32_t)t->part_play_sms;        // the playouts have their own SMs: the grid is sized for them
    else p.leave_room = t->overlap ? 1u : 0u;                                      // shared SMs: the build that leaves a selection warp room
    return mctsb_internal_launch_rollouts(b, MCTSB_GAME_QUORIDOR, t->policy, false, p);
}

int launch_backprop(mctsb_tree *t, int slot, uint32_t n, cudaStream_t st) {
    mctsb_backend *b = t->b;
    tree_results_kernel<<<(n + TREE_THREADS - 1) / TREE_THREADS, TREE_THREADS, 0, st>>>(t->ctx[slot], n, t->d_sums[slot], t->d_w[slot]);
    const uint32_t nch = (n + t->ctx[slot].chunk - 1) / t->ctx[slot].chunk;
    tree_backprop_kernel<<<t->select_blocks / 2, TREE_THREADS, 0, st>>>(t->ctx[slot], t->d_w[slot], nch);
    TCK(cudaGetLastError());
    b->launches += 2;
    return MCTSB_OK;
}

}  // namespace

int mctsb_tree_grow(mctsb_tree *t, uint32_t iterations, uint32_t leaves_per_flush, uint64_t first_rollout_id) {
    if (!t || leaves_per_flush == 0 || leaves_per_flush > t->K) return MCTSB_ERR_BAD_ARG;
    mctsb_backend *b = t->b;
    if (b->pending) return MCTSB_ERR_BUSY;
    TCK(cudaSetDevice(b->device));
    b->res_game = -1; b->res_leaves = 0;
    t->ms[0] = t->ms[1] = t->ms[2] = t->ms_total = 0.0f;
    cudaStream_t first_stream = !t->overlap ? b->stream : t->part_ok ? t->part_tree_stream : t->tree_stream;      // both modes start with a Selection and end with a Backpropagation
    TCK(cudaEventRecord(t->ev_begin, first_stream));
    const uint32_t n_flushes = (iterations + leaves_per_flush - 1) / leaves_per_flush;
    int status = MCTSB_OK;
    uint32_t done = 0, f = 0;
    std::vector<cudaEvent_t> evs;           // synchronous mode: per-flush phase marks, read after the last flush
    if (!t->overlap) {
        // Selection(f) -> playouts(f) -> Backpropagation(f), all on the backend's stream
        const bool timed = n_flushes <= 4096;
        if (timed) { evs.resize((size_t)n_flushes * 4); for (auto &e : evs) TCK(cudaEventCreate(&e)); }
        while (done < iterations && status == MCTSB_OK) {
            const uint32_t n = iterations - done < leaves_per_flush ? iterations - done : leaves_per_flush;
            if (timed) cudaEventRecord(evs[4 * f + 0], b->stream);
            status = launch_select(t, 0, n, b->stream);
            if (timed) cudaEventRecord(evs[4 * f + 1], b->stream);
            if (status == MCTSB_OK) status = launch_playouts(t, 0, n, first_rollout_id + (uint64_t)done * t->R);
            if (timed) cudaEventRecord(evs[4 * f + 2], b->stream);
            if (status == MCTSB_OK) status = launch_backprop(t, 0, n, b->stream);
            if (timed) cudaEventRecord(evs[4 * f + 3], b->stream);
            if (status != MCTSB_OK) break;
            done += n; f++;
            t->flushes++; t->rollouts += (uint64_t)n * t->R;
        }
        cudaError_t e = cudaStreamSynchronize(b->stream);
        if (e != cudaSuccess && status == MCTSB_OK) status = fail(b, e, "tree grow");
        if (timed) {
            if (status == MCTSB_OK)
                for (uint32_t i = 0; i < f; i++)
                    for (int k = 0; k < 3; k++) { float ms = 0; cudaEventElapsedTime(&ms, evs[4 * i + k], evs[4 * i + k + 1]); t->ms[k] += ms; }
            for (auto &ev : evs) cudaEventDestroy(ev);
        }
    } else {
        // One flush in flight, the order of the host tree with MCTS_batching::overlap (host/mcts.cpp grow_tree):
        //   S(1)  S(2) B(1)  S(3) B(2) ...  S(n) B(n-1)  B(n)
        // Selection and Backpropagation on the tree's own stream, playouts on the backend's: S(f+1) runs while the
        // playouts of flush f are on the SMs, B(f) as soon as they are done.  Flush f uses slot f & 1 of the per-flush arrays.
        cudaStream_t ts = t->part_ok ? t->part_tree_stream : t->tree_stream;
        cudaStream_t own_stream = b->stream;
        if (t->part_ok) b->stream = t->part_play_stream;   // the rollout launcher works on the handle's stream: for this grow, the partition's
        uint32_t n_prev = 0;
        cudaError_t e = cudaSuccess;
        while (done < iterations && status == MCTSB_OK) {
            const uint32_t n = iterations - done < leaves_per_flush ? iterations - done : leaves_per_flush;
            const int slot = (int)(f & 1u);
            status = launch_select(t, slot, n, ts);
            if (status != MCTSB_OK) break;
            if ((e = cudaEventRecord(t->ev_selected[slot], ts)) != cudaSuccess) { status = fail(b, e, "cudaEventRecord"); break; }
            if (f > 0) {                                   // the flush before: its playouts must be done
                if ((e = cudaStreamWaitEvent(ts, t->ev_played[slot ^ 1], 0)) != cudaSuccess) { status = fail(b, e, "cudaStreamWaitEvent"); break; }
                status = launch_backprop(t, slot ^ 1, n_prev, ts);
                if (status != MCTSB_OK) break;
            }
            if ((e = cudaStreamWaitEvent(b->stream, t->ev_selected[slot], 0)) != cudaSuccess) { status = fail(b, e, "cudaStreamWaitEvent"); break; }
            status = launch_playouts(t, slot, n, first_rollout_id + (uint64_t)done * t->R);
            if (status != MCTSB_OK) break;
            if ((e = cudaEventRecord(t->ev_played[slot], b->stream)) != cudaSuccess) { status = fail(b, e, "cudaEventRecord"); break; }
            n_prev = n; done += n; f++;
            t->flushes++; t->rollouts += (uint64_t)n * t->R;
        }
        if (status == MCTSB_OK && f > 0) {
            const int last = (int)((f - 1) & 1u);
            if ((e = cudaStreamWaitEvent(ts, t->ev_played[last], 0)) != cudaSuccess) status = fail(b, e, "cudaStreamWaitEvent");
            else status = launch_backprop(t, last, n_prev, ts);
        }
        e = cudaStreamSynchronize(b->stream);
        if (e != cudaSuccess && status == MCTSB_OK) status = fail(b, e, "tree grow");
        e = cudaStreamSynchronize(ts);
        if (e != cudaSuccess && status == MCTSB_OK) status = fail(b, e, "tree grow");
        b->stream = own_stream;
    }
    if (status != MCTSB_OK) return status;
    TCK(cudaEventRecord(t->ev_end, first_stream));
    TCK(cudaEventSynchronize(t->ev_end));
    TCK(cudaEventElapsedTime(&t->ms_total, t->ev_begin, t->ev_end));
    TreeCtl ctl;
    int s = read_ctl(t, &ctl);
    if (s != MCTSB_OK) return s;
    if (ctl.error & (TR_ERR_POOL | TR_ERR_ITEMS | TR_ERR_DEPTH | TR_ERR_COUNT)) {
        snprintf(b->cuda_err, sizeof b->cuda_err, "device tree: %s%s%s%s", (ctl.error & TR_ERR_POOL) ? "node pool exhausted " : "",
                 (ctl.error & TR_ERR_DEPTH) ? "selection path deeper than the level buffers " : "", (ctl.error & TR_ERR_ITEMS) ? "item list full " : "",
                 (ctl.error & TR_ERR_COUNT) ? "a node has more than 2^31 simulations" : "");
        return MCTSB_ERR_TREE_FULL;
    }
    return MCTSB_OK;
}

int mctsb_tree_get_info(mctsb_tree *t, mctsb_tree_info *info) {
    if (!t || !info) return MCTSB_ERR_BAD_ARG;
    mctsb_backend *b = t->b;
    TCK(cudaSetDevice(b->device));
    TreeCtl ctl;
    int s = read_ctl(t, &ctl);
    if (s != MCTSB_OK) return s;
    memset(info, 0, sizeof *info);
    const TreePool &P = t->ctx[0].pool;
    unsigned long long sims = 0; uint16_t nc = 0;
    TCK(cudaMemcpy(&sims, P.sims + ctl.root, 8, cudaMemcpyDeviceToHost));
    TCK(cudaMemcpy(&info->root_score, P.score + ctl.root, 8, cudaMemcpyDeviceToHost));
    TCK(cudaMemcpy(&info->root_size, P.size + ctl.root, 4, cudaMemcpyDeviceToHost));
    TCK(cudaMemcpy(&nc, P.n_children + ctl.root, 2, cudaMemcpyDeviceToHost));
    info->root_sims = sims; info->root_children = nc; info->nodes_used = ctl.pool_top; info->deepest_level = ctl.depth_max;
    info->exact_evals = ctl.exact_evals; info->flags = ctl.error; info->flushes = t->flushes; info->rollouts = t->rollouts;
    if (getenv("MCTSB_TREE_DEBUG"))
        fprintf(stderr, "tree: steps %u (unique or alike %u, same-statistics ties %u, settled by the exact expression %u) | level ms: %.3f %.3f %.3f %.3f %.3f %.3f %.3f %.3f\n", ctl.steps,
                ctl.steps - ctl.ties_same - ctl.exact_evals, ctl.ties_same, ctl.exact_evals, ctl.level_ns[0] * 1e-6, ctl.level_ns[1] * 1e-6, ctl.level_ns[2] * 1e-6,
                ctl.level_ns[3] * 1e-6, ctl.level_ns[4] * 1e-6, ctl.level_ns[5] * 1e-6, ctl.level_ns[6] * 1e-6, ctl.level_ns[7] * 1e-6);
    info->ms_select = t->ms[0]; info->ms_rollout = t->ms[1]; info->ms_backprop = t->ms[2]; info->ms_total = t->ms_total;
    return MCTSB_OK;
}

int mctsb_tree_root_children(mctsb_tree *t, int32_t *move_ids, uint64_t *sims, double *scores, int max_children, int *n) {
    if (!t || !n || max_children < 0) return MCTSB_ERR_BAD_ARG;
    mctsb_backend *b = t->b;
    TCK(cudaSetDevice(b->device));
    RootView v;
    int s = read_root(t, v);
    if (s != MCTSB_OK) return s;
    *n = v.nc;
    for (int i = 0; i < v.nc && i < max_children; i++) {
        if (move_ids) move_ids[i] = v.move[i];
        if (sims) sims[i] = v.sims[i];
        if (scores) scores[i] = v.score[i];
    }
    return MCTSB_OK;
}

int mctsb_tree_best_move(mctsb_tree *t, int32_t *move_id) {
    if (!t || !move_id) return MCTSB_ERR_BAD_ARG;
    mctsb_backend *b = t->b;
    TCK(cudaSetDevice(b->device));
    RootView v;
    int s = read_root(t, v);
    if (s != MCTSB_OK) return s;
    *move_id = -1;
    if (v.nc == 0) return MCTSB_OK;
    if (v.nc == 1) { *move_id = v.move[0]; return MCTSB_OK; }
    // select_best_child(0.0): win rate of the side to move at the root, first strict maximum (mcts.cpp:104-129)
    const bool p1 = v.state.turn == 0;
    double best = -1; int arg = -1;
    for (int i = 0; i < v.nc; i++) {
        const double exact = d_add(p1 ? 0.0 : 1.0, d_mul(p1 ? 1.0 : -1.0, d_div(v.score[i], (double)v.sims[i])));
        if (exact > best) { best = exact; arg = i; }
    }
    if (arg >= 0) *move_id = v.move[arg];
    return MCTSB_OK;
}

int mctsb_tree_advance(mctsb_tree *t, int32_t move_id) {
    if (!t || move_id < 0 || move_id >= MCTSB_Q_NUM_MOVES) return MCTSB_ERR_BAD_ARG;
    mctsb_backend *b = t->b;
    TCK(cudaSetDevice(b->device));
    RootView v;
    int s = read_root(t, v);
    if (s != MCTSB_OK) return s;
    int32_t next = -1;
    for (int i = 0; i < v.nc; i++) if (v.move[i] == move_id) next = v.first + i;
    if (next >= 0) {
        const int32_t minus1 = -1;
        TCK(cudaMemcpy(t->ctx[0].pool.parent + next, &minus1, 4, cudaMemcpyHostToDevice));
    } else {                                              // "Didn't find child node. Had to start over." (mcts.cpp:149-152)
        QState q;
        q_unpack(q, v.state);
        if (!q_legal_move(q, move_id)) return MCTSB_ERR_BAD_ARG;
        q_play_id(q, move_id);
        mctsb_qstate st; q_pack(q, st);
        if ((uint64_t)v.ctl.pool_top + 1 > t->pool_cap) return MCTSB_ERR_TREE_FULL;
        next = (int32_t)v.ctl.pool_top;
        s = write_node(t, (uint32_t)next, st);
        if (s != MCTSB_OK) return s;
        const uint32_t top = v.ctl.pool_top + 1;
        TCK(cudaMemcpy(&t->ctx[0].ctl->pool_top, &top, 4, cudaMemcpyHostToDevice));
    }
    uint32_t top = next >= (int32_t)v.ctl.pool_top ? v.ctl.pool_top + 1 : v.ctl.pool_top;
    // the dropped subtrees are reclaimed by copying the kept one into the spare pool — once the pool is a quarter full
    // (MCTSB_TREE_GC=always / never for tests and measurements)
    const char *gc = getenv("MCTSB_TREE_GC");
    const bool want = gc && gc[0] == 'a' ? true : gc && gc[0] == 'n' ? false : (uint64_t)top * 4 > t->pool_cap;
    if (want) {
        uint32_t new_top = 0;
        s = collect(t, next, &new_top);
        if (s != MCTSB_OK) return s;
        next = 0; top = new_top;
        TCK(cudaMemcpy(&t->ctx[0].ctl->pool_top, &top, 4, cudaMemcpyHostToDevice));
    }
    TCK(cudaMemcpy(&t->ctx[0].ctl->root, &next, 4, cudaMemcpyHostToDevice));
    return MCTSB_OK;
}

// Root-parallel search: trees of the same root grown on several GPUs add their root-child statistics
// before the final choice, as MCTS_tree::merge_root_stats does on the host — slot k = the k-th move of the root's
// generate_all_moves order = the root's k-th child, the same on every rank.  After the sum every rank holds the same
// (simulations, score) per child and the root holds their totals (MCTS_tree::import_root_stats).
int mctsb_tree_merge_root(mctsb_tree *t) {
    if (!t) return MCTSB_ERR_BAD_ARG;
    mctsb_backend *b = t->b;
    int world = 1;
    mctsb_comm_info(b, &world, nullptr);
    if (world <= 1) return MCTSB_OK;                       // a lone backend: nothing to add (as MCTS_tree::merge_root_stats)
    TCK(cudaSetDevice(b->device));
    RootView v;
    int s = read_root(t, v);
    if (s != MCTSB_OK) return s;
    uint16_t n_moves = 0;
    TCK(cudaMemcpy(&n_moves, t->ctx[0].pool.n_moves + v.ctl.root, 2, cudaMemcpyDeviceToHost));
    // every rank must own the same slots: the root has to be fully expanded (true after n_moves iterations)
    if (n_moves == TR_UNLISTED || v.nc != (int)n_moves || v.nc == 0) return MCTSB_ERR_BAD_STATE;
    std::vector<double> stats(2 * (size_t)v.nc);
    for (int i = 0; i < v.nc; i++) { stats[2 * i] = (double)v.sims[i]; stats[2 * i + 1] = v.score[i]; }
    s = mctsb_allreduce_root_stats(b, stats.data(), 2 * v.nc);
    if (s != MCTSB_OK) return s;
    double sims = 0, score = 0;
    for (int i = 0; i < v.nc; i++) { v.sims[i] = (unsigned long long)stats[2 * i]; v.score[i] = stats[2 * i + 1]; sims += stats[2 * i]; score += stats[2 * i + 1]; }
    const TreePool &P = t->ctx[0].pool;
    const unsigned long long root_sims = (unsigned long long)sims;
    TCK(cudaMemcpy(P.sims + v.first, v.sims.data(), (size_t)v.nc * 8, cudaMemcpyHostToDevice));
    TCK(cudaMemcpy(P.score + v.first, v.score.data(), (size_t)v.nc * 8, cudaMemcpyHostToDevice));
    TCK(cudaMemcpy(P.sims + v.ctl.root, &root_sims, 8, cudaMemcpyHostToDevice));
    TCK(cudaMemcpy(P.score + v.ctl.root, &score, 8, cudaMemcpyHostToDevice));
    return MCTSB_OK;
}

int mctsb_tree_root_state(mctsb_tree *t, mctsb_qstate *out) {
    if (!t || !out) return MCTSB_ERR_BAD_ARG;
    mctsb_backend *b = t->b;
    TCK(cudaSetDevice(b->device));
    TreeCtl ctl;
    int s = read_ctl(t, &ctl);
    if (s != MCTSB_OK) return s;
    TCK(cudaMemcpy(out, t->ctx[0].pool.state + ctl.root, sizeof *out, cudaMemcpyDeviceToHost));
    return MCTSB_OK;
}

int mctsb_tree_dump(mctsb_tree *t, uint64_t cap, int32_t *depth, int32_t *move_id, uint64_t *sims, double *score,
                    uint32_t *size, uint32_t *n_children, uint64_t *n) {
    if (!t || !n) return MCTSB_ERR_BAD_ARG;
    mctsb_backend *b = t->b;
    TCK(cudaSetDevice(b->device));
    TreeCtl ctl;
    int s = read_ctl(t, &ctl);
    if (s != MCTSB_OK) return s;
    const size_t N = ctl.pool_top;
    const TreePool &P = t->ctx[0].pool;
    std::vector<int32_t> first(N); std::vector<uint16_t> nc(N); std::vector<uint8_t> mv(N);
    std::vector<unsigned long long> sm(N); std::vector<double> sc(N); std::vector<uint32_t> sz(N);
    TCK(cudaMemcpy(first.data(), P.first_child, N * 4, cudaMemcpyDeviceToHost));
    TCK(cudaMemcpy(nc.data(), P.n_children, N * 2, cudaMemcpyDeviceToHost));
    TCK(cudaMemcpy(mv.data(), P.move, N, cudaMemcpyDeviceToHost));
    TCK(cudaMemcpy(sm.data(), P.sims, N * 8, cudaMemcpyDeviceToHost));
    TCK(cudaMemcpy(sc.data(), P.score, N * 8, cudaMemcpyDeviceToHost));
    TCK(cudaMemcpy(sz.data(), P.size, N * 4, cudaMemcpyDeviceToHost));
    // iterative pre-order walk, children in creation order
    struct Frame { int32_t node; int32_t depth; };
    std::vector<Frame> stack;
    stack.push_back({ctl.root, 0});
    uint64_t k = 0;
    while (!stack.empty()) {
        const Frame fr = stack.back(); stack.pop_back();
        if (k < cap) {
            if (depth) depth[k] = fr.depth;
            if (move_id) move_id[k] = mv[fr.node] == 0xff ? -1 : mv[fr.node];
            if (sims) sims[k] = sm[fr.node];
            if (score) score[k] = sc[fr.node];
            if (size) size[k] = sz[fr.node];
            if (n_children) n_children[k] = nc[fr.node];
        }
        k++;
        for (int i = nc[fr.node] - 1; i >= 0; i--) stack.push_back({first[fr.node] + i, fr.depth + 1});
    }
    *n = k;
    return MCTSB_OK;
}

}  // extern "C"
