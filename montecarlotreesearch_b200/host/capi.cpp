// capi.cpp — flat C entry points over the C++ host side, for ctypes-driven tests and bench scripts.
// (The C++ classes are the interface a reference user programs against; this file only lets Python
// drive them.)  A caller may install its own RolloutBackend through a callback — the tests use that
// to exercise the tree logic on a machine without a GPU; nothing in the product installs one.
#include "mcts.h"
#include "Quoridor.h"
#include "TicTacToe.h"
#include "device_tree.h"
#include "mctsb.h"
#include "../csrc/qcore.cuh"
#include <cstring>
#include <deque>
#include <memory>

extern "C" {
typedef void (*mcts_rollout_cb)(int game, int policy, const void *states, int n_leaves, uint32_t rollouts_per_leaf,
                                uint64_t first_rollout_id, double *score_sum, void *user);
}

namespace {
class CallbackBackend : public RolloutBackend {
    mcts_rollout_cb cb; void *user;
    struct Held { int g, p, n; uint32_t R; uint64_t first; std::vector<unsigned char> states; };
    std::deque<Held> held;
    bool expand = false, have_legal = false;
    std::vector<uint64_t> legal;
    // the masks the move-generation kernel would return, from the same shared rule code compiled for the host:
    // lets a machine without a GPU exercise the tree's mask -> move list path
    void masks(int game, const void *states, int n) {
        have_legal = expand && game == MCTSB_GAME_QUORIDOR;
        if (!have_legal) return;
        legal.assign((size_t)n * 4, 0);
        for (int i = 0; i < n; i++) {
            mctsb::QState q;
            mctsb::q_unpack(q, ((const mctsb_qstate *)states)[i]);
            mctsb::q_legal_moves(q, &legal[4 * (size_t)i]);
        }
    }
public:
    CallbackBackend(mcts_rollout_cb cb, void *user) : cb(cb), user(user) {}
    void set_expand(bool on) override { expand = on; }
    const uint64_t *last_legal() const override { return have_legal ? legal.data() : nullptr; }
    void rollout_batch(int game, int policy, const void *states, int n_leaves, uint32_t rpl, uint64_t first_id, double *out) override {
        cb(game, policy, states, n_leaves, rpl, first_id, out, user);
        masks(game, states, n_leaves);
    }
    void submit(int game, int policy, const void *states, int n_leaves, uint32_t rpl, uint64_t first_id) override {
        size_t rec = game == MCTSB_GAME_QUORIDOR ? sizeof(mctsb_qstate) : sizeof(mctsb_tstate);
        Held h{game, policy, n_leaves, rpl, first_id, {}};
        h.states.assign((const unsigned char *)states, (const unsigned char *)states + rec * n_leaves);
        held.push_back(std::move(h));
    }
    void wait(double *out) override {
        Held h = std::move(held.front()); held.pop_front();
        cb(h.g, h.p, h.states.data(), h.n, h.R, h.first, out, user);
        masks(h.g, h.states.data(), h.n);
    }
    int max_in_flight() const override { return 4; }
};

struct HostSession {
    std::unique_ptr<RolloutBackend> backend;
    std::unique_ptr<MCTS_tree> tree;
    int game = 0;
};

MCTS_state *make_state(int game, const void *packed, int policy) {
    if (game == MCTSB_GAME_QUORIDOR) return new Quoridor_state(*(const mctsb_qstate *)packed, policy);
    const mctsb_tstate *t = (const mctsb_tstate *)packed;
    return new TicTacToe_state(t->x_mask & 0x1ff, t->o_mask & 0x1ff, (t->o_mask & 0x8000) ? 'o' : 'x');
}

int move_index(int game, const MCTS_move *m) {
    if (game == MCTSB_GAME_QUORIDOR) return ((const Quoridor_move *)m)->canonical_id();
    const TicTacToe_move *t = (const TicTacToe_move *)m;
    return t->x * 3 + t->y;
}

MCTS_move *make_move(int game, int id, const MCTS_state *s) {
    if (game == MCTSB_GAME_QUORIDOR) return Quoridor_move::from_canonical_id(id, ((const Quoridor_state *)s)->whose_turn());
    return new TicTacToe_move(id / 3, id % 3, ((const TicTacToe_state *)s)->get_turn());
}
}  // namespace

extern "C" {

// ---- rule-level entry points of the host game class (CPU only, no rollouts) ----------------------
int mcts_host_q_actions(const mctsb_qstate *s, int32_t *ids, int max_ids) {
    Quoridor_state q(*s);
    queue<MCTS_move *> *Q = q.actions_to_try();
    int n = 0;
    while (!Q->empty()) { if (n < max_ids) ids[n] = ((Quoridor_move *)Q->front())->canonical_id(); n++; delete Q->front(); Q->pop(); }
    delete Q;
    return n;
}

int mcts_host_q_good_actions(const mctsb_qstate *s, int32_t *ids, int max_ids) {
    Quoridor_state q(*s);
    q.set_good_moves_only(true);
    queue<MCTS_move *> *Q = q.actions_to_try();
    int n = 0;
    while (!Q->empty()) { if (n < max_ids) ids[n] = ((Quoridor_move *)Q->front())->canonical_id(); n++; delete Q->front(); Q->pop(); }
    delete Q;
    return n;
}

int mcts_host_q_next(const mctsb_qstate *s, int id, mctsb_qstate *out) {
    Quoridor_state q(*s);
    Quoridor_move *m = Quoridor_move::from_canonical_id(id, q.whose_turn());
    bool ok = q.legal_move(m);
    if (ok) { MCTS_state *n = q.next_state(m); ((Quoridor_state *)n)->pack(*out); delete n; } else *out = *s;
    delete m;
    return ok ? 1 : 0;
}

int mcts_host_q_info(const mctsb_qstate *s, int32_t *info6, double *eval) {
    Quoridor_state q(*s);
    char w = q.check_winner();
    info6[0] = w == 'W' ? 1 : w == 'B' ? 2 : 0;
    info6[1] = q.get_shortest_path('W'); info6[2] = q.get_shortest_path('B');
    info6[3] = q.force_playwall() ? 1 : 0; info6[4] = q.is_terminal() ? 1 : 0; info6[5] = q.player1_turn() ? 1 : 0;
    if (eval) *eval = q.evaluate();
    Quoridor_move *b = q.is_terminal() ? NULL : q.get_best_step_move(q.whose_turn());
    int r = b ? b->canonical_id() : -1;
    delete b;
    return r;
}

int mcts_host_q_sprint(const mctsb_qstate *s, int id, char *buf, int cap) {
    Quoridor_state q(*s);
    Quoridor_move *m = Quoridor_move::from_canonical_id(id, q.whose_turn());
    string t = m->sprint();
    delete m;
    strncpy(buf, t.c_str(), cap - 1); buf[cap - 1] = 0;
    return (int)t.size();
}

// ---- tree sessions ---------------------------------------------------------------------------------
// backend: cb == NULL -> CudaRolloutBackend(device, seed) (fails without a GPU); else the caller's callback.
void *mcts_host_session_create(int game, const void *packed_root, int policy, int device, uint64_t seed,
                               uint32_t rollouts_per_leaf, uint32_t leaves_per_flush, mcts_rollout_cb cb, void *user,
                               char *err, int err_cap) {
    try {
        mcts_set_verbose(false);
        std::unique_ptr<HostSession> h(new HostSession());
        if (cb) h->backend.reset(new CallbackBackend(cb, user));
        else h->backend.reset(new CudaRolloutBackend(device, seed));
        MCTS_batching b; b.rollouts_per_leaf = rollouts_per_leaf; b.leaves_per_flush = leaves_per_flush;
        h->game = game;
        h->tree.reset(new MCTS_tree(make_state(game, packed_root, policy), h->backend.get(), b));
        return h.release();
    } catch (const std::exception &e) {
        if (err && err_cap > 0) { strncpy(err, e.what(), err_cap - 1); err[err_cap - 1] = 0; }
        return NULL;
    }
}

void mcts_host_session_destroy(void *hs) { delete (HostSession *)hs; }

void mcts_host_set_overlap(void *hs, int on) {
    HostSession *h = (HostSession *)hs;
    MCTS_batching b = h->tree->get_batching();
    b.overlap = on != 0; b.flushes_in_flight = on > 1 ? (uint32_t)on : 1;     // on = number of flushes in flight (0 = synchronous)
    h->tree->set_batching(b);
}

// Quoridor only, before the first grow: the tree expands generate_good_moves() instead of generate_all_moves()
int mcts_host_set_good_moves_only(void *hs, int on) {
    HostSession *h = (HostSession *)hs;
    if (h->game != MCTSB_GAME_QUORIDOR || h->tree->get_size() > 0) return -1;
    ((Quoridor_state *)h->tree->get_current_state())->set_good_moves_only(on != 0);
    return 0;
}

void mcts_host_set_gpu_expand(void *hs, int on) {
    HostSession *h = (HostSession *)hs;
    MCTS_batching b = h->tree->get_batching();
    b.gpu_expand = on != 0;
    h->tree->set_batching(b);
}

// wall seconds inside grow_tree so far, and the part of them inside backend calls (the rest is the host tree)
void mcts_host_times(void *hs, double *total, double *backend) {
    HostSession *h = (HostSession *)hs;
    *total = h->tree->get_seconds_total(); *backend = h->tree->get_seconds_backend();
}

// ---- multi-GPU exchange: NCCL behind the C ABI, reached through the session's CUDA backend -----------------
int mcts_host_comm_unique_id(void *id128) { return mctsb_comm_unique_id(id128); }

static int comm_fail(const std::exception &e, char *err, int err_cap) {
    if (err && err_cap > 0) { strncpy(err, e.what(), err_cap - 1); err[err_cap - 1] = 0; }
    return -1;
}

int mcts_host_comm_init_rank(void *hs, const void *id128, int world, int rank, char *err, int err_cap) {
    try {
        CudaRolloutBackend *be = dynamic_cast<CudaRolloutBackend *>(((HostSession *)hs)->backend.get());
        if (!be) throw std::runtime_error("the session does not run on the CUDA backend");
        be->comm_init_rank(id128, world, rank);
        return 0;
    } catch (const std::exception &e) { return comm_fail(e, err, err_cap); }
}

// root-parallel trees: sum the root-child statistics over the ranks (MCTS_tree::merge_root_stats)
int mcts_host_merge_root(void *hs, char *err, int err_cap) {
    try { ((HostSession *)hs)->tree->merge_root_stats(); return 0; }
    catch (const std::exception &e) { return comm_fail(e, err, err_cap); }
}

// the same sum over any small vector of doubles (move agreement checks, totals)
int mcts_host_allreduce(void *hs, double *v, int n, char *err, int err_cap) {
    try { ((HostSession *)hs)->backend->allreduce_sum(v, n); return 0; }
    catch (const std::exception &e) { return comm_fail(e, err, err_cap); }
}

int mcts_host_grow(void *hs, int max_iter, double max_seconds, char *err, int err_cap) {
    try { ((HostSession *)hs)->tree->grow_tree(max_iter, max_seconds); return 0; }
    catch (const std::exception &e) { if (err && err_cap > 0) { strncpy(err, e.what(), err_cap - 1); err[err_cap - 1] = 0; } return -1; }
}

// out4 = {root simulations, root score, tree size, #children}; children: {move id, simulations, score} each
int mcts_host_root_stats(void *hs, double *out4, double *children, int max_children) {
    HostSession *h = (HostSession *)hs;
    const MCTS_node *root = h->tree->get_root();
    out4[0] = root->get_number_of_simulations(); out4[1] = root->get_score(); out4[2] = root->get_size();
    out4[3] = (double)root->get_children().size();
    int n = 0;
    for (const MCTS_node *c : root->get_children()) {
        if (n >= max_children) break;
        children[3 * n] = move_index(h->game, c->get_move()); children[3 * n + 1] = c->get_number_of_simulations();
        children[3 * n + 2] = c->get_score(); n++;
    }
    return n;
}

// pre-order walk from the root, children in their current order: one row per node (tests: the device-resident tree
// of csrc/tree_body.cuh must hold the same nodes with the same statistics)
static void dump_walk(int game, const MCTS_node *n, int depth, uint64_t cap, uint64_t &k, int32_t *o_depth, int32_t *o_move,
                      uint64_t *o_sims, double *o_score, uint32_t *o_size, uint32_t *o_nchildren) {
    if (k < cap) {
        o_depth[k] = depth; o_move[k] = n->get_move() ? move_index(game, n->get_move()) : -1;
        o_sims[k] = n->get_number_of_simulations(); o_score[k] = n->get_score(); o_size[k] = n->get_size();
        o_nchildren[k] = (uint32_t)n->get_children().size();
    }
    k++;
    for (const MCTS_node *c : n->get_children()) dump_walk(game, c, depth + 1, cap, k, o_depth, o_move, o_sims, o_score, o_size, o_nchildren);
}

uint64_t mcts_host_dump_tree(void *hs, uint64_t cap, int32_t *o_depth, int32_t *o_move, uint64_t *o_sims, double *o_score,
                             uint32_t *o_size, uint32_t *o_nchildren) {
    HostSession *h = (HostSession *)hs;
    uint64_t k = 0;
    dump_walk(h->game, h->tree->get_root(), 0, cap, k, o_depth, o_move, o_sims, o_score, o_size, o_nchildren);
    return k;
}

int mcts_host_best_move(void *hs) {
    HostSession *h = (HostSession *)hs;
    MCTS_node *b = h->tree->select_best_child();
    return b ? move_index(h->game, b->get_move()) : -1;
}

int mcts_host_advance(void *hs, int move_id) {
    HostSession *h = (HostSession *)hs;
    MCTS_move *m = make_move(h->game, move_id, h->tree->get_current_state());
    h->tree->advance_tree(m);
    delete m;    // the tree keeps its own copy (the child's move) or built a fresh root
    return h->tree->get_current_state()->is_terminal() ? 1 : 0;
}

int mcts_host_current_state(void *hs, void *packed_out) {
    HostSession *h = (HostSession *)hs;
    return h->tree->get_current_state()->gpu_pack(packed_out) ? 0 : -1;
}

// ---- the C++ face of the device-resident tree (MCTS_device_tree): same calls as a host-tree session ------------------
struct DeviceTreeSession { std::unique_ptr<CudaRolloutBackend> backend; std::unique_ptr<MCTS_device_tree> tree; };

void *mcts_host_dtree_create(const mctsb_qstate *root, int policy, int device, uint64_t seed, uint32_t rollouts_per_leaf,
                             uint32_t leaves_per_flush, int overlap, uint64_t max_nodes, char *err, int err_cap) {
    try {
        mcts_set_verbose(false);
        std::unique_ptr<DeviceTreeSession> h(new DeviceTreeSession());
        h->backend.reset(new CudaRolloutBackend(device, seed));
        MCTS_batching b; b.rollouts_per_leaf = rollouts_per_leaf; b.leaves_per_flush = leaves_per_flush; b.overlap = overlap != 0;
        Quoridor_state start(*root, policy);
        h->tree.reset(new MCTS_device_tree(start, h->backend.get(), b, max_nodes));
        return h.release();
    } catch (const std::exception &e) { comm_fail(e, err, err_cap); return NULL; }
}
void mcts_host_dtree_destroy(void *hs) { delete (DeviceTreeSession *)hs; }
int mcts_host_dtree_grow(void *hs, int max_iter, double max_seconds, char *err, int err_cap) {
    try { ((DeviceTreeSession *)hs)->tree->grow_tree(max_iter, max_seconds); return 0; }
    catch (const std::exception &e) { return comm_fail(e, err, err_cap); }
}
int mcts_host_dtree_best_move(void *hs) {
    try {
        Quoridor_move *m = ((DeviceTreeSession *)hs)->tree->select_best_child();
        if (m == NULL) return -1;
        const int id = m->canonical_id();
        delete m;
        return id;
    } catch (const std::exception &) { return -2; }
}
int mcts_host_dtree_advance(void *hs, int move_id) {
    try {
        DeviceTreeSession *h = (DeviceTreeSession *)hs;
        Quoridor_move *m = Quoridor_move::from_canonical_id(move_id, ((const Quoridor_state *)h->tree->get_current_state())->whose_turn());
        h->tree->advance_tree(m);
        delete m;
        return h->tree->get_current_state()->is_terminal() ? 1 : 0;
    } catch (const std::exception &) { return -1; }
}
unsigned int mcts_host_dtree_size(void *hs) { return ((DeviceTreeSession *)hs)->tree->get_size(); }
void mcts_host_dtree_print_stats(void *hs) { ((DeviceTreeSession *)hs)->tree->print_stats(); }

// the Philox stream id the next flush of ANY tree in this process will start at (ids are never reused)
uint64_t mcts_host_peek_rollout_id(void) { return next_rollout_ids(0); }
// tests only: two trees that must draw the same streams are started from the same id
void mcts_host_set_next_rollout_id(uint64_t id) { set_next_rollout_id(id); }

void mcts_host_counters(void *hs, uint64_t *flushes, uint64_t *rollouts) {
    HostSession *h = (HostSession *)hs;
    *flushes = h->tree->get_flushes(); *rollouts = h->tree->get_rollouts_played();
}

int mcts_host_export_root(void *hs, double *buf, int cap) {
    std::vector<double> v; ((HostSession *)hs)->tree->export_root_stats(v);
    if ((int)v.size() > cap) return -(int)v.size();
    memcpy(buf, v.data(), v.size() * sizeof(double));
    return (int)v.size();
}

int mcts_host_import_root(void *hs, const double *buf, int n) {
    try { std::vector<double> v(buf, buf + n); ((HostSession *)hs)->tree->import_root_stats(v); return 0; }
    catch (const std::exception &) { return -1; }
}

}  // extern "C"
