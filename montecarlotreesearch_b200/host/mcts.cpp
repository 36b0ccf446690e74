// mcts.cpp — host search tree (Selection, Expansion, Backpropagation) with the Simulation phase
// flushed to the GPU rollout backend.  Behavioural contract: the reference's mcts/src/mcts.cpp
// (cited per function); the code is written from scratch.
#include "mcts.h"
#include "mctsb.h"
#include <algorithm>
#include <cassert>
#include <cmath>
#include <chrono>
#include <ctime>

static bool g_verbose = true;
void mcts_set_verbose(bool on) { g_verbose = on; }

/*** MCTS_node ***/

// reference mcts.cpp:15-21; the node owns state, move, children and whatever stays in untried_actions
// (mcts.cpp:23-35).  The reference fills untried_actions in the constructor; here the list is generated the
// first time the node is asked for it — same list, same order, but a leaf at the fringe of the tree (most
// nodes: they are played out once and never selected again) never pays for move generation.
MCTS_node::MCTS_node(MCTS_node *parent, MCTS_state *state, const MCTS_move *move, MCTS_tree *tree)
    : terminal(false), size(0), number_of_simulations(0), score(0.0), state(state), move(move),
      children(new vector<MCTS_node *>()), index_in_parent(0), parent(parent), untried_actions(NULL), next_untried_id(0), listed(false), listed_as_ids(false), tree(tree), virtual_visits(0) {
    terminal = state->is_terminal();
    p1_to_move = state->player1_turn();
}

void MCTS_node::list_actions() const {
    if (listed) return;
    listed = true;
    listed_as_ids = state->actions_to_try_ids(untried_ids);
    if (!listed_as_ids) untried_actions = state->actions_to_try();
    children->reserve(STARTING_NUMBER_OF_CHILDREN);
    MCTS_node *self = const_cast<MCTS_node *>(this);
    self->child_stats.reserve(STARTING_NUMBER_OF_CHILDREN);
    self->rank_exploit.reserve(STARTING_NUMBER_OF_CHILDREN); self->rank_inv_sqrt.reserve(STARTING_NUMBER_OF_CHILDREN);
}

bool MCTS_node::no_untried_action() const {
    list_actions();
    return listed_as_ids ? next_untried_id >= untried_ids.size() : untried_actions->empty();
}

MCTS_node::~MCTS_node() {
    delete state;
    delete move;
    for (MCTS_node *c : *children) delete c;
    delete children;
    if (untried_actions != NULL) {
        while (!untried_actions->empty()) { delete untried_actions->front(); untried_actions->pop(); }
        delete untried_actions;
    }
}

bool MCTS_node::is_fully_expanded() const { return terminal || no_untried_action(); }   // mcts.cpp:92-94
bool MCTS_node::is_terminal() const { return terminal; }
unsigned int MCTS_node::get_size() const { return size; }
const MCTS_move *MCTS_node::get_move() const { return move; }
const MCTS_state *MCTS_node::get_current_state() const { return state; }

// Expansion only: pop the FRONT of the FIFO (children appear in actions_to_try order, mcts.cpp:46-54)
MCTS_node *MCTS_node::expand_no_rollout() {
    if (terminal || no_untried_action()) return NULL;
    MCTS_move *next_move;
    if (listed_as_ids) next_move = state->move_of_id(untried_ids[next_untried_id++]);
    else { next_move = untried_actions->front(); untried_actions->pop(); }
    MCTS_state *next = state->next_state_of_listed_move(next_move);    // the move comes from this state's own list
    MCTS_node *child = new MCTS_node(this, next, next_move, tree);
    child->index_in_parent = (uint32_t)children->size();
    children->push_back(child);
    child_stats.push_back(MCTS_child_stat{0.0, 0.0, 0.0});
    rank_exploit.push_back(0.0f); rank_inv_sqrt.push_back(0.0f);
    return child;
}

// reference mcts.cpp:37-55
void MCTS_node::expand() {
    if (is_terminal()) { rollout(); return; }          // terminal nodes are re-scored (mcts.cpp:38-40)
    if (is_fully_expanded()) { cerr << "Warning: Cannot expanded this node any more!" << endl; return; }
    MCTS_node *child = expand_no_rollout();
    child->rollout();
}

// Simulation phase of ONE leaf.  Replaces mcts.cpp:57-81: instead of NUMBER_OF_THREADS RolloutJobs on
// the pthread pool, `rollouts_per_leaf` playouts of this node's state run on the GPU and their sum
// is back-propagated with the matching count.
void MCTS_node::rollout() {
    uint32_t R = tree ? tree->batching.rollouts_per_leaf : 4;
    int game = state->gpu_game_id();
    if (game >= 0) {
        unsigned char packed[32];
        if (!state->gpu_pack(packed)) throw std::runtime_error("gpu_pack failed for a state that advertises a gpu_game_id");
        RolloutBackend *be = tree ? tree->get_backend() : default_rollout_backend();
        double sum = 0.0;
        be->rollout_batch(game, state->gpu_policy_id(), packed, 1, R, next_rollout_ids(R), &sum);
        if (tree) { tree->flushes++; tree->rollouts_played += R; }
        backpropagate(sum, R);
        return;
    }
    // a game the backend does not know: its own virtual rollout(), validated like mcts.cpp:68-75
    double sum = 0.0;
    for (uint32_t i = 0; i < R; i++) {
        double r = state->rollout();
        if (r >= 0.0 && r <= 1.0) sum += r;
        else cerr << "Warning: Invalid result when aggregating parallel rollouts" << endl;
    }
    backpropagate(sum, R);
}

// the two numbers the ranking pass of select_best_child keeps per child (float: the pass only has to find the
// children that CAN be the maximum)
static inline void rank_terms(const MCTS_child_stat &c, bool p1, float &exploit, float &inv_sqrt) {
    const double s = c.score + (p1 ? 0.0 : 1.0) * c.virtual_visits;
    exploit = (float)((p1 ? 0.0 : 1.0) + (p1 ? 1.0 : -1.0) * (s / c.visits));
    inv_sqrt = (float)(1.0 / sqrt(c.visits));
}

void MCTS_node::mirror_to_parent() {
    if (parent == NULL) return;
    MCTS_child_stat &c = parent->child_stats[index_in_parent];
    c.score = score; c.visits = (double)(number_of_simulations + virtual_visits); c.virtual_visits = (double)virtual_visits;
    rank_terms(c, parent->p1_to_move, parent->rank_exploit[index_in_parent], parent->rank_inv_sqrt[index_in_parent]);
}

void MCTS_node::rebuild_child_stats() {
    child_stats.resize(children->size());
    rank_exploit.resize(children->size()); rank_inv_sqrt.resize(children->size());
    for (size_t i = 0; i < children->size(); i++) {
        MCTS_node *c = children->at(i);
        c->index_in_parent = (uint32_t)i;
        child_stats[i] = MCTS_child_stat{c->score, (double)(c->number_of_simulations + c->virtual_visits), (double)c->virtual_visits};
        rank_terms(child_stats[i], p1_to_move, rank_exploit[i], rank_inv_sqrt[i]);
    }
}

// reference mcts.cpp:83-90: score += w; n += n; every ancestor's `size` counts the event
void MCTS_node::backpropagate(double w, uint64_t n) {
    for (MCTS_node *node = this; node != NULL; node = node->parent) {
        node->score += w;
        node->number_of_simulations += n;
        node->mirror_to_parent();
        if (node->parent != NULL) node->parent->size++;
    }
}

void MCTS_node::add_virtual(int64_t n) {
    for (MCTS_node *node = this; node != NULL; node = node->parent) {
        node->virtual_visits = (uint64_t)((int64_t)node->virtual_visits + n);
        node->mirror_to_parent();
    }
}

// reference mcts.cpp:104-130.  UCT on the win rate of the side to move AT THIS NODE; first strict
// maximum wins; c <= 0 means pure win rate.  Simulations still in flight (virtual_visits) count as
// losses for the side choosing here — zero unless leaves_per_flush > 1.  Reads the parent's own array of
// child statistics, never the children.
//
// The reference's expression costs two divisions and a square root per child; a Quoridor node has ~131 children and
// a selection passes three such nodes, each with its statistics in a different corner of memory.  Here a first pass
// ranks the children with one multiply-add each over two dense float arrays (exploit + c*sqrt(log N) * n^-1/2: the
// reference's value up to ~1e-6), and only the children within 1e-5 of the best — almost always one — are
// evaluated with the reference's own double-precision expression, in index order.  Every child whose exact value
// could be the maximum is in that set, so the first strict maximum is the reference's.
__attribute__((target_clones("avx2", "default")))
MCTS_node *MCTS_node::select_best_child(double c) const {
    if (children->empty()) return NULL;
    if (children->size() == 1) return children->at(0);
    const bool p1 = p1_to_move;
    const double n_parent = (double)(number_of_simulations + virtual_visits);
    const double log_parent = c > 0 ? log(n_parent) : 0.0;      // the same value for every child (mcts.cpp:121 computes it per child)
    const MCTS_child_stat *cs = child_stats.data();
    const size_t e = children->size();
    const double virt_as_loss = p1 ? 0.0 : 1.0, sign = p1 ? 1.0 : -1.0, base = p1 ? 0.0 : 1.0;
    // pass 1: the maximum of the approximate values
    const float coef = c > 0 ? (float)(c * sqrt(log_parent)) : 0.0f;
    const float *re = rank_exploit.data(), *ri = rank_inv_sqrt.data();
    float amax = -1.0f;
    for (size_t i = 0; i < e; i++) {
        const float a = re[i] + coef * ri[i];
        amax = a > amax ? a : amax;
    }
    double best = -1;
    size_t argmax = e;
    const bool finite = amax - amax == 0.0f;                     // false when a child has no visit yet (n = 0): evaluate everybody
    const float threshold = finite ? amax - 1e-5f * (1.0f + fabsf(amax)) : 0.0f;
    // pass 2: the reference's expression, term by term, for the children that can be the maximum
    for (size_t i = 0; i < e; i++) {
        if (finite && !(re[i] + coef * ri[i] >= threshold)) continue;
        const double n = cs[i].visits;
        const double s = cs[i].score + virt_as_loss * cs[i].virtual_visits;
        const double exact = c > 0 ? (base + sign * (s / n)) + c * sqrt(log_parent / n) : base + sign * (s / n);
        if (exact > best) { best = exact; argmax = i; }
    }
    return argmax < e ? children->at(argmax) : NULL;
}

// reference mcts.cpp:132-157
MCTS_node *MCTS_node::advance_tree(const MCTS_move *m) {
    MCTS_node *next = NULL;
    for (MCTS_node *child : *children) {
        if (*(child->move) == *m) next = child;
        else delete child;
    }
    children->clear();
    child_stats.clear(); rank_exploit.clear(); rank_inv_sqrt.clear();
    if (next == NULL) {
        if (g_verbose) cout << "INFO: Didn't find child node. Had to start over." << endl;
        next = new MCTS_node(NULL, state->next_state(m), NULL, tree);
    } else {
        next->parent = NULL;
    }
    return next;
}

double MCTS_node::calculate_winrate(bool player1turn) const {   // mcts.cpp:252-258
    return player1turn ? score / number_of_simulations : 1.0 - score / number_of_simulations;
}

// reference mcts.cpp:222-250 (note: sorts the children in place)
void MCTS_node::print_stats() const {
    const unsigned TOPK = 10;
    if (number_of_simulations == 0) { cout << "Tree not expanded yet" << endl; return; }
    cout << "___ INFO _______________________" << endl
         << "Tree size: " << size << endl
         << "Number of simulations: " << number_of_simulations << endl
         << "Branching factor at root: " << children->size() << endl
         << "Chances of P1 winning: " << setprecision(4) << 100.0 * (score / number_of_simulations) << "%" << endl;
    const bool p1 = state->player1_turn();
    std::sort(children->begin(), children->end(), [p1](const MCTS_node *a, const MCTS_node *b) {
        return a->calculate_winrate(p1) > b->calculate_winrate(p1);
    });
    const_cast<MCTS_node *>(this)->rebuild_child_stats();
    cout << "Best moves:" << endl;
    for (unsigned i = 0; i < children->size() && i < TOPK; i++)
        cout << "  " << i + 1 << ". " << children->at(i)->move->sprint() << "  -->  "
             << setprecision(4) << 100.0 * children->at(i)->calculate_winrate(p1) << "%" << endl;
    cout << "________________________________" << endl;
}


/*** MCTS_tree ***/

MCTS_tree::MCTS_tree(MCTS_state *starting_state) : root(NULL), backend(NULL), flushes(0), rollouts_played(0), seconds_total(0), seconds_backend(0) {
    assert(starting_state != NULL);
    root = new MCTS_node(NULL, starting_state, NULL, this);
}

MCTS_tree::MCTS_tree(MCTS_state *starting_state, RolloutBackend *backend, MCTS_batching batching)
    : root(NULL), backend(backend), batching(batching), flushes(0), rollouts_played(0), seconds_total(0), seconds_backend(0) {
    assert(starting_state != NULL);
    root = new MCTS_node(NULL, starting_state, NULL, this);
}

MCTS_tree::~MCTS_tree() { delete root; }

RolloutBackend *MCTS_tree::get_backend() { return backend ? backend : default_rollout_backend(); }

// reference mcts.cpp:161-171
MCTS_node *MCTS_tree::select(double c) {
    MCTS_node *node = root;
    while (!node->is_terminal()) {
        if (!node->is_fully_expanded()) return node;
        node = node->select_best_child(c);
    }
    return node;
}

namespace {
struct BackendTimer {                        // adds the lifetime of the object to a seconds counter
    double &acc; std::chrono::steady_clock::time_point t0;
    explicit BackendTimer(double &acc) : acc(acc), t0(std::chrono::steady_clock::now()) {}
    ~BackendTimer() { acc += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};
}

// Expansion prepared on the device (MCTS_batching::gpu_expand): the backend's move-generation kernel ran over
// the leaves of the submission that was just collected; every leaf keeps its mask for the day it is expanded.
void MCTS_tree::apply_legal_masks(const vector<MCTS_node *> &leaves, size_t first, size_t count) {
    if (!batching.gpu_expand) return;
    const uint64_t *legal = get_backend()->last_legal();
    if (legal == NULL) return;
    for (size_t k = 0; k < count; k++) {
        MCTS_node *leaf = leaves[first + k];
        if (!leaf->listed && !leaf->terminal) leaf->state->gpu_set_legal(legal + 4 * k);
    }
}

// collect everything that is still on the GPU and give the unflushed leaves their virtual loss back: the tree
// is consistent again (used when a backend call throws in the middle of grow_tree)
void MCTS_tree::drain(deque<vector<MCTS_node *>> &inflight, vector<MCTS_node *> &leaves) {
    const uint32_t R = batching.rollouts_per_leaf;
    while (!inflight.empty()) {
        try { collect_leaves(inflight.front()); }
        catch (...) { for (MCTS_node *l : inflight.front()) l->add_virtual(-(int64_t)R); }
        inflight.pop_front();
    }
    for (MCTS_node *l : leaves) l->add_virtual(-(int64_t)R);
    leaves.clear();
}

// One kernel launch for all the leaves collected by grow_tree: every leaf gets rollouts_per_leaf
// playouts; results come back as one exact sum per leaf and are back-propagated in selection order.
void MCTS_tree::flush_leaves(vector<MCTS_node *> &leaves) {
    if (leaves.empty()) return;
    const uint32_t R = batching.rollouts_per_leaf;
    size_t i = 0;
    while (i < leaves.size()) {
        // group consecutive leaves that the backend can take in one call (same game/policy); others fall back
        int game = leaves[i]->state->gpu_game_id();
        if (game < 0) { leaves[i]->add_virtual(-(int64_t)R); leaves[i]->rollout(); i++; continue; }
        int policy = leaves[i]->state->gpu_policy_id();
        size_t rec = game == MCTSB_GAME_QUORIDOR ? sizeof(mctsb_qstate) : sizeof(mctsb_tstate), j = i;
        vector<unsigned char> packed;
        while (j < leaves.size() && leaves[j]->state->gpu_game_id() == game && leaves[j]->state->gpu_policy_id() == policy) {
            packed.resize((j - i + 1) * rec);
            if (!leaves[j]->state->gpu_pack(&packed[(j - i) * rec])) throw std::runtime_error("gpu_pack failed");
            j++;
        }
        vector<double> sums(j - i, 0.0);
        {
            BackendTimer t(seconds_backend);
            get_backend()->rollout_batch(game, policy, packed.data(), (int)(j - i), R, next_rollout_ids((uint64_t)R * (j - i)), sums.data());
        }
        flushes++; rollouts_played += (uint64_t)R * (j - i);
        apply_legal_masks(leaves, i, j - i);
        for (size_t k = i; k < j; k++) {
            leaves[k]->add_virtual(-(int64_t)R);
            leaves[k]->backpropagate(sums[k - i], R);
        }
        i = j;
    }
    leaves.clear();
}

// Asynchronous flush: all leaves must belong to one backend game/policy (true for a game tree).
bool MCTS_tree::submit_leaves(const vector<MCTS_node *> &leaves) {
    if (leaves.empty()) return false;
    const int game = leaves[0]->state->gpu_game_id(), policy = leaves[0]->state->gpu_policy_id();
    if (game < 0) return false;
    const size_t rec = game == MCTSB_GAME_QUORIDOR ? sizeof(mctsb_qstate) : sizeof(mctsb_tstate);
    vector<unsigned char> packed(leaves.size() * rec);
    for (size_t k = 0; k < leaves.size(); k++) {
        if (leaves[k]->state->gpu_game_id() != game || leaves[k]->state->gpu_policy_id() != policy) return false;
        if (!leaves[k]->state->gpu_pack(&packed[k * rec])) return false;
    }
    const uint32_t R = batching.rollouts_per_leaf;
    {
        BackendTimer t(seconds_backend);
        get_backend()->submit(game, policy, packed.data(), (int)leaves.size(), R, next_rollout_ids((uint64_t)R * leaves.size()));
    }
    flushes++; rollouts_played += (uint64_t)R * leaves.size();
    return true;
}

void MCTS_tree::collect_leaves(vector<MCTS_node *> &leaves) {
    vector<double> sums(leaves.size(), 0.0);
    {
        BackendTimer t(seconds_backend);
        get_backend()->wait(sums.data());
    }
    apply_legal_masks(leaves, 0, leaves.size());
    const uint32_t R = batching.rollouts_per_leaf;
    for (size_t k = 0; k < leaves.size(); k++) {
        leaves[k]->add_virtual(-(int64_t)R);
        leaves[k]->backpropagate(sums[k], R);
    }
    leaves.clear();
}

// reference mcts.cpp:182-210: up to max_iter Selection+Expansion+Simulation+Backpropagation rounds,
// stopped early when difftime(now, start) > max_time (whole seconds, strict).
void MCTS_tree::grow_tree(int max_iter, double max_time_in_seconds) {
    if (g_verbose) cout << "Growing tree..." << endl;
    BackendTimer whole(seconds_total);
    get_backend()->set_expand(batching.gpu_expand);
    time_t start_t, now_t;
    time(&start_t);
    double dt = 0;
    const uint32_t K = batching.leaves_per_flush == 0 ? 1 : batching.leaves_per_flush;
    int i = 0;
    if (K == 1) {
        for (; i < max_iter; i++) {
            MCTS_node *node = select();
            node->expand();
            time(&now_t);
            dt = difftime(now_t, start_t);
            if (dt > max_time_in_seconds) {
                if (g_verbose) cout << "Early stopping: Made " << (i + 1) << " iterations in " << dt << " seconds." << endl;
                break;
            }
        }
    } else {
        vector<MCTS_node *> leaves;
        deque<vector<MCTS_node *>> inflight;                          // flushes on the GPU, oldest first
        size_t depth = 0;
        if (batching.overlap) {
            depth = batching.flushes_in_flight == 0 ? 1 : batching.flushes_in_flight;
            const size_t cap = (size_t)get_backend()->max_in_flight();
            if (depth > cap) depth = cap;
        }
        try {
        while (i < max_iter || !inflight.empty()) {
            for (uint32_t k = 0; k < K && i < max_iter; k++, i++) {
                MCTS_node *node = select();
                MCTS_node *leaf = node->is_terminal() ? node : node->expand_no_rollout();
                if (leaf == NULL) leaf = node;                        // cannot happen: select() returns expandable or terminal nodes
                leaf->add_virtual((int64_t)batching.rollouts_per_leaf);   // keep the next selection away from this path
                leaves.push_back(leaf);
            }
            // the GPU ran the queued flushes while the host selected: take the oldest results when the queue is
            // full (or nothing is left to select), then queue the new batch behind whatever is still running
            if (!inflight.empty() && (inflight.size() >= depth || leaves.empty())) {
                collect_leaves(inflight.front());
                inflight.pop_front();
            }
            if (!leaves.empty()) {
                if (depth > 0 && submit_leaves(leaves)) { inflight.emplace_back(); inflight.back().swap(leaves); }
                else {
                    while (!inflight.empty()) { collect_leaves(inflight.front()); inflight.pop_front(); }
                    flush_leaves(leaves);
                }
            }
            time(&now_t);
            dt = difftime(now_t, start_t);
            if (dt > max_time_in_seconds) {
                if (g_verbose) cout << "Early stopping: Made " << i << " iterations in " << dt << " seconds." << endl;
                while (!inflight.empty()) { collect_leaves(inflight.front()); inflight.pop_front(); }
                break;
            }
        }
        } catch (...) {
            // a backend call failed: no virtual loss may stay behind and nothing may stay queued on the GPU
            drain(inflight, leaves);
            throw;
        }
    }
    if (g_verbose) {
        time(&now_t);
        cout << "Finished in " << difftime(now_t, start_t) << " seconds." << endl;
    }
}

unsigned int MCTS_tree::get_size() const { return root->get_size(); }
const MCTS_state *MCTS_tree::get_current_state() const { return root->get_current_state(); }
MCTS_node *MCTS_tree::select_best_child() { return root->select_best_child(0.0); }   // mcts.cpp:268-270
void MCTS_tree::print_stats() const { root->print_stats(); }

void MCTS_tree::advance_tree(const MCTS_move *move) {   // mcts.cpp:260-264
    MCTS_node *old_root = root;
    root = root->advance_tree(move);
    delete old_root;
}

// ---- root-parallel merge ---------------------------------------------------------------------------
// Slot k belongs to the k-th move of root->state->actions_to_try() (the same list on every rank), NOT to
// the k-th child: children exist only as far as expanded and print_stats() re-sorts them (mcts.cpp:234-242).
static vector<MCTS_move *> root_action_list(const MCTS_state *s) {
    vector<MCTS_move *> v;
    queue<MCTS_move *> *q = s->actions_to_try();
    while (!q->empty()) { v.push_back(q->front()); q->pop(); }
    delete q;
    return v;
}

void MCTS_tree::export_root_stats(vector<double> &out) const {
    vector<MCTS_move *> acts = root_action_list(root->state);
    out.assign(2 * acts.size(), 0.0);
    for (MCTS_node *c : *root->children)
        for (size_t k = 0; k < acts.size(); k++)
            if (*acts[k] == *c->move) { out[2 * k] = (double)c->number_of_simulations; out[2 * k + 1] = c->score; break; }
    for (MCTS_move *m : acts) delete m;
}

void MCTS_tree::import_root_stats(const vector<double> &in) {
    vector<MCTS_move *> acts = root_action_list(root->state);
    if (in.size() != 2 * acts.size()) { for (MCTS_move *m : acts) delete m; throw std::invalid_argument("root stats size mismatch"); }
    // children are created in action order; make sure every slot that carries statistics has its child
    size_t last = 0;
    for (size_t k = 0; k < acts.size(); k++) if (in[2 * k] > 0) last = k + 1;
    while (root->children->size() < last && root->expand_no_rollout() != NULL) {}
    double sims = 0, score = 0;
    for (MCTS_node *c : *root->children)
        for (size_t k = 0; k < acts.size(); k++)
            if (*acts[k] == *c->move) {
                c->number_of_simulations = (uint64_t)in[2 * k]; c->score = in[2 * k + 1];
                sims += in[2 * k]; score += in[2 * k + 1];
                break;
            }
    root->number_of_simulations = (uint64_t)sims;
    root->score = score;
    root->rebuild_child_stats();
    for (MCTS_move *m : acts) delete m;
}


void MCTS_tree::merge_root_stats() {
    if (get_backend()->world() == 1) return;
    vector<double> v;
    export_root_stats(v);
    if (v.empty()) return;
    get_backend()->allreduce_sum(v.data(), (int)v.size());
    import_root_stats(v);
}

/*** MCTS_agent ***/

MCTS_agent::MCTS_agent(MCTS_state *starting_state, int max_iter, int max_seconds)
    : tree(new MCTS_tree(starting_state)), max_iter(max_iter), max_seconds(max_seconds) {}

MCTS_agent::MCTS_agent(MCTS_state *starting_state, int max_iter, int max_seconds, RolloutBackend *backend, MCTS_batching batching)
    : tree(new MCTS_tree(starting_state, backend, batching)), max_iter(max_iter), max_seconds(max_seconds) {}

MCTS_agent::~MCTS_agent() { delete tree; }

// reference mcts.cpp:281-306
const MCTS_move *MCTS_agent::genmove(const MCTS_move *enemy_move) {
    if (enemy_move != NULL) tree->advance_tree(enemy_move);
    if (tree->get_current_state()->is_terminal()) return NULL;
    if (g_verbose) cout << "___ DEBUG ______________________" << endl << "Growing tree..." << endl;
    tree->grow_tree(max_iter, max_seconds);
    if (g_verbose) cout << "Tree size: " << tree->get_size() << endl << "________________________________" << endl;
    MCTS_node *best_child = tree->select_best_child();
    if (best_child == NULL) {
        cerr << "Warning: Tree root has no children! Possibly terminal node!" << endl;
        return NULL;
    }
    const MCTS_move *best_move = best_child->get_move();
    tree->advance_tree(best_move);
    return best_move;
}

const MCTS_state *MCTS_agent::get_current_state() const { return tree->get_current_state(); }
