/*
 * harness.cpp — TEST INFRASTRUCTURE.  extern "C" driver around the UNMODIFIED reference
 * sources (compiled where they lie under /root/reference by oracle/Makefile, output only into
 * oracle/_ref/).  Built twice:
 *   libref_replay.so : with oracle/shim.h's stream-driven RNG  -> bit-exact replay + distributions
 *   libref_native.so : with -DORACLE_SHIM_NATIVE_RNG (stock RNG) -> the CPU baseline that is timed
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load these libraries.  Nothing here is on the product path.
 */
#include "../include/mctsb.h"
#include "examples/Quoridor/Quoridor.h"
#include "examples/TicTacToe/TicTacToe.h"
#include "mcts/include/mcts.h"
#include <chrono>
#include <atomic>
#include <thread>
#include <random>

/* friend functions of Quoridor_state (Quoridor.h:78-80) re-declared at namespace scope */
bool force_playwall(Quoridor_state &s);
Quoridor_move *pick_semirandom_move(Quoridor_state &s, std::uniform_real_distribution<double> &dist,
                                    std::default_random_engine &gen);
double evaluate_position(Quoridor_state &s, bool cheap);

namespace refshim {
static std::atomic<uint64_t> g_thread_ids{1ull << 40};
thread_local DrawSource g_src = {NULL, 0, 0x2026101900000001ull, g_thread_ids.fetch_add(1), 0};
}

namespace {

struct Silence {   /* the reference chats on cout/cerr (mcts.cpp:185-209, Quoridor.cpp:346,492,833) */
    std::streambuf *o, *e;
    Silence() : o(std::cout.rdbuf(NULL)), e(std::cerr.rdbuf(NULL)) {}
    ~Silence() { std::cout.rdbuf(o); std::cerr.rdbuf(e); }
};

void load_q(Quoridor_state &q, const mctsb_qstate *s) {
    q.wx = s->wx; q.wy = s->wy; q.bx = s->bx; q.by = s->by;
    q.wwallsno = s->wwalls; q.bwallsno = s->bwalls;
    q.turn = s->turn ? 'B' : 'W';
    q.move_counter = s->move_counter;
    for (int i = 0; i < 81; i++) q.walls[i / 9][i % 9] = ' ';
    for (int i = 0; i < 64; i++) q.wall_connections[i / 8][i % 8] = false;
    for (int a = 0; a < 64; a++) {
        int x = a / 8, y = a % 8;
        if ((s->h_anchor >> a) & 1) { q.add_wall(x, y, true); q.add_wall(x, y + 1, true); q.wall_connections[x][y] = true; }
    }
    for (int a = 0; a < 64; a++) {
        int x = a / 8, y = a % 8;
        if ((s->v_anchor >> a) & 1) { q.add_wall(x, y, false); q.add_wall(x + 1, y, false); q.wall_connections[x][y] = true; }
    }
}

void raw_q(const Quoridor_state &q, int32_t *out) {
    for (int i = 0; i < 81; i++) out[i] = q.walls[i / 9][i % 9];
    for (int i = 0; i < 64; i++) out[81 + i] = q.wall_connections[i / 8][i % 8];
    out[145] = q.wx; out[146] = q.wy; out[147] = q.bx; out[148] = q.by;
    out[149] = q.wwallsno; out[150] = q.bwallsno; out[151] = q.turn; out[152] = (int32_t)q.move_counter;
}

int move_id(const Quoridor_move *m) {
    if (m->type == 'h') return MCTSB_Q_MOVE_HWALL(m->x, m->y);
    if (m->type == 'v') return MCTSB_Q_MOVE_VWALL(m->x, m->y);
    return MCTSB_Q_MOVE_STEP(m->x, m->y);
}

Quoridor_move make_move(int id, char p) {
    if (id < 81) return Quoridor_move(id / 9, id % 9, p, ' ');
    if (id < 145) return Quoridor_move((id - 81) / 8, (id - 81) % 8, p, 'h');
    return Quoridor_move((id - 145) / 8, (id - 145) % 8, p, 'v');
}

void load_t(TicTacToe_state &t, uint32_t x_mask, uint32_t o_mask) {
    for (int c = 0; c < 9; c++)
        t.board[c / 3][c % 3] = ((x_mask >> c) & 1) ? 'x' : ((o_mask >> c) & 1) ? 'o' : ' ';
    t.turn = (o_mask & 0x8000) ? 'o' : 'x';
    t.winner = t.calculate_winner();
}

}  // namespace

extern "C" {

int ref_has_stream_rng(void) {
#ifdef ORACLE_SHIM_NATIVE_RNG
    return 0;
#else
    return 1;
#endif
}

int ref_sizeof_qstate(void) { return (int)sizeof(Quoridor_state); }

/* ---------------- Quoridor primitives ---------------- */

int ref_q_raw(const mctsb_qstate *s, int32_t *out153) {
    Silence q_;
    Quoridor_state q; load_q(q, s); raw_q(q, out153);
    return 0;
}

/* info[0]=winner(0 none,1 W,2 B) [1]=sp W [2]=sp B [3]=force_playwall [4]=is_terminal [5]=player1_turn */
int ref_q_primitives(const mctsb_qstate *s, int32_t *info, uint8_t *legal209, int32_t *actions,
                     int32_t *n_actions, int32_t *steps2, int32_t *n_steps2, double *eval,
                     uint8_t *cheap128) {
    Silence q_;
    Quoridor_state q; load_q(q, s);
    char p = q.turn;
    char w = q.check_winner();
    info[0] = (w == 'W') ? 1 : (w == 'B') ? 2 : 0;
    info[1] = q.get_shortest_path('W');
    info[2] = q.get_shortest_path('B');
    info[3] = force_playwall(q) ? 1 : 0;
    info[4] = q.is_terminal() ? 1 : 0;
    info[5] = q.player1_turn() ? 1 : 0;
    if (eval) *eval = evaluate_position(q, false);
    if (legal209) {
        for (int id = 0; id < MCTSB_Q_NUM_MOVES; id++) {
            Quoridor_move m = make_move(id, p);
            legal209[id] = q.legal_move(&m) ? 1 : 0;
        }
    }
    if (actions) {
        queue<MCTS_move *> *Q = q.actions_to_try();
        int n = 0;
        while (!Q->empty()) { actions[n++] = move_id((Quoridor_move *)Q->front()); delete Q->front(); Q->pop(); }
        delete Q;
        *n_actions = n;
    }
    if (steps2) {
        vector<MCTS_move *> v = q.get_legal_step_moves2(p);
        *n_steps2 = (int)v.size();
        for (size_t i = 0; i < v.size(); i++) { steps2[i] = move_id((Quoridor_move *)v[i]); delete v[i]; }
    }
    if (cheap128) {
        for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) for (int k = 0; k < 2; k++)
            cheap128[2 * (8 * i + j) + k] = q.legal_wall(i, j, p, k == 0, false) ? 1 : 0;
    }
    return 0;
}

/* get_shortest_path(player, wall move_id by the side to move) — Quoridor.cpp:108-174 */
int ref_q_sp_with_wall(const mctsb_qstate *s, int player_is_black, int wall_move_id) {
    Silence q_;
    Quoridor_state q; load_q(q, s);
    Quoridor_move m = make_move(wall_move_id, q.turn);
    return q.get_shortest_path(player_is_black ? 'B' : 'W', &m);
}

int ref_q_sp_from(const mctsb_qstate *s, int player_is_black, int x, int y) {
    Silence q_;
    Quoridor_state q; load_q(q, s);
    return q.get_shortest_path(player_is_black ? 'B' : 'W', NULL, x, y);
}

/* play_move through the reference; raw153 (nullable) gets the resulting reference state */
int ref_q_play(const mctsb_qstate *s, int id, int32_t *raw153) {
    Silence q_;
    Quoridor_state q; load_q(q, s);
    Quoridor_move m = make_move(id, q.turn);
    bool ok = q.play_move(&m);
    if (raw153) raw_q(q, raw153);
    return ok ? 1 : 0;
}

/* get_best_step_move (Quoridor.cpp:476-497): canonical id or -1 */
int ref_q_best_step(const mctsb_qstate *s) {
    Silence q_;
    Quoridor_state q; load_q(q, s);
    Quoridor_move *m = q.get_best_step_move(q.turn);
    if (!m) return -1;
    int id = move_id(m); delete m; return id;
}

/* the reference's public generate_good_moves() (Quoridor.cpp:518-592): canonical ids in queue order */
extern "C" int ref_q_good_moves(const mctsb_qstate *st, int32_t *ids) {
    Silence q_;
    Quoridor_state s; load_q(s, st);
    queue<MCTS_move *> *Q = s.generate_good_moves();
    int n = 0;
    while (!Q->empty()) { Quoridor_move *m = (Quoridor_move *)Q->front(); Q->pop(); ids[n++] = move_id(m); delete m; }
    delete Q;
    return n;
}

/* ---------------- Quoridor rollouts (stream RNG build only) ---------------- */
#ifndef ORACLE_SHIM_NATIVE_RNG

static void set_stream(const uint32_t *stream, uint64_t len) {
    refshim::g_src.data = stream; refshim::g_src.len = len; refshim::g_src.pos = 0;
}
static void set_philox(uint64_t seed, uint64_t id) {
    refshim::g_src.data = NULL; refshim::g_src.len = 0; refshim::g_src.seed = seed;
    refshim::g_src.rollout_id = id; refshim::g_src.pos = 0;
}

/* the reference's own rollout() (Quoridor.cpp:809-855), untouched */
static double q_rollout_plain(const mctsb_qstate *s, uint64_t *draws) {
    Quoridor_state q; load_q(q, s);
    double r = q.rollout();
    if (draws) *draws = refshim::g_src.pos;
    return r;
}

/* The same loop driven from outside through the reference's friend functions, so that the move
 * sequence and the final position become observable (rollout() keeps its working copy private,
 * Quoridor.cpp:815).  Must return exactly what rollout() returns on the same stream; the
 * callers below check that on every call.  kind: 0 W won, 1 B won, 2 eval after 50 plies,
 * 3 eval after a failed pick. */
static double q_rollout_traced(const mctsb_qstate *st, int32_t *raw_final, int32_t *moves, int32_t *n_moves,
                               int32_t *kind, uint64_t *draws) {
    std::uniform_real_distribution<double> dist(0.0, 1.0);
    Quoridor_state s; load_q(s, st);
    int nm = 0; int k = 2; double result = -1.0; bool have = false;
    for (int i = 0; i < 50; i++) {
        if (s.is_terminal()) { result = (s.check_winner() == 'W') ? 1.0 : 0.0; k = (result == 1.0) ? 0 : 1; have = true; break; }
        double eval = evaluate_position(s, true);
        if (eval <= 1.0 - 0.8 && eval >= 0.8) break;
        Quoridor_move *m = pick_semirandom_move(s, dist, Quoridor_state::generator);
        bool ok = s.play_move(m);
        if (!ok) { k = 3; delete m; break; }
        if (moves) moves[nm] = move_id(m);
        nm++;
        delete m;
    }
    if (!have) result = evaluate_position(s, false);
    if (raw_final) raw_q(s, raw_final);
    if (n_moves) *n_moves = nm;
    if (kind) *kind = k;
    if (draws) *draws = refshim::g_src.pos;
    return result;
}

/* returns score; *mismatch set to 1 if traced loop and plain rollout() disagree (never expected) */
double ref_q_rollout_stream(const mctsb_qstate *s, const uint32_t *stream, uint64_t len, uint64_t *draws,
                            int32_t *raw_final, int32_t *moves, int32_t *n_moves, int32_t *kind, int32_t *mismatch) {
    Silence q_;
    set_stream(stream, len);
    uint64_t d0 = 0, d1 = 0;
    double plain = q_rollout_plain(s, &d0);
    set_stream(stream, len);
    double traced = q_rollout_traced(s, raw_final, moves, n_moves, kind, &d1);
    if (mismatch) *mismatch = (plain != traced || d0 != d1) ? 1 : 0;
    if (draws) *draws = d0;
    return plain;
}

int ref_q_rollout_philox(const mctsb_qstate *s, uint64_t seed, uint64_t first_id, int n, double *scores,
                         uint32_t *draws, int32_t *raw_finals, int32_t *kinds, int32_t *plies, int traced) {
    Silence q_;
    int bad = 0;
    for (int i = 0; i < n; i++) {
        uint64_t d0 = 0;
        set_philox(seed, first_id + i);
        double plain = q_rollout_plain(s, &d0);
        if (traced) {
            uint64_t d1 = 0; int32_t k = 0, nm = 0;
            set_philox(seed, first_id + i);
            double t = q_rollout_traced(s, raw_finals ? raw_finals + 153 * (size_t)i : NULL, NULL, &nm, &k, &d1);
            if (t != plain || d0 != d1) bad++;
            if (kinds) kinds[i] = k;
            if (plies) plies[i] = nm;
        }
        scores[i] = plain;
        if (draws) draws[i] = (uint32_t)d0;
    }
    return bad;
}

/* UNI policy oracle: uniformly random legal move over the reference's generate_all_moves order
 * (Quoridor.cpp:594-616) via the reference's public API; idx = mulhi32(draw, |L|). */
static double q_uniform_one(const mctsb_qstate *st, int max_plies, int32_t *raw_final, int32_t *plies, int32_t *kind) {
    Quoridor_state s; load_q(s, st);
    int n = 0; int k = 4; double result = 0.5;
    while (true) {
        if (s.is_terminal()) { bool w = s.check_winner() == 'W'; result = w ? 1.0 : 0.0; k = w ? 0 : 1; break; }
        if (n >= max_plies) break;
        queue<MCTS_move *> *Q = s.actions_to_try();
        uint32_t L = (uint32_t)Q->size();
        if (L == 0) { delete Q; k = 3; break; }
        uint32_t idx = mulhi32(refshim::next_u32(), L);
        Quoridor_move *chosen = NULL;
        for (uint32_t i = 0; !Q->empty(); i++) {
            Quoridor_move *m = (Quoridor_move *)Q->front(); Q->pop();
            if (i == idx) chosen = m; else delete m;
        }
        delete Q;
        s.play_move(chosen);
        delete chosen;
        n++;
    }
    if (raw_final) raw_q(s, raw_final);
    if (plies) *plies = n;
    if (kind) *kind = k;
    return result;
}

double ref_q_uniform_stream(const mctsb_qstate *s, const uint32_t *stream, uint64_t len, int max_plies,
                            uint64_t *draws, int32_t *raw_final, int32_t *plies, int32_t *kind) {
    Silence q_;
    set_stream(stream, len);
    double r = q_uniform_one(s, max_plies, raw_final, plies, kind);
    if (draws) *draws = refshim::g_src.pos;
    return r;
}

int ref_q_uniform_philox(const mctsb_qstate *s, uint64_t seed, uint64_t first_id, int n, int max_plies,
                         double *scores, uint32_t *draws, int32_t *raw_finals, int32_t *kinds, int32_t *plies) {
    Silence q_;
    for (int i = 0; i < n; i++) {
        set_philox(seed, first_id + i);
        int32_t k = 0, nm = 0;
        scores[i] = q_uniform_one(s, max_plies, raw_finals ? raw_finals + 153 * (size_t)i : NULL, &nm, &k);
        if (draws) draws[i] = (uint32_t)refshim::g_src.pos;
        if (kinds) kinds[i] = k;
        if (plies) plies[i] = nm;
    }
    return 0;
}

/* ---------------- Tic-Tac-Toe (stream RNG build) ---------------- */

double ref_t_rollout_stream(uint32_t x_mask, uint32_t o_mask, const uint32_t *stream, uint64_t len, uint64_t *draws) {
    Silence q_;
    set_stream(stream, len);
    TicTacToe_state t; load_t(t, x_mask, o_mask);
    double r = t.rollout();
    if (draws) *draws = refshim::g_src.pos;
    return r;
}

int ref_t_rollout_philox(uint32_t x_mask, uint32_t o_mask, uint64_t seed, uint64_t first_id, int n,
                         double *scores, uint32_t *draws) {
    Silence q_;
    TicTacToe_state t; load_t(t, x_mask, o_mask);
    for (int i = 0; i < n; i++) {
        set_philox(seed, first_id + i);
        scores[i] = t.rollout();
        if (draws) draws[i] = (uint32_t)refshim::g_src.pos;
    }
    return 0;
}
#endif /* stream RNG build */

/* info[0]=winner(0 none,1 x,2 o,3 draw) [1]=is_terminal [2]=player1_turn; actions = cells in actions_to_try order */
int ref_t_primitives(uint32_t x_mask, uint32_t o_mask, int32_t *info, int32_t *actions, int32_t *n_actions) {
    Silence q_;
    TicTacToe_state t; load_t(t, x_mask, o_mask);
    char w = t.get_winner();
    info[0] = (w == 'x') ? 1 : (w == 'o') ? 2 : (w == 'd') ? 3 : 0;
    info[1] = t.is_terminal() ? 1 : 0;
    info[2] = t.player1_turn() ? 1 : 0;
    queue<MCTS_move *> *Q = t.actions_to_try();
    int n = 0;
    while (!Q->empty()) { TicTacToe_move *m = (TicTacToe_move *)Q->front(); actions[n++] = m->x * 3 + m->y; delete m; Q->pop(); }
    delete Q;
    *n_actions = n;
    return 0;
}

/* ---------------- the reference's host tree, stock (mcts.cpp) ---------------- */

/* grow_tree on the stock MCTS_tree (4 pool threads, mcts.cpp:57-81,182-210); fills root stats:
 * out[0]=root sims, out[1]=root score, out[2]=tree size, out[3]=#children; child_stats = n_children x
 * {move id, sims, score}.  game 0 = Quoridor from `qs`, 1 = Tic-Tac-Toe from masks. */
double ref_grow_tree(int game, const mctsb_qstate *qs, uint32_t x_mask, uint32_t o_mask, int iters,
                     double *out4, double *child_stats, int max_children) {
    Silence q_;
    MCTS_state *st;
    if (game == 0) { Quoridor_state *q = new Quoridor_state(); load_q(*q, qs); st = q; }
    else { TicTacToe_state *t = new TicTacToe_state(); load_t(*t, x_mask, o_mask); st = t; }
    MCTS_tree tree(st);
    auto t0 = std::chrono::steady_clock::now();
    tree.grow_tree(iters, 1e9);
    auto t1 = std::chrono::steady_clock::now();
    MCTS_node *root = tree.root;
    out4[0] = root->number_of_simulations; out4[1] = root->score; out4[2] = root->size;
    out4[3] = (double)root->children->size();
    int n = 0;
    for (MCTS_node *c : *root->children) {
        if (n >= max_children) break;
        int id = (game == 0) ? move_id((const Quoridor_move *)c->move)
                             : ((const TicTacToe_move *)c->move)->x * 3 + ((const TicTacToe_move *)c->move)->y;
        child_stats[3 * n] = id; child_stats[3 * n + 1] = c->number_of_simulations; child_stats[3 * n + 2] = c->score;
        n++;
    }
    return std::chrono::duration<double>(t1 - t0).count();
}

#ifdef ORACLE_WITH_DROPIN
/* drop-in build only: a reference Quoridor_state through integration/gpu_jobscheduler.cpp's packer */
extern "C" int mctsb_dropin_pack_quoridor(const void *reference_state, mctsb_qstate *out);
int ref_q_pack_roundtrip(const mctsb_qstate *in, mctsb_qstate *out) {
    Quoridor_state q; load_q(q, in);
    return mctsb_dropin_pack_quoridor(&q, out);
}
#endif

/* ---------------- CPU baseline: RolloutJobs on the reference JobScheduler ---------------- */

/* JobScheduler sched(n_threads); n_jobs x schedule(new RolloutJob(&state,&res[i])); wait
 * (mcts.cpp:60-66, JobScheduler.cpp:36-46,67-73).  Returns wall seconds; *mean_score = mean result. */
double ref_jobscheduler_bench2(int game, const mctsb_qstate *qs, int n_states, uint32_t x_mask, uint32_t o_mask,
                               int n_threads, int n_jobs, double *mean_score, double *mean_square);
double ref_jobscheduler_bench(int game, const mctsb_qstate *qs, int n_states, uint32_t x_mask, uint32_t o_mask,
                              int n_threads, int n_jobs, double *mean_score) {
    return ref_jobscheduler_bench2(game, qs, n_states, x_mask, o_mask, n_threads, n_jobs, mean_score, NULL);
}

/* the same, also returning the mean of the squared results (for a standard error of the mean) */
double ref_jobscheduler_bench2(int game, const mctsb_qstate *qs, int n_states, uint32_t x_mask, uint32_t o_mask,
                               int n_threads, int n_jobs, double *mean_score, double *mean_square) {
    Silence q_;
    std::vector<MCTS_state *> states;
    if (game == 0) {
        for (int i = 0; i < n_states; i++) { Quoridor_state *q = new Quoridor_state(); load_q(*q, qs + i); states.push_back(q); }
    } else {
        TicTacToe_state *t = new TicTacToe_state(); load_t(*t, x_mask, o_mask); states.push_back(t);
    }
    std::vector<double> res((size_t)n_jobs, -1.0);
    double secs;
    {
        JobScheduler sched((unsigned)n_threads);
        auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < n_jobs; i++) sched.schedule(new RolloutJob(states[(size_t)i % states.size()], &res[(size_t)i]));
        sched.waitUntilJobsHaveFinished();
        auto t1 = std::chrono::steady_clock::now();
        secs = std::chrono::duration<double>(t1 - t0).count();
    }
    double sum = 0.0, sq = 0.0;
    for (int i = 0; i < n_jobs; i++) { sum += res[(size_t)i]; sq += res[(size_t)i] * res[(size_t)i]; }
    if (mean_score) *mean_score = sum / n_jobs;
    if (mean_square) *mean_square = sq / n_jobs;
    for (MCTS_state *s : states) delete s;
    return secs;
}

/* CPU baseline of the UNI policy: uniformly random legal playouts driven through the reference's PUBLIC API
 * (actions_to_try -> generate_all_moves, play_move; Quoridor.cpp:594-616,344-392) on n_threads std::threads, each
 * with its own std::mt19937_64.  The reference has no such function — this is the loop the UNI oracle restates —
 * so it is timed as "the reference's rules doing the UNI work".  Returns wall seconds. */
double ref_uniform_bench(const mctsb_qstate *qs, int n_states, int n_threads, int n_playouts, int max_plies,
                         double *mean_score, double *mean_plies) {
    Silence q_;
    std::vector<double> score((size_t)n_playouts, 0.0);
    std::vector<int> plies((size_t)n_playouts, 0);
    std::atomic<int> next(0);
    auto worker = [&](int tid) {
        std::mt19937_64 gen(0x9E3779B97F4A7C15ull * (uint64_t)(tid + 1));
        for (;;) {
            int i = next.fetch_add(1);
            if (i >= n_playouts) break;
            Quoridor_state s; load_q(s, qs + (i % n_states));
            int n = 0; double result = 0.5;
            while (true) {
                if (s.is_terminal()) { result = s.check_winner() == 'W' ? 1.0 : 0.0; break; }
                if (n >= max_plies) break;
                queue<MCTS_move *> *Q = s.actions_to_try();
                size_t L = Q->size();
                if (L == 0) { delete Q; break; }
                size_t idx = (size_t)(gen() % L);
                Quoridor_move *chosen = NULL;
                for (size_t k = 0; !Q->empty(); k++) {
                    Quoridor_move *m = (Quoridor_move *)Q->front(); Q->pop();
                    if (k == idx) chosen = m; else delete m;
                }
                delete Q;
                s.play_move(chosen);
                delete chosen;
                n++;
            }
            score[(size_t)i] = result; plies[(size_t)i] = n;
        }
    };
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++) th.emplace_back(worker, t);
    for (auto &t : th) t.join();
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    double ss = 0, sp = 0;
    for (int i = 0; i < n_playouts; i++) { ss += score[(size_t)i]; sp += plies[(size_t)i]; }
    if (mean_score) *mean_score = ss / n_playouts;
    if (mean_plies) *mean_plies = sp / n_playouts;
    return secs;
}

}  // extern "C"
