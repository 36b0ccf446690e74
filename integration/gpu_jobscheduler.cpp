// gpu_jobscheduler.cpp — link-time drop-in for the reference's mcts/src/JobScheduler.cpp.
//
// The reference runs its Simulation phase by scheduling NUMBER_OF_THREADS RolloutJobs on a function-local
// static JobScheduler and waiting for them (mcts/src/mcts.cpp:57-81, mcts/include/JobScheduler.h:42-63).
// This file implements that SAME class — the declaration is the reference's own header — with a body that
// collects the jobs and, on waitUntilJobsHaveFinished(), plays them on the B200 through the C ABI
// (include/mctsb.h -> libmctsb.so).  Link it INSTEAD of JobScheduler.cpp, next to the reference's unmodified
// mcts.cpp / Quoridor.cpp / TicTacToe.cpp, and MCTS_tree::grow_tree, MCTS_agent::genmove and the example
// mains run their rollouts on the GPU without one edited line:
//
//   g++ -O2 -std=c++11 -DMCTSB_DROPIN_ONLY_QUORIDOR -I$REF -Iinclude -include integration/open_classes.h -c integration/gpu_jobscheduler.cpp
//   g++ -O2 -std=c++11 -I$REF $REF/mcts/src/mcts.cpp $REF/examples/Quoridor/{Quoridor,main}.cpp \
//       gpu_jobscheduler.o -Lmontecarlotreesearch_b200 -lmctsb -o quoridor_gpu
//
// Jobs whose state is a Quoridor_state or TicTacToe_state become one mctsb_rollout_outcomes call per group of
// jobs that share a state (result i -> job i's score slot, exactly what RolloutJob::run would have stored);
// any other Job (a user's own game, or a non-rollout job) simply run()s on the calling thread — the library
// has no CPU implementation of the two games it knows.  Without a CUDA device the first rollout aborts.
//
// Environment: MCTSB_DEVICE (default 0), MCTSB_SEED (Philox key, default 0x2026101900000001).
// A program that links only one of the two games (the reference Makefile builds `quoridor` and `tictactoe`
// separately) compiles this file with -DMCTSB_DROPIN_ONLY_QUORIDOR or -DMCTSB_DROPIN_ONLY_TICTACTOE.
#include "mctsb.h"
#include "mcts/include/JobScheduler.h"
#include "mcts/include/mcts.h"
#if !defined(MCTSB_DROPIN_ONLY_TICTACTOE)
#define DROPIN_QUORIDOR 1
#include "examples/Quoridor/Quoridor.h"
#endif
#if !defined(MCTSB_DROPIN_ONLY_QUORIDOR)
#define DROPIN_TICTACTOE 1
#include "examples/TicTacToe/TicTacToe.h"
#endif

namespace {

mctsb_backend *g_backend = NULL;
uint64_t g_next_id = 0;                  // Philox stream ids handed out so far: every rollout gets its own
uint64_t g_flushes = 0, g_rollouts = 0;

mctsb_backend *backend() {
    if (g_backend) return g_backend;
    const char *dev = getenv("MCTSB_DEVICE"), *seed = getenv("MCTSB_SEED");
    int st = mctsb_create(&g_backend, dev ? atoi(dev) : 0, seed ? strtoull(seed, NULL, 0) : 0x2026101900000001ull);
    if (st != MCTSB_OK) {
        fprintf(stderr, "gpu_jobscheduler: %s — the rollout backend has no CPU fallback\n", mctsb_strerror(st));
        abort();
    }
    return g_backend;
}

#if defined(DROPIN_QUORIDOR)
// walls[][] + wall_connections[][] -> anchor masks.  Each connection point carries one wall, horizontal
// (segments 'h' at (x,y),(x,y+1)) or vertical ('v' at (x,y),(x+1,y)); each segment belongs to exactly one
// wall.  The arrays do not name the orientation, so find an assignment that uses every segment once
// (depth-first, <= 20 walls).  Any consistent assignment describes the same position: the rules only ever
// read these two arrays (Quoridor.cpp:200-329).
struct WallDecoder {
    const Quoridor_state &q;
    bool h_used[9][9], v_used[9][9];
    int px[64], py[64], n;
    uint64_t h_anchor, v_anchor;
    explicit WallDecoder(const Quoridor_state &q) : q(q), n(0), h_anchor(0), v_anchor(0) {
        memset(h_used, 0, sizeof h_used); memset(v_used, 0, sizeof v_used);
        for (int x = 0; x < 8; x++) for (int y = 0; y < 8; y++) if (q.wall_connections[x][y]) { px[n] = x; py[n] = y; n++; }
    }
    bool all_segments_used() const {
        for (int x = 0; x < 9; x++) for (int y = 0; y < 9; y++)
            if ((q.horizontal_wall(x, y) && !h_used[x][y]) || (q.vertical_wall(x, y) && !v_used[x][y])) return false;
        return true;
    }
    bool solve(int i) {
        if (i == n) return all_segments_used();
        const int x = px[i], y = py[i];
        if (q.horizontal_wall(x, y) && q.horizontal_wall(x, y + 1) && !h_used[x][y] && !h_used[x][y + 1]) {
            h_used[x][y] = h_used[x][y + 1] = true; h_anchor |= 1ull << (8 * x + y);
            if (solve(i + 1)) return true;
            h_used[x][y] = h_used[x][y + 1] = false; h_anchor &= ~(1ull << (8 * x + y));
        }
        if (q.vertical_wall(x, y) && q.vertical_wall(x + 1, y) && !v_used[x][y] && !v_used[x + 1][y]) {
            v_used[x][y] = v_used[x + 1][y] = true; v_anchor |= 1ull << (8 * x + y);
            if (solve(i + 1)) return true;
            v_used[x][y] = v_used[x + 1][y] = false; v_anchor &= ~(1ull << (8 * x + y));
        }
        return false;
    }
};

bool pack(const Quoridor_state &q, mctsb_qstate &p) {
    WallDecoder d(q);
    if (!d.solve(0)) return false;
    memset(&p, 0, sizeof p);
    p.h_anchor = d.h_anchor; p.v_anchor = d.v_anchor;
    p.wx = (uint8_t)q.wx; p.wy = (uint8_t)q.wy; p.bx = (uint8_t)q.bx; p.by = (uint8_t)q.by;
    p.wwalls = (uint8_t)q.wwallsno; p.bwalls = (uint8_t)q.bwallsno;
    p.turn = q.turn == 'B'; p.move_counter = q.move_counter;
    return true;
}

#endif  // DROPIN_QUORIDOR

#if defined(DROPIN_TICTACTOE)
void pack(const TicTacToe_state &t, mctsb_tstate &p) {
    p.x_mask = p.o_mask = 0;
    for (int c = 0; c < 9; c++) {
        if (t.board[c / 3][c % 3] == 'x') p.x_mask |= (uint16_t)(1u << c);
        if (t.board[c / 3][c % 3] == 'o') p.o_mask |= (uint16_t)(1u << c);
    }
    if (t.turn == 'o') p.o_mask |= 0x8000u;
}
#endif

// jobs[a..b) are RolloutJobs of ONE state: play them in one launch, store result i where job i asked for it
void play_group(std::vector<Job *> &jobs, size_t a, size_t b) {
    const MCTS_state *state = static_cast<RolloutJob *>(jobs[a])->state;
    const uint32_t R = (uint32_t)(b - a);
    int game = -1; unsigned char packed[32];
#if defined(DROPIN_QUORIDOR)
    if (const Quoridor_state *q = dynamic_cast<const Quoridor_state *>(state)) {
        game = MCTSB_GAME_QUORIDOR;
        if (!pack(*q, *reinterpret_cast<mctsb_qstate *>(packed))) { fprintf(stderr, "gpu_jobscheduler: inconsistent wall arrays\n"); abort(); }
    }
#endif
#if defined(DROPIN_TICTACTOE)
    if (game < 0) if (const TicTacToe_state *t = dynamic_cast<const TicTacToe_state *>(state)) {
        game = MCTSB_GAME_TICTACTOE;
        pack(*t, *reinterpret_cast<mctsb_tstate *>(packed));
    }
#endif
    if (game < 0) {                            // not one of the two games the kernels play: the job's own run()
        for (size_t i = a; i < b; i++) jobs[i]->run();
        return;
    }
    std::vector<mctsb_outcome> out(R);
    int st = mctsb_rollout_outcomes(backend(), game, MCTSB_POLICY_REF, packed, 1, R, g_next_id, out.data(), NULL);
    if (st != MCTSB_OK) {
        fprintf(stderr, "gpu_jobscheduler: %s %s\n", mctsb_strerror(st), st == MCTSB_ERR_CUDA ? mctsb_last_cuda_error(g_backend) : "");
        abort();
    }
    g_next_id += R; g_flushes++; g_rollouts += R;
    for (size_t i = a; i < b; i++) *static_cast<RolloutJob *>(jobs[i])->score = out[i - a].score;
}

}  // namespace

// ---- the reference's class, re-implemented (JobScheduler.cpp:10-34: the pool threads are simply not created) ----
JobScheduler::JobScheduler(unsigned int _number_of_threads)
    : threads(NULL), number_of_threads(_number_of_threads), threads_must_exit(false), t_args(NULL), jobs_running(0) {}

JobScheduler::~JobScheduler() { waitUntilJobsHaveFinished(); }   // "Waits until all jobs have finished!" (JobScheduler.h:59)

void JobScheduler::schedule(Job *job) { job_queue.push(job); }   // JobScheduler.cpp:36-46: the scheduler owns the job from here

bool JobScheduler::JobsHaveFinished(int) { return job_queue.empty(); }

// JobScheduler.cpp:67-73 — here the wait IS the work: flush everything queued so far
void JobScheduler::waitUntilJobsHaveFinished(int) {
    std::vector<Job *> jobs;
    while (!job_queue.empty()) { jobs.push_back(job_queue.front()); job_queue.pop(); }
    size_t a = 0;
    while (a < jobs.size()) {
        RolloutJob *first = dynamic_cast<RolloutJob *>(jobs[a]);
        if (!first) { jobs[a]->run(); a++; continue; }
        size_t b = a + 1;
        while (b < jobs.size()) {
            RolloutJob *next = dynamic_cast<RolloutJob *>(jobs[b]);
            if (!next || next->state != first->state) break;
            b++;
        }
        play_group(jobs, a, b);
        a = b;
    }
    for (Job *j : jobs) delete j;                                   // JobScheduler.cpp:117: jobs are deleted after run()
}

void *thread_code(void *) { return NULL; }                          // declared in JobScheduler.h:39; no pool threads here

// ---- test / bench hooks (not part of the reference interface) ---------------------------------------------
extern "C" {
void mctsb_dropin_reset(uint64_t first_rollout_id) { g_next_id = first_rollout_id; g_flushes = g_rollouts = 0; }
void mctsb_dropin_counters(uint64_t *flushes, uint64_t *rollouts) { *flushes = g_flushes; *rollouts = g_rollouts; }
int mctsb_dropin_device_ok(void) { int n = 0; return mctsb_device_count(&n) == MCTSB_OK && n > 0; }
// the packer alone (CPU): `reference_state` is a `const Quoridor_state *`; 1 = packed
#if defined(DROPIN_QUORIDOR)
int mctsb_dropin_pack_quoridor(const void *reference_state, mctsb_qstate *out) {
    return pack(*static_cast<const Quoridor_state *>(reference_state), *out) ? 1 : 0;
}
#endif
}
