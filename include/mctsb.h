/*
 * mctsb.h — C ABI of the B200 rollout backend (the MCTS *Simulation* phase).
 *
 * This is the drop-in boundary.  In the reference there is no FFI; the boundary is the pure
 * virtual `double MCTS_state::rollout() const` (reference mcts/include/state.h:29) and its
 * single call site `MCTS_node::rollout()` (reference mcts/src/mcts.cpp:57-81), which fans
 * `NUMBER_OF_THREADS` `RolloutJob`s (reference mcts/include/mcts.h:82-91) out to the pthread
 * `JobScheduler` (reference mcts/include/JobScheduler.h:42-63) and sums the doubles.  Every
 * entry point below names the reference code it replaces.
 *
 * Rules of the boundary: plain C types, caller owns every host buffer, the backend owns all
 * device memory and streams, no exceptions, nothing is printed, every function returns an
 * `mctsb_status` (0 = ok).  There is NO CPU fallback: without a CUDA device every compute
 * entry point returns MCTSB_ERR_NO_DEVICE.
 */
#ifndef MCTSB_H
#define MCTSB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCTSB_ABI_VERSION 1

typedef enum {
    MCTSB_OK = 0,
    MCTSB_ERR_NO_DEVICE = 1,   /* no CUDA device / driver: the product path refuses to run        */
    MCTSB_ERR_CUDA = 2,        /* a CUDA runtime call failed, see mctsb_last_cuda_error           */
    MCTSB_ERR_BAD_ARG = 3,     /* NULL pointer, negative size, unknown game / policy id           */
    MCTSB_ERR_BAD_STATE = 4,   /* a packed state record is not a reachable position               */
    MCTSB_ERR_NOT_SUBMITTED = 5, /* wait without a pending submit                                 */
    MCTSB_ERR_BUSY = 6,        /* submit while a previous submit has not been waited for          */
    MCTSB_ERR_OOM = 7,
    MCTSB_ERR_COMM = 8,        /* an NCCL call failed / libnccl.so.2 is not present: multi-GPU exchange only       */
    MCTSB_ERR_TREE_FULL = 9    /* device-resident tree: node pool exhausted or a selection path deeper than the level buffers */
} mctsb_status;

/* ---- games and rollout policies ------------------------------------------------------- */
#define MCTSB_GAME_QUORIDOR  0   /* reference examples/Quoridor/Quoridor.{h,cpp}                   */
#define MCTSB_GAME_TICTACTOE 1   /* reference examples/TicTacToe/TicTacToe.{h,cpp}                 */

/* Quoridor policies.  REF = bug-faithful `Quoridor_state::rollout()` (Quoridor.cpp:809-855):
 * <= 50 plies of pick_semirandom_move (Quoridor.cpp:680-803), then evaluate_position
 * (Quoridor.cpp:629-664).  UNI = uniformly random legal move per ply over the reference's
 * generate_all_moves order (Quoridor.cpp:594-616); no reference function implements it, the
 * oracle is a loop over the reference's public API. */
#define MCTSB_POLICY_REF 0
#define MCTSB_POLICY_UNI 1
#define MCTSB_UNI_MAX_PLIES 4096   /* UNI ply cap; a capped playout scores 0.5                     */

/* ---- wire formats ------------------------------------------------------------------------
 * Quoridor position, 32 bytes, little endian.  Replaces the 192-byte `Quoridor_state` object
 * (Quoridor.h:32-50).  Coordinates as in the reference: x = row 0..8, y = column 0..8; White
 * (player 1) starts on (0,4) and wins on row 8, Black starts on (8,4) and wins on row 0.
 * A 2-long wall is stored once, by its anchor (x,y), 0 <= x,y < 8, bit 8*x+y:
 *   horizontal anchor (x,y): segments under cells (x,y),(x,y+1)  => walls[x][y], walls[x][y+1] get 'h'
 *   vertical   anchor (x,y): segments right of cells (x,y),(x+1,y) => walls[x][y], walls[x+1][y] get 'v'
 *   wall_connections[x][y] = h_anchor bit | v_anchor bit            (Quoridor.cpp:351-357)        */
typedef struct mctsb_qstate {
    uint64_t h_anchor;
    uint64_t v_anchor;
    uint8_t  wx, wy, bx, by;      /* pawn rows / columns                                          */
    uint8_t  wwalls, bwalls;      /* walls left in hand (wwallsno, bwallsno)                      */
    uint8_t  turn;                /* 0 = White ('W', player 1) to move, 1 = Black ('B')           */
    uint8_t  flags;               /* must be 0                                                    */
    uint32_t move_counter;        /* Quoridor_state::move_counter (Quoridor.h:48)                 */
    uint32_t reserved;            /* must be 0                                                    */
} mctsb_qstate;

/* Tic-Tac-Toe position, 4 bytes: bit c (c = 3*row+col) of x_mask / o_mask.  Replaces
 * `TicTacToe_state` (TicTacToe.h:7-24). */
typedef struct mctsb_tstate {
    uint16_t x_mask;              /* low 9 bits                                                   */
    uint16_t o_mask;              /* low 9 bits; bit 15 = 1 when 'o' is to move                   */
} mctsb_tstate;

/* Canonical Quoridor move ids (209 slots): 0..80 pawn to cell 9*x+y; 81..144 horizontal wall
 * anchored at 8*x+y; 145..208 vertical wall anchored at 8*x+y. */
#define MCTSB_Q_NUM_MOVES 209
#define MCTSB_Q_MOVE_STEP(x, y)  ((x) * 9 + (y))
#define MCTSB_Q_MOVE_HWALL(x, y) (81 + (x) * 8 + (y))
#define MCTSB_Q_MOVE_VWALL(x, y) (145 + (x) * 8 + (y))

/* Per-rollout outcome record (parity mode), 16 bytes.  `score` is the double the reference
 * rollout() would return, computed on the device with round-to-nearest IEEE ops only.
 *   kind: 0 White reached row 8 (1.0) | 1 Black reached row 0 (0.0) | 2 heuristic eval after
 *         the ply budget (Quoridor.cpp:854) | 3 eval after a failed move pick (Quoridor.cpp:846-849)
 *         | 4 UNI ply cap (0.5).  Tic-Tac-Toe: 0 x won, 1 o won, 2 draw.                         */
typedef struct mctsb_outcome {
    double   score;
    uint16_t plies;               /* moves actually played in the playout                         */
    uint8_t  kind;
    uint8_t  reserved;
    uint32_t draws;               /* u32 random draws consumed                                    */
} mctsb_outcome;

/* Exact per-leaf accumulator.  Every rollout score is a multiple of 2^-57 in [0,1], so the
 * device accumulates the integer score*2^57 split in two 29-bit limbs; sums are exact and
 * independent of thread / block / GPU order.  score_sum = (hi * 2^29 + lo) * 2^-57. */
typedef struct mctsb_leaf_sum {
    uint64_t lo;                  /* sum of (score*2^57) & (2^29-1)                               */
    uint64_t hi;                  /* sum of (score*2^57) >> 29                                    */
    uint64_t n;                   /* rollouts accumulated                                         */
} mctsb_leaf_sum;

typedef struct mctsb_backend mctsb_backend;   /* opaque; not thread-safe, distinct handles are independent */

/* ---- lifetime ------------------------------------------------------------------------------
 * Replaces the function-local `static JobScheduler scheduler` (mcts.cpp:60) and its
 * constructor/destructor (JobScheduler.cpp:10-34).  `seed` keys the Philox4x32-10 streams. */
int mctsb_create(mctsb_backend **out, int device, uint64_t seed);
int mctsb_destroy(mctsb_backend *b);
int mctsb_abi_version(void);
const char *mctsb_strerror(int status);
const char *mctsb_last_cuda_error(const mctsb_backend *b);
int mctsb_device_count(int *count);

/* ---- the hot path ---------------------------------------------------------------------------
 * mctsb_submit replaces `scheduler.schedule(new RolloutJob(state, &results[i]))` for
 * i < NUMBER_OF_THREADS (mcts.cpp:62-64): `n_leaves` packed leaf states (host memory, copied
 * before return), `rollouts_per_leaf` independent playouts of each.  Rollout j of leaf i draws
 * from Philox stream id `first_rollout_id + i*rollouts_per_leaf + j`, so a result does not
 * depend on how rollouts are placed on threads, blocks or GPUs.
 * mctsb_wait replaces `scheduler.waitUntilJobsHaveFinished()` plus the aggregation loop
 * (mcts.cpp:66-75): it blocks, then fills `score_sum[i]` (double, correctly rounded exact sum)
 * and/or `sums[i]` (exact integers); both may be NULL.  The caller then runs
 * backpropagate(score_sum[i], rollouts_per_leaf) (mcts.cpp:76,83-90) on the host. */
int mctsb_submit(mctsb_backend *b, int game, int policy,
                 const void *leaf_states, int n_leaves,
                 uint32_t rollouts_per_leaf, uint64_t first_rollout_id);
int mctsb_wait(mctsb_backend *b, double *score_sum, mctsb_leaf_sum *sums);

/* The same pair for a tree that also wants the NEXT Expansion prepared on the device: with
 * MCTSB_SUBMIT_LEGAL (Quoridor only) the leaf states that were just uploaded for their rollouts are also run
 * through the move-generation kernel, and mctsb_wait_x returns their 209-bit legal-move masks (4 words per
 * leaf, canonical move ids) next to the sums.  Replaces the `state->actions_to_try()` the reference's node
 * constructor runs on the CPU for every new node (mcts.cpp:19 -> generate_all_moves, Quoridor.cpp:594-616):
 * when a leaf is selected again, its move list is already there.  `legal` may be NULL. */
#define MCTSB_SUBMIT_LEGAL 1u
int mctsb_submit_x(mctsb_backend *b, int game, int policy, const void *leaf_states, int n_leaves,
                   uint32_t rollouts_per_leaf, uint64_t first_rollout_id, uint32_t flags);
int mctsb_wait_x(mctsb_backend *b, double *score_sum, mctsb_leaf_sum *sums, uint64_t *legal /* n_leaves*4 or NULL */);

/* Synchronous convenience: submit + wait (the shape of MCTS_node::rollout(), mcts.cpp:57-81). */
int mctsb_rollout_batch(mctsb_backend *b, int game, int policy,
                        const void *leaf_states, int n_leaves,
                        uint32_t rollouts_per_leaf, uint64_t first_rollout_id,
                        double *score_sum, mctsb_leaf_sum *sums);

/* Parity mode: one outcome record per rollout (n_leaves*rollouts_per_leaf records, leaf-major),
 * and optionally the final position of every Quoridor playout.  Replaces `*score =
 * state->rollout()` of RolloutJob::run (mcts.h:87-90) one to one. */
int mctsb_rollout_outcomes(mctsb_backend *b, int game, int policy,
                           const void *leaf_states, int n_leaves,
                           uint32_t rollouts_per_leaf, uint64_t first_rollout_id,
                           mctsb_outcome *outcomes, void *final_states /* nullable */);

/* Replay mode: the random draws come from caller-supplied u32 streams instead of Philox
 * (stream i has `stream_len` words starting at streams + i*stream_stride, read modulo
 * stream_len).  One rollout per (state, stream) pair; used for the bit-exact comparison with the
 * reference compiled against the same pre-drawn stream. */
int mctsb_rollout_replay(mctsb_backend *b, int game, int policy,
                         const void *states, int n,
                         const uint32_t *streams, uint32_t stream_len, uint32_t stream_stride,
                         mctsb_outcome *outcomes, void *final_states /* nullable */);

/* ---- rule primitives (corpus parity; later: batched node expansion) ----------------------
 * For each Quoridor state: legal[i*4..i*4+3] = 209-bit mask of legal canonical moves for the
 * side to move (legal_move, Quoridor.cpp:331-342 with check_blocking=true), i.e. the set
 * generate_all_moves enumerates (Quoridor.cpp:594-616); info[i] = {winner 0 none/1 W/2 B
 * (check_winner, Quoridor.cpp:40-44), shortest path W, shortest path B (get_shortest_path,
 * Quoridor.cpp:108-174; -1 unreachable), force_playwall (Quoridor.cpp:666-678)};
 * eval[i] = evaluate_position (Quoridor.cpp:629-664).  Any output pointer may be NULL. */
typedef struct mctsb_qinfo {
    int8_t winner;
    int8_t sp_white;
    int8_t sp_black;
    int8_t force_playwall;
} mctsb_qinfo;
int mctsb_q_movegen(mctsb_backend *b, const mctsb_qstate *states, int n,
                    uint64_t *legal /* n*4 */, mctsb_qinfo *info /* n */, double *eval /* n */);

/* The reference's other action generator, generate_good_moves (Quoridor.cpp:518-592; selected instead of
 * generate_all_moves when TEST_ALL_MOVES is not defined, Quoridor.cpp:7,618-627): moves[i*MCTSB_Q_GOOD_MAX ..]
 * = canonical move ids in the reference's queue order, counts[i] = how many.  One warp per state, one lane
 * per candidate wall for the "shortest path of both players with this wall" flood fills. */
#define MCTSB_Q_GOOD_MAX 144
int mctsb_q_goodmoves(mctsb_backend *b, const mctsb_qstate *states, int n,
                      uint8_t *moves /* n*MCTSB_Q_GOOD_MAX */, int32_t *counts /* n */);

/* Batched next_state (Quoridor.cpp:506-510 -> play_move, Quoridor.cpp:344-392): out[i] =
 * states[i] after canonical move ids[i]; ok[i] = 0 when the move is illegal (state unchanged). */
int mctsb_q_play(mctsb_backend *b, const mctsb_qstate *states, const int32_t *move_ids, int n,
                 mctsb_qstate *out, uint8_t *ok);

/* ---- device-resident throughput path (bench `value`, no host copies in the timed region) ---
 * Runs `n_leaves*rollouts_per_leaf` rollouts from states already uploaded with
 * mctsb_upload_states; results stay on the device until mctsb_wait.  `elapsed_ms` (nullable)
 * receives the CUDA-event time of the kernel(s) on the backend stream. */
int mctsb_upload_states(mctsb_backend *b, int game, const void *leaf_states, int n_leaves);
int mctsb_run_resident(mctsb_backend *b, int game, int policy, uint32_t rollouts_per_leaf,
                       uint64_t first_rollout_id, float *elapsed_ms);

/* Integer-issue peak microbenchmark (the roofline denominator): independent
 * LOP3/IADD3 chains on every SM.  Returns thread-level integer ops per second. */
int mctsb_int_issue_peak(mctsb_backend *b, double *thread_ops_per_sec, float *elapsed_ms);
/* The same with LOP3 only: the ALU pipe's own ceiling (the flood fill is LOP3-bound). */
int mctsb_lop3_peak(mctsb_backend *b, double *thread_ops_per_sec, float *elapsed_ms);
/* Bench hygiene: overwrite a 256 MiB scratch buffer (twice the L2) on the backend stream and wait. */
int mctsb_flush_l2(mctsb_backend *b);

/* Counters of the last kernel launch(es): launches issued by this backend since creation, and
 * the executed-work counters the rollout kernel accumulates (flood-fill steps, plies). */
typedef struct mctsb_counters {
    uint64_t kernel_launches;
    uint64_t rollouts;
    uint64_t plies;
    uint64_t flood_steps;      /* frontier-expansion iterations executed                         */
    uint64_t flood_queries;    /* shortest-path / reachability queries executed                  */
} mctsb_counters;
int mctsb_get_counters(mctsb_backend *b, mctsb_counters *out);

/* Helpers shared by hosts: exact sum -> double, and the reference's score for one record. */
double mctsb_leaf_sum_to_double(const mctsb_leaf_sum *s);

/* ---- multi-GPU exchange -------------------------------------------------------
 * The path shards by rollout id with no data-path collective; the one exchange is a sum over GPUs of a few
 * hundred bytes: the root-child statistics of root-parallel trees before the reference's final choice rule
 * (MCTS_tree::select_best_child with c = 0, mcts.cpp:104-129,268-270) runs on identical numbers on every
 * rank, or the exact per-leaf accumulators of a sharded batch.  NCCL over NVLink; libnccl.so.2 is loaded on
 * the first call (MCTSB_ERR_COMM when absent), a single-GPU user never needs it.
 *   one process per GPU : rank 0 calls mctsb_comm_unique_id, ships the 128 bytes to the others by any
 *                         means, every rank calls mctsb_comm_init_rank on its handle;
 *   one process, n GPUs : mctsb_comm_init_all over the n handles (one per device); reductions issued from
 *                         one thread go through mctsb_allreduce_root_stats_group.
 * The reductions are in place over host buffers, blocking, on the handle's stream; they fail with
 * MCTSB_ERR_NOT_SUBMITTED without a communicator and MCTSB_ERR_BUSY while a submit is pending. */
#define MCTSB_COMM_ID_BYTES 128
int mctsb_comm_unique_id(void *id128);
int mctsb_comm_init_rank(mctsb_backend *b, const void *id128, int world, int rank);
int mctsb_comm_init_all(mctsb_backend **handles, int n);
int mctsb_comm_destroy(mctsb_backend *b);
int mctsb_comm_info(const mctsb_backend *b, int *world, int *rank);
int mctsb_allreduce_root_stats(mctsb_backend *b, double *stats /* in/out */, int n);
int mctsb_allreduce_leaf_sums(mctsb_backend *b, mctsb_leaf_sum *sums /* in/out */, int n_leaves);
int mctsb_allreduce_root_stats_group(mctsb_backend **handles, double **stats, int n_handles, int n);

/* ---- the search tree resident in HBM (tree-parallel selection, virtual loss, batched backprop) ----
 * Replaces the loop of MCTS_tree::grow_tree (mcts.cpp:182-210) — select (mcts.cpp:161-171), select_best_child
 * (mcts.cpp:104-130), expand (mcts.cpp:37-55, children in generate_all_moves order, Quoridor.cpp:594-616), rollout
 * (mcts.cpp:57-81) and backpropagate (mcts.cpp:83-90) — for Quoridor trees whose Simulation phase is batched as
 * `leaves_per_flush` leaves x `rollouts_per_leaf` playouts: the nodes live in device memory, one launch selects and
 * expands a whole flush, the rollout kernel reads the leaf positions where they lie, one launch back-propagates.
 * The host queues launches and never sees a node.  The tree is the host tree's (host/mcts.cpp, same batching) node
 * for node and double for double: selections are made in the same order with the same virtual losses, UCT ties are
 * settled by the reference's double expression (ln N from the host's libm), scores are summed in selection order.
 * Leaf d of a flush draws from Philox streams first_rollout_id + d*rollouts_per_leaf + r, as mctsb_submit numbers them.
 *   max_nodes: node pool (94 bytes per node, allocated twice; a node reserves ids for all its children when it is first
 *   expanded; mctsb_tree_advance reclaims the subtrees it drops by copying the kept one into the second pool).
 * Not thread-safe; uses the stream of `b`, which must outlive the tree and must not have a submit pending. */
typedef struct mctsb_tree mctsb_tree;
typedef struct mctsb_tree_info {
    uint64_t root_sims;            /* MCTS_node::number_of_simulations of the root                                */
    double   root_score;           /* its score                                                                   */
    uint32_t root_size;            /* MCTS_tree::get_size()                                                       */
    uint32_t root_children;        /* children expanded so far                                                    */
    uint32_t nodes_used;           /* node ids handed out (pool fill)                                             */
    uint32_t deepest_level;        /* deepest selection path of any flush so far                                  */
    uint32_t exact_evals;          /* selections settled by the double-precision expression                       */
    uint32_t flags;                /* bit 3: some ln N came from the device's log() instead of the host's table   */
    uint64_t flushes, rollouts;    /* flushes played, playouts played                                             */
    float    ms_select, ms_rollout, ms_backprop;   /* CUDA-event time of the three phases, summed over the last mctsb_tree_grow (synchronous mode) */
    float    ms_total;             /* CUDA-event time from the first launch of the last mctsb_tree_grow to the end of its last one */
} mctsb_tree_info;
int mctsb_tree_create(mctsb_backend *b, mctsb_tree **out, const mctsb_qstate *root, int policy,
                      uint32_t rollouts_per_leaf, uint32_t max_leaves_per_flush, uint64_t max_nodes, double uct_c);
int mctsb_tree_destroy(mctsb_tree *t);
/* `iterations` Selection+Expansion+Simulation+Backpropagation rounds in flushes of `leaves_per_flush` (the last one
 * may be shorter); returns after the last flush has been back-propagated. */
int mctsb_tree_grow(mctsb_tree *t, uint32_t iterations, uint32_t leaves_per_flush, uint64_t first_rollout_id);
/* on: flush f+1 is selected while the playouts of flush f run (its leaves keep their virtual losses until the
 * results arrive) — the order of the host tree with MCTS_batching::overlap, one flush in flight; off (default):
 * every flush is back-propagated before the next one is selected. */
int mctsb_tree_set_overlap(mctsb_tree *t, int on);
/* on: the tree expands generate_good_moves (Quoridor.cpp:518-592, the reference's actions_to_try without TEST_ALL_MOVES,
 * :618-627) instead of generate_all_moves — Quoridor_state::set_good_moves_only of the host tree.  Before the first grow. */
int mctsb_tree_set_good_moves(mctsb_tree *t, int on);
int mctsb_tree_get_info(mctsb_tree *t, mctsb_tree_info *info);
/* root children in creation (= actions_to_try) order: canonical move id, simulations, score; returns the count in *n */
int mctsb_tree_root_children(mctsb_tree *t, int32_t *move_ids, uint64_t *sims, double *scores, int max_children, int *n);
/* the reference's final choice: select_best_child with c = 0 (mcts.cpp:268-270); -1 when the root has no child */
int mctsb_tree_best_move(mctsb_tree *t, int32_t *move_id);
/* advance_tree (mcts.cpp:132-157,260-264): the child that carries the move becomes the root (its subtree is kept);
 * a move without a child starts a fresh root.  MCTSB_ERR_BAD_ARG for an illegal move. */
int mctsb_tree_advance(mctsb_tree *t, int32_t move_id);
/* Root-parallel trees (one per GPU, same root, different Philox keys): sum the root children's statistics over the
 * GPUs `b` is joined with (mctsb_comm_*), as MCTS_tree::merge_root_stats does for host trees; afterwards
 * mctsb_tree_best_move applies the reference's final rule to the same numbers on every rank.  The root must be fully
 * expanded (MCTSB_ERR_BAD_STATE otherwise).  A backend that stands alone (no communicator, or a world of one): no-op. */
int mctsb_tree_merge_root(mctsb_tree *t);
int mctsb_tree_root_state(mctsb_tree *t, mctsb_qstate *out);
/* the whole tree in pre-order (children in creation order), one row per node, at most `cap` rows written; *n = nodes */
int mctsb_tree_dump(mctsb_tree *t, uint64_t cap, int32_t *depth, int32_t *move_id, uint64_t *sims, double *score,
                    uint32_t *size, uint32_t *n_children, uint64_t *n);

#ifdef __cplusplus
}
#endif
#endif /* MCTSB_H */
