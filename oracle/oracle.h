/*
 * oracle.h — TEST INFRASTRUCTURE.  Plain-C restatement of the reference's rollout path
 * (Quoridor + Tic-Tac-Toe), array based like the reference, no bitboards, no CUDA.
 *
 * PARITY PINNING: the reference ships no tests, golden vectors or fixtures,
 * so this restatement is pinned against the reference ITSELF, compiled unmodified into
 * oracle/_ref/libref_replay.so (oracle/Makefile, oracle/shim.h, oracle/harness.cpp):
 * tests/test_oracle_vs_ref.py compares every function below with the reference on a random
 * mid-game corpus and on replayed playouts, and tests/golden/ holds vectors generated from the
 * reference by tests/golden/make_golden.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may link or load this
 * code.  The product (montecarlotreesearch_b200/) never does.
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stdint.h>
#include "../include/mctsb.h"

typedef struct oq_draws {
    const uint32_t *data;   /* explicit stream, read modulo len; NULL => Philox(seed, rollout_id) */
    uint64_t len;
    uint64_t seed;
    uint64_t rollout_id;
    uint64_t pos;           /* draws consumed                                                    */
} oq_draws;

typedef struct oq_work {    /* algorithmic work counters                          */
    uint64_t bfs_calls;
    uint64_t bfs_pops;
    uint64_t flood_steps;   /* sum over BFS calls of the number of BFS levels expanded           */
    uint64_t cheap_wall_tests;
    uint64_t plies;
} oq_work;

#ifdef __cplusplus
extern "C" {
#endif

/* info[0]=winner(0 none,1 W,2 B) [1]=sp W [2]=sp B [3]=force_playwall [4]=terminal [5]=player1_turn */
int oracle_q_primitives(const mctsb_qstate *s, int32_t *info, uint8_t *legal209, int32_t *actions,
                        int32_t *n_actions, int32_t *steps2, int32_t *n_steps2, double *eval,
                        uint8_t *cheap128);
int oracle_q_raw(const mctsb_qstate *s, int32_t *out153);
int oracle_q_sp_with_wall(const mctsb_qstate *s, int player_is_black, int wall_move_id);
int oracle_q_sp_from(const mctsb_qstate *s, int player_is_black, int x, int y);
int oracle_q_play(const mctsb_qstate *s, int move_id, mctsb_qstate *out);
int oracle_q_best_step(const mctsb_qstate *s);
int oracle_q_good_moves(const mctsb_qstate *s, int32_t *ids /* room for MCTSB_Q_NUM_MOVES */);

/* policy 0 = REF (Quoridor.cpp:809-855), 1 = UNI.  final/moves/work nullable. */
double oracle_q_rollout(const mctsb_qstate *s, int policy, oq_draws *src, mctsb_qstate *final_state,
                        int32_t *moves, int32_t *n_moves, int32_t *kind, oq_work *work);
int oracle_q_rollout_philox(const mctsb_qstate *s, int policy, uint64_t seed, uint64_t first_id, int n,
                            mctsb_outcome *outcomes, mctsb_qstate *finals, oq_work *work);
int oracle_q_rollout_streams(const mctsb_qstate *states, int policy, int n, const uint32_t *streams,
                             uint32_t stream_len, uint32_t stream_stride, mctsb_outcome *outcomes,
                             mctsb_qstate *finals);

int oracle_t_primitives(uint32_t x_mask, uint32_t o_mask, int32_t *info, int32_t *actions, int32_t *n_actions);
double oracle_t_rollout(uint32_t x_mask, uint32_t o_mask, oq_draws *src, int32_t *kind, int32_t *plies);
int oracle_t_rollout_philox(uint32_t x_mask, uint32_t o_mask, uint64_t seed, uint64_t first_id, int n,
                            mctsb_outcome *outcomes);
int oracle_t_rollout_streams(const mctsb_tstate *states, int n, const uint32_t *streams, uint32_t stream_len,
                             uint32_t stream_stride, mctsb_outcome *outcomes);

/* exact integer accumulation of scores, the same arithmetic include/mctsb.h documents */
void oracle_leaf_sum_add(mctsb_leaf_sum *acc, double score);
double oracle_leaf_sum_to_double(const mctsb_leaf_sum *acc);

#ifdef __cplusplus
}
#endif
#endif
