/*
 * ttt_oracle.c — TEST INFRASTRUCTURE.  Plain-C restatement of the reference's Tic-Tac-Toe state
 * and its uniformly random rollout (examples/TicTacToe/TicTacToe.cpp).  Pinned against the
 * compiled reference (oracle/_ref/libref_replay.so) by tests/test_oracle_vs_ref.py; the
 * reference has no golden vectors of its own.  Not product code.
 */
#include "oracle.h"
#include "philox_ref.h"
#include <string.h>

/* player_won, TicTacToe.cpp:26-34: three rows, three columns, two diagonals */
static int t_won(const char *b, char p) {
    for (int i = 0; i < 3; i++) {
        if (b[3 * i] == p && b[3 * i + 1] == p && b[3 * i + 2] == p) return 1;
        if (b[i] == p && b[3 + i] == p && b[6 + i] == p) return 1;
    }
    return (b[0] == p && b[4] == p && b[8] == p) || (b[2] == p && b[4] == p && b[6] == p);
}

/* calculate_winner, TicTacToe.cpp:116-128: 'x' before 'o' before draw. 0 none, 1 x, 2 o, 3 draw */
static int t_winner(const char *b) {
    if (t_won(b, 'x')) return 1;
    if (t_won(b, 'o')) return 2;
    for (int i = 0; i < 9; i++) if (b[i] == ' ') return 0;
    return 3;
}

static void t_load(char *b, int *turn, uint32_t x_mask, uint32_t o_mask) {
    for (int c = 0; c < 9; c++) b[c] = ((x_mask >> c) & 1) ? 'x' : ((o_mask >> c) & 1) ? 'o' : ' ';
    *turn = (o_mask & 0x8000) ? 1 : 0;
}

static uint32_t t_draw(oq_draws *d) {
    uint32_t v = d->data ? d->data[d->pos % d->len] : philox_draw(d->seed, d->rollout_id, (uint32_t)d->pos);
    d->pos++;
    return v;
}

int oracle_t_primitives(uint32_t x_mask, uint32_t o_mask, int32_t *info, int32_t *actions, int32_t *n_actions) {
    char b[9]; int turn; t_load(b, &turn, x_mask, o_mask);
    info[0] = t_winner(b);
    info[1] = info[0] != 0;               /* is_terminal, TicTacToe.cpp:36-38 */
    info[2] = turn == 0;                  /* player1_turn, TicTacToe.h:23     */
    int n = 0;                            /* actions_to_try, TicTacToe.cpp:63-71: ascending cells */
    for (int c = 0; c < 9; c++) if (b[c] == ' ') actions[n++] = c;
    *n_actions = n;
    return 0;
}

/* TicTacToe_state::rollout, TicTacToe.cpp:73-107.  `available` is built with push_front, i.e. in
 * DESCENDING cell order; r = rand() % size with rand() = u32 & 0x7fffffff; srand is a no-op. */
double oracle_t_rollout(uint32_t x_mask, uint32_t o_mask, oq_draws *src, int32_t *kind, int32_t *plies) {
    char b[9]; int turn; t_load(b, &turn, x_mask, o_mask);
    int w = t_winner(b), n = 0, avail[9], na = 0;
    if (!w) {
        for (int c = 8; c >= 0; c--) if (b[c] == ' ') avail[na++] = c;
        do {
            if (na == 0) { w = 2; break; }          /* :89-92 returns 0.0 like an 'o' win        */
            uint32_t r = (t_draw(src) & 0x7fffffffu) % (uint32_t)na;
            int cell = avail[r];
            memmove(avail + r, avail + r + 1, (size_t)(na - 1 - (int)r) * sizeof(int)); na--;
            b[cell] = turn ? 'o' : 'x'; turn = 1 - turn; n++;
            w = t_winner(b);
        } while (!w);
    }
    if (kind) *kind = w == 1 ? 0 : w == 2 ? 1 : 2;
    if (plies) *plies = n;
    return w == 1 ? 1.0 : w == 3 ? 0.5 : 0.0;
}

int oracle_t_rollout_philox(uint32_t x_mask, uint32_t o_mask, uint64_t seed, uint64_t first_id, int n,
                            mctsb_outcome *outcomes) {
    for (int i = 0; i < n; i++) {
        oq_draws d = {NULL, 0, seed, first_id + (uint64_t)i, 0};
        int32_t k = 0, p = 0;
        double r = oracle_t_rollout(x_mask, o_mask, &d, &k, &p);
        memset(outcomes + i, 0, sizeof *outcomes);
        outcomes[i].score = r; outcomes[i].kind = (uint8_t)k; outcomes[i].plies = (uint16_t)p; outcomes[i].draws = (uint32_t)d.pos;
    }
    return 0;
}

int oracle_t_rollout_streams(const mctsb_tstate *states, int n, const uint32_t *streams, uint32_t stream_len,
                             uint32_t stream_stride, mctsb_outcome *outcomes) {
    for (int i = 0; i < n; i++) {
        oq_draws d = {streams + (size_t)i * stream_stride, stream_len, 0, 0, 0};
        int32_t k = 0, p = 0;
        double r = oracle_t_rollout(states[i].x_mask, states[i].o_mask, &d, &k, &p);
        memset(outcomes + i, 0, sizeof *outcomes);
        outcomes[i].score = r; outcomes[i].kind = (uint8_t)k; outcomes[i].plies = (uint16_t)p; outcomes[i].draws = (uint32_t)d.pos;
    }
    return 0;
}
