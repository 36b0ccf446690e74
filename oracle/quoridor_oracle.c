/*
 * quoridor_oracle.c — TEST INFRASTRUCTURE.  Plain-C, array-based restatement of the reference's
 * Quoridor rollout path.  See oracle/oracle.h for how it is pinned (against the compiled,
 * unmodified reference; the reference has no golden vectors of its own).  Each function cites
 * the reference lines it restates (paths relative to /root/reference).
 *
 * Not product code: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg use it.
 */
#include "oracle.h"
#include "philox_ref.h"
#include <string.h>

/* ---------------------------------------------------------------------------------------------
 * State.  Mirrors Quoridor_state's members (examples/Quoridor/Quoridor.h:32-48) with the
 * walls[9][9] character table split in two boolean planes:
 *   hseg[x][y] <=> walls[x][y] in {'h','b'}  (segment under cell (x,y), blocks x <-> x+1)
 *   vseg[x][y] <=> walls[x][y] in {'v','b'}  (segment right of cell (x,y), blocks y <-> y+1)
 *   hanc/vanc[x][y]: a 2-long wall of that orientation is anchored at lattice point (x,y);
 *   wall_connections[x][y] == hanc[x][y] || vanc[x][y]  (Quoridor.cpp:357).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int px[2], py[2];          /* [0] = White, [1] = Black  (wx,wy / bx,by)                   */
    int nwalls[2];             /* wwallsno, bwallsno                                           */
    int turn;                  /* 0 = 'W', 1 = 'B'                                             */
    uint32_t move_counter;
    uint8_t hseg[9][9], vseg[9][9];
    uint8_t hanc[8][8], vanc[8][8];
    /* lazily cached shortest paths, Quoridor.h:45-46 (only the scalar is ever consumed)       */
    int sp_valid[2], sp_cache[2];
} qs_t;

static oq_work *g_work = NULL;          /* optional work counters for the call in flight        */

static void qs_load(qs_t *q, const mctsb_qstate *s) {
    memset(q, 0, sizeof *q);
    q->px[0] = s->wx; q->py[0] = s->wy; q->px[1] = s->bx; q->py[1] = s->by;
    q->nwalls[0] = s->wwalls; q->nwalls[1] = s->bwalls;
    q->turn = s->turn ? 1 : 0;
    q->move_counter = s->move_counter;
    for (int a = 0; a < 64; a++) {
        int x = a >> 3, y = a & 7;
        if ((s->h_anchor >> a) & 1) { q->hanc[x][y] = 1; q->hseg[x][y] = 1; q->hseg[x][y + 1] = 1; }
        if ((s->v_anchor >> a) & 1) { q->vanc[x][y] = 1; q->vseg[x][y] = 1; q->vseg[x + 1][y] = 1; }
    }
}

static void qs_store(const qs_t *q, mctsb_qstate *s) {
    memset(s, 0, sizeof *s);
    for (int a = 0; a < 64; a++) {
        if (q->hanc[a >> 3][a & 7]) s->h_anchor |= 1ull << a;
        if (q->vanc[a >> 3][a & 7]) s->v_anchor |= 1ull << a;
    }
    s->wx = (uint8_t)q->px[0]; s->wy = (uint8_t)q->py[0]; s->bx = (uint8_t)q->px[1]; s->by = (uint8_t)q->py[1];
    s->wwalls = (uint8_t)q->nwalls[0]; s->bwalls = (uint8_t)q->nwalls[1];
    s->turn = (uint8_t)q->turn; s->move_counter = q->move_counter;
}

/* check_winner, Quoridor.cpp:40-44: White is tested first. 0 none, 1 W, 2 B */
static int qs_winner(const qs_t *q) {
    if (q->px[0] == 8) return 1;
    if (q->px[1] == 0) return 2;
    return 0;
}

/* is there a wall segment between cell (x,y) and its neighbour in direction (dx,dy)?  The
 * four cases are the four neighbour tests of calculate_dists_from, Quoridor.cpp:86-103. */
static int wall_between(const qs_t *q, int x, int y, int dx, int dy) {
    if (dx == -1) return q->hseg[x - 1][y];
    if (dx == +1) return q->hseg[x][y];
    if (dy == -1) return q->vseg[x][y - 1];
    return q->vseg[x][y];
}
static int on_board(int x, int y) { return x >= 0 && x < 9 && y >= 0 && y < 9; }
/* step from (x,y) in direction d impossible: off board or walled */
static int blocked(const qs_t *q, int x, int y, int dx, int dy) {
    if (!on_board(x + dx, y + dy)) return 1;
    return wall_between(q, x, y, dx, dy);
}

/* BFS distance from (x,y) to `player`'s goal row ignoring pawns; -1 if unreachable.
 * calculate_dists_from with stop_at_goal=true + the end-zone scan of get_shortest_path
 * (Quoridor.cpp:46-106,161-169).  Neighbour order up, down, left, right; the search stops at the
 * first goal-row cell popped or discovered, whose distance is the minimum over the row. */
static int qs_bfs(const qs_t *q, int x, int y, int player) {
    int goal = player == 0 ? 8 : 0;
    int16_t dist[9][9];
    int qx[81], qy[81], head = 0, tail = 0;
    static const int DX[4] = {-1, +1, 0, 0}, DY[4] = {0, 0, -1, +1};
    int result = -1, levels = 0;
    memset(dist, 0xff, sizeof dist);
    dist[x][y] = 0; qx[tail] = x; qy[tail] = y; tail++;
    if (g_work) g_work->bfs_calls++;
    while (head < tail && result < 0) {
        int cx = qx[head], cy = qy[head]; head++;
        if (g_work) g_work->bfs_pops++;
        if (cx == goal) { result = dist[cx][cy]; break; }
        if (dist[cx][cy] + 1 > levels) levels = dist[cx][cy] + 1;
        for (int d = 0; d < 4; d++) {
            int nx = cx + DX[d], ny = cy + DY[d];
            if (!on_board(nx, ny) || wall_between(q, cx, cy, DX[d], DY[d]) || dist[nx][ny] >= 0) continue;
            dist[nx][ny] = (int16_t)(dist[cx][cy] + 1);
            if (d < 2 && nx == goal) { result = dist[nx][ny]; break; }
            qx[tail] = nx; qy[tail] = ny; tail++;
        }
    }
    if (g_work) g_work->flood_steps += (uint64_t)levels;
    return result;
}

/* legal_wall(..., check_blocking=false), Quoridor.cpp:291-303 (turn is implied: p == mover) */
static int qs_cheap_wall(const qs_t *q, int x, int y, int horizontal) {
    if (g_work) g_work->cheap_wall_tests++;
    if (x < 0 || y < 0 || x >= 8 || y >= 8) return 0;
    if (q->nwalls[q->turn] <= 0) return 0;
    if (horizontal) {
        if (q->hseg[x][y]) return 0;                                   /* same type or 'b'       */
        if (q->vseg[x][y] && (q->hanc[x][y] || q->vanc[x][y])) return 0; /* crossing at the anchor */
        if (q->hseg[x][y + 1]) return 0;                               /* second half taken      */
    } else {
        if (q->vseg[x][y]) return 0;
        if (q->hseg[x][y] && (q->hanc[x][y] || q->vanc[x][y])) return 0;
        if (q->vseg[x + 1][y]) return 0;
    }
    return 1;
}

static void qs_put_wall(qs_t *q, int x, int y, int horizontal, int v) {
    if (horizontal) { q->hseg[x][y] = (uint8_t)v; q->hseg[x][y + 1] = (uint8_t)v; }
    else            { q->vseg[x][y] = (uint8_t)v; q->vseg[x + 1][y] = (uint8_t)v; }
}

/* get_shortest_path(player, extra wall), Quoridor.cpp:138-156: temporary wall, BFS from pawn */
static int qs_sp_wall(qs_t *q, int player, int x, int y, int horizontal) {
    if (!qs_cheap_wall(q, x, y, horizontal)) return -1;
    qs_put_wall(q, x, y, horizontal, 1);
    int d = qs_bfs(q, q->px[player], q->py[player], player);
    qs_put_wall(q, x, y, horizontal, 0);
    return d;
}

/* get_shortest_path(player) with the wdists/bdists cache, Quoridor.cpp:111-131 */
static int qs_sp(qs_t *q, int player) {
    if (!q->sp_valid[player]) {
        q->sp_cache[player] = qs_bfs(q, q->px[player], q->py[player], player);
        q->sp_valid[player] = 1;
    }
    return q->sp_cache[player];
}

/* legal_wall(..., check_blocking=true), Quoridor.cpp:305-326, including the "isolated" shortcut */
static int qs_legal_wall(qs_t *q, int x, int y, int horizontal) {
    if (!qs_cheap_wall(q, x, y, horizontal)) return 0;
    int isolated = 1;
    for (int i = x - 1; i <= x + 1 + !horizontal && isolated; i++) {
        if (i < 0 || i >= 9) continue;
        for (int j = y - 1; j <= y + 1 + !!horizontal; j++) {
            if (j < 0 || j >= 9) continue;
            if (q->hseg[i][j] || q->vseg[i][j]) { isolated = 0; break; }
        }
    }
    if (isolated) return 1;
    if (qs_sp_wall(q, 0, x, y, horizontal) < 0) return 0;
    if (qs_sp_wall(q, 1, x, y, horizontal) < 0) return 0;
    return 1;
}

/* legal_step, Quoridor.cpp:200-287, for the side to move */
static int qs_legal_step(const qs_t *q, int x, int y) {
    int me = q->turn, en = 1 - me;
    int mx = q->px[me], my = q->py[me], ex = q->px[en], ey = q->py[en];
    if (!on_board(x, y)) return 0;
    if ((x == ex && y == ey) || (x == mx && y == my)) return 0;
    int dx = x - mx, dy = y - my;
    int adx = dx < 0 ? -dx : dx, ady = dy < 0 ? -dy : dy;
    if (adx + ady == 1) return !wall_between(q, mx, my, dx, dy);                 /* plain step      */
    if ((adx == 2 && ady == 0) || (adx == 0 && ady == 2)) {                      /* straight jump   */
        int sx = dx / 2, sy = dy / 2;
        return ex == mx + sx && ey == my + sy && !wall_between(q, mx, my, sx, sy) &&
               !wall_between(q, ex, ey, sx, sy);
    }
    if (adx == 1 && ady == 1) {                                                  /* diagonal jump   */
        /* around a vertically adjacent enemy (Quoridor.cpp:247-250 and its three mirror images) */
        if (ex == mx + dx && ey == my && blocked(q, ex, ey, dx, 0) && !blocked(q, ex, ey, 0, dy) &&
            !wall_between(q, mx, my, dx, 0)) return 1;
        /* around a horizontally adjacent enemy (Quoridor.cpp:251-254 and mirrors) */
        if (ex == mx && ey == my + dy && blocked(q, ex, ey, 0, dy) && !blocked(q, ex, ey, dx, 0) &&
            !wall_between(q, mx, my, 0, dy)) return 1;
    }
    return 0;
}

/* the fixed candidate order of get_legal_step_moves2, Quoridor.cpp:461-472 */
static const int STEP_DX[12] = {-1, -2, +1, +2, 0, 0, 0, 0, -1, -1, +1, +1};
static const int STEP_DY[12] = {0, 0, 0, 0, -1, -2, +1, +2, -1, +1, -1, +1};

static int qs_steps_forward(const qs_t *q, int *ids) {   /* get_legal_step_moves2 order */
    int n = 0, mx = q->px[q->turn], my = q->py[q->turn];
    for (int k = 0; k < 12; k++)
        if (qs_legal_step(q, mx + STEP_DX[k], my + STEP_DY[k])) ids[n++] = MCTSB_Q_MOVE_STEP(mx + STEP_DX[k], my + STEP_DY[k]);
    return n;
}
static int qs_steps_reversed(const qs_t *q, int *ids) {  /* get_legal_step_moves: push_front => reversed, Quoridor.cpp:442-453 */
    int n = 0, mx = q->px[q->turn], my = q->py[q->turn];
    for (int k = 11; k >= 0; k--)
        if (qs_legal_step(q, mx + STEP_DX[k], my + STEP_DY[k])) ids[n++] = MCTSB_Q_MOVE_STEP(mx + STEP_DX[k], my + STEP_DY[k]);
    return n;
}

/* get_best_step_move, Quoridor.cpp:476-497: first strict minimum over the reversed list; -1 = NULL */
static int qs_best_step(qs_t *q) {
    int ids[12], n = qs_steps_reversed(q, ids), best = -1, min = 9999999;
    for (int i = 0; i < n; i++) {
        int path = qs_bfs(q, ids[i] / 9, ids[i] % 9, q->turn);
        if (path >= 0 && path < min) { min = path; best = ids[i]; }
    }
    return best;
}

static int qs_legal_move(qs_t *q, int id) {              /* legal_move, Quoridor.cpp:331-342 */
    if (id < 0 || id >= MCTSB_Q_NUM_MOVES) return 0;
    if (id < 81) return qs_legal_step(q, id / 9, id % 9);
    if (id < 145) return qs_legal_wall(q, (id - 81) >> 3, (id - 81) & 7, 1);
    return qs_legal_wall(q, (id - 145) >> 3, (id - 145) & 7, 0);
}

static int qs_play(qs_t *q, int id) {                    /* play_move, Quoridor.cpp:344-392 */
    if (!qs_legal_move(q, id)) return 0;
    if (id >= 81) {
        int horizontal = id < 145, a = horizontal ? id - 81 : id - 145, x = a >> 3, y = a & 7;
        qs_put_wall(q, x, y, horizontal, 1);
        if (horizontal) q->hanc[x][y] = 1; else q->vanc[x][y] = 1;
        q->nwalls[q->turn]--;
        q->sp_valid[0] = q->sp_valid[1] = 0;
    } else {
        q->px[q->turn] = id / 9; q->py[q->turn] = id % 9;
        q->sp_valid[q->turn] = 0;
    }
    q->turn = 1 - q->turn;
    q->move_counter++;
    return 1;
}

/* generate_all_moves, Quoridor.cpp:594-616 */
static int qs_all_moves(qs_t *q, int *ids) {
    int n = qs_steps_reversed(q, ids);
    if (q->nwalls[q->turn] > 0)
        for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) for (int k = 0; k < 2; k++)
            if (qs_legal_wall(q, i, j, k == 0)) ids[n++] = k == 0 ? MCTSB_Q_MOVE_HWALL(i, j) : MCTSB_Q_MOVE_VWALL(i, j);
    return n;
}

/* generate_good_moves, Quoridor.cpp:518-592 (the alternative actions_to_try, compiled out by TEST_ALL_MOVES in the
 * shipped build but public): all legal steps, then for every cheap-legal wall in (row, column, h before v) order:
 * the wall itself when it lengthens the enemy's path by more than ours and closes nobody in; and, when the wall
 * would hurt US by at least 3 more than the enemy and the enemy still has walls, the walls that pre-empt it — the
 * crossing wall on the same anchor and the same-orientation walls one square to either side — each at most once
 * (already_used), fully legal.  A closed-in path counts as length -1 in the differences, as in the reference. */
static int qs_good_moves(qs_t *q, int *ids) {
    int n = qs_steps_reversed(q, ids);
    int p = q->turn, e = 1 - p;
    if (q->nwalls[p] <= 0) return n;
    int our_path = qs_sp(q, p), enemy_path = qs_sp(q, e);
    uint8_t used[8][8][2];
    memset(used, 0, sizeof used);
    for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) for (int k = 0; k < 2; k++) {
        if (!qs_cheap_wall(q, i, j, k == 0)) continue;
        int enemy_with = qs_sp_wall(q, e, i, j, k == 0), enemy_enc = enemy_with - enemy_path;
        int our_with = qs_sp_wall(q, p, i, j, k == 0), our_enc = our_with - our_path;
        if (!used[i][j][k] && enemy_enc > our_enc && enemy_with >= 0 && our_with >= 0) {
            used[i][j][k] = 1;
            ids[n++] = k == 0 ? MCTSB_Q_MOVE_HWALL(i, j) : MCTSB_Q_MOVE_VWALL(i, j);
        }
        if (q->nwalls[e] > 0 && our_enc - enemy_enc >= 3) {
            if (!used[i][j][1 - k] && qs_legal_wall(q, i, j, k != 0)) {          /* same anchor, crossing orientation */
                used[i][j][1 - k] = 1;
                ids[n++] = k == 0 ? MCTSB_Q_MOVE_VWALL(i, j) : MCTSB_Q_MOVE_HWALL(i, j);
            }
            if (k == 0) {
                if (j - 1 >= 0 && !used[i][j - 1][k] && qs_legal_wall(q, i, j - 1, 1)) { used[i][j - 1][k] = 1; ids[n++] = MCTSB_Q_MOVE_HWALL(i, j - 1); }
                if (j + 1 < 8 && !used[i][j + 1][k] && qs_legal_wall(q, i, j + 1, 1)) { used[i][j + 1][k] = 1; ids[n++] = MCTSB_Q_MOVE_HWALL(i, j + 1); }
            } else {
                if (i - 1 >= 0 && !used[i - 1][j][k] && qs_legal_wall(q, i - 1, j, 0)) { used[i - 1][j][k] = 1; ids[n++] = MCTSB_Q_MOVE_VWALL(i - 1, j); }
                if (i + 1 < 8 && !used[i + 1][j][k] && qs_legal_wall(q, i + 1, j, 0)) { used[i + 1][j][k] = 1; ids[n++] = MCTSB_Q_MOVE_VWALL(i + 1, j); }
            }
        }
    }
    return n;
}

/* evaluate_position, Quoridor.cpp:629-664 (MAX instead of MIN at :661 kept) */
static double qs_eval(qs_t *q) {
    int wp = qs_sp(q, 0), bp = qs_sp(q, 1), w = q->nwalls[0], b = q->nwalls[1];
    int white_ahead = wp + (q->turn != 0) <= bp - 1;
    int black_ahead = bp + (q->turn != 1) <= wp - 1;
    if (b <= 0 && white_ahead) return 0.95;
    if (w <= 0 && black_ahead) return 1.0 - 0.95;
    if (b <= 1 && w >= 2 && white_ahead) return 0.95 - 0.1;
    if (w <= 1 && b >= 2 && black_ahead) return 1.0 - (0.95 - 0.1);
    double wallsdiff = 0.0, mx = w > b ? w : b;
    if (mx > 0) wallsdiff = ((w > b) ? +1 : -1) * (((double)(w - b) * (double)(w - b)) / (mx * mx));
    double path_diff = bp - wp + (q->turn == 0 ? +0.5 : -0.5);
    double distance = ((path_diff > 10.0) ? path_diff : 10.0) / 10.0;
    return 0.5 + 0.2 * wallsdiff + 0.2 * distance;
}

/* force_playwall, Quoridor.cpp:666-678 */
static int qs_force_playwall(qs_t *q) {
    int p = q->turn, e = 1 - p;
    int ours = qs_sp(q, p), theirs = qs_sp(q, e), ow = q->nwalls[p], ew = q->nwalls[e];
    if (theirs <= 1 && ours > theirs) return 1;
    if (theirs <= 2 && ours > theirs + 1 && ow >= ew) return 1;
    if (theirs <= 3 && ours > theirs + 2 && ow > ew) return 1;
    return 0;
}

/* ---- draws (conventions: oracle/philox_ref.h) ---- */
static uint32_t draw_u32(oq_draws *d) {
    uint32_t v = d->data ? d->data[d->pos % d->len] : philox_draw(d->seed, d->rollout_id, (uint32_t)d->pos);
    d->pos++;
    return v;
}
static double draw_unit(oq_draws *d) { return (double)draw_u32(d) * (1.0 / 4294967296.0); }

/* "does this wall hurt the enemy more than us" — the test shared by the best-wall and guided
 * branches, Quoridor.cpp:728-733 / :756-761.  Returns enc difference (>0) or 0 when rejected. */
static int wall_gain(qs_t *q, int id) {
    int p = q->turn, e = 1 - p, horizontal = id < 145, a = horizontal ? id - 81 : id - 145;
    int ep = qs_sp_wall(q, e, a >> 3, a & 7, horizontal);
    int ee = ep - qs_sp(q, e);
    if (!(ep >= 0 && ee > 0)) return 0;
    int op = qs_sp_wall(q, p, a >> 3, a & 7, horizontal);
    int oe = op - qs_sp(q, p);
    if (!(op >= 0 && ee > oe)) return 0;
    return ee - oe;
}

/* pick_semirandom_move, Quoridor.cpp:680-803, bug-faithful.  -1 = NULL. */
static int qs_pick(qs_t *q, oq_draws *d) {
    int p = q->turn, e = 1 - p;
    /* :696 — force_playwall is always evaluated; one draw is consumed iff it is false; the result
     * of the || is cast to char 0/1 and handed to remaining_walls(), which returns bwallsno for
     * anything that is not 'W' (Quoridor.h:68): the branch is taken iff Black has walls left. */
    if (!qs_force_playwall(q)) (void)draw_unit(d);
    if (q->nwalls[1] != 0) {
        int pool[128], n = 0;
        for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) {      /* :709-718 */
            if (qs_cheap_wall(q, i, j, 1)) pool[n++] = MCTSB_Q_MOVE_HWALL(i, j);
            if (qs_cheap_wall(q, i, j, 0)) pool[n++] = MCTSB_Q_MOVE_VWALL(i, j);
        }
        if (n >= 2) {                                                      /* :720, our shuffle (philox_ref.h) */
            uint32_t key[128]; int rest[128], m = 0;
            int j0 = (int)mulhi32(draw_u32(d), (uint32_t)n), head = pool[j0];
            for (int i = 0; i < n; i++) if (i != j0) { key[m] = draw_u32(d); rest[m] = pool[i]; m++; }
            for (int i = 1; i < m; i++) {                                  /* stable insertion sort on the key */
                uint32_t k = key[i]; int v = rest[i], j = i - 1;
                while (j >= 0 && key[j] > k) { key[j + 1] = key[j]; rest[j + 1] = rest[j]; j--; }
                key[j + 1] = k; rest[j + 1] = v;
            }
            pool[0] = head;
            for (int i = 0; i < m; i++) pool[i + 1] = rest[i];
        }
        if (draw_unit(d) < 0.1) {                                          /* :721-746 best wall */
            int best = -1, maxdiff = 0;
            for (int i = 0; i < n; i++) { int g = wall_gain(q, pool[i]); if (g > maxdiff) { maxdiff = g; best = pool[i]; } }
            if (best >= 0) return best;
        } else if (draw_unit(d) < 0.75) {                                  /* :748-772 guided    */
            for (int i = 0; i < n; i++) if (wall_gain(q, pool[i]) > 0) return pool[i];
        } else {                                                           /* :773-788 random    */
            for (int i = 0; i < n; i++) if (qs_legal_move(q, pool[i])) return pool[i];
        }
    }
    if (q->nwalls[e] == 0 || draw_unit(d) < 0.8) return qs_best_step(q);  /* :792-793 */
    int ids[12], n = qs_steps_forward(q, ids);                             /* :795-800 */
    uint32_t r = draw_u32(d) & 0x7fffffffu;
    if (n == 0) return -1;              /* the reference divides by zero here (rand() % 0)       */
    return ids[r % (uint32_t)n];
}

/* Quoridor_state::rollout, Quoridor.cpp:809-855.  kind as in include/mctsb.h */
static double qs_rollout_ref(qs_t *q, oq_draws *d, int32_t *moves, int32_t *n_moves, int32_t *kind) {
    int nm = 0, k = 2;
    for (int i = 0; i < 50; i++) {
        int w = qs_winner(q);
        if (w) { *n_moves = nm; *kind = w == 1 ? 0 : 1; if (g_work) g_work->plies += (uint64_t)nm; return w == 1 ? 1.0 : 0.0; }
        (void)qs_eval(q);               /* :826-829 — the early-exit test can never be true      */
        int m = qs_pick(q, d);
        if (m < 0 || !qs_play(q, m)) { k = 3; break; }
        if (moves) moves[nm] = m;
        nm++;
    }
    *n_moves = nm; *kind = k;
    if (g_work) g_work->plies += (uint64_t)nm;
    return qs_eval(q);
}

/* UNI policy: uniformly random legal move over generate_all_moves order (no reference function) */
static double qs_rollout_uni(qs_t *q, oq_draws *d, int32_t *moves, int32_t *n_moves, int32_t *kind) {
    int nm = 0, k = 4; double r = 0.5;
    for (;;) {
        int w = qs_winner(q);
        if (w) { k = w == 1 ? 0 : 1; r = w == 1 ? 1.0 : 0.0; break; }
        if (nm >= MCTSB_UNI_MAX_PLIES) break;
        int ids[MCTSB_Q_NUM_MOVES], n = qs_all_moves(q, ids);
        if (n == 0) { k = 3; break; }
        int m = ids[mulhi32(draw_u32(d), (uint32_t)n)];
        qs_play(q, m);
        if (moves && nm < 64) moves[nm] = m;
        nm++;
    }
    *n_moves = nm; *kind = k;
    if (g_work) g_work->plies += (uint64_t)nm;
    return r;
}

/* ============================== exported API ============================== */

int oracle_q_raw(const mctsb_qstate *s, int32_t *out) {
    qs_t q; qs_load(&q, s);
    for (int i = 0; i < 81; i++) {
        int h = q.hseg[i / 9][i % 9], v = q.vseg[i / 9][i % 9];
        out[i] = h && v ? 'b' : h ? 'h' : v ? 'v' : ' ';
    }
    for (int i = 0; i < 64; i++) out[81 + i] = q.hanc[i >> 3][i & 7] || q.vanc[i >> 3][i & 7];
    out[145] = q.px[0]; out[146] = q.py[0]; out[147] = q.px[1]; out[148] = q.py[1];
    out[149] = q.nwalls[0]; out[150] = q.nwalls[1]; out[151] = q.turn ? 'B' : 'W'; out[152] = (int32_t)q.move_counter;
    return 0;
}

int oracle_q_primitives(const mctsb_qstate *s, int32_t *info, uint8_t *legal209, int32_t *actions,
                        int32_t *n_actions, int32_t *steps2, int32_t *n_steps2, double *eval, uint8_t *cheap128) {
    qs_t q; qs_load(&q, s);
    g_work = NULL;
    info[0] = qs_winner(&q);
    info[1] = qs_sp(&q, 0);
    info[2] = qs_sp(&q, 1);
    info[3] = qs_force_playwall(&q);
    info[4] = info[0] != 0;
    info[5] = q.turn == 0;
    if (eval) *eval = qs_eval(&q);
    if (legal209) for (int id = 0; id < MCTSB_Q_NUM_MOVES; id++) legal209[id] = (uint8_t)qs_legal_move(&q, id);
    if (actions) { int ids[MCTSB_Q_NUM_MOVES]; int n = qs_all_moves(&q, ids); for (int i = 0; i < n; i++) actions[i] = ids[i]; *n_actions = n; }
    if (steps2) { int ids[12]; int n = qs_steps_forward(&q, ids); for (int i = 0; i < n; i++) steps2[i] = ids[i]; *n_steps2 = n; }
    if (cheap128)
        for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) for (int k = 0; k < 2; k++)
            cheap128[2 * (8 * i + j) + k] = (uint8_t)qs_cheap_wall(&q, i, j, k == 0);
    return 0;
}

/* ids[] needs room for MCTSB_Q_NUM_MOVES entries; returns the count */
int oracle_q_good_moves(const mctsb_qstate *s, int32_t *ids) {
    qs_t q; qs_load(&q, s); g_work = NULL;
    int tmp[MCTSB_Q_NUM_MOVES], n = qs_good_moves(&q, tmp);
    for (int i = 0; i < n; i++) ids[i] = tmp[i];
    return n;
}

int oracle_q_sp_with_wall(const mctsb_qstate *s, int player_is_black, int id) {
    qs_t q; qs_load(&q, s); g_work = NULL;
    if (id < 81 || id >= MCTSB_Q_NUM_MOVES) return -1;
    int horizontal = id < 145, a = horizontal ? id - 81 : id - 145;
    return qs_sp_wall(&q, player_is_black ? 1 : 0, a >> 3, a & 7, horizontal);
}

int oracle_q_sp_from(const mctsb_qstate *s, int player_is_black, int x, int y) {
    qs_t q; qs_load(&q, s); g_work = NULL;
    return qs_bfs(&q, x, y, player_is_black ? 1 : 0);
}

int oracle_q_play(const mctsb_qstate *s, int id, mctsb_qstate *out) {
    qs_t q; qs_load(&q, s); g_work = NULL;
    int ok = qs_play(&q, id);
    if (out) qs_store(&q, out);
    return ok;
}

int oracle_q_best_step(const mctsb_qstate *s) {
    qs_t q; qs_load(&q, s); g_work = NULL;
    return qs_best_step(&q);
}

double oracle_q_rollout(const mctsb_qstate *s, int policy, oq_draws *src, mctsb_qstate *final_state,
                        int32_t *moves, int32_t *n_moves, int32_t *kind, oq_work *work) {
    qs_t q; qs_load(&q, s);
    int32_t nm = 0, k = 0;
    g_work = work;
    double r = policy == MCTSB_POLICY_UNI ? qs_rollout_uni(&q, src, moves, &nm, &k) : qs_rollout_ref(&q, src, moves, &nm, &k);
    g_work = NULL;
    if (final_state) qs_store(&q, final_state);
    if (n_moves) *n_moves = nm;
    if (kind) *kind = k;
    return r;
}

static void fill_outcome(mctsb_outcome *o, double score, int plies, int kind, uint64_t draws) {
    memset(o, 0, sizeof *o);
    o->score = score; o->plies = (uint16_t)plies; o->kind = (uint8_t)kind; o->draws = (uint32_t)draws;
}

int oracle_q_rollout_philox(const mctsb_qstate *s, int policy, uint64_t seed, uint64_t first_id, int n,
                            mctsb_outcome *outcomes, mctsb_qstate *finals, oq_work *work) {
    for (int i = 0; i < n; i++) {
        oq_draws d = {NULL, 0, seed, first_id + (uint64_t)i, 0};
        int32_t nm = 0, k = 0;
        double r = oracle_q_rollout(s, policy, &d, finals ? finals + i : NULL, NULL, &nm, &k, work);
        fill_outcome(outcomes + i, r, nm, k, d.pos);
    }
    return 0;
}

int oracle_q_rollout_streams(const mctsb_qstate *states, int policy, int n, const uint32_t *streams,
                             uint32_t stream_len, uint32_t stream_stride, mctsb_outcome *outcomes, mctsb_qstate *finals) {
    for (int i = 0; i < n; i++) {
        oq_draws d = {streams + (size_t)i * stream_stride, stream_len, 0, 0, 0};
        int32_t nm = 0, k = 0;
        double r = oracle_q_rollout(states + i, policy, &d, finals ? finals + i : NULL, NULL, &nm, &k, NULL);
        fill_outcome(outcomes + i, r, nm, k, d.pos);
    }
    return 0;
}

/* exact accumulation: every score is a multiple of 2^-57 in [0,1] (include/mctsb.h) */
void oracle_leaf_sum_add(mctsb_leaf_sum *acc, double score) {
    uint64_t fx = (uint64_t)(score * 144115188075855872.0);   /* 2^57, exact */
    acc->lo += fx & ((1ull << 29) - 1);
    acc->hi += fx >> 29;
    acc->n += 1;
}

double oracle_leaf_sum_to_double(const mctsb_leaf_sum *acc) {
    unsigned __int128 t = ((unsigned __int128)acc->hi << 29) + acc->lo;
    /* correctly rounded (round-to-nearest-even) conversion of t * 2^-57 */
    if (t == 0) return 0.0;
    int msb = 127;
    while (!((t >> msb) & 1)) msb--;
    if (msb <= 52) return (double)(uint64_t)t * (1.0 / 144115188075855872.0);
    int shift = msb - 52;
    uint64_t mant = (uint64_t)(t >> shift);
    unsigned __int128 rem = t & (((unsigned __int128)1 << shift) - 1), half = (unsigned __int128)1 << (shift - 1);
    if (rem > half || (rem == half && (mant & 1))) mant++;
    double r = (double)mant;                 /* exact: mant <= 2^53 */
    for (int i = 0; i < shift; i++) r *= 2.0;
    return r * (1.0 / 144115188075855872.0);
}
