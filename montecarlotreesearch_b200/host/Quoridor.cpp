// Quoridor.cpp — host game plugin (see Quoridor.h).  All rule logic is csrc/qcore.cuh compiled for
// the host; this file only adapts it to the reference's object interface.
#include "Quoridor.h"
#include "rollout_backend.h"
#include "mctsb.h"
#include <cstdio>
#include <cstring>

using namespace mctsb;

string Quoridor_move::sprint() const {
    string what = (type == 'h') ? "places horizontal wall at" : (type == 'v') ? "places vertical wall at" : "moves to";
    string who = (player == 'W') ? "White" : "Black";
    return who + " " + what + " " + string(1, (char)('A' + y)) + to_string(x + 1);
}

int Quoridor_move::canonical_id() const {
    if (type == 'h') return MCTSB_Q_MOVE_HWALL(x, y);
    if (type == 'v') return MCTSB_Q_MOVE_VWALL(x, y);
    return MCTSB_Q_MOVE_STEP(x, y);
}

Quoridor_move *Quoridor_move::from_canonical_id(int id, char player) {
    if (id < 81) return new Quoridor_move((short)(id / 9), (short)(id % 9), player, ' ');
    if (id < 145) return new Quoridor_move((short)((id - 81) / 8), (short)((id - 81) % 8), player, 'h');
    return new Quoridor_move((short)((id - 145) / 8), (short)((id - 145) % 8), player, 'v');
}

Quoridor_state::Quoridor_state() : rollout_policy(MCTSB_POLICY_REF), good_moves_only(false), have_legal_mask(false) {   // Quoridor.cpp:17-23
    mctsb_qstate p;
    memset(&p, 0, sizeof p);
    p.wx = 0; p.wy = 4; p.bx = 8; p.by = 4; p.wwalls = 10; p.bwalls = 10;
    q_unpack(s, p);
}

Quoridor_state::Quoridor_state(const mctsb_qstate &packed, int rollout_policy) : rollout_policy(rollout_policy), good_moves_only(false), have_legal_mask(false) {
    if (!q_unpack(s, packed)) throw std::invalid_argument("Quoridor_state: packed record is not a position");
}

void Quoridor_state::pack(mctsb_qstate &out) const { q_pack(s, out); }
bool Quoridor_state::gpu_pack(void *dst) const { mctsb_qstate p; q_pack(s, p); memcpy(dst, &p, sizeof p); return true; }

char Quoridor_state::check_winner() const { int w = q_winner(s); return w == 1 ? 'W' : w == 2 ? 'B' : ' '; }
bool Quoridor_state::is_terminal() const { return q_winner(s) != 0; }

bool Quoridor_state::legal_move(const Quoridor_move *move) const {
    if (move == NULL) return false;
    if (move->player != 'W' && move->player != 'B') return false;
    if (move->player != whose_turn()) return false;            // legal_step / legal_wall both start with the turn test
    bool wall = move->type == 'h' || move->type == 'v';
    if (wall && (move->x < 0 || move->y < 0 || move->x >= 8 || move->y >= 8)) return false;
    if (!wall && (move->x < 0 || move->y < 0 || move->x >= 9 || move->y >= 9)) return false;
    return q_legal_move(s, move->canonical_id());
}

bool Quoridor_state::play_move(const Quoridor_move *move) {
    if (move == NULL || !legal_move(move)) {
        cout << "Invalid command: Illegal move: " << ((move != NULL) ? move->sprint() : "NULL") << endl << endl;
        return false;
    }
    q_play_id(s, move->canonical_id());
    return true;
}

int Quoridor_state::get_shortest_path(char player, const Quoridor_move *extra, short int posx, short int posy) const {
    int pl = player == 'W' ? 0 : player == 'B' ? 1 : -1;
    if (pl < 0) return -1;
    if (extra != NULL) {
        if (extra->type != 'h' && extra->type != 'v') return -1;
        // the reference re-checks cheap legality for extra->player (Quoridor.cpp:144)
        if (extra->player != whose_turn() || extra->x < 0 || extra->y < 0 || extra->x >= 8 || extra->y >= 8) return -1;
        uint64_t lh, lv;
        q_cheap_wall_masks(s, lh, lv);
        int a = extra->x * 8 + extra->y;
        if (!(((extra->type == 'h' ? lh : lv) >> a) & 1ull)) return -1;
        QState t = s;
        if (posx != -1 && posy != -1) t.cell[pl] = (int8_t)(posx * 9 + posy);
        return q_sp_with_wall(t, pl, 2 * a + (extra->type == 'v'), nullptr);
    }
    if (posx == -1 || posy == -1) return q_sp(s, pl, nullptr);
    if (posx < 0 || posx >= 9 || posy < 0 || posy >= 9) return -1;
    return q_flood_dist(s.openS, s.openE, posx * 9 + posy, pl, nullptr);
}

vector<MCTS_move *> Quoridor_state::get_legal_step_moves2(char p) const {
    vector<MCTS_move *> v;
    if (p != whose_turn()) return v;
    uint32_t m = q_step_mask(s);
    int me = s.cell[s.turn];
    for (int i = 0; i < 12; i++)
        if ((m >> i) & 1u) { int t = me + q_step_delta(i); v.push_back(new Quoridor_move((short)(t / 9), (short)(t % 9), p, ' ')); }
    return v;
}

forward_list<MCTS_move *> Quoridor_state::get_legal_step_moves(char p) const {
    forward_list<MCTS_move *> l;
    for (MCTS_move *m : get_legal_step_moves2(p)) l.push_front(m);
    return l;
}

Quoridor_move *Quoridor_state::get_best_step_move(char player) const {
    if (player != whose_turn()) return NULL;
    int t = q_best_step(s, q_step_mask(s), nullptr);
    if (t < 0) return NULL;
    return new Quoridor_move((short)(t / 9), (short)(t % 9), player, ' ');
}

// generate_all_moves as canonical ids (Quoridor.cpp:594-616): steps in reversed candidate order, then anchors row-major
// with the horizontal wall before the vertical one
int Quoridor_state::all_move_ids(uint8_t *ids) const {
    uint32_t steps; uint64_t lh, lv;
    if (have_legal_mask) {
        // the move-generation kernel already answered (gpu_set_legal): ids 81..144 = horizontal anchors,
        // 145..208 = vertical anchors (inverse of q_parts_to_mask); the step candidates need no search
        steps = q_step_mask(s);
        lh = (legal_mask[1] >> 17) | (legal_mask[2] << 47);
        lv = (legal_mask[2] >> 17) | (legal_mask[3] << 47);
    } else q_legal_parts(s, steps, lh, lv, nullptr);
    int n = 0;
    for (int i = 11; i >= 0; i--)
        if ((steps >> i) & 1u) ids[n++] = (uint8_t)(s.cell[s.turn] + q_step_delta(i));
    for (uint64_t any = lh | lv; any; any &= any - 1) {              // anchors that have at least one legal wall, ascending
        const int a = __builtin_ctzll(any);
        if ((lh >> a) & 1ull) ids[n++] = (uint8_t)(81 + a);
        if ((lv >> a) & 1ull) ids[n++] = (uint8_t)(145 + a);
    }
    return n;
}

queue<MCTS_move *> *Quoridor_state::generate_all_moves() const {
    queue<MCTS_move *> *Q = new queue<MCTS_move *>();
    uint8_t ids[MCTSB_Q_NUM_MOVES];
    const int n = all_move_ids(ids);
    const char p = whose_turn();
    for (int i = 0; i < n; i++) Q->push(Quoridor_move::from_canonical_id(ids[i], p));
    return Q;
}

bool Quoridor_state::actions_to_try_ids(std::vector<uint8_t> &ids) const {
    ids.resize(MCTSB_Q_NUM_MOVES);
    const int n = good_moves_only ? q_good_moves(s, ids.data(), nullptr) : all_move_ids(ids.data());
    ids.resize((size_t)n);
    return true;
}

MCTS_move *Quoridor_state::move_of_id(int id) const { return Quoridor_move::from_canonical_id(id, whose_turn()); }

// generate_good_moves, Quoridor.cpp:518-592: the shared rule code does the work (q_good_moves)
queue<MCTS_move *> *Quoridor_state::generate_good_moves() const {
    queue<MCTS_move *> *Q = new queue<MCTS_move *>();
    uint8_t ids[MCTSB_Q_GOOD_MAX];
    const int n = q_good_moves(s, ids, nullptr);
    const char p = whose_turn();
    for (int i = 0; i < n; i++) Q->push(Quoridor_move::from_canonical_id(ids[i], p));
    return Q;
}

// Quoridor.cpp:618-627: the shipped build defines TEST_ALL_MOVES, so generate_all_moves is the default here too
queue<MCTS_move *> *Quoridor_state::actions_to_try() const { return good_moves_only ? generate_good_moves() : generate_all_moves(); }   // TEST_ALL_MOVES build, Quoridor.cpp:618-627

MCTS_state *Quoridor_state::next_state(const MCTS_move *move) const {   // Quoridor.cpp:506-510
    Quoridor_state *n = new Quoridor_state(*this);
    n->have_legal_mask = false;
    n->play_move((const Quoridor_move *) move);
    return n;
}

// a move out of generate_all_moves(): play it without asking legal_move again (two flood fills for a wall)
MCTS_state *Quoridor_state::next_state_of_listed_move(const MCTS_move *move) const {
    Quoridor_state *n = new Quoridor_state(*this);
    n->have_legal_mask = false;
    q_play_id(n->s, ((const Quoridor_move *) move)->canonical_id());
    return n;
}

void Quoridor_state::gpu_set_legal(const uint64_t *mask4) {
    if (good_moves_only) return;             // the mask answers generate_all_moves only
    memcpy(legal_mask, mask4, sizeof legal_mask);
    have_legal_mask = true;
}

double Quoridor_state::evaluate() const { return q_eval(s, nullptr); }
bool Quoridor_state::force_playwall() const { return q_force_playwall(s, nullptr); }

// One playout of this position.  The reference runs pick_semirandom_move for up to 50 plies on the
// calling thread (Quoridor.cpp:809-855); here the playout runs on the GPU through the backend.
double Quoridor_state::rollout() const {
    mctsb_qstate p; q_pack(s, p);
    double sum = 0.0;
    default_rollout_backend()->rollout_batch(MCTSB_GAME_QUORIDOR, rollout_policy, &p, 1, 1, next_rollout_ids(1), &sum);
    return sum;
}

void Quoridor_state::print() const {
    mctsb_qstate p; q_pack(s, p);
    printf("\n      A   B   C   D   E   F   G   H   I        turn: %s   walls W:%d B:%d\n", s.turn ? "Black" : "White", s.nwalls[0], s.nwalls[1]);
    for (int x = 0; x < 9; x++) {
        printf("  %d  ", x + 1);
        for (int y = 0; y < 9; y++) {
            char c = (s.cell[1] == 9 * x + y) ? 'B' : (s.cell[0] == 9 * x + y) ? 'W' : '.';
            bool east_open = y < 8 && b81_test(s.openE, 9 * x + y);
            printf(" %c %c", c, (y < 8 && !east_open) ? '|' : ' ');
        }
        printf("\n     ");
        for (int y = 0; y < 9; y++) printf("%s ", (x < 8 && !b81_test(s.openS, 9 * x + y)) ? "---" : "   ");
        printf("\n");
    }
}
