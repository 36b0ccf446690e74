#!/usr/bin/env python
"""match_vs_stock.py — the B200 agent against the reference's own Quoridor program, over its text protocol.

The opponent is oracle/_ref/quoridor_stock: examples/Quoridor/main.cpp + mcts.cpp + JobScheduler.cpp + Quoridor.cpp
exactly as shipped (`make -C oracle stock`), i.e. MAXITER 20 000 iterations x 4 RolloutJobs on its pthread pool,
at most 15 s per move (main.cpp:6-7,129-146).  This script speaks the REPL's commands (main.cpp:39-205):
`autoprint` once, then `genmove` for the stock side and `m <col><row>` / `w <h|v> <col><row>` for ours; moves come
back in Quoridor_move::sprint notation (Quoridor.h:24).  Our side is the C++ host tree of this repo with every
rollout on the GPU.  One JSON line: winner and length per game, seconds per move of both sides.

Test infrastructure: it runs the compiled reference as a subprocess; nothing in the product imports it.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
MOVE_RE = re.compile(r"^(White|Black) (moves to|places horizontal wall at|places vertical wall at) ([A-I])([1-9])\s*$", re.M)


def move_id_of(text):
    """last move line of a REPL answer -> canonical move id (include/mctsb.h)"""
    found = MOVE_RE.findall(text)
    if not found:
        return None
    _, what, col, row = found[-1]
    x, y = int(row) - 1, ord(col) - ord("A")
    return 9 * x + y if what == "moves to" else (81 if "horizontal" in what else 145) + 8 * x + y


def command_of(move_id):
    if move_id < 81:
        return "m %s%d" % (chr(ord("A") + move_id % 9), move_id // 9 + 1)
    a = move_id - (81 if move_id < 145 else 145)
    return "w %s %s%d" % ("h" if move_id < 145 else "v", chr(ord("A") + a % 8), a // 8 + 1)


class StockRepl:
    def __init__(self, exe):
        self.p = subprocess.Popen([exe], stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, bufsize=0)
        self._until_prompt()
        self.ask("autoprint")

    def _until_prompt(self):
        buf = ""
        while not buf.endswith("move:\n> "):
            ch = self.p.stdout.read(1)
            if ch == "":
                raise RuntimeError("the stock program closed its output:\n" + buf[-400:])
            buf += ch
        return buf

    def ask(self, line):
        self.p.stdin.write(line + "\n")
        self.p.stdin.flush()
        return self._until_prompt()

    def close(self):
        try:
            self.p.stdin.write("q\n")
            self.p.stdin.flush()
            self.p.wait(timeout=10)
        except Exception:
            self.p.kill()


class DeviceTreePlayer:
    """the tree resident in HBM (mctsb_tree_*) behind the three calls the match makes"""

    def __init__(self, m, args, seed):
        self.be = m.RolloutBackend(0, seed)
        self.tree = m.DeviceTree(self.be, m.quoridor_opening(), args.rollouts_per_leaf, args.leaves_per_flush, max_nodes=1 << 26)
        self.tree.set_overlap(True)
        self.R, self.next_id = args.rollouts_per_leaf, 0

    def grow(self, iters):
        self.tree.grow(iters, self.next_id)
        self.next_id += iters * self.R

    def best_move(self):
        return self.tree.best_move()

    def advance(self, mv):
        self.tree.advance(mv)
        s = self.tree.root_state()
        return int(s["wx"]) == 8 or int(s["bx"]) == 0

    def close(self):
        self.tree.close(); self.be.close()


def play(args, host, m, Session, ours_is_white, game_no):
    stock = StockRepl(args.stock)
    if args.device_tree:
        tree = DeviceTreePlayer(m, args, args.seed + game_no)
    else:
        tree = Session(host, m.GAME_QUORIDOR, m.quoridor_opening(), rollouts_per_leaf=args.rollouts_per_leaf,
                       leaves_per_flush=args.leaves_per_flush, device=0, seed=args.seed + game_no)
        tree.set_overlap(2)
        tree.set_gpu_expand(True)
    moves, t_ours, t_stock, winner = [], [], [], None
    white_to_move = True
    while winner is None and len(moves) < args.max_moves:
        if white_to_move == ours_is_white:
            t0 = time.perf_counter()
            tree.grow(args.iters)
            mv = tree.best_move()
            t_ours.append(time.perf_counter() - t0)
            ans = stock.ask(command_of(mv))
            assert move_id_of(ans) == mv, "the stock program refused our move %d:\n%s" % (mv, ans[-300:])
        else:
            t0 = time.perf_counter()
            ans = stock.ask("genmove")
            t_stock.append(time.perf_counter() - t0)
            mv = move_id_of(ans)
            assert mv is not None, "no move in the stock program's answer:\n" + ans[-400:]
        moves.append(mv)
        terminal = tree.advance(mv)
        white_to_move = not white_to_move
        w = stock.ask("winner")
        if "TRUE" in w:
            winner = "White" if "TRUE W" in w else "Black"
            assert terminal, "the stock program sees a winner, our tree does not"
    tree.close()
    stock.close()
    return {"ours": "White" if ours_is_white else "Black", "winner": winner, "ours_won": winner == ("White" if ours_is_white else "Black"),
            "plies": len(moves), "our_seconds_per_move": sum(t_ours) / max(1, len(t_ours)),
            "stock_seconds_per_move": sum(t_stock) / max(1, len(t_stock)), "move_ids": moves}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--games", type=int, default=2, help="colours alternate, ours is White in game 0")
    ap.add_argument("--iters", type=int, default=8192, help="our tree iterations per move")
    ap.add_argument("--rollouts-per-leaf", type=int, default=1024)
    ap.add_argument("--leaves-per-flush", type=int, default=64)
    ap.add_argument("--max-moves", type=int, default=300)
    ap.add_argument("--device-tree", type=int, default=0, help="1 = our tree lives on the GPU (mctsb_tree_*), one flush in flight")
    ap.add_argument("--seed", type=lambda x: int(x, 0), default=0x2026101900000001)
    ap.add_argument("--stock", default=os.path.join(ROOT, "oracle", "_ref", "quoridor_stock"))
    args = ap.parse_args()
    import montecarlotreesearch_b200 as m
    from montecarlotreesearch_b200.hosttree import Session, load
    host = load()
    games = [play(args, host, m, Session, g % 2 == 0, g) for g in range(args.games)]
    print(json.dumps({
        "what": "B200 %s (%d iterations x %d rollouts per move, flushes of %d leaves) vs the "
                "reference's stock Quoridor program (20 000 iterations x 4 rollouts or 15 s per move) over its own text protocol"
                % ("tree in HBM (mctsb_tree_*, one flush in flight)" if args.device_tree else "host tree (2 flushes in flight, GPU-prepared expansion)",
                   args.iters, args.rollouts_per_leaf, args.leaves_per_flush),
        "games": games, "ours_won": sum(1 for g in games if g["ours_won"]), "played": len(games)}))


if __name__ == "__main__":
    main()
