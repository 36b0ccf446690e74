"""Round-2 additions to the golden fixtures, FROM THE COMPILED, UNMODIFIED REFERENCE like make_golden.py (kept apart
so that the round-1 files stay byte-identical):

  golden_q_goodmoves.npz      the reference's public generate_good_moves() (Quoridor.cpp:518-592) on the opening
                              position and on every state of corpus_q.npy: move ids in queue order, padded with -1
  golden_q_distribution64.npz free-running REF-policy scores of 64 states (the opening + 63 corpus states, SURVEY.md 8d
                              config 3) x 2000 Philox-keyed rollouts each, for the bit-exact and the 3-sigma tests
Run from the repo root:  python tests/golden/make_golden_r2.py
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.reflib import RefLib, opening_state  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
R = RefLib("replay")
SEED = 0x2026101900000001


def main():
    t0 = time.time()
    corpus = np.load(os.path.join(OUT, "corpus_q.npy"))
    states = [opening_state()] + [corpus[i] for i in range(len(corpus))]
    lists = [R.q_good_moves(s) for s in states]
    width = max(len(v) for v in lists)
    ids = np.full((len(lists), width), -1, dtype=np.int16)
    for j, v in enumerate(lists):
        ids[j, :len(v)] = v
    np.savez_compressed(os.path.join(OUT, "golden_q_goodmoves.npz"), index=np.arange(-1, len(corpus)), ids=ids,
                        count=np.array([len(v) for v in lists], dtype=np.int16))
    print("good moves", ids.shape, "%.0fs" % (time.time() - t0))

    didx = np.concatenate([[-1], np.arange(0, 4096, 65)[:63]])
    n = 2000
    dist = {"index": didx, "n": n, "first_id": np.uint64(1 << 33), "scores": np.zeros((len(didx), n))}
    for j, i in enumerate(didx):
        s = opening_state() if i < 0 else corpus[i]
        dist["scores"][j] = R.q_rollout_philox(s, SEED, 1 << 33, n, traced=False)["score"]
    np.savez_compressed(os.path.join(OUT, "golden_q_distribution64.npz"), seed=np.uint64(SEED), **dist)
    print("distribution64", "%.0fs" % (time.time() - t0))


if __name__ == "__main__":
    main()
