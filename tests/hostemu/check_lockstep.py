"""Developer tool: quick bit-exact check of the lockstep kernel body (one-lane host emulation) against the
oracle.  Usage: PYTHONPATH=. python tests/hostemu/check_lockstep.py [n_rollouts].  The same checks run under
pytest in tests/test_hostemu.py."""
import numpy as np, time, sys
from oracle.reflib import *
from oracle.portlib import PortLib
from tests.hostemu.emulib import EmuLib
from tests.util import golden
P = PortLib(); E = EmuLib()
o = opening_state()
N=int(sys.argv[1]) if len(sys.argv)>1 else 3000
out, fin, wk = P.q_rollout_philox(o, 99, 0, N, finals=True, work=True)
eo, ef, sums, ctr = E.lockstep(o, N, 99, 0)
bad=np.nonzero((out['score']!=eo['score'])|(out['draws']!=eo['draws'])|(out['plies']!=eo['plies']))[0]
print("lockstep mismatches", len(bad), bad[:8], ctr/N)
assert out.tobytes()==eo.tobytes() and fin.tobytes()==ef.tobytes()
corpus = golden("corpus_q.npy")
rng=np.random.default_rng(5)
idx = rng.choice(len(corpus), 2000, replace=False)
streams = rng.integers(0,2**32,size=(2000,8192),dtype=np.uint32)
po,pf = P.q_rollout_streams(corpus[idx], streams)
eo, ef, _, _ = E.lockstep(corpus[idx], 1, 0, 0, streams=streams)
assert po.tobytes()==eo.tobytes() and pf.tobytes()==ef.tobytes(); print("replay corpus ok")
st = corpus[idx[:60]]
eo, ef, sums, _ = E.lockstep(st, 7, 123, 1000)
for j in range(60):
    o2,f2,_ = P.q_rollout_philox(st[j], 123, 1000+7*j, 7, finals=True)
    assert o2.tobytes()==eo[7*j:7*j+7].tobytes() and f2.tobytes()==ef[7*j:7*j+7].tobytes()
print("multi-leaf ok")
