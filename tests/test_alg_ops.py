"""The roofline numerators of bench.py are counted, not asserted: the oracle's work counters (BFS levels, cheap wall
tests, plies of the REFERENCE algorithm) over the sample bench.ALG_SAMPLES names must reproduce the literals."""
import pytest

import bench


@pytest.mark.parametrize("workload", ["opening", "corpus", "opening_uni"])
def test_algorithmic_work_literals_are_what_the_oracle_counts(workload):
    assert bench.recount(workload) == bench.ALG_SAMPLES[workload]


def test_alg_ops_formula():
    c = bench.ALG_SAMPLES["opening"]
    assert bench.ALG_OPS_PER_ROLLOUT == (c["flood_steps"] * 24 + c["cheap_wall_tests"] * 12 + c["plies"] * 200) / 4000
    assert abs(bench.ALG_OPS_PER_ROLLOUT - 379410.4) < 5.0          # the figure DESIGN.md and round 1 quote
