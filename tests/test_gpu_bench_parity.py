"""Parity on the bench workload itself (the beams bench.py times: BASELINE config 2, C = 64, R = 10,
N = 10 000): the best-of-R loss of the drop-in against the oracle (scipy BFGS over numpy columns,
pinned to the unmodified reference) from the same starting points, in the FD-gradient parity mode
and in the default dual mode.  Every candidate is classified (tests/_parity.py):

  artefact  the oracle's loss is negative / complex (the reference's complex sub-trees): both lose it
  fragile   a restart of the oracle ended at scipy's iteration cap or in a failed line search: where
            such a run stops depends on the rounding of a 10 000-term sum, so no re-run with another
            order of summation reproduces it -- reported, and bounded by "nobody materially worse"
  clean     every restart converged: the drop-in must land in the same basin (>= 95 % in
            FD mode), and NEVER be materially worse (a real divergence)
"""
import os

import pytest

from _bench_parity import statistic

pytestmark = pytest.mark.gpu
N_BEAMS = int(os.environ.get("VSR_PARITY_BEAMS", "20"))


def test_best_of_restarts_agrees_with_the_oracle_on_bench_beams():
    tallies = statistic(nb=N_BEAMS, verbose=True)
    for mode, t in tallies.items():
        s = t.summary()
        n = len(t.rows)
        assert t.count("clean") >= 0.4 * n, (mode, s)
        # same basin for >= 95 % of the candidates whose restarts all converged
        assert t.rate("clean") >= (0.95 if mode == "fd" else 0.90), (mode, s)
        # real divergences: a clean candidate on which the drop-in is worse by more than the prune
        # tolerance of the reference (5 %, bfgs.py:144) -- none
        real = [r for r in t.rows if r["cls"] == "clean" and not r["ok"] and r["gap"] == r["gap"] and r["gap"] > 0.05]
        assert not real, (mode, real)
        # nobody materially worse: over ALL classes (the fragile ones included) at most 3.5 % of the
        # candidates lose more than 1e-3 (relative) against the oracle.  Measured on 24 beams = 1536
        # candidates: 34 (2.2 %, FD) / 41 (2.7 %, dual), every one of them fragile
        assert len(t.worse(1e-3)) <= 0.035 * n, (mode, s, len(t.worse(1e-3)))
        clean_worse = [r for r in t.worse(1e-3) if r["cls"] != "fragile"]
        if mode == "fd":      # the parity mode follows scipy's trajectories: every loss of ground is a fragile run
            assert not clean_worse, (mode, clean_worse)
        else:                 # exact gradients take other trajectories: a converged restart may land in another
            assert len(clean_worse) <= 0.005 * n, (mode, clean_worse)   # basin (measured: 1 of 1280, gap 7.6e-3)
        # artefacts and dropped candidates are lost on both sides
        assert t.rate("artefact") == 1.0 and t.rate("dropped") == 1.0, (mode, s)
