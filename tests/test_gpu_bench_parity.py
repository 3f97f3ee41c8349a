"""Parity on the bench workload itself (BASELINE config 2 beams, R = 10, N = 10 000): the
best-of-R loss of the drop-in against the oracle (scipy BFGS over numpy columns) from the same
starting points, in the FD-gradient parity mode and in the default dual mode.

At this size a tenth of the candidates have restarts that stop at scipy's iteration cap without
converging; which valley such a run drifts into depends on the rounding of a 10 000-term sum,
so neither the drop-in nor any re-run of scipy with another summation order reproduces them.
The test pins the measured rate with a margin and requires every
mismatch that is NOT of that kind to be absent: a candidate the oracle fits to < 1e-8 must be
fitted by the drop-in too."""
import pytest

from _bench_parity import statistic

pytestmark = pytest.mark.gpu


def test_best_of_restarts_agrees_with_the_oracle_on_bench_beams():
    rates, mism = statistic(nb=2, verbose=False)
    assert rates["fd"] >= 0.85 and rates["dual"] >= 0.85, rates
    for name, mode, got, ref in mism:
        if ref is not None and 0 <= ref < 1e-8:
            assert got is not None and got < 1e-6, (name, mode, got, ref)
