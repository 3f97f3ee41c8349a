"""Driver-side scoring on the device against the drivers' own lines
(scripts/Feynman_test.py:81-97): sympy lambdify over numpy, np.nan_to_num, sklearn r2_score."""
import numpy as np
import pytest
import sympy as sp
from sklearn.metrics import r2_score

from src.visymre import scoring

pytestmark = pytest.mark.gpu


def _driver_r2(expr, X, y):
    """The reference's block, verbatim in behaviour."""
    pre_expr = sp.sympify(expr)
    vars_ = scoring.get_variable_names(str(pre_expr))
    func = sp.lambdify(vars_, pre_expr, modules="numpy")
    with np.errstate(all="ignore"):
        y_pre = func(**{v: X[:, i] for i, v in enumerate(vars_)})
    y_pre = np.broadcast_to(y_pre, y.shape)
    y_pre = np.nan_to_num(y_pre.real if np.iscomplexobj(y_pre) else y_pre, nan=0.0)
    with np.errstate(all="ignore"):
        return r2_score(y, y_pre)


CASES = [
    "1.5*x_1*sin(0.7*x_2) + 0.3*x_3**2",
    "x_1 + x_3",                               # by-rank pairing: reads columns 0 and 1
    "exp(-x_2**2/2)/sqrt(2*pi)",
    "log(x_1) + sqrt(x_2)",                    # nan where x <= 0 -> counts as prediction 0
    "exp(1000*x_1)",                           # overflows to inf for x_1 > 0.71 -> DBL_MAX -> -inf score
    "0.25",                                    # no variables at all
    "x_4*x_5/(x_6 + 1.3) - tan(x_1)",
]


@pytest.mark.parametrize("n", [1000, 200_000])
@pytest.mark.parametrize("expr", CASES)
def test_r2_matches_the_driver(expr, n):
    rng = np.random.RandomState(len(expr) + n)
    X = rng.normal(size=(n, 10))
    y = 1.5 * X[:, 0] * np.sin(0.7 * X[:, 1]) + 0.3 * X[:, 2] ** 2 + rng.normal(scale=0.1, size=n)
    ref = _driver_r2(expr, X, y)
    got = scoring.r2_on_device(expr, X, y)
    if np.isfinite(ref) and abs(ref) < 1e12:
        assert got == pytest.approx(ref, rel=1e-9, abs=1e-9), (expr, got, ref)
    else:
        assert got == 0.0 or got == pytest.approx(ref, rel=1e-6)


def test_constant_target_and_true_index_pairing():
    rng = np.random.RandomState(1)
    X = rng.normal(size=(5000, 10))
    y = np.full(5000, 2.0)
    assert scoring.r2_on_device("2.0", X, y) == 1.0            # sklearn: perfect on a constant target
    assert scoring.r2_on_device("x_1", X, y) == 0.0            # imperfect on a constant target
    y2 = X[:, 0] + X[:, 2]
    assert scoring.r2_on_device("x_1 + x_3", X, y2, by_rank=False) == pytest.approx(1.0, abs=1e-12)
    assert scoring.r2_on_device("x_1 + x_3", X, y2) < 0.9      # the driver's pairing reads column 1
