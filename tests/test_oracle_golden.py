"""Pins the oracle (oracle/vectorised.py) to the UNMODIFIED reference.

tests/golden/ref_bfgs.json holds what the reference's own bfgs()/bfgs_wrapper()
returned here (oracle/make_golden.py), including every scipy ``minimize`` call it
made.  The vectorised restatement must reproduce them restart by restart from the
same x0: same basin, (almost) the same evaluation counts, same final answer.
"""
import numpy as np
import pytest
import sympy as sp

from conftest import make_cfg
from oracle import vectorised


def _case_inputs(case):
    X = np.zeros((1, case["n"], 10), dtype=case["dtype"])
    cols = np.asarray(case["X"], dtype=case["dtype"])
    X[0, :, :cols.shape[1]] = cols
    y = np.asarray(case["y"], dtype=case["dtype"])
    return X, y


def _ids(golden_cases):
    return [c["name"] for c in golden_cases]


def test_golden_file_is_from_the_reference(golden):
    assert golden["reference"].startswith("aidalee123/Vision-SR")
    assert len(golden["cases"]) >= 15
    assert golden["word2id"]["c"] == 3 and golden["word2id"]["x_10"] == 38


@pytest.mark.parametrize("idx", range(17))
def test_restatement_matches_reference(golden, test_data, idx):
    case = golden["cases"][idx]
    X, y = _case_inputs(case)
    cfg = make_cfg(case["R"], case["norm"], case["idx_remove"])
    calls = case["minimize_calls"]
    if case["raised"]:
        out = vectorised.bfgs_wrapper((case["tokens"], X, y, cfg, test_data))
        assert out[0] is None and np.isnan(out[1])
        assert case["wrapper"] == [None, None]
        return
    expr, k = vectorised.skeleton_string(case["tokens"], test_data.id2word)
    assert expr == case["skeleton"]
    x0 = [c["x0"] for c in calls[:case["R"]]] if k else None
    rec = vectorised.Recorder()
    best_expr, best_consts, best_loss, skel = vectorised.bfgs(
        case["tokens"], X, y, cfg, test_data, x0=x0, record=rec)
    assert skel == case["skeleton"]
    # per restart: same basin, same effort
    if k:
        same = 0
        for mine, ref in zip(rec.restarts, calls[:case["R"]]):
            tol = 1e-6 * max(1.0, abs(ref["fun"])) + 1e-9
            if abs(mine["fun"] - ref["fun"]) <= tol:
                same += 1
                assert abs(mine["nfev"] - ref["nfev"]) <= max(12, 0.2 * ref["nfev"]), (mine, ref)
        assert same == case["R"], f"{same}/{case['R']} restarts in the reference's basin"
    # final answer
    ref_loss = case["best_loss"]
    if ref_loss is None or not np.isfinite(ref_loss):
        assert not np.isfinite(best_loss) or best_loss >= 1e8
    else:
        assert abs(best_loss - ref_loss) <= 1e-6 * max(1.0, abs(ref_loss)) + 1e-9
        ref_c = np.asarray(case["best_consts"], dtype=float)
        mine_c = np.asarray([float(c) for c in best_consts], dtype=float)
        assert mine_c.shape == ref_c.shape
        if ref_loss < 1e-6 and k:   # identifiable fit: constants agree
            assert np.all(np.abs(mine_c - ref_c) <= 1e-4 * np.maximum(1.0, np.abs(ref_c)))
        # the printed winner evaluates to the same function
        xs = sp.symbols("x_1:11")
        f_ref = sp.lambdify(xs, sp.sympify(case["best_expr_str"]), modules=vectorised.MODULES)
        f_me = sp.lambdify(xs, sp.sympify(best_expr), modules=vectorised.MODULES)
        cols = [X[0, :, j].astype(float) for j in range(10)]
        with np.errstate(all="ignore"):
            a = np.broadcast_to(f_ref(*cols), y.shape)
            b = np.broadcast_to(f_me(*cols), y.shape)
        if ref_loss < 1e-6:
            assert np.allclose(a, b, rtol=1e-4, atol=1e-5, equal_nan=True)
