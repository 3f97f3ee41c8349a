"""GPU parity: the interpreter kernel (through the C ABI) against the oracle.

fp64: 1e-12 relative; fp32: 1e-5 relative at well-conditioned points; gradients vs sympy
diff 1e-10 relative.  Covers ragged / tiny / split point counts.
"""
import numpy as np
import pytest
import sympy as sp
import torch

from oracle import vectorised
from src.visymre.engine import fitter, isa
from src.visymre.engine.compiler import compile_skeleton
from test_compiler import SKELETONS, VARS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    return fitter.Engine("cuda:0")


def _data(n, seed=0, dtype=np.float64):
    rng = np.random.RandomState(seed)
    X = rng.uniform(0.3, 2.0, size=(n, 10)).astype(dtype)
    y = rng.normal(size=n).astype(dtype)
    return X, y


def _oracle_loss_grad(expr, k, X, y, c):
    cs = [sp.Symbol(f"c{i}", real=True) for i in range(k)]
    xs = [sp.Symbol(v, real=True) for v in VARS]
    e = sp.sympify(expr, locals={str(s): s for s in cs + xs})
    f = sp.lambdify(cs + xs, e, modules=vectorised.MODULES)
    with np.errstate(all="ignore"):
        r = np.broadcast_to(f(*c, *X.T.astype(np.float64)), y.shape) - y
        loss = np.mean(r * r)
        grad = []
        for j in range(k):
            dj = sp.lambdify(cs + xs, sp.diff(e, cs[j]), modules="numpy")
            grad.append(np.mean(2 * r * np.broadcast_to(dj(*c, *X.T.astype(np.float64)), y.shape)))
    return loss, np.asarray(grad)


@pytest.mark.parametrize("n", [1, 31, 257, 5000])
def test_loss_and_gradient_fp64(eng, n):
    X, y = _data(n, seed=n)
    eng.set_points(X, y, dtypes=(fitter.F64,), n_vars=10)
    progs = [compile_skeleton(e, k, VARS) for e, k in SKELETONS]
    eng.set_programs(progs)
    kmax = max(p.k for p in progs)
    rng = np.random.RandomState(7)
    consts = np.zeros((len(progs), kmax))
    for i, p in enumerate(progs):
        consts[i, :p.k] = rng.uniform(0.5, 1.5, size=p.k)
    loss, grad = eng.eval(np.arange(len(progs)), consts, dtype=fitter.F64, grad=True)
    loss, grad = loss.cpu().numpy(), grad.cpu().numpy()
    loss_v, _ = eng.eval(np.arange(len(progs)), consts, dtype=fitter.F64, grad=False)
    np.testing.assert_array_equal(np.isnan(loss), np.isnan(loss_v.cpu().numpy()))
    for i, (expr, k) in enumerate(SKELETONS):
        want_l, want_g = _oracle_loss_grad(expr, k, X, y, consts[i, :k])
        assert loss[i] == pytest.approx(want_l, rel=1e-12), expr
        assert loss_v[i].item() == pytest.approx(want_l, rel=1e-12), expr
        for j in range(k):
            assert grad[i, j] == pytest.approx(want_g[j], rel=1e-10, abs=1e-13), (expr, j)


def test_loss_fp32(eng):
    X, y = _data(4096, seed=3, dtype=np.float32)
    eng.set_points(X, y, dtypes=(fitter.F32, fitter.F64), n_vars=10)
    progs = [compile_skeleton(e, k, VARS) for e, k in SKELETONS]
    eng.set_programs(progs)
    kmax = max(p.k for p in progs)
    consts = np.zeros((len(progs), kmax))
    rng = np.random.RandomState(9)
    for i, p in enumerate(progs):
        consts[i, :p.k] = rng.uniform(0.5, 1.5, size=p.k)
    l32, _ = eng.eval(np.arange(len(progs)), consts, dtype=fitter.F32)
    l64, _ = eng.eval(np.arange(len(progs)), consts, dtype=fitter.F64)
    l32, l64 = l32.cpu().numpy(), l64.cpu().numpy()
    for i, (expr, k) in enumerate(SKELETONS):
        want, _ = _oracle_loss_grad(expr, k, X.astype(np.float64), y.astype(np.float64), consts[i, :k])
        assert l64[i] == pytest.approx(want, rel=1e-12), expr
        # mean of squares of well-scaled residuals: fp32 evaluation error enters once
        assert l32[i] == pytest.approx(want, rel=2e-4), expr


def test_split_sweeps_equal_single_sweep(eng):
    """few pairs x many points: grid.y splits the points; result equals the oracle."""
    X, y = _data(200_003, seed=5)
    eng.set_points(X, y, dtypes=(fitter.F64,), n_vars=10)
    expr, k = SKELETONS[0]
    eng.set_programs([compile_skeleton(expr, k, VARS)])
    c = np.array([[0.7, 1.1, 0.9, 1.3]])
    loss, grad = eng.eval([0], c, grad=True)
    want_l, want_g = _oracle_loss_grad(expr, k, X, y, c[0])
    assert loss.item() == pytest.approx(want_l, rel=1e-12)
    np.testing.assert_allclose(grad.cpu().numpy()[0], want_g, rtol=1e-10)


def test_non_finite_losses_are_reported_raw(eng):
    X = np.zeros((8, 10))
    X[:, 0] = np.linspace(-1, 1, 8)
    y = np.zeros(8)
    eng.set_points(X, y, dtypes=(fitter.F64,), n_vars=1)
    eng.set_programs([compile_skeleton("ln(x_1) + c0", 1, VARS),
                      compile_skeleton("exp(1000*x_1**2) + c0", 1, VARS),
                      compile_skeleton("c0*x_1", 1, VARS)])
    loss, _ = eng.eval([0, 1, 2], np.ones((3, 1)))
    loss = loss.cpu().numpy()
    assert np.isnan(loss[0]) and np.isinf(loss[1]) and np.isfinite(loss[2])


def test_errors_are_loud(eng):
    from src.visymre.engine.native import VsrError
    X, y = _data(16)
    eng.set_points(X, y, n_vars=10)
    eng.set_programs([compile_skeleton("c0*x_1", 1, VARS)])
    with pytest.raises(VsrError):
        eng.eval([5], np.ones((1, 1)))           # program out of range
    bad = compile_skeleton("c0*x_1", 1, VARS)
    bad.insns = bad.insns.copy()
    bad.insns[0] = isa.encode(isa.OP["VSR_LOAD"], isa.SRC["VSR_SRC_CONST"], 3)  # slot 3 of 1
    with pytest.raises(VsrError):
        eng.set_programs([bad])


@pytest.mark.parametrize("n,dtype", [(4999, np.float64), (20011, np.float64), (20011, np.float32)])
def test_shared_tile_kernel_equals_the_per_pair_kernel(monkeypatch, n, dtype):
    """vsr_eval / vsr_score over shared tiles (eval_tile_kernel: the points staged once per CTA, every
    pair run over them -- the path for N >= 5e5) against the per-pair kernel on the same inputs:
    values, gradients and nan_to_num scores; ragged chunk ends; fp32 and fp64."""
    X, y = _data(n, seed=3, dtype=dtype)
    X[:, 4:] = 0.0
    progs = [compile_skeleton(e, k, VARS) for e, k in SKELETONS]
    progs = [p for p in progs if p.var_mask < (1 << 4)] * 3
    assert len(progs) >= 8
    kmax = max(p.k for p in progs)
    rng = np.random.RandomState(11)
    consts = np.zeros((len(progs), kmax))
    for i, p in enumerate(progs):
        consts[i, :p.k] = rng.uniform(0.5, 1.5, size=p.k)
    dt = fitter.F32 if dtype == np.float32 else fitter.F64
    outs = []
    for min_points in ("1000000000", "1000"):
        monkeypatch.setenv("VSR_TILE_MIN_POINTS", min_points)      # read by vsr_create
        e = fitter.Engine("cuda:0")
        e.set_points(X, y, dtypes=(dt,), n_vars=4)
        e.set_programs(progs)
        n0 = e.launches
        loss, grad = e.eval(np.arange(len(progs)), consts, dtype=dt, grad=True)
        score = e.score(np.arange(len(progs)), consts, dtype=dt)
        outs.append((loss.cpu().numpy(), grad.cpu().numpy(), score.cpu().numpy(), e.launches - n0))
        e.close()
    (l0, g0, s0, _), (l1, g1, s1, _) = outs
    tol = 1e-12 if dtype == np.float64 else 2e-6
    np.testing.assert_array_equal(np.isnan(l0), np.isnan(l1))
    ok = np.isfinite(l0)
    np.testing.assert_allclose(l1[ok], l0[ok], rtol=tol)
    np.testing.assert_allclose(s1, s0, rtol=tol)
    fin = np.isfinite(g0) & np.isfinite(g1)
    np.testing.assert_allclose(g1[fin], g0[fin], rtol=100 * tol, atol=100 * tol * np.abs(g0[fin]).max())
    np.testing.assert_array_equal(np.isfinite(g0), np.isfinite(g1))
