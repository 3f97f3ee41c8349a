"""Skeleton compiler: bytecode vs sympy.lambdify (what the reference evaluates,
bfgs.py:104/:128), through the numpy VM (oracle/vm.py) and the g++ host build of the
CUDA interpreter (oracle/hostsim).  fp64 1e-12 / fp32 1e-5 relative."""
import ctypes

import numpy as np
import pytest
import sympy as sp

from oracle import vectorised, vm
from src.visymre.engine import isa
from src.visymre.engine.compiler import CompileError, compile_skeleton

VARS = [f"x_{i}" for i in range(1, 11)]

SKELETONS = [
    ("c0 + c1*cos(c2*x_1**3)*sin(c3*x_2)", 4),
    ("c0*x_1 + c1*x_1**2 + c2*x_1**3", 3),
    ("cos(c0*x_1)*sin(c1*x_1)", 2),
    ("c0*exp(-(x_1-c1)**2/(2*c2**2))/sqrt(2*pi*c2**2) + x_2/(x_3*x_4)", 3),
    ("c0*exp(c1*x_1) + ln(c2 + sqrt(x_2))", 3),
    ("c0*x_1**c1", 2),
    ("c0**x_1 + 2**x_2", 1),
    ("Abs(c0*x_1 - x_2)**1.5 + asin(x_3/2)*tan(c1*x_4)", 2),
    ("(x_1 - x_2)/(x_3 - c0) - 1/(c1 + x_4**2)", 2),
    ("sin(x_1) + x_1*x_2", 0),
    ("x_1/3 + c0/x_2**2 + x_3**(-1/2) + x_4**(3/2)", 1),
    ("sqrt(c0*x_1)*ln(x_2)/exp(x_3) - pi*E", 1),
    ("((c0 + x_1)*(c1 + x_2))*((c2 + x_3)*(c3 + x_4)) + ((x_5+c4)*(x_6+c5))/((x_7+c6)*(x_8+c7))", 8),
    ("x_9*x_10 + c0", 1),
    ("atan(c0*x_1) + tan(x_2)**2 - cos(x_3)**3", 1),
    ("c0", 1),
    ("3", 0),
]


def _points(n, dtype, seed=0):
    rng = np.random.RandomState(seed)
    return rng.uniform(0.3, 2.0, size=(n, 10)).astype(dtype)


def _lambdified(expr, k):
    cs = [sp.Symbol(f"c{i}") for i in range(k)]
    xs = [sp.Symbol(v) for v in VARS]
    return sp.lambdify(cs + xs, sp.sympify(expr), modules=vectorised.MODULES)


@pytest.mark.parametrize("expr,k", SKELETONS)
def test_vm_matches_lambdify_fp64(expr, k):
    prog = compile_skeleton(expr, k, VARS)
    X = _points(257, np.float64)
    c = np.random.RandomState(1).uniform(0.5, 1.5, size=k)
    want = np.broadcast_to(_lambdified(expr, k)(*c, *X.T), (X.shape[0],)).astype(float)
    got = vm.run(prog, X, c, isa)
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=0)
    assert prog.stack_depth <= isa.MAX_STACK
    assert prog.insns[-1] & 0xFF == isa.OP["VSR_END"]


def _hostsim_values(hostsim, prog, X, c):
    N = X.shape[0]
    Xc = np.ascontiguousarray(X.T)  # column-major [d][N]
    out = np.empty(N, dtype=np.float64)
    cc = np.ascontiguousarray(np.asarray(c, dtype=np.float64))
    if cc.size == 0:
        cc = np.zeros(1)
    rc = hostsim.hostsim_values(
        prog.insns.ctypes.data_as(ctypes.c_void_p), prog.imms.ctypes.data_as(ctypes.c_void_p),
        cc.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(prog.k),
        Xc.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(N),
        ctypes.c_int(isa.DTYPE["VSR_F64"] if X.dtype == np.float64 else isa.DTYPE["VSR_F32"]),
        out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return out


@pytest.mark.parametrize("expr,k", SKELETONS)
@pytest.mark.parametrize("dtype,rtol", [(np.float64, 1e-12), (np.float32, 1e-5)])
def test_interpreter_core_matches_lambdify(hostsim, expr, k, dtype, rtol):
    prog = compile_skeleton(expr, k, VARS)
    X = _points(300, dtype, seed=2)
    c = np.random.RandomState(3).uniform(0.5, 1.5, size=k)
    f = _lambdified(expr, k)
    want = np.broadcast_to(f(*c, *X.astype(np.float64).T), (X.shape[0],)).astype(float)
    got = _hostsim_values(hostsim, prog, X, c)
    if dtype == np.float64:
        np.testing.assert_allclose(got, want, rtol=rtol, atol=0)
        return
    # fp32: 1e-5 relative "away from poles and cancellations" -- i.e. at the points where
    # numpy's own single-precision evaluation of the lambdified function is well
    # conditioned (agrees with the fp64 value to 2e-6).
    with np.errstate(all="ignore"):
        ref32 = np.broadcast_to(f(*c.astype(np.float32), *X.T), (X.shape[0],)).astype(float)
    ok = np.abs(ref32 - want) <= 2e-6 * np.abs(want)
    assert ok.mean() >= 0.9
    np.testing.assert_allclose(got[ok], want[ok], rtol=rtol, atol=0)


def _hostsim_loss_grad(hostsim, prog, X, y, c, want_grad=True):
    N = X.shape[0]
    Xc = np.ascontiguousarray(X.T)
    loss = ctypes.c_double()
    grad = np.zeros(max(prog.k, 1))
    cc = np.ascontiguousarray(np.asarray(c, dtype=np.float64))
    rc = hostsim.hostsim_loss_grad(
        prog.insns.ctypes.data_as(ctypes.c_void_p), prog.imms.ctypes.data_as(ctypes.c_void_p),
        cc.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(prog.k),
        Xc.ctypes.data_as(ctypes.c_void_p), np.ascontiguousarray(y).ctypes.data_as(ctypes.c_void_p),
        ctypes.c_long(N), ctypes.c_int(0 if X.dtype == np.float64 else 1), ctypes.c_int(int(want_grad)),
        ctypes.byref(loss), grad.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return loss.value, grad[:prog.k]


@pytest.mark.parametrize("expr,k", [s for s in SKELETONS if s[1] > 0])
def test_dual_gradient_matches_sympy_diff(hostsim, expr, k):
    """d(mean r^2)/dc_j from the tangents vs the symbolic derivative: 1e-10 relative."""
    prog = compile_skeleton(expr, k, VARS)
    X = _points(200, np.float64, seed=4)
    rng = np.random.RandomState(5)
    c = rng.uniform(0.5, 1.5, size=k)
    y = rng.normal(size=X.shape[0])
    # real symbols so that sympy can differentiate Abs
    cs = [sp.Symbol(f"c{i}", real=True) for i in range(k)]
    xs = [sp.Symbol(v, real=True) for v in VARS]
    e = sp.sympify(expr, locals={str(s): s for s in cs + xs})
    f = sp.lambdify(cs + xs, e, modules=vectorised.MODULES)
    r = np.broadcast_to(f(*c, *X.T), y.shape) - y
    loss, grad = _hostsim_loss_grad(hostsim, prog, X, y, c)
    assert loss == pytest.approx(np.mean(r * r), rel=1e-12)
    for j in range(k):
        dj = sp.lambdify(cs + xs, sp.diff(e, cs[j]), modules="numpy")
        want = np.mean(2 * r * np.broadcast_to(dj(*c, *X.T), y.shape))
        assert grad[j] == pytest.approx(want, rel=1e-10, abs=1e-13)


def test_numpy_semantics_for_domain_errors(hostsim):
    """nan for domain violations, inf for overflow (bfgs.py:38-40 via numpy)."""
    X = np.zeros((4, 10))
    X[:, 0] = [-1.0, 0.0, 2.0, 800.0]
    for expr, want in [("sqrt(x_1)", [np.nan, 0.0, np.sqrt(2.0), np.sqrt(800.0)]),
                       ("ln(x_1)", [np.nan, -np.inf, np.log(2.0), np.log(800.0)]),
                       ("exp(x_1)", [np.exp(-1.0), 1.0, np.exp(2.0), np.inf]),
                       ("x_1**0.5", [np.nan, 0.0, 2 ** 0.5, 800 ** 0.5]),
                       ("asin(x_1)", [np.arcsin(-1.0), 0.0, np.nan, np.nan]),
                       ("1/x_1", [-1.0, np.inf, 0.5, 1 / 800.0])]:
        prog = compile_skeleton(expr, 0, VARS)
        got = _hostsim_values(hostsim, prog, X, [])
        np.testing.assert_allclose(got, want, rtol=1e-15, equal_nan=True)


def test_zero_tangent_survives_infinite_derivative(hostsim):
    """sqrt'(0) = inf must not poison a structurally or numerically zero tangent."""
    prog = compile_skeleton("c0*sqrt(x_1) + sqrt(c1*x_1)", 2, VARS)
    X = np.zeros((3, 10))
    X[:, 0] = [0.0, 1.0, 4.0]
    y = np.zeros(3)
    loss, grad = _hostsim_loss_grad(hostsim, prog, X, y, [1.0, 1.0])
    assert np.isfinite(loss) and np.all(np.isfinite(grad))


def test_rejects_what_the_reference_cannot_evaluate():
    for bad in ["c0*I + x_1", "zoo*x_1", "gamma(x_1)", "sqrt(-2)*x_1"]:
        with pytest.raises(CompileError):
            compile_skeleton(bad, 1, VARS)
    with pytest.raises(CompileError):
        compile_skeleton("+".join(f"c{i}*x_1" for i in range(40)), 40, VARS)


def test_sethi_ullman_keeps_the_stack_small():
    # a left-deep and a right-deep chain need no stack at all
    assert compile_skeleton("sin(cos(exp(c0*x_1 + c1)))*x_2 + x_3", 2, VARS).stack_depth == 0
    # a balanced product of sums needs log2 depth
    bal = "((c0+x_1)*(c1+x_2))*((c2+x_3)*(c3+x_4))"
    assert compile_skeleton(bal, 4, VARS).stack_depth <= 2


def test_isa_constants_come_from_the_c_header():
    text = open(isa.ISA_HEADER).read()
    for name, val in isa.OP.items():
        assert f"{name} = {val}" in text or name == "VSR_OP_COUNT" or f"{name} =" in text
    assert isa.encode(isa.OP["VSR_MUL"], 2, 5, 0x3, 0x20) == (6 | 2 << 8 | 5 << 16 | 3 << 32 | 0x20 << 48)
    assert isa.decode(isa.encode(26, 0, (-3) & 0xFFFF, 1, 0))[2] == 0xFFFD
