"""GPU parity: the fit kernel and the drop-in bfgs()/refine_hypotheses() against the
reference's golden vectors and the oracle, through the C ABI.

Tolerances: a restart is "in the same basin" when
|dloss| <= 1e-6*max(1,|loss|) + 1e-9; constants of identifiable fits agree to
1e-4*max(1,|c|).  FD-gradient mode must reproduce the reference restart by restart; the
default dual-gradient mode must reach the same best-of-R loss.
"""
import numpy as np
import pytest
import sympy as sp
import torch

from conftest import make_cfg
from oracle import vectorised
from src.visymre.architectures import bfgs as vbfgs
from src.visymre.architectures import model as vmodel
from src.visymre.engine import fitter, isa
from src.visymre.engine.compiler import compile_skeleton
from test_bfgs_core import FITTED, SKELS, _numpy_objective

pytestmark = pytest.mark.gpu
VARS = [f"x_{i}" for i in range(1, 11)]
FD, DUAL = isa.GRAD_MODE["VSR_GRAD_FD"], isa.GRAD_MODE["VSR_GRAD_DUAL"]


@pytest.fixture(scope="module")
def eng():
    return fitter.Engine("cuda:0")


def _case_tensors(case, device="cuda:0"):
    X = np.zeros((1, case["n"], 10), dtype=case["dtype"])
    cols = np.asarray(case["X"], dtype=case["dtype"])
    X[0, :, :cols.shape[1]] = cols
    y = np.asarray(case["y"], dtype=case["dtype"])
    return torch.tensor(X, device=device), torch.tensor(y, device=device)


def _same(a, b):
    return abs(a - b) <= 1e-6 * max(1.0, abs(b)) + 1e-9


@pytest.mark.parametrize("name", FITTED)
def test_fit_kernel_fd_mode_matches_reference_restarts(eng, golden, test_data, name):
    case = next(c for c in golden["cases"] if c["name"] == name)
    X, y = _case_tensors(case)
    expr, k = vectorised.skeleton_string(case["tokens"], test_data.id2word)
    eng.set_points(X[0].double(), y.double(), dtypes=(fitter.F64,), n_vars=10)
    eng.set_programs([compile_skeleton(expr, k, VARS)])
    calls = case["minimize_calls"][:case["R"]]
    x0 = np.array([c["x0"] for c in calls])
    scale = 1.0 / float(y.double().mean()) if case["norm"] == "NMSE" else 1.0
    opts = fitter.default_opts(grad_mode=FD, loss_scale=scale)
    res = eng.fit([0] * len(calls), list(range(len(calls))), x0, opts)
    loss, info = res.loss.cpu().numpy(), res.info.cpu().numpy()
    cx, lx = res.consts.cpu().numpy(), res.lastx.cpu().numpy()
    for r, ref in enumerate(calls):
        assert _same(loss[r], ref["fun"]), (name, r, loss[r], ref["fun"])
        assert abs(info[r, 2] - ref["nfev"]) <= max(12, 0.25 * ref["nfev"])
        assert abs(info[r, 1] - ref["nit"]) <= max(3, 0.25 * ref["nit"])
        assert info[r, 0] == ref["status"] or ref["fun"] < 1e-10
        if ref["fun"] < 1e-8:
            c_ref = np.asarray(ref["res_x"])
            assert np.all(np.abs(cx[r, :k] - c_ref) <= 1e-4 * np.maximum(1, np.abs(c_ref)))
        assert np.max(np.abs(lx[r, :k] - cx[r, :k])) <= 2e-8 * max(1.0, np.max(np.abs(cx[r, :k])))


@pytest.mark.parametrize("mode", ["fd", "dual"])
def test_drop_in_bfgs_matches_the_reference_outputs(golden, test_data, mode):
    """bfgs() end to end on every golden case: loss, constants, skeleton string, printed winner."""
    n_ok = n_tot = 0
    for case in golden["cases"]:
        X, y = _case_tensors(case)
        cfg = make_cfg(case["R"], case["norm"], case["idx_remove"], grad_mode=mode)
        if case["raised"]:
            out = vmodel.bfgs_wrapper((case["tokens"], X, y, cfg, test_data))
            assert out[0] is None and np.isnan(out[1]) and out[2] == case["tokens"]
            continue
        k = case["skeleton"].count("c")  # not used for logic, only to slice x0
        calls = case["minimize_calls"][:case["R"]]
        x0 = np.array([c["x0"] for c in calls]) if calls else None
        expr_str, consts, loss, skel = vbfgs.bfgs(case["tokens"], X, y, cfg, test_data, x0=x0)
        assert skel == case["skeleton"]
        ref_loss = case["best_loss"]
        n_tot += 1
        if ref_loss is None or not np.isfinite(ref_loss):
            assert not np.isfinite(loss) or loss >= 1e8, (case["name"], loss)
            n_ok += 1
            continue
        good = _same(float(loss), ref_loss)
        if mode == "fd":
            assert good, (case["name"], float(loss), ref_loss)
        n_ok += good
        if good and ref_loss < 1e-6:
            ref_c = np.asarray(case["best_consts"], dtype=float)
            got_c = np.asarray([float(c) for c in consts], dtype=float)
            assert got_c.shape == ref_c.shape
            if len(ref_c):
                assert np.all(np.abs(got_c - ref_c) <= 1e-4 * np.maximum(1.0, np.abs(ref_c))), case["name"]
            xs = sp.symbols("x_1:11")
            f_ref = sp.lambdify(xs, sp.sympify(case["best_expr_str"]), modules=vectorised.MODULES)
            f_me = sp.lambdify(xs, sp.sympify(expr_str), modules=vectorised.MODULES)
            cols = [X[0, :, j].double().cpu().numpy() for j in range(10)]
            with np.errstate(all="ignore"):
                a = np.broadcast_to(f_ref(*cols), (case["n"],))
                b = np.broadcast_to(f_me(*cols), (case["n"],))
            assert np.allclose(a, b, rtol=1e-4, atol=1e-5, equal_nan=True), case["name"]
    assert n_ok >= 0.95 * n_tot, f"{mode}: {n_ok}/{n_tot} candidates reach the reference's best loss"


@pytest.mark.parametrize("expr,k,fn", SKELS)
def test_fit_kernel_against_scipy_at_1e4_points(eng, expr, k, fn):
    """Oracle = scipy on the vectorised loss, N = 10 000, 16 restarts, both gradient modes."""
    rng = np.random.RandomState(21)
    N, R = 10_000, 16
    X = np.zeros((N, 10))
    X[:, 0] = rng.uniform(0.5, 2.5, N)
    X[:, 1] = rng.uniform(0.5, 3.0, N)
    y = fn(X[:, 0], X[:, 1])
    eng.set_points(X, y, dtypes=(fitter.F64,), n_vars=10)
    eng.set_programs([compile_skeleton(expr, k, VARS)])
    loss, grad = _numpy_objective(expr, k, X, y)
    from scipy.optimize import minimize
    x0 = np.stack([np.random.RandomState(500 + r).randn(k) * 10 for r in range(R)])
    for mode, jac in ((FD, None), (DUAL, grad)):
        res = eng.fit([0] * R, list(range(R)), x0, fitter.default_opts(grad_mode=mode))
        got = res.loss.cpu().numpy()
        info = res.info.cpu().numpy()
        same = 0
        for r in range(R):
            ref = minimize(loss, x0[r], jac=jac, method="BFGS")
            if _same(got[r], ref.fun):
                same += 1
                assert abs(info[r, 1] - ref.nit) <= max(3, 0.25 * ref.nit)
        assert same >= R - 2, f"mode {mode}: {same}/{R} restarts in scipy's basin"
        # final score = plain MSE at the last evaluated point
        fm = res.final_mse.cpu().numpy()
        lx = res.lastx.cpu().numpy()
        for r in range(0, R, 5):
            assert fm[r] == pytest.approx(loss(lx[r, :k]), rel=1e-9) or not np.isfinite(fm[r])


def test_idempotent_and_closed_form(eng):
    """Size-independent properties at N = 100 000: an affine skeleton has a closed-form
    least-squares answer, and restarting from the optimum stops at once."""
    rng = np.random.RandomState(33)
    N = 100_000
    X = np.zeros((N, 10))
    X[:, 0] = rng.uniform(-3, 3, N)
    y = 1.25 - 0.5 * X[:, 0] + rng.normal(scale=0.1, size=N)
    eng.set_points(X, y, dtypes=(fitter.F64,), n_vars=1)
    eng.set_programs([compile_skeleton("c0 + c1*x_1", 2, VARS)])
    A = np.stack([np.ones(N), X[:, 0]], axis=1)
    sol, *_ = np.linalg.lstsq(A, y, rcond=None)
    res = eng.fit([0, 0], [0, 1], np.array([[9.0, -7.0], [-3.0, 4.0]]))
    c = res.consts.cpu().numpy()
    np.testing.assert_allclose(c[0], sol, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(c[1], sol, rtol=1e-6, atol=1e-7)
    again = eng.fit([0], [0], c[:1])
    assert again.info.cpu().numpy()[0, 1] == 0  # nit
    assert again.loss.item() == pytest.approx(res.loss.cpu().numpy()[0], rel=1e-12)


def test_fp32_sweeps_reach_the_fp64_answer(eng):
    rng = np.random.RandomState(44)
    N = 20_000
    X = np.zeros((N, 10), dtype=np.float32)
    X[:, 0] = rng.uniform(-2, 2, N)
    y = (0.5 - 1.25 * np.cos(1.5 * X[:, 0])).astype(np.float32)
    eng.set_points(X, y, dtypes=(fitter.F32, fitter.F64), n_vars=1)
    eng.set_programs([compile_skeleton("c0 + c1*cos(c2*x_1)", 3, VARS)])
    x0 = np.array([[0.3, -1.0, 1.4]] * 2)
    r64 = eng.fit([0], [0], x0[:1], fitter.default_opts(eval_dtype=fitter.F64, score_dtype=fitter.F32))
    r32 = eng.fit([0], [0], x0[:1], fitter.default_opts(eval_dtype=fitter.F32, score_dtype=fitter.F32))
    np.testing.assert_allclose(r32.consts.cpu().numpy()[0], r64.consts.cpu().numpy()[0], rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(r64.consts.cpu().numpy()[0], [0.5, -1.25, 1.5], rtol=1e-4)


def test_refine_hypotheses_dict_and_error_conventions(golden, test_data):
    w = golden["word2id"]
    tok = lambda s: [w["S"]] + [w[t] for t in s.split()] + [w["F"]]
    rng = np.random.RandomState(2)
    X = torch.zeros((1, 120, 10), dtype=torch.float64)
    X[0, :, 0] = torch.tensor(rng.uniform(-2, 2, 120))
    y = (0.75 + 2.5 * torch.sin(X[0, :, 0])).reshape(1, -1, 1)
    cfg = make_cfg(4)
    cfg.beam_size, cfg.no_c_in_pow = 4, False
    hyps = [(-0.1, tok("add c mul c sin x_1")),      # the right skeleton
            (-0.2, tok("add c mul c x_1")),          # a worse one
            (-0.3, tok("add c")),                    # incomplete tree: filtered out
            (-0.4, torch.tensor(tok("mul c cos x_1") + [0, 0]))]  # tensor with padding
    np.random.seed(0)
    out = vmodel.refine_hypotheses(hyps, X.cuda(), y.cuda(), cfg, test_data)
    assert set(out) == {"pred_target", "all_bfgs_preds", "all_bfgs_loss", "best_bfgs_preds",
                        "best_bfgs_loss", "best_token"}
    assert len(out["all_bfgs_preds"]) == 3 and len(out["all_bfgs_loss"]) == 3
    assert out["best_bfgs_loss"][0] < 1e-10
    best = sp.sympify(out["best_bfgs_preds"][0])
    assert abs(float(best.subs("x_1", 0.3)) - (0.75 + 2.5 * np.sin(0.3))) < 1e-5
    assert out["best_token"][0] == hyps[0][1]
    assert test_data.id2word[3] == "constant"
    # nothing valid and the fallback fails too -> [None] / [nan]
    out = vmodel.refine_hypotheses([(-1.0, tok("add c"))], X.cuda(), y.cuda(), cfg, test_data)
    assert out["best_bfgs_preds"] == [None] and np.isnan(out["best_bfgs_loss"][0])
    assert out["all_bfgs_preds"] == []


def test_fit_host_equals_fit_device(eng):
    rng = np.random.RandomState(8)
    X = np.zeros((500, 10))
    X[:, 0] = rng.uniform(-1, 1, 500)
    y = X[:, 0] + X[:, 0] ** 2 + X[:, 0] ** 3
    eng.set_points(X, y, dtypes=(fitter.F64,), n_vars=1)
    eng.set_programs([compile_skeleton("c0*x_1 + c1*x_1**2 + c2*x_1**3", 3, VARS)])
    x0 = rng.randn(6, 3) * 10
    dev = eng.fit([0] * 6, list(range(6)), x0)
    host = eng.fit_host([0] * 6, list(range(6)), x0)
    np.testing.assert_array_equal(host["consts"], dev.consts.cpu().numpy())
    np.testing.assert_array_equal(host["loss"], dev.loss.cpu().numpy())
    np.testing.assert_array_equal(host["info"], dev.info.cpu().numpy())
    assert np.all(host["final_mse"] < 1e-10)


def test_stop_time_zero_flattens_the_objective_like_timedfun(eng):
    """TimedFun (bfgs.py:23-36): the clock starts at the first loss call; once stop_time has
    passed every later call is the 1e6 penalty.  stop_time = 0 makes that deterministic."""
    from conftest import make_cfg, make_test_data
    rng = np.random.RandomState(3)
    X = np.zeros((1, 80, 10))
    X[0, :, 0] = rng.uniform(-2, 2, 80)
    y = 1.0 + 2.0 * X[0, :, 0]
    expr, k = "c0 + c1*x_1", 2
    eng.set_points(X[0], y, dtypes=(fitter.F64,), n_vars=1)
    eng.set_programs([compile_skeleton(expr, k, VARS)])
    x0 = np.array([[3.0, -4.0], [0.5, 0.25]])
    res = eng.fit([0, 0], [0, 1], x0, fitter.default_opts(grad_mode=FD, stop_time=0.0))
    loss, info = res.loss.cpu().numpy(), res.info.cpu().numpy()
    # the oracle with the same rule
    import sympy as sp
    from scipy.optimize import minimize
    f = sp.lambdify(sp.symbols("c0 c1 x_1"), sp.sympify(expr), modules=vectorised.MODULES)
    for r in range(2):
        calls = {"n": 0}

        def timed(c):
            calls["n"] += 1
            if calls["n"] > 1:
                return 1e6
            return float(np.mean((f(*c, X[0, :, 0]) - y) ** 2))
        ref = minimize(timed, x0[r], method="BFGS")
        assert loss[r] == pytest.approx(ref.fun, rel=1e-12)
        assert info[r, 1] == ref.nit and info[r, 2] == ref.nfev


def test_duplicate_skeletons_are_fitted_once_when_asked(golden, test_data):
    """with cfg.bfgs.collapse_duplicates, candidates that compile to the same
    bytecode share one fit (from the first one's starting points) and keep their own skeleton."""
    w2i = golden["word2id"]
    cands = ["add c mul c x_1", "add mul c x_1 c", "add c mul c x_1", "mul c sin x_1"]
    toks = [[w2i["S"]] + [w2i[w] for w in c.split()] + [w2i["F"]] for c in cands]
    rng = np.random.RandomState(12)
    N, R = 300, 3
    X = np.zeros((1, N, 10))
    X[0, :, 0] = rng.uniform(-2, 2, N)
    y = 0.4 + 1.7 * X[0, :, 0]
    Xt, yt = torch.tensor(X, device="cuda:0"), torch.tensor(y, device="cuda:0")
    x0 = [rng.randn(R, 2) * 3, rng.randn(R, 2) * 3, rng.randn(R, 2) * 3, rng.randn(R, 1) * 3]
    cfg = make_cfg(R, grad_mode="dual")
    plain = vbfgs.bfgs_batch(toks, Xt, yt, cfg, test_data, x0=x0)
    cfg.bfgs.collapse_duplicates = True
    coll = vbfgs.bfgs_batch(toks, Xt, yt, cfg, test_data, x0=x0)
    # the first of a group and candidates without a twin are untouched
    assert coll[0][0] == plain[0][0] and float(coll[0][2]) == float(plain[0][2])
    assert coll[3][0] == plain[3][0] and float(coll[3][2]) == float(plain[3][2])
    # the exact twin shares the first one's fit; every candidate keeps its own skeleton string
    assert coll[2][0] == coll[0][0] and float(coll[2][2]) == float(coll[0][2])
    assert [c[3] for c in coll] == [p[3] for p in plain]
    # all variants of the affine skeleton reach the exact fit either way
    for c in coll[:3]:
        assert float(c[2]) < 1e-12


def _tok(test_data, words):
    w2i = test_data.word2id
    return [w2i["S"]] + [w2i[w] for w in words.split()] + [w2i["F"]]


def test_all_pruned_singular_skeleton_keeps_the_unpruned_fit_and_the_beam(test_data):
    """`c0 + x_1/c1` with both constants below the prune threshold substitutes to
    zoo*x_1; the reference scores that prune 1e9 (bfgs.py:196-202) and keeps the fit.  One such
    candidate must not fail the other candidates of the batch."""
    rng = np.random.RandomState(3)
    X = np.zeros((1, 200, 10))
    X[0, :, 0] = rng.uniform(1, 2, 200)
    y = np.zeros(200)                           # exact fit needs c0 = 0 and a huge c1
    toks = [_tok(test_data, "add c div x_1 c"), _tok(test_data, "add c mul c x_1")]
    # restart starts at (1e-4, 1e-4): loss is finite there, gradient pushes c1 up ... start INSIDE the
    # prune band and cap the iterations by giving the exact zero-gradient point for the line
    cfg = make_cfg(1, grad_mode="fd", prune_threshold=1e9)   # everything counts as "small"
    outs = vbfgs.bfgs_batch(toks, torch.tensor(X, device="cuda:0"), torch.tensor(y, device="cuda:0"), cfg,
                            test_data, x0=[np.array([[0.5, 3.0]]), np.array([[0.5, 0.25]])])
    assert not any(isinstance(o, Exception) for o in outs), outs
    ref = [vectorised.bfgs(t, X, y, cfg, test_data, x0=s) for t, s in
           zip(toks, [np.array([[0.5, 3.0]]), np.array([[0.5, 0.25]])])]
    for o, r in zip(outs, ref):
        assert _same(float(o[2]), float(r[2])), (o, r)


def test_nmse_scale_uses_the_full_y_when_rows_are_removed(test_data):
    """with idx_remove the objective is divided by the mean of the FULL y
    (bfgs.py:85-90); every final score is 1e9 then (shape mismatch), restart 0 is picked and its
    last evaluated point depends on the objective's scale."""
    rng = np.random.RandomState(4)
    X = np.zeros((1, 60, 10))
    X[0, :, 0] = rng.uniform(-2, 2, 60)
    X[0, :7, 0] = 500.0                          # rows idx_remove drops
    y = 3.0 + 2.0 * X[0, :, 0]
    tok = _tok(test_data, "add c mul c x_1")
    cfg = make_cfg(2, norm="NMSE", idx_remove=True, grad_mode="fd")
    x0 = np.array([[1.0, 1.0], [-2.0, 0.5]])
    got = vbfgs.bfgs(tok, torch.tensor(X, device="cuda:0"), torch.tensor(y, device="cuda:0"), cfg, test_data, x0=x0)
    ref = vectorised.bfgs(tok, X, y, cfg, test_data, x0=x0)
    assert float(got[2]) == float(ref[2]) == 1e9
    # both stop where |grad|_inf <= gtol: the constants agree to the width of that valley floor
    np.testing.assert_allclose(np.asarray(got[1], dtype=float), np.asarray(ref[1], dtype=float), rtol=1e-3, atol=1e-5)
