"""Seeded synthetic workloads for the refinement path.

The Drive checkpoint and the external datasets are not available offline, so beams are
synthesised: for each row of an in-repo equation table the ground-truth formula gives
the points and its constant-stripped skeleton (with fitted constants injected) is
candidate 0; the other candidates of the beam are seeded structural mutations of it,
kept only when the prefix tree is complete (``analyze_prefix_tree_context``) and the
skeleton compiles.  Starting points are ``RandomState(1_000_000 + 1000*e + j).randn(R, k)
* 10`` -- the reference's distribution (``bfgs.py:103``), made reproducible.

Everything here is host-side data preparation; no numeric fitting happens in this
module.
"""
import json
import os
from dataclasses import dataclass, field
from types import SimpleNamespace

import numpy as np
import sympy as sp

from ..architectures import data as tok
from ..architectures.bfgs import skeleton_string
from ..architectures.refine import BINARY_NAMES, UNARY_NAMES, analyze_prefix_tree_context
from ..dataset.generator import Generator
from ..engine.compiler import CompileError, compile_sympy

_HERE = os.path.dirname(os.path.abspath(__file__))
TABLES = os.path.join(_HERE, "tables.json")
_NP = {"sqrt": np.sqrt, "exp": np.exp, "log": np.log, "ln": np.log, "sin": np.sin, "cos": np.cos,
       "tan": np.tan, "asin": np.arcsin, "arcsin": np.arcsin, "arccos": np.arccos,
       "acos": np.arccos, "atan": np.arctan, "tanh": np.tanh, "Abs": np.abs, "pi": np.pi}


def load_tables():
    with open(TABLES) as fh:
        return json.load(fh)


def make_test_data(tables=None):
    """A DatasetDetails-like record with the shipped 47-word vocabulary."""
    t = tables or load_tables()
    w2i = dict(t["word2id"])
    i2w = {v: k for k, v in w2i.items()}
    i2w[3] = "constant"  # what fitfunc2 sets before fitting (model.py:452)
    return SimpleNamespace(word2id=w2i, id2word=i2w, total_variables=list(t["total_variables"]),
                           total_coefficients=[], una_ops=list(UNARY_NAMES),
                           bin_ops=list(BINARY_NAMES), rewrite_functions=[])


def make_cfg(n_restarts=10, beam_size=64, **bfgs_extra):
    """An object shaped like ``cfg.inference`` of scripts/config.yaml:111-125."""
    b = SimpleNamespace(activated=True, n_restarts=n_restarts,
                        add_coefficients_if_not_existing=False, normalization_o=False,
                        idx_remove=False, normalization_type="MSE", stop_time=1e9, **bfgs_extra)
    return SimpleNamespace(beam_size=beam_size, device="cuda", no_c_in_pow=False, bfgs=b)


@dataclass
class Beam:
    """One fitfunc2 call's worth of work: points of one equation and its candidates."""
    name: str
    X: np.ndarray                 # [N, 10] zero-padded
    y: np.ndarray                 # [N]
    tokens: list                  # C token-id lists (S ... F)
    x0: list                      # C arrays [R, k_j]
    truth: str = ""
    programs: list = field(default_factory=list)   # filled by compile_beam


# ---- skeleton construction ---------------------------------------------------------------
def _strip_constants(expr):
    """Floats and integers beyond +-9 become the placeholder (data.py:160-169)."""
    return tok.constants_to_placeholder(expr, symbol="c")[0]


def _inject_constants(expr, rng, p_term=0.9, p_unary=0.6, p_root=0.8):
    """Fitted constants where a beam candidate typically has them: a factor on each
    additive term, a factor inside transcendental arguments, an offset at the root."""
    c = sp.Symbol("c", real=True, nonzero=True)
    trans = (sp.sin, sp.cos, sp.tan, sp.exp, sp.log, sp.asin)

    def inside(e):
        if isinstance(e, trans) and rng.rand() < p_unary:
            return e.func(c * inside_args(e.args[0]))
        if e.args:
            return e.func(*[inside(a) for a in e.args])
        return e

    def inside_args(e):
        return inside(e)
    e = inside(expr)
    terms = sp.Add.make_args(e)
    e = sp.Add(*[(c * t if rng.rand() < p_term else t) for t in terms])
    if rng.rand() < p_root:
        e = e + c
    return e


def _to_words(expr):
    words = tok.sanitize_prefix(Generator.sympy_to_prefix(expr))
    return words


def _valid(words, td):
    w2i = td.word2id
    try:
        ids = [w2i[w] for w in words]
    except KeyError:
        return None
    a1 = {w2i[n] for n in UNARY_NAMES}
    a2 = {w2i[n] for n in BINARY_NAMES}
    valency, _ = analyze_prefix_tree_context(ids, a1, a2, set(), w2i["pow"], None, w2i["S"])
    if valency != 0 or len(ids) + 2 > 100:
        return None
    return [w2i["S"]] + ids + [w2i["F"]]


def _subtree_end(words, i):
    need = 1
    while need:
        need += Generator.OPERATORS.get(words[i], 0) - 1
        i += 1
    return i


def _mutate(words, rng, n_vars):
    """One structural mutation of a prefix word list."""
    words = list(words)
    kind = rng.randint(5)
    idx = rng.randint(len(words))
    un = [w for w in UNARY_NAMES if w not in ("asin",)]
    bi = ["add", "sub", "mul", "div"]
    w = words[idx]
    if kind == 0 and w in UNARY_NAMES:       # swap a unary operator
        words[idx] = un[rng.randint(len(un))]
    elif kind == 0 and w in bi:              # swap a binary operator
        words[idx] = bi[rng.randint(len(bi))]
    elif kind == 1 and w not in Generator.OPERATORS:   # replace a leaf
        words[idx] = f"x_{rng.randint(n_vars) + 1}" if rng.rand() < 0.7 else "c"
    elif kind == 2:                          # wrap a subtree: mul c (.)
        words[idx:idx] = ["mul", "c"]
    elif kind == 3:                          # wrap a subtree: add c (.)
        words[idx:idx] = ["add", "c"]
    elif kind == 4 and w in ("mul", "add") and idx + 1 < len(words) and words[idx + 1] == "c":
        del words[idx:idx + 2]               # drop a `mul c` / `add c`
    else:                                    # wrap in a unary operator
        words[idx:idx] = [un[rng.randint(len(un))]]
    return words


def _points(formula, variables, n, rng, log_uniform=False):
    """visymre_utils.py:220-236 style sampling: uniform per variable, non-finite y dropped."""
    syms = [sp.Symbol(f"x_{i + 1}") for i in range(len(variables))]
    f = sp.lambdify(syms, sp.sympify(formula, locals=dict(_SYM_LOCALS)), modules=[_NP, "numpy"])
    X = np.zeros((0, 10))
    y = np.zeros(0)
    tries = 0
    while X.shape[0] < n and tries < 20:
        m = int((n - X.shape[0]) * 1.3) + 16
        cols = []
        for v in variables:
            lo, hi = float(v["low"]), float(v["high"])
            if log_uniform:
                a = np.log10(max(abs(lo), 1e-3))
                cols.append(np.sign(lo if lo != 0 else 1.0) * 10 ** rng.uniform(a, a + 2, m))
            else:
                cols.append(rng.uniform(lo, hi, m))
        with np.errstate(all="ignore"):
            yy = np.broadcast_to(np.asarray(f(*cols), dtype=np.float64), (m,))
        ok = np.isfinite(yy)
        XX = np.zeros((int(ok.sum()), 10))
        for j, c in enumerate(cols):
            XX[:, j] = c[ok]
        X = np.concatenate([X, XX])
        y = np.concatenate([y, yy[ok]])
        tries += 1
    if X.shape[0] < n:
        raise ValueError("too few valid samples")
    return X[:n], y[:n]


_SYM_LOCALS = {f"x_{i}": sp.Symbol(f"x_{i}") for i in range(1, 11)}
_SYM_LOCALS.update({"ln": sp.log, "arcsin": sp.asin, "arccos": sp.acos})


def build_beam(e, name, formula, variables, n_points, n_cand, n_restarts, td, log_uniform=False,
               kmax=8):
    """Points + candidates + starting points of equation number ``e``; None if the ground
    truth does not tokenise with the shipped vocabulary (arcsin/tanh/... rows)."""
    rng = np.random.RandomState(e)
    truth = sp.sympify(formula, locals=dict(_SYM_LOCALS))
    try:
        base = _strip_constants(truth)
        cfg = make_cfg(n_restarts)
        tokens, x0 = [], []
        seen = set()
        # j = 0: ground-truth skeleton with injected constants
        for attempt in range(20):
            words = _to_words(_inject_constants(base, np.random.RandomState(1000 * e + attempt)))
            ids = _valid(words, td)
            if ids is not None:
                break
        else:
            return None
        X, y = _points(formula, variables, n_points, rng, log_uniform)
    except Exception:  # noqa: BLE001 -- unknown operator / word not in the vocabulary
        return None
    base_words = words
    j = 0
    pool = [base_words]
    guard = 0
    while len(tokens) < n_cand and guard < 200 * n_cand:
        guard += 1
        if j == 0:
            cand_words = base_words
        else:
            mrng = np.random.RandomState(1000 * e + j + 7919 * guard)
            cand_words = pool[mrng.randint(len(pool))]
            for _ in range(1 + mrng.randint(3)):
                cand_words = _mutate(cand_words, mrng, len(variables))
        ids = _valid(cand_words, td)
        key = tuple(ids) if ids else None
        if ids is None or key in seen:
            j += (j == 0)
            continue
        try:
            expr, k = skeleton_string(ids, cfg, td)
            if k > kmax:
                raise CompileError("too many constants for the workload")
            compile_sympy(sp.sympify(expr), k, td.total_variables)
        except Exception:  # noqa: BLE001
            j += (j == 0)
            continue
        seen.add(key)
        tokens.append(ids)
        x0.append(np.random.RandomState(1_000_000 + 1000 * e + len(tokens) - 1).randn(n_restarts, k) * 10)
        if len(pool) < 16:
            pool.append(cand_words)
        j += 1
    if len(tokens) < n_cand:
        return None
    return Beam(name=name, X=X, y=y, tokens=tokens, x0=x0, truth=str(formula))


def feynman_beams(n_points=10_000, n_cand=64, n_restarts=10, limit=None, bonus=False,
                  log_uniform=False):
    """BASELINE config 2: the Feynman table rows that tokenise, one beam each."""
    t = load_tables()
    td = make_test_data(t)
    beams = []
    for e, row in enumerate(t["feynman"]):
        if row["name"].startswith("test_") and not bonus:
            continue
        formula = row["replaced"] or row["formula"]
        b = build_beam(e, row["name"], formula, row["variables"], n_points, n_cand, n_restarts, td,
                       log_uniform=log_uniform)
        if b is not None:
            beams.append(b)
        if limit and len(beams) >= limit:
            break
    return beams, td


def low_beam(e, row, n_points, n_cand, n_restarts, td):
    """One low_benchmarks.csv row -> beam (the table gives one range for all variables)."""
    variables = [dict(low=row["range"][0], high=row["range"][1]) for _ in range(max(1, row["n_vars"]))]
    nv = max([int(s.name.split("_")[1]) for s in sp.sympify(row["formula"], locals=dict(_SYM_LOCALS)).free_symbols
              if s.name.startswith("x_")] + [1])
    while len(variables) < nv:
        variables.append(dict(variables[0]))
    return build_beam(10_000 + e, row["name"], row["formula"], variables, n_points, n_cand, n_restarts, td)


def low_beams(n_points=500, n_cand=16, n_restarts=10, limit=None):
    """BASELINE config 1: low_benchmarks.csv rows (Nguyen-style, 1-2 variables)."""
    t = load_tables()
    td = make_test_data(t)
    beams = []
    for e, row in enumerate(t["low"]):
        b = low_beam(e, row, n_points, n_cand, n_restarts, td)
        if b is not None:
            beams.append(b)
        if limit and len(beams) >= limit:
            break
    return beams, td


def ode_beams(n_points=10_000, n_cand=128, n_restarts=32, limit=None):
    """BASELINE config 3: ode.xlsx rows; the file has no ranges, x_1, x_2 ~ U(0.1, 5)."""
    t = load_tables()
    td = make_test_data(t)
    beams = []
    for e, row in enumerate(t["ode"]):
        variables = [dict(low=0.1, high=5.0), dict(low=0.1, high=5.0)]
        b = build_beam(20_000 + e, row["name"], row["formula"], variables, n_points, n_cand,
                       n_restarts, td)
        if b is not None:
            beams.append(b)
        if limit and len(beams) >= limit:
            break
    return beams, td


def blackbox_beam(n_points, n_cand=1024, n_restarts=10, seed=0, dtype=np.float32):
    """BASELINE config 5: X ~ N(0,1) in 3 variables, y = 1.5 x1 sin(0.7 x2) + 0.3 x3^2 + noise."""
    t = load_tables()
    td = make_test_data(t)
    rng = np.random.RandomState(seed)
    formula = "1.5*x_1*sin(0.7*x_2) + 0.3*x_3**2"
    variables = [dict(low=-2.0, high=2.0)] * 3
    b = build_beam(30_000 + seed, "blackbox", formula, variables, 64, n_cand, n_restarts, td, kmax=6)
    X = np.zeros((n_points, 10), dtype=dtype)
    X[:, :3] = rng.standard_normal((n_points, 3)).astype(dtype)
    y = (1.5 * X[:, 0] * np.sin(0.7 * X[:, 1]) + 0.3 * X[:, 2] ** 2
         + rng.normal(scale=0.1, size=n_points)).astype(dtype)
    b.X, b.y = X, y
    return b, td


def compile_beam(beam, td, cfg=None):
    """Token lists -> compiled programs (host side of bfgs_batch, done once per beam)."""
    cfg = cfg or make_cfg()
    progs = []
    for ids in beam.tokens:
        expr, k = skeleton_string(ids, cfg, td)
        progs.append(compile_sympy(sp.sympify(expr), k, td.total_variables))
    beam.programs = progs
    return progs
