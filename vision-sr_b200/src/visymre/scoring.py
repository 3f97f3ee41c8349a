"""Driver-side scoring on the device.

Every driver of the reference re-lambdifies the winner after each ``fitfunc`` call and scores
it on the FULL train / test set (``scripts/Feynman_test.py:81-97``; the same block in
``Low-dimensional_benchmark_test.py``, ``SRSD_test.py``, ``Black-box_test.py`` ...)::

    pre_expr = sp.sympify(output['best_bfgs_preds'][0])
    vars_ = vu.get_variable_names(str(pre_expr))
    func = sp.lambdify(vars_, pre_expr, modules="numpy")
    X_dict = {v: X_full[:, i] for i, v in enumerate(vars_)}
    y_pre = np.nan_to_num(func(**X_dict)...)            # nan -> 0, +-inf -> +-DBL_MAX
    r2 = r2_score(y_full, y_pre)

Here the expression is compiled once to bytecode and one ``vsr_score`` launch sweeps the set
(SURVEY 8f, row 2).  Two of the reference's behaviours are kept on purpose:

* the i-th USED variable is paired with the i-th COLUMN (``enumerate(vars_)``), not with the
  column of its own index: ``x_1 + x_3`` reads columns 0 and 1.  ``by_rank=False`` pairs
  ``x_j`` with column j-1 instead;
* ``r2_score``'s defaults: a constant target gives 1.0 for a perfect and 0.0 for an imperfect
  prediction, and a non-finite score is reported as 0.0.
"""
import re

import numpy as np
import sympy as sp
import torch

from .engine import fitter
from .engine.compiler import compile_sympy

_VARS = [f"x_{i}" for i in range(1, 11)]


def get_variable_names(expr_str):
    """scripts/visymre_utils.py:38-40."""
    names = re.findall(r"x_\d+", str(expr_str))
    return sorted(set(names), key=lambda v: int(v.split("_")[1]))


def compile_for_scoring(expr, by_rank=True):
    """Bytecode of ``expr`` (no fitted constants) reading the columns the driver would."""
    e = sp.sympify(expr)
    if by_rank:
        used = get_variable_names(str(e))
        # two steps so that x_3 -> x_2 and x_2 -> x_1 cannot collide
        tmp = {sp.Symbol(v): sp.Symbol(f"__col{i}") for i, v in enumerate(used)}
        e = e.xreplace({s: t for s, t in tmp.items()})
        e = e.xreplace({sp.Symbol(f"__col{i}"): sp.Symbol(_VARS[i]) for i in range(len(used))})
    return compile_sympy(e, 0, _VARS)


def mse_and_r2(expr, X, y, engine=None, dtype=None, by_rank=True):
    """(mse, r2) of ``expr`` on (X [N, d], y [N]) with the drivers' rule; one device sweep."""
    X = torch.as_tensor(X)
    y = torch.as_tensor(y).reshape(-1)
    if X.dim() == 3:
        X = X[0]
    eng = engine if engine is not None else fitter.get_engine(X.device if X.is_cuda else None)
    dt = dtype if dtype is not None else (fitter.F32 if X.dtype == torch.float32 else fitter.F64)
    prog = compile_for_scoring(expr, by_rank=by_rank)
    eng.set_points(X, y, dtypes=(dt,), n_vars=max(1, prog.var_mask.bit_length()))
    eng.set_programs([prog])
    mse = float(eng.score([0], dtype=dt).item())
    yd = y.to(eng.device, torch.float64)
    ss_tot = float(((yd - yd.mean()) ** 2).sum().item())
    ss_res = mse * y.shape[0]
    if ss_tot == 0.0:                      # sklearn: constant target
        r2 = 1.0 if ss_res == 0.0 else 0.0
    else:
        r2 = 1.0 - ss_res / ss_tot
        if not np.isfinite(r2):
            r2 = 0.0                       # sklearn force_finite=True
    return mse, r2


def r2_on_device(expr, X, y, engine=None, dtype=None, by_rank=True):
    return mse_and_r2(expr, X, y, engine=engine, dtype=dtype, by_rank=by_rank)[1]
