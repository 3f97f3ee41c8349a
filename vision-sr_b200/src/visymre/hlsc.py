"""Batched candidate evaluation for the HLSC reinforcement loop.

The reference scores every sampled sequence of an RL iteration with one serial
``bfgs.bfgs`` call each (``src/visymre/hlsc.py:392-443``, called 64 times per iteration at
``:647-648``), one restart in coarse mode, behind a per-sequence ``expression_cache``.  Here the
uncached sequences of the iteration go through ONE ``bfgs_batch`` launch; the per-sequence
result rules are the reference's:

* ``loss`` complex -> ``abs(loss) + 1e6``; nan / inf / None -> ``1e9``;
* the winner string must ``sympify(..., evaluate=False)``, else ``(1e9, None)``;
* any exception of the fit -> ``(1e9, None)``;
* results are cached by the token tuple, cache hits cost nothing.
"""
import numpy as np
import sympy as sp
import torch

from .architectures.bfgs import bfgs_batch


class _Restarts:
    """cfg.bfgs.n_restarts = 1 for the duration of a coarse evaluation (hlsc.py:402-404, :439-441)."""

    def __init__(self, cfg, coarse):
        self.cfg, self.coarse = cfg, coarse

    def __enter__(self):
        self.saved = self.cfg.bfgs.n_restarts
        if self.coarse:
            self.cfg.bfgs.n_restarts = 1

    def __exit__(self, *exc):
        if self.coarse:
            self.cfg.bfgs.n_restarts = self.saved


def _result_of(out):
    """One ``bfgs()`` outcome (4-tuple or the exception it raised) -> (loss_val, sympy_expr)."""
    if isinstance(out, Exception):
        return 1e9, None
    pred_str, _, loss, _ = out
    if loss is not None:
        if isinstance(loss, complex) or np.iscomplexobj(loss):
            loss = float(abs(loss)) + 1e6
        if np.isnan(loss) or np.isinf(loss):
            loss = 1e9
        else:
            loss = float(loss)
    else:
        loss = 1e9
    expr = None
    if isinstance(pred_str, str):
        try:
            expr = sp.sympify(pred_str, evaluate=False)
        except Exception:  # noqa: BLE001 -- the reference's bare except (hlsc.py:428-431)
            expr = None
            loss = 1e9
    return (loss, expr) if expr is not None else (1e9, None)


def evaluate_smart_batch(token_seqs, X_padded, y_raw, cfg, test_data, cache=None, coarse=True,
                         x0=None, engine=None):
    """``[evaluate_smart(seq, X, y, coarse) for seq in token_seqs]`` in one device launch.

    ``token_seqs``: iterable of 1-D token id tensors / lists (leading ``S`` included, as the
    reference passes them); ``cache``: the caller's ``expression_cache`` dict (updated in
    place); ``x0``: optional list of ``[R, k]`` starting points, one per sequence (the
    reference draws them from the global numpy RNG).
    """
    cache = cache if cache is not None else {}
    seqs = [tuple(int(t) for t in (s.tolist() if hasattr(s, "tolist") else s)) for s in token_seqs]
    X_in = X_padded.unsqueeze(0) if torch.as_tensor(X_padded).dim() == 2 else X_padded
    todo, index = [], {}
    for i, s in enumerate(seqs):
        if s not in cache and s not in index:
            index[s] = len(todo)
            todo.append(i)
    if todo:
        with _Restarts(cfg, coarse):
            try:
                outs = bfgs_batch([list(seqs[i]) for i in todo], X_in, y_raw, cfg, test_data,
                                  x0=None if x0 is None else [x0[i] for i in todo], engine=engine)
            except Exception as exc:  # noqa: BLE001 -- a failure of the whole launch fails each sequence
                outs = [exc] * len(todo)
        for i, out in zip(todo, outs):
            cache[seqs[i]] = _result_of(out)
    return [cache[s] for s in seqs]
