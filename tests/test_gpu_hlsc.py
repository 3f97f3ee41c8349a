"""Batched HLSC candidate evaluation against the reference's per-sequence rule
(src/visymre/hlsc.py:392-443) driven by the oracle's bfgs()."""
from types import SimpleNamespace as NS

import numpy as np
import pytest
import sympy as sp
import torch

from conftest import make_cfg
from oracle import vectorised
from src.visymre import hlsc

pytestmark = pytest.mark.gpu


def _reference_evaluate_smart(seq, X, y, cfg, td, x0):
    """hlsc.py:392-443 with the oracle's bfgs in place of the reference's."""
    saved = cfg.bfgs.n_restarts
    cfg.bfgs.n_restarts = 1
    try:
        pred, _, loss, _ = vectorised.bfgs(list(seq), X, y, cfg, td, x0=x0)
        if loss is None or np.isnan(loss) or np.isinf(loss):
            loss = 1e9
        expr = sp.sympify(pred, evaluate=False) if isinstance(pred, str) else None
        return (float(loss), expr) if expr is not None else (1e9, None)
    except Exception:  # noqa: BLE001
        return 1e9, None
    finally:
        cfg.bfgs.n_restarts = saved


def test_batch_equals_the_serial_rule(golden, test_data):
    w2i = golden["word2id"]
    cands = ["add c mul mul c sin x_1 x_2", "add mul c x_1 mul c x_2", "mul c exp mul c x_1",
             "mul x_1 x_2", "add c mul mul c sin x_1 x_2", "mul c ln x_1"]
    seqs = [[w2i["S"]] + [w2i[w] for w in c.split()] + [w2i["F"]] for c in cands]
    seqs.append([w2i["S"], w2i["add"], w2i["x_1"], w2i["F"]])       # incomplete tree: the fit raises
    rng = np.random.RandomState(2)
    N = 400
    X = np.zeros((1, N, 10))
    X[0, :, 0] = rng.uniform(0.2, 2.0, N)
    X[0, :, 1] = rng.uniform(0.5, 3.0, N)
    y = 0.75 + 2.5 * np.sin(X[0, :, 0]) * X[0, :, 1]
    cfg = make_cfg(10, "MSE", False, grad_mode="dual")
    ks = [c.split().count("c") for c in cands] + [0]
    x0 = [np.random.RandomState(40 + i).randn(1, max(k, 1))[:, :k] * 2 for i, k in enumerate(ks)]
    cache = {}
    got = hlsc.evaluate_smart_batch([torch.tensor(s) for s in seqs], torch.tensor(X[0], device="cuda:0"),
                                    torch.tensor(y, device="cuda:0"), cfg, test_data, cache=cache,
                                    coarse=True, x0=x0)
    assert cfg.bfgs.n_restarts == 10                      # restored
    assert len(got) == len(seqs) and len(cache) == len(set(map(tuple, seqs)))
    assert got[0] is got[4]                               # duplicate sequence: one fit, one cache entry
    for s, g, s0 in zip(seqs, got, x0):
        ref = _reference_evaluate_smart(s, X, y, cfg, test_data, s0)
        if ref[1] is None:
            assert g == (1e9, None)
            continue
        assert g[1] is not None
        assert abs(g[0] - ref[0]) <= 1e-6 * max(1.0, abs(ref[0])) + 1e-9, (s, g[0], ref[0])
    # cache hit: no new fit, same objects
    again = hlsc.evaluate_smart_batch([torch.tensor(seqs[1])], torch.tensor(X[0], device="cuda:0"),
                                      torch.tensor(y, device="cuda:0"), cfg, test_data, cache=cache)
    assert again[0] is got[1]
