"""Host logic of the SURVEY 8f rows (no GPU): the scoring compiler's column pairing, the HLSC
per-sequence result rules, the beam-mask token sets."""
import numpy as np
import pytest
import sympy as sp

from src.visymre import hlsc, scoring
from src.visymre.architectures import model as vmodel
from src.visymre.architectures import refine as vrefine
from src.visymre.engine import isa


def test_scoring_pairs_the_ith_used_variable_with_the_ith_column():
    """Feynman_test.py:84: {v: X[:, i] for i, v in enumerate(vars_)}."""
    assert scoring.get_variable_names("x_10*x_2 + sin(x_2) - x_7") == ["x_2", "x_7", "x_10"]
    p = scoring.compile_for_scoring("x_3 + 2*x_7")
    assert p.k == 0 and p.var_mask == 0b11           # reads columns 0 and 1
    q = scoring.compile_for_scoring("x_3 + 2*x_7", by_rank=False)
    assert q.var_mask == (1 << 2) | (1 << 6)
    r = scoring.compile_for_scoring("x_2*x_1")       # already dense: unchanged
    assert r.var_mask == 0b11
    assert scoring.compile_for_scoring("0.25").var_mask == 0


def test_hlsc_result_rules():
    """hlsc.py:408-437."""
    assert hlsc._result_of(ValueError("boom")) == (1e9, None)
    loss, expr = hlsc._result_of(("2.5*x_1 + 1", [2.5, 1.0], 0.125, "c0*x_1 + c1"))
    assert loss == 0.125 and expr == sp.sympify("2.5*x_1 + 1", evaluate=False)
    assert hlsc._result_of(("x_1", [], float("nan"), "x_1"))[0] == 1e9
    assert hlsc._result_of(("x_1", [], float("inf"), "x_1"))[0] == 1e9
    assert hlsc._result_of(("x_1", [], None, "x_1"))[0] == 1e9
    assert hlsc._result_of(("x_1", [], complex(3, 4), "x_1"))[0] == 5.0 + 1e6
    assert hlsc._result_of(("x_1 +* 2", [], 0.5, "x_1"))[0] == 1e9 and hlsc._result_of(("x_1 +* 2", [], 0.5, "x_1"))[1] is None
    assert hlsc._result_of((None, [], 0.5, "x_1")) == (1e9, None)


def test_restarts_context_restores_the_knob():
    from types import SimpleNamespace as NS
    cfg = NS(bfgs=NS(n_restarts=10))
    with hlsc._Restarts(cfg, coarse=True):
        assert cfg.bfgs.n_restarts == 1
    assert cfg.bfgs.n_restarts == 10
    with hlsc._Restarts(cfg, coarse=False):
        assert cfg.bfgs.n_restarts == 10


def test_beam_mask_token_sets():
    assert vrefine._bits([0, 3, 63]) == (1 << 0) | (1 << 3) | (1 << 63)
    assert vrefine._bits(None) == 0
    with pytest.raises(ValueError):
        vrefine._bits([64])
    # the ctypes mirror of vsr_beam_rules has the header's layout: 5 x u64 + 6 x i32
    from src.visymre.engine import native
    import ctypes
    assert ctypes.sizeof(native.BeamRules) == 5 * 8 + 6 * 4


def test_widths_cover_every_constant_count_up_to_the_dual_limit():
    ws = isa.DUAL_WIDTHS
    assert list(ws) == sorted(ws) and ws[0] == 0 and ws[-1] == isa.MAX_DUAL
    for k in range(isa.MAX_DUAL + 1):
        assert isa.pick_dual_width(k) >= k
    assert isa.pick_dual_width(isa.MAX_DUAL + 1) is None
