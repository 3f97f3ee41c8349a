"""Host side of the drop-in: token/string rules, workload generator, formatting, caching.
No GPU: everything up to (and after) the fit."""
import numpy as np
import pytest
import sympy as sp

from oracle import vectorised
from src.visymre.architectures import bfgs as vbfgs
from src.visymre.architectures import data as vdata
from src.visymre.architectures.model import analyze_prefix_tree_context, BINARY_NAMES, UNARY_NAMES
from src.visymre.dataset.generator import Generator, InvalidPrefixExpression
from src.visymre.workloads import generator as wg


@pytest.fixture(scope="module")
def beams():
    return wg.feynman_beams(n_points=64, n_cand=24, n_restarts=3, limit=4)


def test_skeleton_string_equals_the_oracles_independent_rules(beams, golden, test_data):
    bs, td = beams
    cfg = wg.make_cfg(3)
    for b in bs:
        for ids in b.tokens:
            assert vbfgs.skeleton_string(ids, cfg, td) == vectorised.skeleton_string(ids, td.id2word)
    for case in golden["cases"]:
        if case["raised"]:
            with pytest.raises(InvalidPrefixExpression):
                vbfgs.skeleton_string(case["tokens"], cfg, test_data)
        else:
            assert vbfgs.skeleton_string(case["tokens"], cfg, test_data)[0] == case["skeleton"]


def test_variable_shift_rule():
    # bfgs.py:11-21: x_i -> x_{i-1} for i = 2..5 when x_{i-1} is absent, applied in order
    assert vbfgs.replace_illegal_variables("(c0)/(x_2)") == "(c0)/(x_1)"
    assert vbfgs.replace_illegal_variables("x_1+x_3") == "x_1+x_2"
    assert vbfgs.replace_illegal_variables("x_2*x_3") == "x_1*x_3"   # only x_2 moves (x_2 was present)
    assert vbfgs.replace_illegal_variables("x_7") == "x_7"
    with pytest.raises(ValueError):
        vbfgs.replace_illegal_variables("x_0+1")


def test_one_pass_substitution_prints_like_the_references_sequence(beams):
    bs, td = beams
    cfg = wg.make_cfg(3)
    rng = np.random.RandomState(0)
    for b in bs:
        for ids in b.tokens:
            expr, k, prog = vbfgs._compile_candidate(ids, cfg, td, td.total_variables)
            vals = list(rng.randn(k) * 3)
            syms = [sp.Symbol(f"c{i}") for i in range(k)]
            assert str(vbfgs._substitute(prog.expr, syms, vals)) == \
                str(vbfgs._substitute_like_reference(expr, syms, vals))


def test_compile_cache_hits(beams):
    bs, td = beams
    cfg = wg.make_cfg(3)
    a = vbfgs._compile_candidate(bs[0].tokens[0], cfg, td, td.total_variables)
    b = vbfgs._compile_candidate(list(bs[0].tokens[0]), cfg, td, td.total_variables)
    assert a is b


def test_tokens_roundtrip_and_sanitize(test_data):
    w2i = test_data.word2id
    words = ["add", "c", "mul", "x_1", "sin", "x_10"]
    ids = vdata.tokenize(words, w2i)
    assert ids[0] == w2i["S"] and ids[-1] == w2i["F"]
    back = vdata.de_tokenize(ids[1:], {v: k for k, v in w2i.items()})
    assert back == words
    assert vdata.sanitize_prefix(["3", "12", "-4", "2.5", "1e-3", "I", "x_1"]) == ["3", "c", "-4", "c", "c", "c", "x_1"]
    expr, orig = vdata.constants_to_placeholder("2.5*x_1 + 11*x_2 + 3")
    assert str(expr) == "c*x_1 + c*x_2 + 3"


def test_prefix_infix_and_valency(test_data):
    w2i = test_data.word2id
    a1 = {w2i[n] for n in UNARY_NAMES}
    a2 = {w2i[n] for n in BINARY_NAMES}
    good = [w2i[w] for w in "S add c mul c sin x_1".split()]
    bad = [w2i[w] for w in "S add c mul c".split()]
    assert analyze_prefix_tree_context(good, a1, a2, set(), w2i["pow"], None, w2i["S"])[0] == 0
    assert analyze_prefix_tree_context(bad, a1, a2, set(), w2i["pow"], None, w2i["S"])[0] == 1
    # no_c_in_pow: the exponent slot of pow forbids the constant token
    half = [w2i[w] for w in "S pow x_1".split()]
    val, forb = analyze_prefix_tree_context(half, a1, a2, set(), w2i["pow"], w2i["c"], w2i["S"])
    assert val == 1 and w2i["c"] in forb
    assert Generator.prefix_to_infix(["mul", "c", "pow", "x_1", "2"], coefficients=["c"],
                                     variables=["x_1"]) == "(({c})*((x_1)**(2)))"
    with pytest.raises(InvalidPrefixExpression):
        Generator.prefix_to_infix(["add", "x_1", "x_2", "x_3"], coefficients=[], variables=[])
    assert Generator.sympy_to_prefix(sp.sympify("sqrt(x_1)*3/2")) == ["mul", "div", "3", "2", "sqrt", "x_1"]


def test_workload_is_seeded_and_well_formed(beams):
    bs, td = beams
    again, _ = wg.feynman_beams(n_points=64, n_cand=24, n_restarts=3, limit=4)
    for b, b2 in zip(bs, again):
        assert b.tokens == b2.tokens
        np.testing.assert_array_equal(b.X, b2.X)
        for u, v in zip(b.x0, b2.x0):
            np.testing.assert_array_equal(u, v)
        assert b.X.shape == (64, 10) and np.all(np.isfinite(b.y))
        assert len(set(map(tuple, b.tokens))) == len(b.tokens) == 24
        progs = wg.compile_beam(b, td)
        assert all(p.k == x.shape[1] and p.k <= 8 for p, x in zip(progs, b.x0))


def test_parse_skeleton_builds_sympifys_tree(beams, golden):
    """engine/compiler.py:parse_skeleton against sympy.sympify (what the reference calls, bfgs.py:81)
    on every skeleton string of the workload beams, the golden cases and the odd words of the
    vocabulary; strings outside the closed vocabulary fall through to sympify."""
    import pickle
    from src.visymre.engine.compiler import compile_skeleton, parse_skeleton
    bs, td = beams
    cfg = wg.make_cfg(3)
    texts = [vbfgs.skeleton_string(ids, cfg, td)[0] for b in bs for ids in b.tokens]
    texts += [c["skeleton"] for c in golden["cases"] if c.get("skeleton")]
    texts += ["(1/((x_1)**2))", "((pi)*((x_2)**3))+((1)/(2))", "((E)**(c0))-((3)/(x_1))", "(atan((c0)*(I)))",
              "((x_1)-(x_1))", "((2)*(3))", "(ln(Abs((c0)+(-1))))", "(((x_1)**5)/((-2)+(c1)))"]
    assert len(texts) > 100
    for t in texts:
        assert sp.srepr(parse_skeleton(t)) == sp.srepr(sp.sympify(t)), t
    assert parse_skeleton("x_1 + foo(2)") == sp.sympify("x_1 + foo(2)")      # unknown word: sympify's answer
    assert parse_skeleton("2.5*x_1") == sp.sympify("2.5*x_1")                # a float literal: sympify's answer
    with pytest.raises(Exception):
        parse_skeleton("((x_1)+")
    # a program that crosses a process boundary leaves its tree behind and parses its source again
    prog = compile_skeleton("((c0)*(sin((x_1)+(c1))))", 2, td.total_variables)
    back = pickle.loads(pickle.dumps(prog))
    assert back._expr is None and sp.srepr(back.expr) == sp.srepr(prog.expr)
    assert np.array_equal(back.insns, prog.insns) and back.k == 2


def test_derivative_is_constant_shortcut_agrees_with_sympy():
    """bfgs.py:160 ``diff(expr, c).is_constant()``: the numeric shortcut may only ever answer what
    sympy answers."""
    c0 = sp.Symbol("c0")
    for text in ["(c0)*((x_1)+(x_2))", "(c0)*(sin(x_1)**2+cos(x_1)**2)", "(c0)+(x_1)", "(c0)*(c0)*(x_1)", "sin(c0)",
                 "(c0)*(3)", "(c0)/(x_1)", "(c0)*(((x_1)+1)**2-(x_1)**2-2*(x_1))", "exp((c0)*(x_1))",
                 "(c0)*(tan(cos((x_1)*(x_3)))+((x_2)*(x_4)))/((x_1)+(x_2))"]:
        e = sp.sympify(text)
        assert bool(vbfgs._derivative_is_constant(e, c0)) == bool(sp.diff(e, c0).is_constant()), text


def test_a_dead_worker_does_not_fail_the_beam(beams, monkeypatch):
    """The symbolic work runs in worker processes; when one dies the task is compiled in-process and
    the next call gets a fresh pool (the reference's pool is created per call, model.py:490)."""
    import os
    from src.visymre.engine import hostpool
    bs, td = beams
    cfg = wg.make_cfg(3)
    monkeypatch.setenv("VSR_HOST_WORKERS", "2")
    toks = list(bs[0].tokens)
    vbfgs._COMPILED.clear()
    want = vbfgs._compile_candidates(toks, cfg, td, list(td.total_variables))
    vbfgs._COMPILED.clear()
    pool = hostpool.get_pool(2)
    for pid in list(pool._processes):                 # kill the workers under the pool
        os.kill(pid, 9)
    got = vbfgs._compile_candidates(toks, cfg, td, list(td.total_variables))
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert isinstance(a, Exception) == isinstance(b, Exception)
        if not isinstance(a, Exception):
            assert a[0] == b[0] and np.array_equal(a[2].insns, b[2].insns)
    vbfgs._COMPILED.clear()
    again = vbfgs._compile_candidates(toks, cfg, td, list(td.total_variables))    # a fresh pool serves it
    assert len(again) == len(want) and hostpool._POOL is not None
    hostpool.shutdown()


def test_parse_skeleton_takes_the_direct_path_for_integer_literals(monkeypatch):
    """Regression: ``Integer`` was only bound when the word occurred in the text, so every skeleton with
    an integer literal fell through to ``sympify`` (correct, but the slow path)."""
    import sympy
    from src.visymre.engine import compiler
    calls = []
    real = sympy.sympify
    monkeypatch.setattr(compiler.sp, "sympify", lambda *a, **k: (calls.append(a), real(*a, **k))[1])
    e = compiler.parse_skeleton("((c0)+((x_1)**(2)))*((-1)/((c1)+(3)))")
    assert not calls and sympy.srepr(e) == sympy.srepr(real("((c0)+((x_1)**(2)))*((-1)/((c1)+(3)))"))


def test_stage_plan_of_the_staged_fit():
    """architectures/bfgs.py:stage_goes -- when the next stage of a fit staged behind the compilation
    is launched (the GPU tests run the staged path itself; this is its decision rule)."""
    go = vbfgs.stage_goes
    assert go(64, 64, 0, 0.0, 15) and go(3, 3, 3, 0.0, 15)          # all that is left always goes
    assert not go(14, 64, 0, 1.0, 15) and go(15, 64, 0, 0.0, 15)     # first stage: enough to fill the GPU
    assert not go(40, 49, 1, 0.001, 15)                              # later stage: not before the stragglers
    assert go(40, 49, 1, 0.005, 15) and not go(30, 49, 1, 0.1, 15)   # ... have kept it waiting; and most is there
    assert not go(7, 9, 2, 1.0, 15)                                  # fewer than eight: wait for the rest
    assert not go(40, 49, vbfgs.MAX_STAGES - 1, 1.0, 15)             # the last engine takes everything or nothing
