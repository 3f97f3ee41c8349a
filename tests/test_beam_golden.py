"""The beam search's constraint rules against the UNMODIFIED reference.

tests/golden/ref_beam_mask.json (oracle/make_golden_beam.py): for 27 seeded beams of prefix sequences
-- three settings of the rules (as shipped; ``c`` forbidden in the exponent of ``pow``; nested
transcendentals forbidden), nine prefix lengths from 1 to 61 -- what the reference's own
``Model._analyze_prefix_tree_context`` (model.py:522-560) returned for every beam and the ``-inf``
mask its "Constraint Logic" block (model.py:382-411) produced, the block executed unchanged.

CPU: the Python port (architectures/refine.py).  GPU: ``vsr_beam_mask``, one launch per decode step,
and its incremental form, which carries the walk's state from step to step.  Bit-exact."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from src.visymre.architectures import refine


@pytest.fixture(scope="module")
def beam_golden():
    return json.load(open(os.path.join(GOLDEN, "ref_beam_mask.json")))


def _rules(g, case):
    w = g["word2id"]
    return dict(arity_1_ids=g["arity_1"], arity_2_ids=g["arity_2"], transcendental_ids=case["transcendental"],
                all_op_ids=g["arity_1"] + g["arity_2"], masked_var_ids=case["masked_vars"], pow_id=w["pow"],
                c_id=case["c_id"], start_id=w["S"], finish_id=w["F"], pad_id=w["P"], length_eq=case["length_eq"])


def test_the_file_comes_from_the_reference(beam_golden):
    g = beam_golden
    assert g["reference"].startswith("aidalee123/Vision-SR") and len(g["cases"]) == 27
    assert {c["setting"] for c in g["cases"]} == {"as_shipped", "no_c_in_pow", "nested_transcendentals"}


def test_python_port_equals_the_reference(beam_golden):
    g = beam_golden
    w = g["word2id"]
    n = 0
    for case in g["cases"]:
        gen = np.asarray(case["generated"])
        for i in range(case["beam"]):
            v, f = refine.analyze_prefix_tree_context(gen[i, :case["cur_len"]].tolist(), set(g["arity_1"]), set(g["arity_2"]),
                                                      set(case["transcendental"]), w["pow"], case["c_id"], w["S"])
            assert [int(v), sorted(int(x) for x in f)] == case["context"][i], (case["setting"], case["cur_len"], i)
            n += 1
    assert n == 27 * 40


def _mask_to_bits(mask):
    m = np.asarray(mask)
    assert set(np.unique(m)) <= {0.0, -np.inf}
    return [int(sum(1 << j for j in range(m.shape[1]) if m[i, j] == -np.inf)) for i in range(m.shape[0])]


@pytest.mark.gpu
def test_device_mask_equals_the_reference_block(beam_golden):
    import torch
    g = beam_golden
    for case in g["cases"]:
        gen = torch.tensor(case["generated"], device="cuda:0")
        sc = torch.tensor(case["beam_scores"], device="cuda:0")
        got = refine.beam_constraint_mask(gen, case["cur_len"], sc, g["n_words"], **_rules(g, case))
        assert _mask_to_bits(got.cpu().numpy()) == case["mask_bits"], (case["setting"], case["cur_len"])


@pytest.mark.gpu
def test_incremental_device_mask_equals_the_reference_block(beam_golden):
    """The decode loop's form: the walk's state lives on the device and every step consumes ONE new
    token per beam (O(1) per step instead of re-walking the prefix); beams are re-ordered between
    steps the way the beam search re-orders ``generated`` (an index per surviving beam)."""
    import torch
    g = beam_golden
    for case in g["cases"]:
        if case["cur_len"] < 2:
            continue
        gen = torch.tensor(case["generated"], device="cuda:0")
        sc = torch.tensor(case["beam_scores"], device="cuda:0")
        rules = _rules(g, case)
        state = refine.BeamMaskState(gen.shape[0], g["n_words"], device="cuda:0", **rules)
        rng = np.random.RandomState(case["cur_len"])
        order = torch.arange(gen.shape[0], device="cuda:0")
        live = torch.ones_like(sc)
        for t in range(case["cur_len"]):
            # the search picks, for every slot of the next step, the beam it continues (beam_idx):
            # emulate a shuffle half way through by permuting the beams and un-permuting the tokens
            if t == case["cur_len"] // 2:
                perm = torch.tensor(rng.permutation(gen.shape[0]), device="cuda:0")
                state.reorder(perm)
                order = order[perm]
            mask = state.step(gen[order, t], sc[order] if t == case["cur_len"] - 1 else live)
        bits = _mask_to_bits(mask.cpu().numpy())
        inv = np.argsort(order.cpu().numpy())
        assert [bits[j] for j in inv] == case["mask_bits"], (case["setting"], case["cur_len"])
