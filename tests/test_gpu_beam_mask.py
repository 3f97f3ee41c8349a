"""Device constraint mask of the beam search against the reference's host
loop (model.py:385-411) restated with analyze_prefix_tree_context (model.py:522-560)."""
import numpy as np
import pytest
import torch

from src.visymre.architectures import model as vmodel

pytestmark = pytest.mark.gpu


def _host_mask(gen, cur_len, beam_scores, n_words, ids, length_eq):
    """The reference's block, line for line."""
    mask = np.zeros((gen.shape[0], n_words), dtype=np.float32)
    for i in range(gen.shape[0]):
        if beam_scores[i] < -1e8:
            continue
        seq = gen[i, :cur_len].tolist()
        valency, structural = vmodel.analyze_prefix_tree_context(
            seq, ids["arity_1"], ids["arity_2"], ids["trans"], ids["pow"], ids["c"], ids["start"])
        forbidden = set(structural)
        if valency >= length_eq - cur_len:
            forbidden.update(ids["all_ops"])
        if valency > 0:
            forbidden.add(ids["finish"])
            forbidden.add(ids["pad"])
        forbidden.update(ids["masked_vars"])
        for x in forbidden:
            if x < n_words:
                mask[i, x] = -np.inf
    return mask


def _ids(golden):
    w = golden["word2id"]
    a1 = [w[k] for k in ("abs", "asin", "cos", "exp", "ln", "sin", "sqrt", "tan")]
    a2 = [w[k] for k in ("add", "div", "mul", "pow", "sub")]
    return dict(arity_1=a1, arity_2=a2, trans=[w[k] for k in ("sin", "cos", "tan", "exp", "ln", "asin")],
                pow=w["pow"], c=w["c"], start=w["S"], finish=w["F"], pad=w["P"], all_ops=a1 + a2,
                masked_vars=[w["x_4"], w["x_5"], w["x_9"]]), max(w.values()) + 1


@pytest.mark.parametrize("cur_len", [1, 2, 7, 23, 60])
def test_device_mask_equals_the_host_loop(golden, cur_len):
    ids, n_words = _ids(golden)
    w = golden["word2id"]
    rng = np.random.RandomState(cur_len)
    beam, L, length_eq = 96, 64, 62
    leaves = [w[f"x_{j}"] for j in range(1, 4)] + [w["c"], w["2"], w["pi"]]
    pool = ids["arity_1"] + ids["arity_2"] * 2 + leaves * 3 + [w["F"]]
    gen = np.full((beam, L), w["P"], dtype=np.int64)
    gen[:, 0] = w["S"]
    gen[:, 1:] = rng.choice(pool, size=(beam, L - 1))
    gen[5, 0] = w["add"]                      # a beam that does not start with S
    gen[6, 1:4] = [w["pow"], w["x_1"], w["x_2"]]   # pow: the exponent slot forbids c
    gen[7, 1:3] = [w["pow"], w["x_1"]]
    gen[8, 1:3] = [w["sin"], w["add"]]             # inside a transcendental
    scores = rng.normal(size=beam).astype(np.float32)
    scores[[3, 11]] = -1e9                         # dead beams
    ref = _host_mask(gen, cur_len, scores, n_words, ids, length_eq)
    got = vmodel.beam_constraint_mask(
        torch.tensor(gen, device="cuda:0"), cur_len, torch.tensor(scores, device="cuda:0"), n_words,
        arity_1_ids=ids["arity_1"], arity_2_ids=ids["arity_2"], transcendental_ids=ids["trans"],
        all_op_ids=ids["all_ops"], masked_var_ids=ids["masked_vars"], pow_id=ids["pow"], c_id=ids["c"],
        start_id=ids["start"], finish_id=ids["finish"], pad_id=ids["pad"], length_eq=length_eq)
    assert np.array_equal(got.cpu().numpy(), ref)
    assert (ref[[3, 11]] == 0).all() and np.isinf(ref).any()
