"""Token helpers of the skeleton language (hot-path subset).

Behaviour follows reference ``src/visymre/architectures/data.py``:
``tokenize`` :199-205, ``de_tokenize`` :207-221, ``constants_to_placeholder``
:160-169, ``sanitize_prefix`` :183-197.  The dataset / rendering half of that file
is training-only and out of scope.
"""
import re

import sympy as sp

_NUMBER = re.compile(r"[+-]?(\d+(\.\d*)?|\.\d+)([eE][+-]?\d+)?")
ALLOWED_INTS = {str(i) for i in range(-9, 10)}


def tokenize(prefix_expr, word2id):
    """words -> ids, wrapped in S ... F."""
    return [word2id["S"], *(word2id[w] for w in prefix_expr), word2id["F"]]


def de_tokenize(tokenized_expr, id2word):
    """ids -> words, stopping at the first F (the caller already dropped S)."""
    words = []
    for tok in tokenized_expr:
        idx = tok.item() if hasattr(tok, "item") else tok
        word = id2word[idx]
        if word == "F":
            break
        words.append(word)
    return words


def constants_to_placeholder(s, symbol="c"):
    """Replace Floats and integers beyond +-9 by the placeholder symbol.

    Returns ``(expr_with_placeholders, original_expr)`` like the reference does.
    """
    expr = sp.sympify(s)
    hole = sp.Symbol(symbol, real=True, nonzero=True)

    def is_const(node):
        return isinstance(node, sp.Float) or (isinstance(node, sp.Integer) and abs(node) > 9)

    replaced = expr.xreplace({n: hole for n in expr.atoms(sp.Number) if is_const(n)})
    return replaced, expr


def sanitize_prefix(tokens):
    """Map numeric words outside the vocabulary (and the imaginary unit) to ``c``."""
    out = []
    for t in tokens:
        if t == "I":
            out.append("c")
        elif t.lstrip("-").isdigit():
            out.append(t if t in ALLOWED_INTS else "c")
        elif _NUMBER.fullmatch(t):
            out.append("c")
        else:
            out.append(t)
    return out
