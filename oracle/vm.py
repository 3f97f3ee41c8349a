"""numpy interpreter for the skeleton bytecode (TEST INFRASTRUCTURE).

Executes a compiled program (``vision-sr_b200/src/visymre/engine/compiler.py``) over
numpy columns with the arithmetic the CUDA interpreter uses, so the CPU tests can
check compiler output against ``sympy.lambdify`` -- the function the reference
evaluates (src/visymre/architectures/bfgs.py:104, :128).  Value only; tangents are
checked through the host simulator and on the GPU.
"""
import numpy as np


def run(program, X, consts, isa):
    """X: [N, d] array; consts: [k]; returns f(X; consts) as float64 [N]."""
    OP, SRC = isa.OP, isa.SRC
    X = np.asarray(X)
    N = X.shape[0]
    dt = X.dtype
    acc = np.zeros(N, dtype=dt)
    stack = []
    un = {
        OP["VSR_NEG"]: np.negative, OP["VSR_ABS"]: np.abs, OP["VSR_SQRT"]: np.sqrt,
        OP["VSR_EXP"]: np.exp, OP["VSR_LOG"]: np.log, OP["VSR_SIN"]: np.sin,
        OP["VSR_COS"]: np.cos, OP["VSR_TAN"]: np.tan, OP["VSR_ASIN"]: np.arcsin,
        OP["VSR_ACOS"]: np.arccos, OP["VSR_ATAN"]: np.arctan, OP["VSR_SINH"]: np.sinh,
        OP["VSR_COSH"]: np.cosh, OP["VSR_TANH"]: np.tanh, OP["VSR_SIGN"]: np.sign,
        OP["VSR_INV"]: lambda a: 1.0 / a,
    }
    with np.errstate(all="ignore"):
        for w in program.insns:
            op, src, idx, _, _ = isa.decode(w)
            if op == OP["VSR_END"]:
                break
            if op == OP["VSR_PUSH"]:
                stack.append(acc.copy())
                continue
            if op <= OP["VSR_RPOW"]:
                if src == SRC["VSR_SRC_STACK"]:
                    b = stack.pop()
                elif src == SRC["VSR_SRC_VAR"]:
                    b = X[:, idx]
                elif src == SRC["VSR_SRC_CONST"]:
                    b = np.full(N, consts[idx], dtype=dt)
                else:
                    b = np.full(N, program.imms[idx], dtype=dt)
                if op == OP["VSR_LOAD"]:
                    acc = b.astype(dt, copy=True)
                elif op == OP["VSR_ADD"]:
                    acc = acc + b
                elif op == OP["VSR_SUB"]:
                    acc = acc - b
                elif op == OP["VSR_RSUB"]:
                    acc = b - acc
                elif op == OP["VSR_MUL"]:
                    acc = acc * b
                elif op == OP["VSR_DIV"]:
                    acc = acc / b
                elif op == OP["VSR_RDIV"]:
                    acc = b / acc
                elif op == OP["VSR_POW"]:
                    acc = np.power(acc, b)
                elif op == OP["VSR_RPOW"]:
                    acc = np.power(b, acc)
                continue
            if op == OP["VSR_POWI"]:
                n = idx - 0x10000 if idx & 0x8000 else idx
                r = np.ones(N, dtype=dt)
                for _ in range(abs(n)):
                    r = r * acc
                acc = r if n > 0 else 1.0 / r
                continue
            acc = un[op](acc)
    return acc.astype(np.float64)
