"""Skeleton compiler: sympy expression tree -> accumulator-machine bytecode.

Replaces the reference's per-point symbolic loss build and its per-restart
``lambdify`` (``src/visymre/architectures/bfgs.py:77-92`` and ``:104``): instead of
substituting every data point into a sympy expression, the skeleton is compiled
ONCE to a short program that the CUDA interpreter (``csrc/vsr_interp.h``) runs over
all points.

The input is what ``sp.sympify`` makes of the reference's infix string
(``bfgs.py:81``), so the program inherits sympy's canonical form (n-ary Add/Mul,
``x*x -> x**2``, ``x/x -> 1``, ``sqrt = Pow(., 1/2)`` ...) exactly like the
reference's lambdified functions do.  The arithmetic a node lowers to follows what
sympy's printer would hand to numpy (``bfgs.py:38-40``): rational coefficients as
float literals, negative powers as a division, ``Pow(., +-1/2)`` as sqrt.

Children of binary nodes are ordered Sethi-Ullman style and leaf operands are folded
into the instruction, so the operand stack is only touched when both children of a
node are non-trivial.
"""
import re
from dataclasses import dataclass, field

import numpy as np
import sympy as sp

from . import isa
from .isa import OP, SRC


class CompileError(Exception):
    """The skeleton cannot be lowered (unknown function, complex constant, limits)."""


@dataclass
class Program:
    insns: np.ndarray            # uint64 [n_insns], END-terminated
    imms: np.ndarray             # float64 [n_imms]
    k: int                       # number of fitted constants c0..c{k-1}
    var_mask: int                # bit j set <=> column j (x_{j+1}) is read
    stack_depth: int             # operand stack slots needed
    n_nodes: int                 # sympy nodes lowered (program "length" L)
    flops: int                   # algorithmic flops per point-eval (SURVEY 8d weights)
    sfu: int = 0                 # transcendental nodes per point-eval
    _expr: object = field(default=None, repr=False)   # the sympy expression (see ``expr``)
    source: str = None           # the infix string it was parsed from, when there is one

    @property
    def expr(self):
        """The sympy tree the program was lowered from.  A program that crossed a process boundary
        carries only its ``source`` string (pickled sympy trees cost more than the bytecode) and
        parses it again when the tree is first needed (prune, winner formatting)."""
        if self._expr is None and self.source is not None:
            self._expr = parse_skeleton(self.source)
        return self._expr

    def __getstate__(self):
        d = dict(self.__dict__)
        if d.get("source") is not None:
            d["_expr"] = None
        return d

    @property
    def n_insns(self):
        return int(self.insns.shape[0])

    def disassemble(self):
        return isa.disassemble(self.insns, self.imms)


# ---- infix string -> sympy tree ----------------------------------------------------------------------
# The reference hands the c-named infix string to ``sympify`` (bfgs.py:81): tokenizer, four token
# transformations, ``eval``.  The strings of this path are fully parenthesised expressions over a
# closed vocabulary (dataset/generator.py:_TEMPLATES), so evaluating them directly with the same names
# bound and integer literals wrapped in ``Integer`` builds the same tree through the same operator
# calls (tests/test_host_path.py compares ``srepr`` on the workloads) at ~3/4 of the cost; anything
# outside that vocabulary goes through ``sympify``.
_PARSE_NAMES = {n: getattr(sp, n) for n in ("sin", "cos", "tan", "exp", "sqrt", "Abs", "asin", "atan", "acos",
                                            "sinh", "cosh", "tanh", "log", "pi", "E", "I", "Integer")}
_PARSE_NAMES["ln"] = sp.log
_PARSE_INT = re.compile(r"(?<![\w.])(\d+)(?![\w.])")
_PARSE_WORD = re.compile(r"[A-Za-z_]\w*")
_PARSE_SYMBOL = re.compile(r"(?:x_|c)\d+\Z")
_PARSE_SAFE = re.compile(r"[\w\s()+\-*/]*\Z")
_SYMBOLS = {}


def parse_skeleton(text):
    """``sympy.sympify(text)`` for the skeleton strings of this path (same tree, less overhead)."""
    if not isinstance(text, str) or not _PARSE_SAFE.match(text):
        return sp.sympify(text)
    ns = {"__builtins__": {}, "Integer": sp.Integer}      # (the literals are wrapped below)
    for w in set(_PARSE_WORD.findall(text)):
        if w in _PARSE_NAMES:
            ns[w] = _PARSE_NAMES[w]
        elif _PARSE_SYMBOL.match(w):
            ns[w] = _SYMBOLS.get(w) or _SYMBOLS.setdefault(w, sp.Symbol(w))
        else:
            return sp.sympify(text)
    try:
        return eval(_PARSE_INT.sub(r"Integer(\1)", text), ns)   # noqa: S307 -- closed vocabulary, no builtins
    except Exception:  # noqa: BLE001 -- let sympify raise what the reference would see
        return sp.sympify(text)


_UNARY = {
    sp.exp: "VSR_EXP", sp.log: "VSR_LOG", sp.Abs: "VSR_ABS",
    sp.sin: "VSR_SIN", sp.cos: "VSR_COS", sp.tan: "VSR_TAN",
    sp.asin: "VSR_ASIN", sp.acos: "VSR_ACOS", sp.atan: "VSR_ATAN",
    sp.sinh: "VSR_SINH", sp.cosh: "VSR_COSH", sp.tanh: "VSR_TANH",
    sp.sign: "VSR_SIGN",
}
_TRANSCENDENTAL = {"VSR_EXP", "VSR_LOG", "VSR_SIN", "VSR_COS", "VSR_TAN", "VSR_ASIN",
                   "VSR_ACOS", "VSR_ATAN", "VSR_SINH", "VSR_COSH", "VSR_TANH"}


def _powi_cost(n):
    n = abs(int(n))
    return max(1, n.bit_length() - 1 + bin(n).count("1") - 1)


class _Lowering:
    def __init__(self, k, variables, slots=None):
        self.k = k
        self.slots = slots      # {i: slot} when the constants c_i keep their names (pruned skeletons)
        self.var_index = {v: i for i, v in enumerate(variables)}
        self.code = []
        self.imms = []
        self.imm_index = {}
        self.var_mask = 0
        self.depth = 0
        self.max_depth = 0
        self.n_nodes = 0
        self.flops = 0
        self.sfu = 0
        self._need = {}
        self._tmask = {}
        self._leaf = {}

    # ---- operands ----------------------------------------------------------------
    def imm(self, value):
        value = float(value)
        key = np.float64(value).tobytes()
        if key not in self.imm_index:
            if len(self.imms) >= isa.MAX_IMMS:
                raise CompileError("too many literals")
            self.imm_index[key] = len(self.imms)
            self.imms.append(value)
        return self.imm_index[key]

    def const_slot(self, name):
        """Slot of the fitted constant called `name` (``c<i>``), None if it is not one."""
        if name[0] != "c" or not name[1:].isdigit():
            return None
        j = int(name[1:])
        if self.slots is not None:
            return self.slots.get(j)
        return j if j < self.k else None

    def leaf(self, node):
        """(src, idx, tangent mask) if `node` is a leaf operand, else None (memoised: the
        emitter asks several times per node and free_symbols walks the whole subtree)."""
        try:
            return self._leaf[node]
        except KeyError:
            r = self._leaf[node] = self._leaf_of(node)
            return r

    def _leaf_of(self, node):
        if isinstance(node, sp.Symbol):
            name = node.name
            if name in self.var_index:
                return SRC["VSR_SRC_VAR"], self.var_index[name], 0
            j = self.const_slot(name)
            if j is not None:
                return SRC["VSR_SRC_CONST"], j, (1 << j) if j < isa.MAX_DUAL else 0
            raise CompileError(f"unknown symbol {name!r}")
        if isinstance(node, sp.Number) or node in (sp.pi, sp.E) or isinstance(node, sp.NumberSymbol):
            if not node.is_real or node.is_infinite or node is sp.nan:
                raise CompileError(f"non-real literal {node}")
            if isinstance(node, sp.Rational):
                val = float(int(node.p)) / float(int(node.q)) if node.q != 1 else float(int(node.p))
            else:
                val = float(node)
            return SRC["VSR_SRC_IMM"], self.imm(val), 0
        if not node.free_symbols:
            # a constant-free subtree sympy kept symbolic (sqrt(2), sin(1), ...): fold it
            # to a literal.  numpy would evaluate it in double precision at run time; the
            # correctly rounded value differs from that by at most an ulp.
            try:
                val = complex(node.evalf(20))
            except (TypeError, ValueError) as exc:
                raise CompileError(f"cannot evaluate literal {node}") from exc
            if val.imag != 0.0 or not np.isfinite(val.real):
                raise CompileError(f"non-real literal {node}")
            return SRC["VSR_SRC_IMM"], self.imm(val.real), 0
        return None

    # ---- analysis ------------------------------------------------------------------
    def tmask(self, node):
        """Tangents (constants c_j, j < MAX_DUAL) that occur under `node`."""
        m = self._tmask.get(node)
        if m is None:
            m = 0
            for s in node.free_symbols:
                j = self.const_slot(s.name)
                if j is not None and j < isa.MAX_DUAL:
                    m |= 1 << j
            self._tmask[node] = m
        return m

    def need(self, node):
        """Operand-stack slots needed to evaluate `node` into acc."""
        n = self._need.get(node)
        if n is not None:
            return n
        if self.leaf(node) is not None:
            n = 0
        else:
            kind, parts = self.shape(node)
            if kind == "unary":
                n = self.need(parts[1])
            elif kind == "binary":
                a, b = parts[1], parts[2]
                la, lb = self.leaf(a) is not None, self.leaf(b) is not None
                if lb:
                    n = self.need(a)
                elif la:
                    n = self.need(b)
                else:
                    na, nb = self.need(a), self.need(b)
                    n = max(na, nb + 1) if na >= nb else max(nb, na + 1)
            else:  # fold
                heavy = sorted((t for t in parts[1] if self.leaf(t) is None),
                               key=self.need, reverse=True)
                n = 0
                for i, t in enumerate(heavy):
                    n = max(n, self.need(t) + (1 if i else 0))
        self._need[node] = n
        return n

    def shape(self, node):
        """Classify a non-leaf sympy node.

        ("unary", (opname, child[, n])) | ("binary", (opname, a, b)) -> a op b |
        ("fold", (opname, [terms])) -> left fold with a commutative op
        """
        if isinstance(node, sp.Add):
            return "fold", ("VSR_ADD", list(node.args))
        if isinstance(node, sp.Mul):
            num, den = [], []
            for a in node.args:
                if isinstance(a, sp.Pow) and isinstance(a.exp, sp.Number) and a.exp.is_negative:
                    den.append(a.base if a.exp == -1 else sp.Pow(a.base, -a.exp, evaluate=False))
                else:
                    num.append(a)
            if den:
                d = den[0] if len(den) == 1 else sp.Mul(*den, evaluate=False)
                if not num:
                    return "unary", ("VSR_INV", d)
                n = num[0] if len(num) == 1 else sp.Mul(*num, evaluate=False)
                if n == -1:
                    return "unary", ("VSR_NEG", sp.Pow(d, -1, evaluate=False))
                return "binary", ("VSR_DIV", n, d)
            if len(num) == 2 and num[0] == -1:
                return "unary", ("VSR_NEG", num[1])
            return "fold", ("VSR_MUL", num)
        if isinstance(node, sp.Pow):
            base, ex = node.base, node.exp
            if isinstance(ex, sp.Integer):
                n = int(ex)
                if n == -1:
                    return "unary", ("VSR_INV", base)
                if n == 1:
                    return "unary", ("VSR_ID", base)
                if n == 0:
                    raise CompileError("x**0 survived sympy")
                if abs(n) <= 32767:
                    return "unary", ("VSR_POWI", base, n)
            if ex == sp.Rational(1, 2):
                return "unary", ("VSR_SQRT", base)
            if ex == sp.Rational(-1, 2):
                return "unary", ("VSR_INV", sp.Pow(base, sp.Rational(1, 2), evaluate=False))
            # half-integer powers: sqrt(x)**p, cheaper than pow and as accurate (numpy calls
            # pow(x, p/2); the two agree to an ulp or two)
            half = None
            if isinstance(ex, sp.Rational) and ex.q == 2:
                half = int(ex.p)
            elif isinstance(ex, sp.Float) and float(ex) * 2 == int(float(ex) * 2) and float(ex) != int(float(ex)):
                half = int(float(ex) * 2)
            if half is not None and 1 < abs(half) <= 15:
                return "unary", ("VSR_POWI", sp.Pow(base, sp.Rational(1, 2), evaluate=False), half)
            return "binary", ("VSR_POW", base, ex)
        if isinstance(node, sp.exp):
            return "unary", ("VSR_EXP", node.args[0])
        for cls, name in _UNARY.items():
            if isinstance(node, cls):
                return "unary", (name, node.args[0])
        raise CompileError(f"cannot lower {type(node).__name__}: {node}")

    # ---- emission ------------------------------------------------------------------
    def emit(self, opname, src=0, idx=0, amask=0, bmask=0):
        self.code.append(isa.encode(OP[opname], src, idx, amask, bmask))
        if len(self.code) >= isa.MAX_INSNS:
            raise CompileError("program too long")

    def push(self, amask):
        self.emit("VSR_PUSH", amask=amask)
        self.depth += 1
        self.max_depth = max(self.max_depth, self.depth)
        if self.depth > isa.MAX_STACK:
            raise CompileError("operand stack too deep")

    def use_var(self, src, idx):
        if src == SRC["VSR_SRC_VAR"]:
            self.var_mask |= 1 << idx

    def gen(self, node):
        """Emit code leaving `node` in acc; returns acc's tangent mask."""
        lf = self.leaf(node)
        if lf is not None:
            src, idx, m = lf
            self.use_var(src, idx)
            self.emit("VSR_LOAD", src, idx, 0, m)
            return m
        self.n_nodes += 1
        kind, parts = self.shape(node)
        if kind == "unary":
            name = parts[0]
            m = self.gen(parts[1])
            if name == "VSR_ID":
                return m
            if name == "VSR_POWI":
                self.emit(name, 0, parts[2] & 0xFFFF, m, 0)
                self.flops += _powi_cost(parts[2]) + (1 if parts[2] < 0 else 0)
            else:
                self.emit(name, 0, 0, m, 0)
                self.flops += 1
                if name in _TRANSCENDENTAL:
                    self.sfu += 1
            return m
        if kind == "binary":
            name, a, b = parts
            rname = {"VSR_DIV": "VSR_RDIV", "VSR_POW": "VSR_RPOW", "VSR_SUB": "VSR_RSUB"}[name]
            self.flops += 1
            if name == "VSR_POW":
                self.sfu += 1
            la, lb = self.leaf(a), self.leaf(b)
            if lb is not None:
                ma = self.gen(a)
                self.use_var(lb[0], lb[1])
                self.emit(name, lb[0], lb[1], ma, lb[2])
                return ma | lb[2]
            if la is not None:
                mb = self.gen(b)
                self.use_var(la[0], la[1])
                self.emit(rname, la[0], la[1], mb, la[2])
                return mb | la[2]
            if self.need(a) >= self.need(b):
                ma = self.gen(a)
                self.push(ma)
                mb = self.gen(b)
                self.depth -= 1
                self.emit(rname, SRC["VSR_SRC_STACK"], 0, mb, ma)  # acc = stack (op) acc
            else:
                mb = self.gen(b)
                self.push(mb)
                ma = self.gen(a)
                self.depth -= 1
                self.emit(name, SRC["VSR_SRC_STACK"], 0, ma, mb)   # acc = acc (op) stack
            return ma | mb
        # fold
        name, terms = parts
        heavy = sorted((t for t in terms if self.leaf(t) is None), key=self.need, reverse=True)
        light = [t for t in terms if self.leaf(t) is not None]
        self.flops += len(terms) - 1
        m = None
        for t in heavy:
            if m is None:
                m = self.gen(t)
            else:
                self.push(m)
                mt = self.gen(t)
                self.depth -= 1
                self.emit(name, SRC["VSR_SRC_STACK"], 0, mt, m)
                m |= mt
        for t in light:
            src, idx, mt = self.leaf(t)
            self.use_var(src, idx)
            if m is None:
                self.emit("VSR_LOAD", src, idx, 0, mt)
                m = mt
            else:
                self.emit(name, src, idx, m, mt)
                m |= mt
        return m


def _peephole(code):
    """NEG followed by ADD <operand>  ->  RSUB <operand>  (operand - acc).

    Bit-exact: IEEE negation is exact and b + (-a) == b - a, for the value and for every
    tangent.  sympy's canonical form writes x - y as Add(x, Mul(-1, y)), so this removes one
    dispatch from every difference."""
    out = []
    for w in code:
        op, src, idx, am, bm = isa.decode(w)
        if out and op == OP["VSR_ADD"]:
            pop, psrc, pidx, pam, pbm = isa.decode(out[-1])
            if pop == OP["VSR_NEG"]:
                out[-1] = isa.encode(OP["VSR_RSUB"], src, idx, am, bm)
                continue
        out.append(w)
    return out


def compile_sympy(expr, k, variables, slots=None):
    """Lower a sympy expression in the symbols ``variables`` and ``c0..c{k-1}``.  ``slots``: the
    fitted constants are the symbols ``c<i>`` for i in ``slots``, at slot ``slots[i]`` (the reference
    lambdifies a pruned loss over the REMAINING symbols under their own names, bfgs.py:161-168)."""
    if k > isa.MAX_CONSTS:
        raise CompileError(f"{k} constants exceed the limit of {isa.MAX_CONSTS}")
    expr = sp.sympify(expr)
    if expr.has(sp.I) or expr.has(sp.zoo) or expr.has(sp.nan) or expr.has(sp.oo):
        raise CompileError(f"non-real expression {expr}")
    low = _Lowering(k, list(variables), slots)
    low.gen(expr)
    low.emit("VSR_END")
    low.code = _peephole(low.code)
    return Program(
        insns=np.asarray(low.code, dtype=np.uint64),
        imms=np.asarray(low.imms if low.imms else [0.0], dtype=np.float64),
        k=k, var_mask=low.var_mask, stack_depth=low.max_depth,
        n_nodes=max(low.n_nodes, 1), flops=low.flops + 3, sfu=low.sfu, _expr=expr)


def compile_skeleton(expr_str, k, variables):
    """Compile the reference's c-named infix string (``bfgs.py:69-71``)."""
    prog = compile_sympy(parse_skeleton(expr_str), k, variables)
    prog.source = expr_str
    return prog
