"""Grammar of the skeleton language: prefix words <-> infix string <-> sympy.

Hot-path subset of reference ``src/visymre/dataset/generator.py``: operator arities
(:84-106), ``write_infix`` (:523-595), ``_prefix_to_infix``/``prefix_to_infix``
(:625-680), ``sympy_to_prefix`` (:720-781).  The random expression sampler of that
file generates training data and is out of scope.
"""
import sympy as sp


class InvalidPrefixExpression(Exception):
    pass


class UnknownSymPyOperator(Exception):
    pass


# word -> (arity, infix template)
_TEMPLATES = {
    "add": (2, "({0})+({1})"),
    "sub": (2, "({0})-({1})"),
    "mul": (2, "({0})*({1})"),
    "div": (2, "({0})/({1})"),
    "pow": (2, "({0})**({1})"),
    "pow2": (1, "({0})**2"),
    "pow3": (1, "({0})**3"),
    "pow5": (1, "({0})**5"),
    "inv": (1, "1/({0})"),
    "abs": (1, "Abs({0})"),
}
for _f in ("sqrt", "exp", "ln", "sin", "cos", "tan", "atan", "asin"):
    _TEMPLATES[_f] = (1, _f + "({0})")


class Generator:
    OPERATORS = {w: a for w, (a, _) in _TEMPLATES.items()}
    operators = sorted(OPERATORS)
    constants = ["pi", "E"]

    # sympy node class -> word, for sympy_to_prefix
    SYMPY_OPERATORS = {
        sp.Add: "add",
        sp.Mul: "mul",
        sp.Pow: "pow",
        sp.exp: "exp",
        sp.log: "ln",
        sp.Abs: "abs",
        sp.sin: "sin",
        sp.cos: "cos",
        sp.tan: "tan",
        sp.asin: "asin",
        sp.atan: "atan",
    }

    @classmethod
    def write_infix(cls, token, args):
        tpl = _TEMPLATES.get(token)
        return tpl[1].format(*args) if tpl else token

    @classmethod
    def _prefix_to_infix(cls, expr, coefficients=None, variables=None):
        if len(expr) == 0:
            raise InvalidPrefixExpression("Empty prefix list.")
        head, rest = expr[0], expr[1:]
        if head in cls.OPERATORS:
            args = []
            for _ in range(cls.OPERATORS[head]):
                arg, rest = cls._prefix_to_infix(rest, coefficients, variables)
                args.append(arg)
            return cls.write_infix(head, args), rest
        if coefficients is not None and head in coefficients:
            return "{" + head + "}", rest
        return str(head), rest  # variable, pi/E/I, or an integer word

    @classmethod
    def prefix_to_infix(cls, expr, coefficients=None, variables=None):
        infix, rest = cls._prefix_to_infix(list(expr), coefficients, variables)
        if len(rest) > 0:
            raise InvalidPrefixExpression(
                f'Incorrect prefix expression "{expr}". "{rest}" was not parsed.')
        return f"({infix})"

    @classmethod
    def sympy_to_prefix(cls, expr):
        if isinstance(expr, sp.Symbol):
            return [str(expr)]
        if isinstance(expr, sp.Integer):
            return [str(expr)]
        if isinstance(expr, sp.Rational):
            return ["div", str(expr.p), str(expr.q)]
        if expr == sp.E:
            return ["E"]
        if expr == sp.pi:
            return ["pi"]
        if expr == sp.I:
            return ["I"]
        for node_type, word in cls.SYMPY_OPERATORS.items():
            if isinstance(expr, node_type):
                return cls._sympy_to_prefix(word, expr)
        raise UnknownSymPyOperator(f"Unknown SymPy operator: {expr}")

    @classmethod
    def _sympy_to_prefix(cls, op, expr):
        args = expr.args
        if op == "pow" and args[1] == sp.Rational(1, 2):
            return ["sqrt"] + cls.sympy_to_prefix(args[0])
        # n-ary add/mul become right-nested binary nodes
        out = []
        for i, a in enumerate(args):
            if i == 0 or i < len(args) - 1:
                out.append(op)
            out += cls.sympy_to_prefix(a)
        return out
