"""Scalar expression DAG with hash-consing, symbolic differentiation and C emission.

This is the symbolic core that replaces the part of CasADi's SX the reference's hot path relies on
(reference do_mpc/model/_model.py:937-1060 builds `_rhs_fun`, `_alg_fun`, ... from SX graphs; CasADi then
differentiates the flat NLP inside nlpsol, do_mpc/controller/_mpc.py:1328).  Here the graph is kept at
*stage* granularity and differentiated per stage atom; the result is emitted as straight-line CUDA.

Design: every node is interned (structural hashing), so common sub-expressions are shared by
construction and derivative graphs re-use the primal's nodes.  Nodes are immutable tuples
(op, args) identified by an integer id in a global table; Python-side handles are `Ex` objects.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Sequence

# op codes
CONST, SYM, ADD, SUB, MUL, DIV, NEG, POW, SQ, SQRT, EXP, LOG, SIN, COS, TAN, TANH, FABS, ATAN, ASIN, ACOS, \
    SINH, COSH, SIGN, FMIN, FMAX, ATAN2, LT, LE, EQ, NE, IFELSE, ERF, FLOOR, CEIL, NOT, AND, OR = range(37)

_UNARY_C = {NEG: "-({0})", SQ: "(({0})*({0}))", SQRT: "sqrt({0})", EXP: "exp({0})", LOG: "log({0})", SIN: "sin({0})",
            COS: "cos({0})", TAN: "tan({0})", TANH: "tanh({0})", FABS: "fabs({0})", ATAN: "atan({0})",
            ASIN: "asin({0})", ACOS: "acos({0})", SINH: "sinh({0})", COSH: "cosh({0})",
            SIGN: "((({0})>0.0)-(({0})<0.0))", ERF: "erf({0})", FLOOR: "floor({0})", CEIL: "ceil({0})",
            NOT: "(double)(({0})==0.0)"}
_BINARY_C = {ADD: "({0}+{1})", SUB: "({0}-{1})", MUL: "({0}*{1})", DIV: "({0}/{1})", POW: "pow({0},{1})",
             FMIN: "fmin({0},{1})", FMAX: "fmax({0},{1})", ATAN2: "atan2({0},{1})",
             LT: "(double)({0}<{1})", LE: "(double)({0}<={1})", EQ: "(double)({0}=={1})", NE: "(double)({0}!={1})",
             AND: "(double)(({0}!=0.0)&&({1}!=0.0))", OR: "(double)(({0}!=0.0)||({1}!=0.0))"}
_UNARY_PY = {NEG: lambda a: -a, SQ: lambda a: a * a, SQRT: math.sqrt, EXP: math.exp, LOG: math.log, SIN: math.sin,
             COS: math.cos, TAN: math.tan, TANH: math.tanh, FABS: abs, ATAN: math.atan, ASIN: math.asin,
             ACOS: math.acos, SINH: math.sinh, COSH: math.cosh, SIGN: lambda a: (a > 0) - (a < 0), ERF: math.erf,
             FLOOR: math.floor, CEIL: math.ceil, NOT: lambda a: float(a == 0.0)}
_BINARY_PY = {ADD: lambda a, b: a + b, SUB: lambda a, b: a - b, MUL: lambda a, b: a * b, DIV: lambda a, b: a / b,
              POW: lambda a, b: a ** b, FMIN: min, FMAX: max, ATAN2: math.atan2,
              LT: lambda a, b: float(a < b), LE: lambda a, b: float(a <= b), EQ: lambda a, b: float(a == b),
              NE: lambda a, b: float(a != b), AND: lambda a, b: float(a != 0 and b != 0),
              OR: lambda a, b: float(a != 0 or b != 0)}


class _Graph:
    """Global interned node store."""

    def __init__(self):
        self.ops: List[int] = []
        self.args: List[tuple] = []
        self.table: Dict[tuple, int] = {}
        self.sym_count = 0

    def intern(self, op, args):
        key = (op, args)
        i = self.table.get(key)
        if i is None:
            i = len(self.ops)
            self.ops.append(op)
            self.args.append(args)
            self.table[key] = i
        return i


G = _Graph()


def _const_id(v: float) -> int:
    v = float(v)
    if v == 0.0:
        v = 0.0  # merge -0.0
    return G.intern(CONST, (v,))


ZERO = _const_id(0.0)
ONE = _const_id(1.0)
TWO = _const_id(2.0)
MINUS_ONE = _const_id(-1.0)


def is_const(i): return G.ops[i] == CONST
def const_val(i): return G.args[i][0]


def mk_sym(name: str) -> int:
    G.sym_count += 1
    return G.intern(SYM, (name, G.sym_count))


def mk(op, *a) -> int:
    """Build a node with local algebraic simplification (constant folding, neutral elements)."""
    if all(is_const(x) for x in a):
        try:
            if len(a) == 1:
                return _const_id(_UNARY_PY[op](const_val(a[0])))
            if len(a) == 2:
                return _const_id(_BINARY_PY[op](const_val(a[0]), const_val(a[1])))
            if op == IFELSE:
                return a[1] if const_val(a[0]) != 0 else a[2]
        except (ValueError, ZeroDivisionError, OverflowError):
            pass
    if op == ADD:
        x, y = a
        if x == ZERO: return y
        if y == ZERO: return x
        if G.ops[y] == NEG: return mk(SUB, x, G.args[y][0])
        if G.ops[x] == NEG: return mk(SUB, y, G.args[x][0])
        if x == y: return mk(MUL, TWO, x)
        if x > y: x, y = y, x           # canonical order for commutative ops
        return G.intern(ADD, (x, y))
    if op == SUB:
        x, y = a
        if y == ZERO: return x
        if x == ZERO: return mk(NEG, y)
        if x == y: return ZERO
        if G.ops[y] == NEG: return mk(ADD, x, G.args[y][0])
        return G.intern(SUB, (x, y))
    if op == MUL:
        x, y = a
        if x == ZERO or y == ZERO: return ZERO
        if x == ONE: return y
        if y == ONE: return x
        if x == MINUS_ONE: return mk(NEG, y)
        if y == MINUS_ONE: return mk(NEG, x)
        if x == y: return mk(SQ, x)
        if G.ops[x] == NEG and G.ops[y] == NEG: return mk(MUL, G.args[x][0], G.args[y][0])
        if G.ops[x] == NEG: return mk(NEG, mk(MUL, G.args[x][0], y))
        if G.ops[y] == NEG: return mk(NEG, mk(MUL, x, G.args[y][0]))
        if x > y: x, y = y, x
        return G.intern(MUL, (x, y))
    if op == DIV:
        x, y = a
        if x == ZERO: return ZERO
        if y == ONE: return x
        if y == MINUS_ONE: return mk(NEG, x)
        if x == y: return ONE
        if G.ops[x] == NEG and G.ops[y] == NEG: return mk(DIV, G.args[x][0], G.args[y][0])
        if G.ops[x] == NEG: return mk(NEG, mk(DIV, G.args[x][0], y))
        if G.ops[y] == NEG: return mk(NEG, mk(DIV, x, G.args[y][0]))
        return G.intern(DIV, (x, y))
    if op == NEG:
        (x,) = a
        if G.ops[x] == NEG: return G.args[x][0]
        if G.ops[x] == SUB: return mk(SUB, G.args[x][1], G.args[x][0])
        return G.intern(NEG, (x,))
    if op == POW:
        x, y = a
        if is_const(y):
            v = const_val(y)
            if v == 0.0: return ONE
            if v == 1.0: return x
            if v == 2.0: return mk(SQ, x)
            if v == 0.5: return mk(SQRT, x)
            if v == -1.0: return mk(DIV, ONE, x)
            if v == 3.0: return mk(MUL, mk(SQ, x), x)
            if v == 4.0: return mk(SQ, mk(SQ, x))
        return G.intern(POW, (x, y))
    if op == SQ:
        (x,) = a
        if G.ops[x] == NEG: return mk(SQ, G.args[x][0])
        return G.intern(SQ, (x,))
    if op in (FMIN, FMAX) and a[0] == a[1]:
        return a[0]
    return G.intern(op, tuple(a))


# ----------------------------------------------------------------------------------------------------
# graph utilities
# ----------------------------------------------------------------------------------------------------
def topo_order(roots: Iterable[int]) -> List[int]:
    """Nodes reachable from roots, children before parents (ids are creation-ordered, so sorting works)."""
    seen = set()
    stack = list(roots)
    while stack:
        i = stack.pop()
        if i in seen:
            continue
        seen.add(i)
        if G.ops[i] not in (CONST, SYM):
            stack.extend(G.args[i])
    return sorted(seen)


def depends_on(root: int, syms: Iterable[int]) -> bool:
    s = set(syms)
    return any(i in s for i in topo_order([root]))


def symbols_of(roots: Iterable[int]) -> List[int]:
    return [i for i in topo_order(roots) if G.ops[i] == SYM]


def substitute(roots: Sequence[int], old: Sequence[int], new: Sequence[int]) -> List[int]:
    m = dict(zip(old, new))
    order = topo_order(roots)
    out: Dict[int, int] = {}
    for i in order:
        if i in m:
            out[i] = m[i]
        elif G.ops[i] in (CONST, SYM):
            out[i] = i
        else:
            na = tuple(out[a] for a in G.args[i])
            out[i] = i if na == G.args[i] else mk(G.ops[i], *na)
    return [out[r] for r in roots]


def evaluate(roots: Sequence[int], env: Dict[int, float]) -> List[float]:
    val: Dict[int, float] = {}
    for i in topo_order(roots):
        op = G.ops[i]
        if op == CONST:
            val[i] = const_val(i)
        elif op == SYM:
            val[i] = env[i]
        else:
            a = [val[x] for x in G.args[i]]
            if len(a) == 1:
                val[i] = _UNARY_PY[op](a[0])
            elif len(a) == 2:
                val[i] = _BINARY_PY[op](a[0], a[1])
            else:
                val[i] = a[1] if a[0] != 0 else a[2]
    return [val[r] for r in roots]


# ----------------------------------------------------------------------------------------------------
# differentiation (reverse mode on the DAG; returns expression ids)
# ----------------------------------------------------------------------------------------------------
def _partials(i):
    """d node / d arg_k as expression ids."""
    op = G.ops[i]
    a = G.args[i]
    if op == ADD: return (ONE, ONE)
    if op == SUB: return (ONE, MINUS_ONE)
    if op == MUL: return (a[1], a[0])
    if op == DIV: return (mk(DIV, ONE, a[1]), mk(NEG, mk(DIV, i, a[1])))
    if op == NEG: return (MINUS_ONE,)
    if op == SQ: return (mk(MUL, TWO, a[0]),)
    if op == SQRT: return (mk(DIV, _const_id(0.5), i),)
    if op == EXP: return (i,)
    if op == LOG: return (mk(DIV, ONE, a[0]),)
    if op == SIN: return (mk(COS, a[0]),)
    if op == COS: return (mk(NEG, mk(SIN, a[0])),)
    if op == TAN: return (mk(ADD, ONE, mk(SQ, i)),)
    if op == TANH: return (mk(SUB, ONE, mk(SQ, i)),)
    if op == FABS: return (mk(SIGN, a[0]),)
    if op == ATAN: return (mk(DIV, ONE, mk(ADD, ONE, mk(SQ, a[0]))),)
    if op == ASIN: return (mk(DIV, ONE, mk(SQRT, mk(SUB, ONE, mk(SQ, a[0])))),)
    if op == ACOS: return (mk(NEG, mk(DIV, ONE, mk(SQRT, mk(SUB, ONE, mk(SQ, a[0]))))),)
    if op == SINH: return (mk(COSH, a[0]),)
    if op == COSH: return (mk(SINH, a[0]),)
    if op == ERF: return (mk(MUL, _const_id(2.0 / math.sqrt(math.pi)), mk(EXP, mk(NEG, mk(SQ, a[0])))),)
    if op == POW:
        x, y = a
        dx = mk(MUL, y, mk(POW, x, mk(SUB, y, ONE)))
        dy = ZERO if is_const(y) else mk(MUL, i, mk(LOG, x))
        return (dx, dy)
    if op == FMIN: return (mk(LE, a[0], a[1]), mk(LT, a[1], a[0]))
    if op == FMAX: return (mk(LE, a[1], a[0]), mk(LT, a[0], a[1]))
    if op == ATAN2:
        den = mk(ADD, mk(SQ, a[0]), mk(SQ, a[1]))
        return (mk(DIV, a[1], den), mk(NEG, mk(DIV, a[0], den)))
    if op == IFELSE: return (ZERO, a[0], mk(NOT, a[0]))
    if op in (SIGN, LT, LE, EQ, NE, FLOOR, CEIL, NOT, AND, OR):
        return tuple(ZERO for _ in a)
    raise NotImplementedError(op)


def gradient(root: int, wrt: Sequence[int]) -> List[int]:
    """Reverse-mode gradient of a scalar expression."""
    order = topo_order([root])
    adj: Dict[int, int] = {root: ONE}
    for i in reversed(order):
        ai = adj.get(i)
        if ai is None or ai == ZERO or G.ops[i] in (CONST, SYM):
            continue
        for arg, part in zip(G.args[i], _partials(i)):
            if part == ZERO:
                continue
            contrib = mk(MUL, ai, part)
            prev = adj.get(arg)
            adj[arg] = contrib if prev is None else mk(ADD, prev, contrib)
    return [adj.get(w, ZERO) for w in wrt]


def jacobian(roots: Sequence[int], wrt: Sequence[int]) -> List[List[int]]:
    return [gradient(r, wrt) for r in roots]


def hessian(root: int, wrt: Sequence[int]) -> List[List[int]]:
    g = gradient(root, wrt)
    return [gradient(gi, wrt) for gi in g]


# ----------------------------------------------------------------------------------------------------
# C emission
# ----------------------------------------------------------------------------------------------------
def _fmt_const(v: float) -> str:
    if v != v: return "NAN"
    if v == math.inf: return "INFINITY"
    if v == -math.inf: return "(-INFINITY)"
    s = repr(float(v))
    if "e" not in s and "." not in s and "inf" not in s:
        s += ".0"
    return s


def emit_c(outputs: Sequence[tuple], inputs: Dict[int, str], tmp_prefix: str = "t", indent: str = "  ") -> str:
    """Straight-line C for `lhs = expr;` assignments.

    outputs: sequence of (lhs_string, expr_id).  inputs: symbol id -> C expression (e.g. 'x[3]').
    Single-use interior nodes are inlined; shared ones get a `const double tN`.
    """
    roots = [e for _, e in outputs]
    order = topo_order(roots)
    uses: Dict[int, int] = {}
    for i in order:
        if G.ops[i] not in (CONST, SYM):
            for a in G.args[i]:
                uses[a] = uses.get(a, 0) + (2 if G.ops[i] in (SQ, SIGN) else 1)
    for r in roots:
        uses[r] = uses.get(r, 0) + 1
    text: Dict[int, str] = {}
    lines: List[str] = []
    n_tmp = 0
    for i in order:
        op = G.ops[i]
        if op == CONST:
            v = const_val(i)
            text[i] = _fmt_const(v) if v >= 0 else f"({_fmt_const(v)})"
            continue
        if op == SYM:
            if i not in inputs:
                raise KeyError(f"free symbol {G.args[i][0]} is not an input of the generated function")
            text[i] = inputs[i]
            continue
        a = [text[x] for x in G.args[i]]
        if op in _UNARY_C:
            s = _UNARY_C[op].format(*a)
        elif op in _BINARY_C:
            s = _BINARY_C[op].format(*a)
        elif op == IFELSE:
            s = f"(({a[0]})!=0.0?({a[1]}):({a[2]}))"
        else:
            raise NotImplementedError(op)
        if uses.get(i, 0) > 1 or len(s) > 120:
            name = f"{tmp_prefix}{n_tmp}"
            n_tmp += 1
            lines.append(f"{indent}const double {name} = {s};")
            text[i] = name
        else:
            text[i] = s
    for lhs, e in outputs:
        lines.append(f"{indent}{lhs} = {text[e]};")
    return "\n".join(lines)


def count_ops(roots: Sequence[int]) -> int:
    return sum(1 for i in topo_order(roots) if G.ops[i] not in (CONST, SYM))
