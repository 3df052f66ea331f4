"""Matrix-valued symbolic type `SX`, numeric `DM` and `Function` with the CasADi surface do-mpc uses.

Only the API that the reference's Model/MPC path and its example templates touch is provided
(SURVEY.md App. D): SX.sym, arithmetic with numpy broadcasting of scalars, `@`, `.T`, indexing,
vertcat/horzcat/sum1/sum2/reshape, elementary functions, jacobian/gradient/hessian/substitute, Function.
Column-major flattening like CasADi (`vec`), matrices are dense lists of expression ids.
"""
from __future__ import annotations

import numbers
from typing import List, Sequence

import numpy as np

from . import expr as E


def _is_num(v):
    return isinstance(v, (numbers.Number, np.number, np.bool_))


class SX:
    """Dense matrix of scalar expressions (row-major storage, CasADi semantics on the outside)."""
    __array_priority__ = 1000.0
    __array_ufunc__ = None          # make numpy defer to our __r*__ operators (A @ x with ndarray A)

    def __init__(self, data=0.0, cols=None):
        if isinstance(data, SX):
            self.shape = data.shape
            self._d = list(data._d)
        elif isinstance(data, DM):
            a = data.full()
            self.shape = a.shape
            self._d = [E._const_id(v) for v in a.ravel()]
        elif _is_num(data) and cols is not None:          # SX(n, m) zeros
            self.shape = (int(data), int(cols))
            self._d = [E.ZERO] * (self.shape[0] * self.shape[1])
        elif _is_num(data):
            self.shape = (1, 1)
            self._d = [E._const_id(data)]
        else:
            a = np.asarray(data, dtype=object)
            if a.ndim == 0:
                a = a.reshape(1, 1)
            elif a.ndim == 1:
                a = a.reshape(-1, 1)
            self.shape = a.shape
            self._d = []
            for v in a.ravel():
                if isinstance(v, SX):
                    assert v.shape == (1, 1)
                    self._d.append(v._d[0])
                else:
                    self._d.append(E._const_id(float(v)))

    # -- construction ----------------------------------------------------------------------------
    @staticmethod
    def _from(ids: Sequence[int], shape) -> "SX":
        o = SX.__new__(SX)
        o.shape = (int(shape[0]), int(shape[1]))
        o._d = list(ids)
        assert len(o._d) == o.shape[0] * o.shape[1]
        return o

    @staticmethod
    def sym(name: str, n: int = 1, m: int = 1) -> "SX":
        if isinstance(n, (tuple, list)):
            n, m = n
        n, m = int(n), int(m)
        if n * m == 1:
            ids = [E.mk_sym(name)]
        else:
            ids = [E.mk_sym(f"{name}_{i}") for i in range(n * m)]
            # CasADi numbers column-major; keep element (i,j) <-> name_{i+j*n}
            ids = [ids[i + j * n] for i in range(n) for j in range(m)]
        return SX._from(ids, (n, m))

    @staticmethod
    def zeros(n=1, m=1):
        if isinstance(n, (tuple, list)): n, m = n
        return SX._from([E.ZERO] * (int(n) * int(m)), (n, m))

    @staticmethod
    def ones(n=1, m=1):
        if isinstance(n, (tuple, list)): n, m = n
        return SX._from([E.ONE] * (int(n) * int(m)), (n, m))

    @staticmethod
    def eye(n):
        return SX._from([E.ONE if i == j else E.ZERO for i in range(n) for j in range(n)], (n, n))

    # -- basic properties --------------------------------------------------------------------------
    def size1(self): return self.shape[0]
    def size2(self): return self.shape[1]
    def numel(self): return self.shape[0] * self.shape[1]
    def size(self, *a): return self.shape if not a else self.shape[a[0] - 1]
    def is_scalar(self): return self.shape == (1, 1)
    def is_empty(self): return self.numel() == 0
    def __len__(self): return self.shape[0]
    def elements(self) -> List["SX"]: return [SX._from([i], (1, 1)) for i in self.vec_ids()]
    def nz(self): return self.elements()
    def is_constant(self): return all(E.is_const(i) for i in self._d)
    def is_symbolic(self): return all(E.G.ops[i] == E.SYM for i in self._d)
    def name(self):
        assert self.shape == (1, 1) and E.G.ops[self._d[0]] == E.SYM
        return E.G.args[self._d[0]][0]

    def vec_ids(self) -> List[int]:
        """Column-major list of element ids (CasADi `vec`)."""
        n, m = self.shape
        return [self._d[i * m + j] for j in range(m) for i in range(n)]

    def ids(self) -> List[int]:
        return self.vec_ids()

    @property
    def T(self):
        n, m = self.shape
        return SX._from([self._d[i * m + j] for j in range(m) for i in range(n)], (m, n))

    def __float__(self):
        assert self.shape == (1, 1) and E.is_const(self._d[0]), "not a constant scalar"
        return E.const_val(self._d[0])

    def to_DM(self):
        assert self.is_constant()
        return DM(np.array([E.const_val(i) for i in self._d]).reshape(self.shape))

    def __repr__(self):
        return f"SX({self.shape[0]}x{self.shape[1]})"

    def __hash__(self):
        return hash((self.shape, tuple(self._d)))

    # -- indexing -------------------------------------------------------------------------------------
    def _index(self, key):
        n, m = self.shape
        if not isinstance(key, tuple):
            # single index: linear (column-major) for matrices, row for column vectors
            lin = np.arange(n * m).reshape(m, n).T         # lin[i,j] = i + j*n
            flat = lin.T.ravel()                             # column-major order of linear ids
            sel = np.atleast_1d(flat[key] if not isinstance(key, (list, np.ndarray)) else flat[np.asarray(key)])
            ids = [self._d[(k % n) * m + (k // n)] for k in sel]
            return ids, (len(ids), 1)
        r, c = key
        rows = np.atleast_1d(np.arange(n)[r])
        cols = np.atleast_1d(np.arange(m)[c])
        ids = [self._d[i * m + j] for i in rows for j in cols]
        return ids, (len(rows), len(cols))

    def __getitem__(self, key):
        ids, shp = self._index(key)
        return SX._from(ids, shp)

    def __setitem__(self, key, value):
        n, m = self.shape
        value = value if isinstance(value, SX) else SX(value)
        if not isinstance(key, tuple):
            key = (key, slice(None)) if m == 1 else key
        if isinstance(key, tuple):
            rows = np.atleast_1d(np.arange(n)[key[0]]); cols = np.atleast_1d(np.arange(m)[key[1]])
            pos = [i * m + j for i in rows for j in cols]
        else:
            sel = np.atleast_1d(np.arange(n * m)[key])
            pos = [(k % n) * m + (k // n) for k in sel]
        vals = value._d if value.numel() == len(pos) else value._d * len(pos)
        assert len(vals) == len(pos)
        for p, v in zip(pos, vals):
            self._d[p] = v

    def __iter__(self):
        for i in range(self.shape[0]):
            yield self[i, :] if self.shape[1] > 1 else self[i]

    # -- arithmetic -------------------------------------------------------------------------------------
    @staticmethod
    def _coerce(o):
        if isinstance(o, SX):
            return o
        if isinstance(o, DM):
            return SX(o)
        if _is_num(o):
            return SX(o)
        a = np.asarray(o)
        if a.dtype == object:
            return SX(a)
        return SX(DM(a))

    def _binary(self, other, op, swap=False):
        other = SX._coerce(other)
        a, b = (other, self) if swap else (self, other)
        if a.shape == b.shape:
            return SX._from([E.mk(op, x, y) for x, y in zip(a._d, b._d)], a.shape)
        if a.numel() == 1:
            return SX._from([E.mk(op, a._d[0], y) for y in b._d], b.shape)
        if b.numel() == 1:
            return SX._from([E.mk(op, x, b._d[0]) for x in a._d], a.shape)
        raise ValueError(f"shape mismatch {a.shape} vs {b.shape}")

    def _unary(self, op):
        return SX._from([E.mk(op, x) for x in self._d], self.shape)

    def __add__(self, o): return self._binary(o, E.ADD)
    def __radd__(self, o): return self._binary(o, E.ADD, True)
    def __sub__(self, o): return self._binary(o, E.SUB)
    def __rsub__(self, o): return self._binary(o, E.SUB, True)
    def __mul__(self, o): return self._binary(o, E.MUL)
    def __rmul__(self, o): return self._binary(o, E.MUL, True)
    def __truediv__(self, o): return self._binary(o, E.DIV)
    def __rtruediv__(self, o): return self._binary(o, E.DIV, True)
    def __pow__(self, o): return self._binary(o, E.POW)
    def __rpow__(self, o): return self._binary(o, E.POW, True)
    def __neg__(self): return self._unary(E.NEG)
    def __pos__(self): return self
    def __abs__(self): return self._unary(E.FABS)
    def __lt__(self, o): return self._binary(o, E.LT)
    def __le__(self, o): return self._binary(o, E.LE)
    def __gt__(self, o): return self._binary(o, E.LT, True)
    def __ge__(self, o): return self._binary(o, E.LE, True)
    def __eq__(self, o):
        try:
            return self._binary(o, E.EQ)
        except Exception:
            return NotImplemented
    def __ne__(self, o): return self._binary(o, E.NE)

    def __matmul__(self, o):
        return mtimes(self, o)

    def __rmatmul__(self, o):
        return mtimes(o, self)

    # CasADi-style method aliases used by templates (exp(x) resolves through numpy-like free functions below)
    def exp(self): return self._unary(E.EXP)
    def log(self): return self._unary(E.LOG)
    def sqrt(self): return self._unary(E.SQRT)
    def sin(self): return self._unary(E.SIN)
    def cos(self): return self._unary(E.COS)
    def tan(self): return self._unary(E.TAN)
    def tanh(self): return self._unary(E.TANH)
    def fabs(self): return self._unary(E.FABS)
    def arctan(self): return self._unary(E.ATAN)
    def arcsin(self): return self._unary(E.ASIN)
    def arccos(self): return self._unary(E.ACOS)
    def sinh(self): return self._unary(E.SINH)
    def cosh(self): return self._unary(E.COSH)


class DM:
    """Numeric dense matrix (thin numpy wrapper) with the bits of casadi.DM the reference touches."""
    __array_priority__ = 900.0

    def __init__(self, data=0.0, cols=None):
        if isinstance(data, DM):
            self._a = data._a.copy()
        elif _is_num(data) and cols is not None:
            self._a = np.zeros((int(data), int(cols)))
        else:
            a = np.array(data, dtype=float)
            if a.ndim == 0: a = a.reshape(1, 1)
            elif a.ndim == 1: a = a.reshape(-1, 1)
            self._a = a

    @property
    def shape(self): return self._a.shape
    def full(self): return self._a.copy()
    def __array__(self, dtype=None, copy=None): return self._a if dtype is None else self._a.astype(dtype)
    def __float__(self): return float(self._a.reshape(-1)[0]) if self._a.size == 1 else float(self._a)
    def __len__(self): return self._a.shape[0]
    def numel(self): return self._a.size
    def size1(self): return self._a.shape[0]
    def size2(self): return self._a.shape[1]
    @property
    def T(self): return DM(self._a.T)
    def __repr__(self): return f"DM({self._a.tolist()})"
    def __getitem__(self, k):
        if not isinstance(k, tuple):
            flat = self._a.T.ravel()[k]      # column-major linear index
            return DM(np.atleast_1d(flat).reshape(-1, 1))
        return DM(np.atleast_2d(self._a[k]))
    def __setitem__(self, k, v):
        v = v.full() if isinstance(v, DM) else np.asarray(v, dtype=float)
        if not isinstance(k, tuple) and self._a.shape[1] == 1:
            self._a[k, 0] = np.asarray(v).reshape(-1) if np.size(v) > 1 else float(np.asarray(v).reshape(-1)[0])
        else:
            self._a[k] = v
    def _b(self, o, f):
        if isinstance(o, SX): return NotImplemented
        o = o._a if isinstance(o, DM) else o
        return DM(f(self._a, o))
    def __add__(self, o): return self._b(o, np.add)
    def __radd__(self, o): return self._b(o, lambda a, b: b + a)
    def __sub__(self, o): return self._b(o, np.subtract)
    def __rsub__(self, o): return self._b(o, lambda a, b: b - a)
    def __mul__(self, o): return self._b(o, np.multiply)
    def __rmul__(self, o): return self._b(o, lambda a, b: b * a)
    def __truediv__(self, o): return self._b(o, np.divide)
    def __rtruediv__(self, o): return self._b(o, lambda a, b: b / a)
    def __pow__(self, o): return self._b(o, np.power)
    def __neg__(self): return DM(-self._a)
    def __matmul__(self, o):
        if isinstance(o, SX): return NotImplemented
        return DM(self._a @ (o._a if isinstance(o, DM) else np.asarray(o)))
    def __eq__(self, o): return self._a == (o._a if isinstance(o, DM) else o)
    def __lt__(self, o): return self._a < (o._a if isinstance(o, DM) else o)
    def __gt__(self, o): return self._a > (o._a if isinstance(o, DM) else o)
    def __le__(self, o): return self._a <= (o._a if isinstance(o, DM) else o)
    def __ge__(self, o): return self._a >= (o._a if isinstance(o, DM) else o)


# ----------------------------------------------------------------------------------------------------
# free functions (casadi namespace)
# ----------------------------------------------------------------------------------------------------
def _sx(o): return SX._coerce(o)


def _un(op, npf):
    def f(x):
        if isinstance(x, SX): return x._unary(op)
        if isinstance(x, DM): return DM(npf(x._a))
        return npf(x)
    return f


exp = _un(E.EXP, np.exp); log = _un(E.LOG, np.log); sqrt = _un(E.SQRT, np.sqrt)
sin = _un(E.SIN, np.sin); cos = _un(E.COS, np.cos); tan = _un(E.TAN, np.tan); tanh = _un(E.TANH, np.tanh)
fabs = _un(E.FABS, np.abs); atan = _un(E.ATAN, np.arctan); asin = _un(E.ASIN, np.arcsin)
acos = _un(E.ACOS, np.arccos); sinh = _un(E.SINH, np.sinh); cosh = _un(E.COSH, np.cosh)
sign = _un(E.SIGN, np.sign); floor = _un(E.FLOOR, np.floor); ceil = _un(E.CEIL, np.ceil)
arctan, arcsin, arccos = atan, asin, acos


def _bin(op, npf):
    def f(x, y):
        if isinstance(x, SX) or isinstance(y, SX): return _sx(x)._binary(y, op)
        return npf(np.asarray(x, dtype=float), np.asarray(y, dtype=float))
    return f


fmin = _bin(E.FMIN, np.minimum); fmax = _bin(E.FMAX, np.maximum); atan2 = _bin(E.ATAN2, np.arctan2)
power = _bin(E.POW, np.power); arctan2 = atan2


def if_else(c, a, b):
    c, a, b = _sx(c), _sx(a), _sx(b)
    n = max(c.numel(), a.numel(), b.numel())
    shp = next(v.shape for v in (a, b, c) if v.numel() == n)
    g = lambda v, i: v._d[i] if v.numel() > 1 else v._d[0]
    return SX._from([E.mk(E.IFELSE, g(c, i), g(a, i), g(b, i)) for i in range(n)], shp)


def mtimes(*args):
    if len(args) == 1 and isinstance(args[0], (list, tuple)):
        args = args[0]
    out = args[0]
    for b in args[1:]:
        out = _mtimes2(out, b)
    return out


def _mtimes2(a, b):
    if not isinstance(a, SX) and not isinstance(b, SX):
        return DM(np.asarray(a) @ np.asarray(b))
    a, b = _sx(a), _sx(b)
    if a.numel() == 1 or b.numel() == 1:
        return a * b
    assert a.shape[1] == b.shape[0], f"mtimes shape mismatch {a.shape} x {b.shape}"
    n, k, m = a.shape[0], a.shape[1], b.shape[1]
    out = []
    for i in range(n):
        for j in range(m):
            acc = E.ZERO
            for l in range(k):
                acc = E.mk(E.ADD, acc, E.mk(E.MUL, a._d[i * k + l], b._d[l * m + j]))
            out.append(acc)
    return SX._from(out, (n, m))


def dot(a, b):
    a, b = _sx(a), _sx(b)
    acc = E.ZERO
    for x, y in zip(a.vec_ids(), b.vec_ids()):
        acc = E.mk(E.ADD, acc, E.mk(E.MUL, x, y))
    return SX._from([acc], (1, 1))


def vertcat(*args):
    if len(args) == 1 and isinstance(args[0], (list, tuple)):
        args = args[0]
    if len(args) == 0:
        return SX.zeros(0, 1)
    if not any(isinstance(a, SX) for a in args):
        return DM(np.vstack([np.atleast_2d(DM(a).full()) for a in args if np.size(DM(a).full())] or [np.zeros((0, 1))]))
    mats = [_sx(a) for a in args]
    mats = [m for m in mats if m.numel() > 0] or [SX.zeros(0, 1)]
    m = mats[0].shape[1]
    ids = []
    for a in mats:
        assert a.shape[1] == m, "vertcat column mismatch"
        ids.extend(a._d)
    return SX._from(ids, (sum(a.shape[0] for a in mats), m))


def horzcat(*args):
    if len(args) == 1 and isinstance(args[0], (list, tuple)):
        args = args[0]
    if not any(isinstance(a, SX) for a in args):
        return DM(np.hstack([np.atleast_2d(DM(a).full()) for a in args]))
    mats = [_sx(a) for a in args]
    return vertcat(*[a.T for a in mats]).T


def veccat(*args):
    if len(args) == 1 and isinstance(args[0], (list, tuple)):
        args = args[0]
    return vertcat(*[vec(a) for a in args])


def vec(a):
    if not isinstance(a, SX):
        return DM(np.asarray(DM(a).full()).T.reshape(-1, 1))
    return SX._from(a.vec_ids(), (a.numel(), 1))


def reshape(a, *shape):
    if len(shape) == 1: shape = shape[0]
    n, m = shape
    a = _sx(a)
    v = a.vec_ids()
    assert n * m == len(v)
    return SX._from([v[i + j * n] for i in range(n) for j in range(m)], (n, m))


def sum1(a):
    if not isinstance(a, SX): return DM(np.sum(DM(a).full(), axis=0, keepdims=True))
    n, m = a.shape
    out = []
    for j in range(m):
        acc = E.ZERO
        for i in range(n):
            acc = E.mk(E.ADD, acc, a._d[i * m + j])
        out.append(acc)
    return SX._from(out, (1, m))


def sum2(a):
    if not isinstance(a, SX): return DM(np.sum(DM(a).full(), axis=1, keepdims=True))
    return sum1(a.T).T


def sumsqr(a):
    a = _sx(a)
    return sum1(vec(a * a))


def norm_2(a):
    return sqrt(sumsqr(a))


def transpose(a): return a.T
def diag(a):
    a = _sx(a)
    if a.shape[1] == 1 or a.shape[0] == 1:
        v = a.vec_ids(); n = len(v)
        return SX._from([v[i] if i == j else E.ZERO for i in range(n) for j in range(n)], (n, n))
    n = min(a.shape)
    return SX._from([a._d[i * a.shape[1] + i] for i in range(n)], (n, 1))


def jacobian(ex, wrt):
    ex, wrt = _sx(ex), _sx(wrt)
    rows = E.jacobian(ex.vec_ids(), wrt.vec_ids())
    return SX._from([v for r in rows for v in r], (ex.numel(), wrt.numel()))


def gradient(ex, wrt):
    ex, wrt = _sx(ex), _sx(wrt)
    assert ex.numel() == 1
    return SX._from(E.gradient(ex._d[0], wrt.vec_ids()), (wrt.numel(), 1))


def hessian(ex, wrt):
    ex, wrt = _sx(ex), _sx(wrt)
    g = E.gradient(ex._d[0], wrt.vec_ids())
    H = [E.gradient(gi, wrt.vec_ids()) for gi in g]
    n = wrt.numel()
    return SX._from([v for r in H for v in r], (n, n)), SX._from(g, (n, 1))


def tangent(ex, wrt):
    return jacobian(ex, wrt)


def substitute(ex, old, new):
    if isinstance(ex, (list, tuple)):
        return [substitute(e, old, new) for e in ex]
    ex = _sx(ex)
    old_ids, new_ids = _sx(old).vec_ids(), _sx(new).vec_ids()
    ids = E.substitute(ex._d, old_ids, new_ids)
    return SX._from(ids, ex.shape)


def depends_on(ex, syms):
    return any(E.depends_on(i, _sx(syms).vec_ids()) for i in _sx(ex)._d)


def symvar(ex):
    return [SX._from([i], (1, 1)) for i in E.symbols_of(_sx(ex)._d)]


def collocation_points(deg, scheme="radau"):
    """casadi.collocation_points (used at reference do_mpc/optimizer.py:845-847)."""
    from numpy.polynomial import legendre as npleg
    one = [0.0] * deg + [1.0]
    if scheme == "legendre":
        c = np.array(one)
    elif scheme == "radau":
        c = np.array(one) - np.array([0.0] * (deg - 1) + [1.0, 0.0])
    else:
        raise Exception("Unknown collocation scheme")
    roots = np.sort(npleg.legroots(c).real)
    return [float(r) for r in (roots + 1.0) / 2.0]


class Function:
    """casadi.Function look-alike: callable on numeric (-> DM) or symbolic (-> SX) arguments."""

    def __init__(self, name, ins, outs, in_names=None, out_names=None, opts=None):
        self._name = name
        self.ins = [_sx(getattr(i, "cat", i)) for i in ins]
        self.outs = [_sx(getattr(o, "cat", o)) for o in outs]
        self.in_names = list(in_names) if in_names else [f"i{k}" for k in range(len(ins))]
        self.out_names = list(out_names) if out_names else [f"o{k}" for k in range(len(outs))]
        self._in_ids = [i for a in self.ins for i in a.vec_ids()]
        free = set(E.symbols_of([i for o in self.outs for i in o._d])) - set(self._in_ids)
        if free:
            names = sorted(E.G.args[i][0] for i in free)
            raise Exception(f"Function {name}: free variables {names} are not inputs")

    def name(self): return self._name
    def n_in(self): return len(self.ins)
    def n_out(self): return len(self.outs)
    def size_in(self, i): return self.ins[i].shape
    def size_out(self, i): return self.outs[i].shape

    def __call__(self, *args, **kwargs):
        if kwargs:
            args = [kwargs[n] for n in self.in_names]
        assert len(args) == len(self.ins), f"{self._name}: expected {len(self.ins)} inputs, got {len(args)}"
        args = [getattr(a, "cat", a) for a in args]
        symbolic = any(isinstance(a, SX) and not a.is_constant() for a in args)
        new_ids = []
        for a, ref in zip(args, self.ins):
            a = _sx(a)
            if a.numel() != ref.numel():
                if a.numel() == 1 and ref.numel() > 1:
                    a = SX._from(a._d * ref.numel(), ref.shape)
                else:
                    raise Exception(f"{self._name}: argument size {a.shape} does not match {ref.shape}")
            new_ids.extend(a.vec_ids())
        res = []
        flat = [i for o in self.outs for i in o._d]
        sub = E.substitute(flat, self._in_ids, new_ids)
        k = 0
        for o in self.outs:
            ids = sub[k:k + o.numel()]; k += o.numel()
            r = SX._from(ids, o.shape)
            res.append(r if symbolic else r.to_DM())
        if kwargs:
            return dict(zip(self.out_names, res))
        return res[0] if len(res) == 1 else res
