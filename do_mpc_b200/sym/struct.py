"""Structured (named, repeated) vectors with CasADi-tools style power indexing.

Provides what do-mpc users see of `casadi.tools.struct_symSX / struct_SX / DMStruct`
(reference do_mpc/tools/_casstructure.py:15-27, do_mpc/model/_model.py:205-330 and the power indices
used in do_mpc/controller/_mpc.py:1016-1026): `s['name']`, `s['name', k, 'sub']`, slices, lists, -1,
`.cat`, `.master`, `.keys()`, `.labels()`, `.shape`, call `s(0)` -> numeric copy.
Flattening is declaration order, repeats row-major, matrices column-major (CasADi convention).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Union

import numpy as np

from .sx import DM, SX


class entry:
    """casadi.tools.entry(name, shape=|struct=|repeat=|expr=|sym=)"""

    def __init__(self, name, shape=None, struct=None, repeat=None, expr=None, sym=None):
        self.name = name
        self.struct = struct.layout if isinstance(struct, _StructBase) else struct
        self.expr = expr
        self.sym = sym
        if repeat is None:
            repeat = []
        elif isinstance(repeat, (int, np.integer)):
            repeat = [int(repeat)]
        self.repeat = [int(r) for r in repeat]
        if self.struct is not None:
            self.shape = (self.struct.size, 1)
        else:
            src = expr if expr is not None else sym
            if src is not None and shape is None:
                if isinstance(src, (list, tuple)):
                    raise NotImplementedError("list-valued entry expressions")
                src = src if isinstance(src, SX) else SX(src)
                shape = src.shape
            if shape is None:
                shape = (1, 1)
            if isinstance(shape, (int, np.integer)):
                shape = (int(shape), 1)
            if len(shape) == 1:
                shape = (shape[0], 1)
            self.shape = (int(shape[0]), int(shape[1]))
        self.item_size = self.shape[0] * self.shape[1]
        self.n_rep = int(np.prod(self.repeat)) if self.repeat else 1
        self.size = self.item_size * self.n_rep


class Layout:
    def __init__(self, entries: Sequence[entry]):
        self.entries = list(entries)
        self.by_name = {}
        off = 0
        for e in self.entries:
            e.offset = off
            self.by_name[e.name] = e
            off += e.size
        self.size = off

    def keys(self):
        return [e.name for e in self.entries]

    def labels(self):
        out = []
        for e in self.entries:
            for r in range(e.n_rep):
                ridx = list(np.unravel_index(r, e.repeat)) if e.repeat else []
                if e.struct is not None:
                    for l in e.struct.labels():
                        out.append("[" + ",".join([e.name] + [str(i) for i in ridx] + [l[1:-1]]) + "]")
                else:
                    for i in range(e.item_size):
                        out.append("[" + ",".join([e.name] + [str(j) for j in ridx] + [str(i)]) + "]")
        return out

    # -- power index -> nested lists of index arrays ------------------------------------------------
    def resolve(self, key):
        if not isinstance(key, tuple):
            key = (key,)
        return self._resolve(list(key), 0)

    def _resolve(self, key: list, base: int):
        if not key:
            return base + np.arange(self.size)
        name = key[0]
        if isinstance(name, (list, tuple)):
            return [self._resolve([n] + key[1:], base) for n in name]
        if isinstance(name, slice):
            return [self._resolve([n] + key[1:], base) for n in self.keys()[name]]
        if name not in self.by_name:
            raise KeyError(f"'{name}' is not in structure with keys {self.keys()}")
        e = self.by_name[name]
        return self._resolve_rep(e, key[1:], 0, base + e.offset, e.n_rep)

    def _resolve_rep(self, e: entry, key: list, dim: int, base: int, stride_total: int):
        if dim < len(e.repeat):
            n = e.repeat[dim]
            stride = e.item_size * int(np.prod(e.repeat[dim + 1:])) if dim + 1 < len(e.repeat) else e.item_size
            if not key:
                sel, scalar = list(range(n)), False
            else:
                k = key[0]
                key = key[1:]
                if isinstance(k, (int, np.integer)):
                    kk = int(k)
                    if kk < 0: kk += n
                    if not 0 <= kk < n:
                        raise IndexError(f"index {k} out of range for '{e.name}' (size {n})")
                    sel, scalar = [kk], True
                elif isinstance(k, slice):
                    sel, scalar = list(range(n))[k], False
                else:
                    sel, scalar = [int(i) % n if int(i) < 0 else int(i) for i in k], False
            res = [self._resolve_rep(e, key, dim + 1, base + i * stride, 0) for i in sel]
            return res[0] if scalar else res
        # past the repeats
        if e.struct is not None:
            return e.struct._resolve(key, base) if key else base + np.arange(e.item_size)
        if not key:
            return base + np.arange(e.item_size)
        # matrix index (column-major)
        n, m = e.shape
        lin = np.arange(n * m).reshape(m, n).T
        if len(key) == 1:
            k = key[0]
            sel = lin.T.ravel()[k] if m > 1 else lin[:, 0][k]
        else:
            sel = lin[key[0], key[1]]
        return base + np.atleast_1d(sel).ravel()


def _map_nested(x, f):
    if isinstance(x, list):
        return [_map_nested(i, f) for i in x]
    return f(x)


def _leaves(x, out):
    if isinstance(x, list):
        for i in x:
            _leaves(i, out)
    else:
        out.append(x)
    return out


class _StructBase:
    layout: Layout

    @property
    def size(self): return self.layout.size
    @property
    def shape(self): return (self.layout.size, 1)
    def keys(self): return self.layout.keys()
    def labels(self): return self.layout.labels()
    def __len__(self): return self.layout.size

    def f_index(self, key):
        """flat indices selected by a power index (casadi `.f[...]`)"""
        return np.concatenate([np.atleast_1d(a) for a in _leaves(self.layout.resolve(key), [])] or [np.zeros(0, int)])

    class _F:
        def __init__(self, s): self.s = s
        def __getitem__(self, key): return list(self.s.f_index(key))

    @property
    def f(self): return _StructBase._F(self)


class SymStruct(_StructBase):
    """struct_symSX (fresh symbols) or struct_SX (given expressions)."""

    def __init__(self, entries: Sequence[entry], symbolic: bool = True):
        self.layout = Layout(entries)
        ids: List[int] = []
        for e in self.layout.entries:
            if e.expr is not None:
                ex = e.expr if isinstance(e.expr, SX) else SX(e.expr)
                assert not e.repeat, "repeated expr entries are not supported"
                ids.extend(ex.vec_ids())
            elif e.sym is not None:
                ids.extend(e.sym.vec_ids())
            else:
                for r in range(e.n_rep):
                    suffix = "" if not e.repeat else "_" + "_".join(str(i) for i in np.unravel_index(r, e.repeat))
                    if e.struct is not None:
                        for l in e.struct.labels():
                            ids.extend(SX.sym(f"{e.name}{suffix}_{l[1:-1].replace(',', '_')}").vec_ids())
                    else:
                        ids.extend(SX.sym(f"{e.name}{suffix}", e.shape[0], e.shape[1]).vec_ids())
        self._ids = ids
        self.master = self.cat

    @property
    def cat(self) -> SX:
        return SX._from(self._ids, (len(self._ids), 1))

    def _item(self, idx, e_shape=None):
        idx = np.atleast_1d(idx)
        return SX._from([self._ids[i] for i in idx], (len(idx), 1))

    def __getitem__(self, key):
        res = self.layout.resolve(key)
        shp = self._leaf_shape(key)
        def mk(idx):
            v = self._item(idx)
            if shp is not None and shp[1] > 1 and shp[0] * shp[1] == v.numel():
                from .sx import reshape
                return reshape(v, shp)
            return v
        return _map_nested(res, mk)

    def _leaf_shape(self, key):
        if not isinstance(key, tuple): key = (key,)
        if isinstance(key[0], str) and key[0] in self.layout.by_name:
            e = self.layout.by_name[key[0]]
            if e.struct is None and len(key) == 1 + len(e.repeat):
                return e.shape
        return None

    def __setitem__(self, key, value):
        value = value if isinstance(value, SX) else SX(value)
        for idx in _leaves(self.layout.resolve(key), []):
            idx = np.atleast_1d(idx)
            v = value.vec_ids()
            if len(v) == 1: v = v * len(idx)
            assert len(v) == len(idx), f"cannot assign {value.shape} to selection of size {len(idx)}"
            for i, x in zip(idx, v):
                self._ids[i] = x

    def __call__(self, arg=0):
        """Numeric structure with the same layout, filled from a scalar / vector."""
        return NumStruct(self.layout, arg)

    def __iter__(self):
        raise TypeError("structures are not iterable")


class NumStruct(_StructBase):
    """DMStruct: numeric values in a layout."""

    def __init__(self, layout: Union[Layout, _StructBase], arg=0):
        self.layout = layout.layout if isinstance(layout, _StructBase) else layout
        if isinstance(arg, NumStruct):
            arg = arg._v
        if isinstance(arg, (DM, SX)):
            arg = np.asarray(arg.full() if isinstance(arg, DM) else arg.to_DM().full()).ravel()
        a = np.asarray(arg, dtype=float).ravel()
        self._v = np.full(self.layout.size, a[0]) if a.size == 1 and self.layout.size != 1 else a.astype(float).copy()
        if self._v.size != self.layout.size:
            raise ValueError(f"expected {self.layout.size} values, got {self._v.size}")

    @property
    def cat(self) -> DM: return DM(self._v.reshape(-1, 1))
    @property
    def master(self) -> DM: return self.cat
    @master.setter
    def master(self, v):
        a = np.asarray(v.full() if isinstance(v, DM) else v, dtype=float).ravel()
        if a.size != self._v.size:
            raise ValueError(f"master expects {self._v.size} values, got {a.size}")
        self._v[:] = a

    def full(self): return self._v.reshape(-1, 1).copy()
    def __array__(self, dtype=None, copy=None): return self._v.reshape(-1, 1)

    def __getitem__(self, key):
        return _map_nested(self.layout.resolve(key), lambda idx: DM(self._v[np.atleast_1d(idx)].reshape(-1, 1)))

    def __setitem__(self, key, value):
        if isinstance(value, (list, tuple)) and len(value) > 1:
            # CasADi structure semantics: a list assigned to a slice over repeated entries goes item by item
            # (`tvp_template['_tvp', :-1] = vertsplit(traj)`), each item broadcast over / matching its entry
            leaves = _leaves(self.layout.resolve(key), [])
            if len(leaves) == len(value):
                for idx, item in zip(leaves, value):
                    idx = np.atleast_1d(idx)
                    a = np.asarray(item.full() if hasattr(item, "full") else item, dtype=float).ravel()
                    if a.size not in (1, idx.size):
                        raise ValueError(f"cannot assign {a.size} values to an entry of size {idx.size}")
                    self._v[idx] = a if a.size > 1 else float(a[0])
                return
        if isinstance(value, (DM, NumStruct)):
            value = value.full()
        elif isinstance(value, SX):
            value = value.to_DM().full()
        val = np.asarray(value, dtype=float)
        leaves = _leaves(self.layout.resolve(key), [])
        if val.size > 1 and len(leaves) > 1 and val.size == sum(np.size(l) for l in leaves) and val.size != np.size(leaves[0]):
            k = 0                               # one value per selected element
            for idx in leaves:
                n = np.size(idx)
                self._v[np.atleast_1d(idx)] = val.ravel()[k:k + n]; k += n
            return
        for idx in leaves:
            idx = np.atleast_1d(idx)
            if val.size == 1:
                self._v[idx] = float(val.ravel()[0])
            else:
                v = val.T.ravel() if val.ndim == 2 and val.shape[1] > 1 else val.ravel()
                if v.size != idx.size:
                    raise ValueError(f"cannot assign {val.shape} to selection of size {idx.size}")
                self._v[idx] = v

    def __call__(self, arg=0):
        return NumStruct(self.layout, arg)

    # arithmetic on the flat vector (x0.cat / scaling style expressions go through .cat)
    def __truediv__(self, o): return DM(self._v.reshape(-1, 1) / np.asarray(o.full() if hasattr(o, "full") else o))
    def __mul__(self, o): return DM(self._v.reshape(-1, 1) * np.asarray(o.full() if hasattr(o, "full") else o))


def struct_symSX(entries): return SymStruct(entries, True)
def struct_SX(entries): return SymStruct(entries, False)
struct_symMX = struct_symSX
struct_MX = struct_SX
