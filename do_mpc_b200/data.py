"""Result store with the do_mpc.data.Data / MPCData surface (reference do_mpc/data.py:35-457).

Row-per-step numpy arrays for `_x,_u,_z,_tvp,_p,_aux,_time` plus the optimiser fields
`opt_p_num,_opt_x_num,_opt_aux_num,_lam_g_num,success,...`; power-index reads `data['_x','C_a']`;
`MPCData.prediction(('_x','C_a'))`; `save_results/load_results` pickle round trip.
Rows are appended into geometrically grown buffers (the reference re-allocates with np.append on
every call, data.py:216).
"""
from __future__ import annotations

import os
import pickle
from typing import Dict

import numpy as np


class Data:
    dtype = "default"

    def __init__(self, model):
        self.model = model
        self.data_fields: Dict[str, int] = {
            "_time": 1, "_x": model.n_x, "_y": model.n_y, "_u": model.n_u, "_z": model.n_z,
            "_tvp": model.n_tvp, "_p": model.n_p, "_aux": model.n_aux,
        }
        self.meta_data = {}
        self.init_storage()

    def init_storage(self):
        self._buf = {k: np.empty((0, n)) for k, n in self.data_fields.items()}
        self._len = {k: 0 for k in self.data_fields}

    def set_meta(self, **kwargs):
        self.meta_data.update(kwargs)

    def update(self, **kwargs):
        for key, value in kwargs.items():
            v = np.asarray(value.full() if hasattr(value, "full") else value, dtype=float).reshape(1, -1)
            if key not in self.data_fields:
                self.data_fields[key] = v.shape[1]
                self._buf[key] = np.empty((0, v.shape[1])); self._len[key] = 0
            n, buf = self._len[key], self._buf[key]
            if buf.shape[1] != v.shape[1]:
                raise ValueError(f"field {key} expects {buf.shape[1]} values per step, got {v.shape[1]}")
            if n == buf.shape[0]:
                grown = np.empty((max(8, 2 * n), buf.shape[1])); grown[:n] = buf[:n]
                self._buf[key] = buf = grown
            buf[n] = v
            self._len[key] = n + 1

    def _field(self, key):
        return self._buf[key][: self._len[key]]

    def __getattr__(self, name):
        if name.startswith("__") or name in ("_buf", "_len", "data_fields", "model"):
            raise AttributeError(name)
        if name in self.__dict__.get("data_fields", {}):
            return self._field(name)
        raise AttributeError(name)

    def __getitem__(self, ind):
        if not isinstance(ind, tuple):
            ind = (ind,)
        name = ind[0]
        arr = self._field(name)
        if len(ind) == 1:
            return arr
        struct = {"_x": "_x", "_u": "_u", "_z": "_z", "_tvp": "_tvp", "_p": "_p", "_aux": "_aux_expression",
                  "_y": "_y_expression"}.get(name)
        if struct is None:
            raise KeyError(f"field {name} has no named sub-entries")
        idx = getattr(self.model, struct).f_index(tuple(ind[1:]))
        return arr[:, idx]

    def export(self):
        return {k: self._field(k).copy() for k in self.data_fields}


class MPCData(Data):
    dtype = "MPC"

    def prepare(self, mpc):
        self._opt_x = mpc._opt_x
        self._opt_aux = mpc.aux_struct
        self._opt_x_scaling = mpc.opt_x_scaling.full().ravel()
        for k, n in (("opt_p_num", mpc.n_opt_p), ("_opt_x_num", mpc.n_opt_x), ("_opt_aux_num", mpc.n_opt_aux),
                     ("_lam_g_num", mpc.n_opt_lagr), ("success", 1)):
            if k not in self.data_fields:
                self.data_fields[k] = n
                self._buf[k] = np.empty((0, n)); self._len[k] = 0

    def prediction(self, ind, t_ind=-1):
        """Predicted trajectory of a state/input/aux from the stored full solution (reference data.py:246-330).
        Returns array (n_elements, n_horizon(+1), n_scenarios)."""
        assert self.meta_data.get("store_full_solution", False), "store_full_solution must be True"
        assert isinstance(ind, tuple) and ind[0] in ("_x", "_u", "_z", "_aux")
        row = self._field("_opt_aux_num" if ind[0] == "_aux" else "_opt_x_num")[t_ind]
        struct = self._opt_aux if ind[0] == "_aux" else self._opt_x
        if ind[0] == "_x":
            key = ("_x", slice(None), slice(None), -1) + tuple(ind[1:])
        elif ind[0] == "_z":
            key = ("_z", slice(None), slice(None), 0) + tuple(ind[1:])
        else:
            key = (ind[0], slice(None), slice(None)) + tuple(ind[1:])
        nested = struct.layout.resolve(key)
        arr = np.array([[np.atleast_1d(row[np.atleast_1d(i)]) for i in per_k] for per_k in nested])  # (K, S, n)
        return np.transpose(arr, (2, 0, 1))


def save_results(save_list, result_name="results", result_path="./results/", overwrite=False):
    os.makedirs(result_path, exist_ok=True)
    results = {}
    for obj in save_list:
        d = obj.data if hasattr(obj, "data") and not isinstance(obj, Data) else obj
        results[{"MPC": "mpc", "default": type(obj).__name__.lower()}.get(d.dtype, d.dtype)] = \
            {"fields": d.export(), "meta": dict(d.meta_data)}
    name = os.path.join(result_path, result_name + ".pkl")
    i = 0
    while os.path.exists(name) and not overwrite:
        i += 1
        name = os.path.join(result_path, f"{i:03d}_{result_name}.pkl")
    with open(name, "wb") as f:
        pickle.dump(results, f)
    return name


def load_results(file_name):
    with open(file_name, "rb") as f:
        return pickle.load(f)
