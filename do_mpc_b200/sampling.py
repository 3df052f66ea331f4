"""Batched data generation for approximate MPC: what the reference's AMPCSampler does with a serial Python loop
(do_mpc/approximateMPC/_ampc_sampler.py:216-330, 387-526: random (x0, u_prev) in the bound box -> reset_history,
set_initial_guess, make_step -> one pickle per sample -> pandas table) as ONE batched solve.

Output table has the reference dataset's columns (`x0, u_prev, id, u0, status, t_make_step, t_wall, iter_count`,
see documentation/.../sampling_notebook/my_dataset/data_my_dataset_all.pkl)."""
from __future__ import annotations

import time

import numpy as np


def sampling_plan_box(mpc, n_samples: int, seed=None, lbx=None, ubx=None, lbu=None, ubu=None):
    """x0 ~ U[lbx, ubx], u_prev ~ U[lbu, ubu] (reference approx_mpc_sampling_plan_box, _ampc_sampler.py:236-283)."""
    rng = np.random.RandomState(seed) if seed is not None else np.random
    lbx = mpc._x_lb.full().ravel() if lbx is None else np.asarray(lbx, dtype=float)
    ubx = mpc._x_ub.full().ravel() if ubx is None else np.asarray(ubx, dtype=float)
    lbu = mpc._u_lb.full().ravel() if lbu is None else np.asarray(lbu, dtype=float)
    ubu = mpc._u_ub.full().ravel() if ubu is None else np.asarray(ubu, dtype=float)
    for name, a in (("state", lbx), ("state", ubx), ("input", lbu), ("input", ubu)):
        if not np.all(np.isfinite(a)):
            raise ValueError(f"box sampling needs finite {name} bounds (pass lbx/ubx/lbu/ubu explicitly)")
    x0 = rng.uniform(lbx, ubx, size=(n_samples, lbx.size))
    up = rng.uniform(lbu, ubu, size=(n_samples, lbu.size))
    return [{"x0": x0[i].reshape(-1, 1), "u_prev": up[i].reshape(-1, 1), "id": f"{i:0{len(str(max(n_samples - 1, 1)))}d}"}
            for i in range(n_samples)]


def sample_open_loop(mpc, plan, as_dataframe: bool = True):
    """Solve every entry of a sampling plan in one batch (cold start per sample, like the reference's sampler)."""
    X0 = np.array([np.asarray(p["x0"], dtype=float).ravel() for p in plan])
    UP = np.array([np.asarray(p["u_prev"], dtype=float).ravel() for p in plan])
    t0 = time.perf_counter()
    out = mpc.make_step_batch(X0, UP)
    dt = time.perf_counter() - t0
    rows = []
    for i, p in enumerate(plan):
        ok = bool(out["success"][i])
        rows.append({"x0": X0[i:i + 1], "u_prev": UP[i:i + 1], "id": p.get("id", str(i)),
                     "u0": out["u0"][i:i + 1] if ok else np.zeros((0,)), "status": ok,
                     "t_make_step": dt / len(plan), "t_wall": float("nan"), "iter_count": int(out["iter_count"][i])})
    if not as_dataframe:
        return rows
    import pandas as pd
    return pd.DataFrame(rows)


# ----------------------------------------------------------------------------------------------------------------------
# Generic sampling tools of the reference (do_mpc/sampling: SamplingPlanner, Sampler, DataHandler) — same files on disk
# (`<data_dir>/<sampling_plan_name>.pkl`, `<data_dir>/<sample_name>_<id>.pkl`), plus a *batch* sample function: the whole
# plan is handed to one call (e.g. MPC.make_step_batch) instead of one Python call per sample (_sampler.py:198-241).
# ----------------------------------------------------------------------------------------------------------------------
import inspect
import itertools
import pickle
from pathlib import Path


class SamplingPlanner:
    def __init__(self, **kwargs):
        self.sampling_vars, self.sampling_var_names, self.sampling_plan = [], [], []
        self.flags = {"gen_sampling_plan": False}
        self.data_dir, self.overwrite, self.id_precision = "./", False, 3
        self.set_param(**kwargs)

    def set_param(self, **kwargs):
        for k, v in kwargs.items():
            if k not in ("overwrite", "id_precision", "data_dir"):
                raise KeyError(f"Cannot set param {k}: unknown parameter")
            setattr(self, k, v)
        Path(self.data_dir).mkdir(parents=True, exist_ok=True)

    def set_sampling_var(self, name, fun_var_pdf=None):
        assert isinstance(name, str), "name must be str"
        assert fun_var_pdf is None or callable(fun_var_pdf), "fun_var_pdf must be None or callable without arguments"
        self.sampling_vars.append({"name": name, "fun_var_pdf": fun_var_pdf})
        self.sampling_var_names.append(name)

    def add_sampling_case(self, **kwargs):
        unknown = set(kwargs) - set(self.sampling_var_names)
        assert not unknown, f"unknown sampling variables {unknown}"
        case = {}
        for v in self.sampling_vars:
            if v["name"] in kwargs:
                case[v["name"]] = kwargs[v["name"]]
            else:
                assert v["fun_var_pdf"] is not None, f"no value and no generating function for {v['name']}"
                case[v["name"]] = v["fun_var_pdf"]()
        case["id"] = str(len(self.sampling_plan)).zfill(self.id_precision)
        self.sampling_plan.append(case)
        return self.sampling_plan

    def gen_sampling_plan(self, n_samples):
        for _ in range(n_samples):
            self.add_sampling_case()
        self.flags["gen_sampling_plan"] = True
        return self.sampling_plan

    def product(self, **kwargs):
        names = list(kwargs)
        for combo in itertools.product(*[kwargs[n] for n in names]):
            self.add_sampling_case(**dict(zip(names, combo)))
        return self.sampling_plan

    def export(self, sampling_plan_name):
        f = Path(self.data_dir) / f"{sampling_plan_name}.pkl"
        if f.exists() and not self.overwrite:
            raise FileExistsError(f"{f} exists; set_param(overwrite=True) to replace it")
        with open(f, "wb") as fh:
            pickle.dump(self.sampling_plan, fh)


class Sampler:
    def __init__(self, sampling_plan, **kwargs):
        assert isinstance(sampling_plan, list) and sampling_plan and all(isinstance(s, dict) for s in sampling_plan), \
            "sampling_plan must be a non-empty list of dicts"
        assert all(list(s.keys()) == list(sampling_plan[0].keys()) for s in sampling_plan), "all samples need the same variables"
        self.sampling_plan = sampling_plan
        self.sampling_vars = [k for k in sampling_plan[0] if k != "id"]
        self.n_samples = len(sampling_plan)
        self.completion_list = []
        self.flags = {"set_sample_function": False}
        self.data_dir, self.sample_name, self.save_format, self.overwrite, self.print_progress = "./", "sample", "pickle", False, False
        self._batch = False
        self.set_param(**kwargs)

    def set_param(self, **kwargs):
        for k, v in kwargs.items():
            if k not in ("overwrite", "sample_name", "save_format", "print_progress", "data_dir"):
                raise KeyError(f"Cannot set param {k}: unknown parameter")
            setattr(self, k, v)
        assert self.save_format == "pickle", "only save_format='pickle' is supported"
        Path(self.data_dir).mkdir(parents=True, exist_ok=True)

    def _check(self, f):
        assert callable(f), "sample_function must be callable"
        unknown = set(inspect.getfullargspec(f).args) - set(self.sampling_vars)
        assert not unknown, f"sample_function has arguments that are no sampling variables: {unknown}"

    def set_sample_function(self, sample_function):
        self._check(sample_function)
        self.sample_function, self._batch = sample_function, False
        self.flags["set_sample_function"] = True

    def set_batch_sample_function(self, batch_sample_function):
        """f(**{var: [values of all pending samples]}) -> list of results, one per sample (one call for the whole plan)."""
        self._check(batch_sample_function)
        self.sample_function, self._batch = batch_sample_function, True
        self.flags["set_sample_function"] = True

    def _file(self, sample_id):
        return Path(self.data_dir) / f"{self.sample_name}_{sample_id}.pkl"

    def _save(self, sample_id, result):
        with open(self._file(sample_id), "wb") as fh:
            pickle.dump(result, fh)
        self.completion_list.append(sample_id)

    def sample_idx(self, idx):
        assert self.flags["set_sample_function"] and not self._batch, "set a (per-sample) sample function first"
        assert 0 <= idx < self.n_samples, f"idx must be in [0, {self.n_samples})"
        s = dict(self.sampling_plan[idx]); sid = s.pop("id")
        if not self._file(sid).exists() or self.overwrite:
            self._save(sid, self.sample_function(**s))

    def sample_data(self):
        assert self.flags["set_sample_function"], "Cannot sample before setting the sample function"
        if not self._batch:
            for i in range(self.n_samples):
                self.sample_idx(i)
            return
        todo = [s for s in self.sampling_plan if self.overwrite or not self._file(s["id"]).exists()]
        if not todo:
            return
        args = inspect.getfullargspec(self.sample_function).args
        res = self.sample_function(**{k: [s[k] for s in todo] for k in args})
        assert len(res) == len(todo), "batch sample function must return one result per sample"
        for s, r_ in zip(todo, res):
            self._save(s["id"], r_)


class DataHandler:
    def __init__(self, sampling_plan, **kwargs):
        self.sampling_plan = sampling_plan
        self.sampling_vars = [k for k in sampling_plan[0] if k != "id"]
        self.data_dir, self.sample_name, self.save_format = "./", "sample", "pickle"
        self.post_processing = {}
        self.flags = {"set_post_processing": False}
        self.set_param(**kwargs)

    def set_param(self, **kwargs):
        for k, v in kwargs.items():
            if k not in ("data_dir", "sample_name", "save_format"):
                raise KeyError(f"Cannot set param {k}: unknown parameter")
            setattr(self, k, v)

    def set_post_processing(self, name, post_processing_function):
        assert isinstance(name, str) and callable(post_processing_function)
        self.post_processing[name] = post_processing_function
        self.flags["set_post_processing"] = True

    def _load(self, sample_id):
        f = Path(self.data_dir) / f"{self.sample_name}_{sample_id}.pkl"
        if not f.exists():
            return None
        with open(f, "rb") as fh:
            return pickle.load(fh)

    def _process(self, sample):
        res = self._load(sample["id"])
        if res is None:
            return None
        out = dict(sample)
        if self.post_processing:
            for name, f in self.post_processing.items():
                n_args = len(inspect.getfullargspec(f).args)
                out[name] = f(sample, res) if n_args == 2 else f(res)
        else:
            out["res"] = res
        return out

    def __getitem__(self, ind):
        plan = self.sampling_plan[ind] if isinstance(ind, slice) else [self.sampling_plan[i] for i in np.atleast_1d(ind)]
        return [r for r in (self._process(s) for s in plan) if r is not None]

    def filter(self, input_filter=None, output_filter=None):
        """Samples whose sampling variables satisfy input_filter(**vars) and whose post-processed outputs satisfy output_filter(**outs)."""
        sel = []
        for s in self.sampling_plan:
            if input_filter is not None and not input_filter(**{k: s[k] for k in inspect.getfullargspec(input_filter).args}):
                continue
            r_ = self._process(s)
            if r_ is None:
                continue
            if output_filter is not None and not output_filter(**{k: r_[k] for k in inspect.getfullargspec(output_filter).args}):
                continue
            sel.append(r_)
        return sel
