"""Approximate MPC on top of the batched solver: sampler -> dataset -> neural-network controller.

Mirrors `do_mpc.approximateMPC` (reference: _ampc.py `FeedforwardNN`/`ApproxMPC`, _ampc_sampler.py `AMPCSampler`,
_trainer.py `Trainer`, _ampcsettings.py).  What changes is where the time goes: the reference's sampler calls
`mpc.make_step` once per sample in a Python loop (_ampc_sampler.py:312-330, 420-470); here every sampling round is ONE
batched GPU solve (`MPC.make_step_batch`, closed loop: one batched simulator step per time step as well).  The dataset
files keep the reference's names and columns (`<data_dir>/<name>/data_<name>_all.pkl`, `data_<name>_opt.pkl`, columns
`x0, u_prev, id, u0, status, t_make_step, t_wall, iter_count`), so they are interchangeable with the reference's trainer.
The network / trainer are plain PyTorch like the reference's (they are not on the solver's hot path).
"""
from __future__ import annotations

import json
import time
from dataclasses import dataclass, field
from pathlib import Path
from typing import Optional

import numpy as np

from .sampling import sampling_plan_box


@dataclass
class ApproximateMPCSettings:           # reference _ampcsettings.py:4-28
    n_hidden_layers: int = 3
    n_neurons: int = 50
    act_fn: str = "tanh"
    output_act_fn: str = "linear"
    device: str = "auto"
    scaling: bool = True
    lbx: Optional[np.ndarray] = None
    ubx: Optional[np.ndarray] = None
    lbu: Optional[np.ndarray] = None
    ubu: Optional[np.ndarray] = None


@dataclass
class SamplerSettings:                  # reference _ampcsettings.py:30-68
    n_samples: Optional[int] = None
    dataset_name: Optional[str] = None
    trajectory_length: int = 1
    closed_loop_flag: bool = False
    data_dir: str = "./sampling"
    overwrite_sampler: bool = True
    seed: Optional[int] = None
    lbx: Optional[np.ndarray] = None
    ubx: Optional[np.ndarray] = None
    lbu: Optional[np.ndarray] = None
    ubu: Optional[np.ndarray] = None


@dataclass
class TrainerSchedulerSettings:         # reference _ampcsettings.py:108-128
    mode: str = "min"
    factor: float = 0.1
    patience: int = 10
    threshold: float = 1e-4
    threshold_mode: str = "rel"
    cooldown: int = 2
    min_lr: float = 1e-7
    eps: float = 1e-8


@dataclass
class TrainerSettings:                  # reference _ampcsettings.py:70-106
    dataset_name: Optional[str] = None
    n_epochs: Optional[int] = None
    data_dir: str = "./sampling"
    results_dir: str = "./training"
    scheduler_flag: bool = False
    val: float = 0.2
    batch_size: int = 1000
    shuffle: bool = True
    learning_rate: float = 1e-3
    save_history: bool = False
    show_fig: bool = False            # accepted for compatibility; plotting is out of scope
    save_fig: bool = False
    print_frequency: int = 10
    seed: int = 0


def _torch():
    import torch
    return torch


def FeedforwardNN(n_in, n_out, n_hidden_layers, n_neurons, act_fn="tanh", output_act_fn="linear"):
    """MLP n_in -> n_hidden_layers x n_neurons -> n_out (reference _ampc.py:32-100)."""
    torch = _torch()
    acts = {"relu": torch.nn.ReLU, "tanh": torch.nn.Tanh, "leaky_relu": torch.nn.LeakyReLU, "sigmoid": torch.nn.Sigmoid}
    if act_fn not in acts or output_act_fn not in (*acts, "linear"):
        raise ValueError(f"unknown activation {act_fn!r} / {output_act_fn!r}")
    layers, n = [], n_in
    for _ in range(n_hidden_layers):
        layers += [torch.nn.Linear(n, n_neurons), acts[act_fn]()]
        n = n_neurons
    layers.append(torch.nn.Linear(n, n_out))
    if output_act_fn != "linear":
        layers.append(acts[output_act_fn]())
    return torch.nn.Sequential(*layers)


def _box(mpc, s):
    """state / input boxes: settings first, else the controller's bounds (reference _ampc.py:160-172)."""
    g = lambda v, d: np.asarray(d.full()).ravel() if v is None else np.asarray(v, dtype=float).ravel()
    return g(s.lbx, mpc._x_lb), g(s.ubx, mpc._x_ub), g(s.lbu, mpc._u_lb), g(s.ubu, mpc._u_ub)


class ApproxMPC:
    """Neural-network controller u0 = net(x0 [, u_prev]) with box scaling and clipping (reference _ampc.py:115-444)."""

    def __init__(self, mpc):
        self.mpc = mpc
        self.settings = ApproximateMPCSettings()
        self.flags = {"setup": False}
        self.with_u_prev = bool(mpc.flags.get("set_rterm", False) or mpc.flags.get("rterm_fun", False))

    def setup(self):
        torch = _torch()
        assert not self.flags["setup"], "Setup can only be once."
        s = self.settings
        if s.device == "auto":
            s.device = "cuda" if torch.cuda.is_available() else "cpu"
        self.torch_device = torch.device(s.device)
        self.torch_data_type = torch.float32
        m = self.mpc.model
        self.n_in = m.n_x + (m.n_u if self.with_u_prev else 0)
        self.net = FeedforwardNN(self.n_in, m.n_u, s.n_hidden_layers, s.n_neurons, s.act_fn, s.output_act_fn).to(self.torch_device)
        lbx, ubx, lbu, ubu = _box(self.mpc, s)
        for name, a in (("lbx", lbx), ("ubx", ubx), ("lbu", lbu), ("ubu", ubu)):
            if not np.all(np.isfinite(a)):
                raise ValueError(f"missing {name}: finite state and input bounds are required for scaling and clipping")
        t = lambda a: torch.tensor(a, dtype=self.torch_data_type, device=self.torch_device).reshape(1, -1)
        lb = np.concatenate([lbx, lbu]) if self.with_u_prev else lbx
        ub = np.concatenate([ubx, ubu]) if self.with_u_prev else ubx
        self.x_shift, self.x_range = t(lb), t(ub - lb)
        self.y_shift, self.y_range = t(lbu), t(ubu - lbu)
        self.lbu_t, self.ubu_t = t(lbu), t(ubu)
        self.u0 = np.asarray(self.mpc._u0.full()).reshape(-1, 1)
        self.flags["setup"] = True

    def scale_inputs(self, x): return ((x - self.x_shift) / self.x_range).to(self.torch_data_type)
    def rescale_outputs(self, y): return y * self.y_range + self.y_shift
    def clip_control_actions(self, y): return _torch().minimum(_torch().maximum(y, self.lbu_t), self.ubu_t)

    def predict(self, x_batch):
        with _torch().no_grad():
            return self.net(x_batch)

    def make_step(self, x0, u_prev=None, clip_to_bounds=True):
        """x0 (n_x, 1) or a batch (B, n_x); returns (n_u, 1) or (B, n_u)."""
        torch = _torch()
        assert self.flags["setup"], "ApproxMPC was not setup yet. Please call ApproxMPC.setup()."
        m = self.mpc.model
        x0 = np.asarray(x0, dtype=float)
        batch = x0.ndim == 2 and x0.shape[1] == m.n_x and x0.shape != (m.n_x, 1)
        X = x0.reshape(-1, m.n_x)
        if self.with_u_prev:
            up = self.u0 if u_prev is None else np.asarray(u_prev, dtype=float)
            X = np.concatenate([X, np.broadcast_to(up.reshape(-1, m.n_u), (X.shape[0], m.n_u))], axis=1)
        with torch.no_grad():
            x = torch.tensor(X, dtype=self.torch_data_type, device=self.torch_device)
            y = self.rescale_outputs(self.net(self.scale_inputs(x))) if self.settings.scaling else self.net(x)
            if clip_to_bounds:
                y = self.clip_control_actions(y)
        y = y.cpu().numpy().astype(float)
        if batch:
            return y
        self.u0 = y.reshape(-1, 1)
        return self.u0

    def save_to_state_dict(self, directory="approx_mpc.pth"): _torch().save(self.net.state_dict(), directory)
    def load_from_state_dict(self, directory="approx_mpc.pth"):
        self.net.load_state_dict(_torch().load(directory, map_location=self.torch_device))


class AMPCSampler:
    """Dataset generation (reference _ampc_sampler.py): open loop = one batched solve of the whole sampling plan;
    closed loop = `trajectory_length` batched controller + plant steps from every sampled initial condition."""

    def __init__(self, mpc):
        self.mpc = mpc
        self.settings = SamplerSettings()
        self.flags = {"setup": False}

    def setup(self):
        s = self.settings
        assert s.n_samples is not None and s.dataset_name is not None, "n_samples and dataset_name must be set"
        if s.closed_loop_flag:
            from .simulator import Simulator
            self.simulator = Simulator(self.mpc.model)
            self.simulator.set_param(t_step=self.mpc.settings.t_step)
            if self.mpc.model.n_p:
                # the plant runs with the controller's nominal parameter scenario (first entry of p_fun, reference _ampc_sampler.py:196-203)
                p_num = self.simulator.get_p_template()
                p0 = self.mpc.p_fun(0.0)["_p"][0]
                p_num.master = np.asarray(p0.full() if hasattr(p0, "full") else p0, dtype=float).reshape(-1, 1)
                self.simulator.set_p_fun(lambda t: p_num)
            if self.mpc.model.n_tvp:
                tvp_num = self.simulator.get_tvp_template()
                self.simulator.set_tvp_fun(lambda t: tvp_num)
            self.simulator.setup()
        self.flags["setup"] = True

    def default_sampling(self):
        import pandas as pd
        assert self.flags["setup"], "call AMPCSampler.setup() first"
        s = self.settings
        lbx, ubx, lbu, ubu = _box(self.mpc, s)
        plan = sampling_plan_box(self.mpc, s.n_samples, seed=s.seed, lbx=lbx, ubx=ubx, lbu=lbu, ubu=ubu)
        X = np.array([p["x0"].ravel() for p in plan]); U = np.array([p["u_prev"].ravel() for p in plan])
        rows = []
        steps = s.trajectory_length if s.closed_loop_flag else 1
        alive = np.ones(len(plan), bool)
        t_start = np.array(self.mpc._t0, dtype=float).copy()
        prev = None
        for k in range(steps):
            t0 = time.perf_counter()
            # closed loop: the controller clock advances (time-varying parameters follow) and every step after the first starts
            # from the previous solution, as mpc.make_step does in the reference's closed-loop sampler (_ampc_sampler.py:420-470)
            self.mpc._t0 = t_start + k * self.mpc.settings.t_step
            out = self.mpc.make_step_batch(X, U, return_full=s.closed_loop_flag, warm_from=prev)
            prev = out if s.closed_loop_flag else None
            dt = (time.perf_counter() - t0) / len(plan)
            ok = np.asarray(out["success"], bool)
            for i in np.where(alive)[0]:
                rows.append({"x0": X[i:i + 1].copy(), "u_prev": U[i:i + 1].copy(), "id": plan[i]["id"] if steps == 1 else f"{plan[i]['id']}_{k}",
                             "u0": out["u0"][i:i + 1].copy(), "status": bool(ok[i]), "t_make_step": dt, "t_wall": float("nan"),
                             "iter_count": int(out["iter_count"][i])})
            if k + 1 < steps:
                # the reference stops a trajectory at the first failed solve (_ampc_sampler.py:440-447)
                alive &= ok
                Xn, _, sim_status = self.simulator.make_step_batch(X, out["u0"])
                alive &= (np.asarray(sim_status) == 0)          # a failed integrator step ends the trajectory as well
                U = np.where(alive[:, None], out["u0"], U)
                X = np.where(alive[:, None], Xn, X)
        self.mpc._t0 = t_start
        df = pd.DataFrame(rows)
        d = Path(s.data_dir) / s.dataset_name
        d.mkdir(parents=True, exist_ok=True)
        df.to_pickle(d / f"data_{s.dataset_name}_all.pkl")
        df[df["status"]].reset_index(drop=True).to_pickle(d / f"data_{s.dataset_name}_opt.pkl")
        self.data = df
        return df


class Trainer:
    """Supervised training of the ApproxMPC network on a sampled dataset (reference _trainer.py:38-506)."""

    def __init__(self, approx_mpc):
        self.approx_mpc = approx_mpc
        self.settings = TrainerSettings()
        self.scheduler_settings = TrainerSchedulerSettings()
        self.flags = {"setup": False}
        self.history = {"epoch": [], "train_loss": [], "val_loss": [], "lr": []}

    def setup(self):
        assert self.approx_mpc.flags["setup"], "ApproxMPC must be set up before the trainer"
        assert self.settings.dataset_name is not None and self.settings.n_epochs is not None
        self.flags["setup"] = True

    def scale_dataset(self, x, u0):
        a = self.approx_mpc
        return (x - a.x_shift) / a.x_range, (u0 - a.y_shift) / a.y_range

    def load_data(self):
        import pandas as pd
        torch = _torch()
        s, a = self.settings, self.approx_mpc
        df = pd.read_pickle(Path(s.data_dir) / s.dataset_name / f"data_{s.dataset_name}_opt.pkl")
        m = a.mpc.model
        st = lambda col, n: torch.tensor(np.stack([np.asarray(v, dtype=float).reshape(n) for v in df[col]]), dtype=a.torch_data_type,
                                         device=a.torch_device)
        x = st("x0", m.n_x)
        if a.with_u_prev:
            x = torch.cat([x, st("u_prev", m.n_u)], dim=1)
        u0 = st("u0", m.n_u)
        if a.settings.scaling:
            x, u0 = self.scale_dataset(x, u0)
        g = torch.Generator(device="cpu").manual_seed(s.seed)
        perm = torch.randperm(x.shape[0], generator=g).to(x.device)
        n_val = int(round(s.val * x.shape[0]))
        return (x[perm[n_val:]], u0[perm[n_val:]]), (x[perm[:n_val]], u0[perm[:n_val]]), g

    def default_training(self):
        torch = _torch()
        assert self.flags["setup"], "call Trainer.setup() first"
        s, a = self.settings, self.approx_mpc
        (xt, yt), (xv, yv), g = self.load_data()
        opt = torch.optim.Adam(a.net.parameters(), lr=s.learning_rate)
        sched = None
        if s.scheduler_flag:
            ss = self.scheduler_settings
            sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode=ss.mode, factor=ss.factor, patience=ss.patience, threshold=ss.threshold,
                                                               threshold_mode=ss.threshold_mode, cooldown=ss.cooldown, min_lr=ss.min_lr * 0.1, eps=ss.eps)
        loss_fn = torch.nn.MSELoss()
        for epoch in range(s.n_epochs):
            a.net.train()
            order = torch.randperm(xt.shape[0], generator=g).to(xt.device) if s.shuffle else torch.arange(xt.shape[0], device=xt.device)
            tot, nb = 0.0, 0
            for i in range(0, xt.shape[0], s.batch_size):
                idx = order[i:i + s.batch_size]
                opt.zero_grad()
                loss = loss_fn(a.net(xt[idx]), yt[idx])
                loss.backward(); opt.step()
                tot += float(loss.detach()); nb += 1
            a.net.eval()
            with torch.no_grad():
                vl = float(loss_fn(a.net(xv), yv)) if xv.shape[0] else float("nan")
            for k, v in (("epoch", epoch), ("train_loss", tot / max(nb, 1)), ("val_loss", vl), ("lr", opt.param_groups[0]["lr"])):
                self.history[k].append(v)
            if (epoch + 1) % s.print_frequency == 0:
                print(f"epoch {epoch + 1}: train_loss {tot / max(nb, 1):.3e} val_loss {vl:.3e}")
            if sched is not None:
                sched.step(vl)
                if opt.param_groups[0]["lr"] < self.scheduler_settings.min_lr:
                    break
        if s.save_history:
            d = Path(s.results_dir) / f"results_{s.dataset_name}"
            d.mkdir(parents=True, exist_ok=True)
            (d / "history.json").write_text(json.dumps(self.history))
        return self.history
