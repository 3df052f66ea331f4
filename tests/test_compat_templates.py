"""The reference's own example templates run unchanged on this package through compat/{casadi,do_mpc}.
Needs the read-only reference tree (build container only) — skipped on the GPU box."""
import importlib.util
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference/examples")
pytestmark = pytest.mark.skipif(not REF.exists(), reason="reference tree not mounted")


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture()
def compat_path():
    sys.path.insert(0, str(ROOT / "compat"))
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "casadi" or k.startswith("casadi.") or k == "do_mpc" or k.startswith("do_mpc.")}
    yield
    sys.path.remove(str(ROOT / "compat"))
    for k in list(sys.modules):
        if k == "casadi" or k.startswith("casadi.") or k == "do_mpc" or k.startswith("do_mpc."):
            del sys.modules[k]
    sys.modules.update(saved)


def _setup_without_gpu(mpc):
    # template_mpc() ends with mpc.setup(); on a CPU-only host stop after the NLP recipe
    mpc.prepare_nlp()


def test_reference_cstr_templates_build_the_same_recipe(compat_path, monkeypatch):
    from do_mpc_b200.controller import MPC
    monkeypatch.setattr(MPC, "setup", _setup_without_gpu)
    tm = _load(REF / "CSTR_approximate_mpc" / "template_model.py", "ref_cstr_model")
    tc = _load(REF / "CSTR_approximate_mpc" / "template_mpc.py", "ref_cstr_mpc")
    model = tm.template_model()
    mpc = tc.template_mpc(model, silence_solver=True)
    r = mpc.recipe
    assert (r.n_w, r.n_g, r.n_p) == (376, 324, 6)
    # same NLP as this repo's own examples/cstr.py: evaluate the generated edge functions numerically
    import cstr
    mine = cstr.cstr_nominal_mpc(setup=False); mine.prepare_nlp()
    from do_mpc_b200.sym import expr as E
    rng = np.random.RandomState(0)
    for a, b in ((r, mine.recipe),):
        q = rng.uniform(0.5, 1.5, size=a.q.numel())
        va = E.evaluate(a.G.ids() + a.lterm.ids(), dict(zip(a.q.ids(), q)))
        vb = E.evaluate(b.G.ids() + b.lterm.ids(), dict(zip(b.q.ids(), q)))
        assert np.allclose(va, vb, rtol=1e-13, atol=1e-13)
    assert np.array_equal(mpc._lb_opt_x.full(), mine._lb_opt_x.full()) and np.array_equal(mpc._ub_opt_x.full(), mine._ub_opt_x.full())


@pytest.mark.parametrize("ex,sizes", [("oscillating_masses_discrete", (39, 32, 5)), ("batch_reactor", (608, 564, 23)),
                                      ("CSTR", (3564, 3064, 24)), ("industrial_poly", (8100, 7210, 31))])
def test_reference_templates_reproduce_golden_problem_sizes(compat_path, monkeypatch, ex, sizes):
    from do_mpc_b200.controller import MPC
    monkeypatch.setattr(MPC, "setup", _setup_without_gpu)
    tm = _load(REF / ex / "template_model.py", f"ref_{ex}_model")
    tc = _load(REF / ex / "template_mpc.py", f"ref_{ex}_mpc")
    mpc = tc.template_mpc(tm.template_model(), silence_solver=True)
    r = mpc.recipe
    assert (r.n_w, r.n_g, r.n_p) == sizes          # SURVEY.md §4 table (n_opt_x / n_g / n_opt_p of the goldens)


def test_reference_dip_and_simulator_templates(compat_path, monkeypatch):
    """double_inverted_pendulum (DAE + obstacle constraints + tvp) and its simulator template."""
    from do_mpc_b200.controller import MPC
    from do_mpc_b200.simulator import Simulator
    monkeypatch.setattr(MPC, "setup", _setup_without_gpu)
    monkeypatch.setattr(Simulator, "setup", lambda self: self.settings.check_for_mandatory_settings())
    ex = "double_inverted_pendulum"
    obstacles = [{"x": 0.0, "y": 0.6, "r": 0.3}]
    tm = _load(REF / ex / "template_model.py", "ref_dip_model")
    tc = _load(REF / ex / "template_mpc.py", "ref_dip_mpc")
    ts = _load(REF / ex / "template_simulator.py", "ref_dip_sim")
    model = tm.template_model(obstacles)
    mpc = tc.template_mpc(model, silence_solver=True)
    r = mpc.recipe
    assert (r.n_w, r.n_g, r.n_p) == (4330, 4506, 126)          # golden sizes of results_dip.pkl
    sim = ts.template_simulator(model)
    assert sim.settings.t_step == mpc.settings.t_step


def test_reference_rotating_masses_templates_with_mhe(compat_path, monkeypatch):
    """model / MPC / simulator / MHE templates of examples/rotating_oscillating_masses_mhe_mpc: golden problem sizes of
    results_rotatingMasses.pkl (mpc 712 variables; estimator 423 variables, 450 constraints, 322 parameters)."""
    from do_mpc_b200.controller import MPC
    from do_mpc_b200.mhe import MHE
    from do_mpc_b200.simulator import Simulator
    monkeypatch.setattr(MPC, "setup", _setup_without_gpu)
    monkeypatch.setattr(MHE, "setup", lambda self: self.prepare_nlp())
    monkeypatch.setattr(Simulator, "setup", lambda self: self.settings.check_for_mandatory_settings())
    ex = "rotating_oscillating_masses_mhe_mpc"
    model = _load(REF / ex / "template_model.py", "ref_rm_model").template_model()
    mpc = _load(REF / ex / "template_mpc.py", "ref_rm_mpc").template_mpc(model, silence_solver=True)
    _load(REF / ex / "template_simulator.py", "ref_rm_sim").template_simulator(model)
    mhe = _load(REF / ex / "template_mhe.py", "ref_rm_mhe").template_mhe(model, silence_solver=True)
    assert mpc.recipe.n_w == 712
    assert (mhe.recipe.n_w, mhe.recipe.n_g, mhe.n_opt_p) == (423, 450, 322)
    # same estimator NLP as this repo's own examples/rotating_masses.py
    import rotating_masses
    mine = rotating_masses.masses_mhe(setup=False); mine.prepare_nlp()
    from do_mpc_b200.sym import expr as E
    rng = np.random.RandomState(0)
    a, b = mhe.recipe, mine.recipe
    q = rng.uniform(0.5, 1.5, size=a.q.numel()); tv = rng.uniform(0.5, 1.5, size=a.tvp.numel()); pp = rng.uniform(1e-4, 3e-4, size=a.pp.numel())
    va = E.evaluate(a.G.ids() + a.lterm.ids() + a.nl.ids(), dict(zip(a.q.ids() + a.tvp.ids() + a.pp.ids(), np.concatenate([q, tv, pp]))))
    vb = E.evaluate(b.G.ids() + b.lterm.ids() + b.nl.ids(), dict(zip(b.q.ids() + b.tvp.ids() + b.pp.ids(), np.concatenate([q, tv, pp]))))
    assert np.allclose(va, vb, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("ex,args,mine", [("kite", (10.0, 6.0, 100.0), "kite.kite_mpc"), ("kinematic_bicycle_model", (), "bicycle.kinematic_mpc"),
                                          ("dynamic_bicycle_model", (), "bicycle.dynamic_mpc")])
def test_reference_templates_without_golden_equal_this_repos_examples(compat_path, monkeypatch, ex, args, mine):
    """examples/kite, kinematic_bicycle_model, dynamic_bicycle_model (no golden in testing/results): the reference's templates
    build, through compat/, the same stage recipe as examples/kite.py / examples/bicycle.py — the controllers that the oracle
    fixtures tests/golden/cfg{kite,kinbike,dynbike}_oracle.npz pin on the CPU emulation and on the GPU."""
    import importlib
    from do_mpc_b200.controller import MPC
    from do_mpc_b200.simulator import Simulator
    monkeypatch.setattr(MPC, "setup", _setup_without_gpu)
    monkeypatch.setattr(Simulator, "setup", lambda self: self.settings.check_for_mandatory_settings())
    sys.path.insert(0, str(REF / ex))                       # kite/template_simulator.py imports its wind_model module
    try:
        model = _load(REF / ex / "template_model.py", f"ref_{ex}_model").template_model()
        mpc = _load(REF / ex / "template_mpc.py", f"ref_{ex}_mpc").template_mpc(model, *args, silence_solver=True)
        ts = _load(REF / ex / "template_simulator.py", f"ref_{ex}_sim")
        sim = ts.template_simulator(model, *args[:2])
    finally:
        sys.path.remove(str(REF / ex))
    assert sim.settings.t_step == mpc.settings.t_step
    mod, fn = mine.split(".")
    own = getattr(importlib.import_module(mod), fn)(setup=False); own.prepare_nlp()
    a, b = mpc.recipe, own.recipe
    assert (a.n_w, a.n_g, a.n_p) == (b.n_w, b.n_g, b.n_p)
    from do_mpc_b200.sym import expr as E
    rng = np.random.RandomState(0)
    syms = lambda r: sum(([] if v is None else v.ids() for v in (r.q, r.tvp, r.pp, r.uprev, r.xN)), [])
    vals = np.concatenate([rng.uniform(0.3, 1.2, size=a.q.numel()), rng.uniform(0.5, 1.5, size=a.tvp.numel()),
                           rng.uniform(5.0, 7.0, size=a.pp.numel()), rng.uniform(0.3, 1.2, size=a.uprev.numel() + a.xN.numel())])
    ids = lambda r: sum((v.ids() for v in (r.G, r.lterm, r.mterm, r.rterm, r.nl) if v is not None and v.numel()), [])
    assert len(syms(a)) == len(syms(b)) == len(vals)
    va = E.evaluate(ids(a), dict(zip(syms(a), vals)))
    vb = E.evaluate(ids(b), dict(zip(syms(b), vals)))
    assert len(va) == len(vb) > 0
    assert np.allclose(va, vb, rtol=1e-13, atol=1e-13)
    assert np.array_equal(mpc._lb_opt_x.full(), own._lb_opt_x.full()) and np.array_equal(mpc._ub_opt_x.full(), own._ub_opt_x.full())
    assert np.array_equal(np.asarray(mpc.nlp_cons_ub), np.asarray(own.nlp_cons_ub))
