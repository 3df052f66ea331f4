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
