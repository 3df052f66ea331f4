"""Approximate-MPC network + trainer (reference do_mpc/approximateMPC/_ampc.py, _trainer.py) on the reference's own sampled
dataset (tests/golden/cstr_dataset.npz = documentation/.../my_dataset), written in the reference's pickle layout."""
import numpy as np
import pandas as pd
import pytest

from conftest import GOLDEN


@pytest.fixture()
def dataset_dir(tmp_path):
    d = np.load(GOLDEN / "cstr_dataset.npz")
    rows = [{"x0": d["x0"][i:i + 1], "u_prev": d["u_prev"][i:i + 1], "id": f"{i:03d}", "u0": d["u0"][i:i + 1],
             "status": bool(d["success"][i]), "t_make_step": 0.0, "t_wall": 0.0, "iter_count": int(d["iter_count"][i])} for i in range(1000)]
    df = pd.DataFrame(rows)
    (tmp_path / "my_dataset").mkdir()
    df.to_pickle(tmp_path / "my_dataset" / "data_my_dataset_all.pkl")
    df[df["status"]].reset_index(drop=True).to_pickle(tmp_path / "my_dataset" / "data_my_dataset_opt.pkl")
    return tmp_path


def test_network_trains_on_the_reference_dataset(dataset_dir):
    import cstr
    from do_mpc_b200.approximateMPC import ApproxMPC, Trainer
    mpc = cstr.cstr_nominal_mpc(setup=False); mpc.prepare_nlp()
    a = ApproxMPC(mpc)
    a.settings.n_hidden_layers, a.settings.n_neurons, a.settings.device = 2, 32, "cpu"
    a.settings.lbx, a.settings.ubx = [0.1, 0.1, 50.0, 50.0], [2.0, 2.0, 140.0, 140.0]
    a.setup()
    assert a.with_u_prev and a.n_in == 6
    t = Trainer(a)
    t.settings.dataset_name, t.settings.n_epochs, t.settings.data_dir = "my_dataset", 60, str(dataset_dir)
    t.settings.batch_size, t.settings.print_frequency, t.settings.learning_rate = 128, 1000, 3e-3
    t.settings.scheduler_flag = True
    t.setup()
    h = t.default_training()
    assert len(h["train_loss"]) == 60 and h["train_loss"][-1] < 0.3 * h["train_loss"][0] and h["val_loss"][-1] < 0.5 * h["val_loss"][0]
    u = a.make_step(np.array([[0.8], [0.4], [134.0], [130.0]]), u_prev=np.array([[5.0], [-8500.0]]))
    assert u.shape == (2, 1) and 5.0 <= u[0, 0] <= 100.0 and -8500.0 <= u[1, 0] <= 0.0
    ub = a.make_step(np.tile([0.8, 0.4, 134.0, 130.0], (7, 1)), u_prev=np.tile([5.0, -8500.0], (7, 1)))
    assert ub.shape == (7, 2) and np.allclose(ub[0], u.ravel(), atol=1e-4)
    a.save_to_state_dict(str(dataset_dir / "net.pth"))
    b = ApproxMPC(mpc); b.settings = a.settings; b.flags = {"setup": False}; b.setup(); b.load_from_state_dict(str(dataset_dir / "net.pth"))
    assert np.allclose(b.make_step(np.array([[0.8], [0.4], [134.0], [130.0]]), u_prev=np.array([[5.0], [-8500.0]])), u, atol=1e-5)


def test_missing_bounds_are_reported():
    import cstr
    from do_mpc_b200.approximateMPC import ApproxMPC
    mpc = cstr.cstr_nominal_mpc(setup=False); mpc.prepare_nlp()
    a = ApproxMPC(mpc); a.settings.device = "cpu"
    a.settings.lbx = [-np.inf, 0.1, 50.0, 50.0]
    with pytest.raises(ValueError, match="lbx"):
        a.setup()


def test_closed_loop_sampler_configures_the_plant_with_the_nominal_scenario(monkeypatch):
    """AMPCSampler.setup() for a model with uncertain parameters (no GPU needed up to the simulator handle)."""
    import sys
    import cstr
    import do_mpc_b200  # noqa: F401
    from do_mpc_b200.approximateMPC import AMPCSampler
    Sim = sys.modules["do_mpc_b200.simulator"].Simulator
    monkeypatch.setattr(Sim, "setup", lambda self: self.flags.__setitem__("setup", True))
    mpc = cstr.cstr_robust_mpc(setup=False); mpc.prepare_nlp()
    s = AMPCSampler(mpc)
    s.settings.n_samples, s.settings.dataset_name, s.settings.closed_loop_flag, s.settings.trajectory_length = 4, "x", True, 2
    s.setup()
    assert np.allclose(s.simulator.p_fun(0.0).full().ravel(), [1.0, 1.0])          # alpha, beta of the nominal scenario
