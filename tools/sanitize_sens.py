import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/examples'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
import cstr
d = np.load('/root/repo/tests/golden/cstr_dataset.npz')
mpc = cstr.cstr_robust_mpc() if len(sys.argv) > 1 and sys.argv[1] == "robust" else cstr.cstr_nominal_mpc()
idx = np.where(d["success"])[0][:6]
if mpc.recipe.n_pm:
    x0 = np.tile([0.8, 0.5, 134.14, 130.0], (4, 1)) * np.linspace(0.99, 1.01, 4)[:, None]
    out = mpc.make_step_batch(x0, np.zeros((4, 2)), return_full=True)
else:
    out = mpc.make_step_batch(d["x0"][idx], d["u_prev"][idx], return_full=True)
print("status", out["status"])
r = mpc.S.sensitivities(out["x"], out["lam_g"], out["p"], status=out["status"], want_dlam=True)
print("sens", r["sens_status"], np.abs(r["dxdp"]).max())
