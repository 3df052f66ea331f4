# quick GPU correctness + speed check: dataset parity (cfg2) and config fixtures
import sys, json
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo/examples')
import numpy as np
import cstr
from config_cases import MAX_ITER, compare_with_oracle, config_mpc, fixture
d=np.load('/root/repo/tests/golden/cstr_dataset.npz')
mpc=cstr.cstr_nominal_mpc()
out=mpc.make_step_batch(d["x0"], d["u_prev"])
ok=d["success"].astype(bool)
print("cfg2 dataset: split equal", np.array_equal(out["success"], ok), "u0 dev", (np.abs(out["u0"][ok]-d["u0"][ok])/mpc.recipe.u_scaling).max(),
      "iter maxdiff", np.abs(out["iter_count"][ok]-d["iter_count"][ok]).max(), "iter equal frac", (out["iter_count"][ok]==d["iter_count"][ok]).mean())
for cfg in [int(a) for a in sys.argv[1:]]:
    m=config_mpc(cfg, setup=True); fx=fixture(cfg); m.S.set_options(max_iter=MAX_ITER)
    o=m.make_step_batch(fx["x0"], fx["u_prev"], fx["z0"] if "z0" in fx else None, tvp=fx.get("tvp"))
    try:
        res=compare_with_oracle(cfg, fx, np.arange(64), o["status"], o["iter_count"], o["u0"], m.recipe.u_scaling)
        print("cfg", cfg, {k:v for k,v in res.items() if k not in ("iter_diff","same_basin")})
    except AssertionError as e:
        print("cfg", cfg, "FAIL", str(e)[:300])
