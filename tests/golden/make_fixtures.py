"""Extract the reference's own golden data for the MPC.make_step -> nlpsol('ipopt') path into small fixtures.

Run HERE (build container; /root/reference is read-only and absent on the GPU box):
    python tests/golden/make_fixtures.py
Sources (reference tree):
  * testing/results/results_{oscillatingMasses,oscillatingMasses_dae,batch_reactor,CSTR,industrial_poly,dip}.pkl — 5 closed-loop steps each,
    fields opt_p_num / _opt_x_num (unscaled primal solution) / _lam_g_num / _u / _x / success (do_mpc/data.py:376-434,
    written by testing/test_*.py).
  * documentation/source/example_gallery/sampling_notebook/my_dataset/{sampling_plan_my_dataset,data_my_dataset_all}.pkl —
    1000 cold-start nominal-CSTR solves (x0, u_prev) -> (u0, status, iter_count) produced by IPOPT through
    do_mpc/approximateMPC/_ampc_sampler.py:420-468.
No CasADi is needed: the pickles are read with a stub unpickler (oracle/golden_io.py).
"""
import pickle
import sys
from pathlib import Path

import numpy as np
import pandas as pd

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.golden_io import load_results  # noqa: E402

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


def main():
    for name in ["oscillatingMasses", "oscillatingMasses_dae", "batch_reactor", "CSTR", "industrial_poly", "dip"]:
        m = load_results(REF / "testing" / "results" / f"results_{name}.pkl")["mpc"]
        np.savez_compressed(OUT / f"kkt_{name}.npz", opt_p_num=m["opt_p_num"], opt_x_num=m["_opt_x_num"],
                            lam_g_num=m["_lam_g_num"], u=m["_u"], x=m["_x"], success=m["success"], **({"z": m["_z"]} if name.endswith("_dae") else {}))
        print(name, m["_opt_x_num"].shape)
    d = REF / "documentation/source/example_gallery/sampling_notebook/my_dataset"
    plan = pickle.load(open(d / "sampling_plan_my_dataset.pkl", "rb"))
    df = pd.read_pickle(d / "data_my_dataset_all.pkl")
    n = len(plan)
    x0 = np.array([p["x0"].ravel() for p in plan]); up = np.array([p["u_prev"].ravel() for p in plan])
    ok = df["status"].to_numpy().astype(bool)
    u0 = np.full((n, 2), np.nan)
    for i in range(n):
        if ok[i]:
            u0[i] = np.asarray(df.iloc[i]["u0"]).ravel()
    np.savez_compressed(OUT / "cstr_dataset.npz", x0=x0, u_prev=up, u0=u0, success=ok,
                        iter_count=df["iter_count"].to_numpy().astype(int), t_make_step=df["t_make_step"].to_numpy())
    print("dataset", n, ok.sum())


if __name__ == "__main__":
    main()
