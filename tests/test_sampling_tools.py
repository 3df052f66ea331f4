"""Generic sampling tools (reference do_mpc/sampling, testing/test_sampling_tools.py): planner -> sampler -> data handler,
per-sample and batch sample functions give the same files."""
import numpy as np

from do_mpc_b200.sampling import DataHandler, Sampler, SamplingPlanner


def test_planner_sampler_datahandler_roundtrip(tmp_path):
    np.random.seed(0)
    sp = SamplingPlanner(overwrite=True, data_dir=str(tmp_path) + "/")
    sp.set_sampling_var("alpha", np.random.randn)
    sp.set_sampling_var("beta", lambda: np.random.randint(0, 5))
    plan = sp.gen_sampling_plan(n_samples=10)
    sp.add_sampling_case(alpha=2.0, beta=3)
    sp.export("plan")
    assert len(plan) == 11 and plan[-1] == {"alpha": 2.0, "beta": 3, "id": "010"} and (tmp_path / "plan.pkl").exists()

    calls = []
    s1 = Sampler(plan, data_dir=str(tmp_path / "single") + "/", sample_name="dummy", overwrite=True)
    s1.set_sample_function(lambda alpha, beta: calls.append(1) or alpha * beta)
    s1.sample_data()
    assert len(calls) == 11 and len(s1.completion_list) == 11

    batch_calls = []
    def batch(alpha, beta):
        batch_calls.append(len(alpha))
        return list(np.asarray(alpha) * np.asarray(beta))
    s2 = Sampler(plan, data_dir=str(tmp_path / "batch") + "/", sample_name="dummy")
    s2.set_batch_sample_function(batch)
    s2.sample_data()
    s2.sample_data()                                    # nothing left to do: existing files are kept (overwrite=False)
    assert batch_calls == [11]

    for d in ("single", "batch"):
        dh = DataHandler(plan, data_dir=str(tmp_path / d) + "/", sample_name="dummy")
        dh.set_post_processing("res_1", lambda x: x)
        dh.set_post_processing("res_2", lambda x: x ** 2)
        rows = dh[:]
        assert len(rows) == 11 and all(np.isclose(r["res_1"], r["alpha"] * r["beta"]) and np.isclose(r["res_2"], r["res_1"] ** 2) for r in rows)
        assert dh[10][0]["res_1"] == 6.0
        sel = dh.filter(input_filter=lambda beta: beta >= 3, output_filter=lambda res_2: res_2 > 1.0)
        assert all(r["beta"] >= 3 and r["res_2"] > 1.0 for r in sel) and 0 < len(sel) < 11

    sp2 = SamplingPlanner(data_dir=str(tmp_path) + "/")
    sp2.set_sampling_var("a"); sp2.set_sampling_var("b")
    assert len(sp2.product(a=[1, 2, 3], b=[10, 20])) == 6
