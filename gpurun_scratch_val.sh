mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/v_tests.log 2>&1
echo tests rc=$? >> gpurun_out/v_tests.log
python tools/latency_bench.py cstr_robust reactor_tree dip kite > gpurun_out/v_latency.json 2> gpurun_out/v_latency.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/v_smoke.log 2>&1
tail -14 gpurun_out/v_tests.log; cut -c1-220 gpurun_out/v_latency.json; tail -1 gpurun_out/v_smoke.log; tail -3 gpurun_out/v_latency.err
