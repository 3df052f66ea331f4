mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/f_tests.log 2>&1
echo tests rc=$? >> gpurun_out/f_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1
python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs --no-traffic > gpurun_out/f_bench_under_ncu.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:dmb_solve -s 2 -c 1 -f -o gpurun_out/r2f_cfg2 python tools/quick_bench.py 2 --steps 2 > gpurun_out/f_ncu_cfg2.log 2>&1
python tools/quick_bench.py 2 3 4 5 --steps 3 > gpurun_out/f_quick.json 2>> gpurun_out/f_bench.err
tail -4 gpurun_out/f_tests.log; tail -2 gpurun_out/f_smoke.log; cut -c1-600 gpurun_out/f_bench.json; cut -c1-150 gpurun_out/f_quick.json; ls -la gpurun_out/*.ncu-rep
python tools/latency_bench.py cstr_robust reactor_tree cstr_nominal > gpurun_out/lat_rule.json 2> gpurun_out/lat_err.txt
DMB_FORCE_THREADS=256 DMB_FORCE_MIN_CTAS=1 python tools/latency_bench.py cstr_robust reactor_tree cstr_nominal > gpurun_out/lat_256.json 2>> gpurun_out/lat_err.txt
DMB_FORCE_THREADS=512 DMB_FORCE_MIN_CTAS=1 python tools/latency_bench.py reactor_tree > gpurun_out/lat_512.json 2>> gpurun_out/lat_err.txt
for f in rule 256 512; do echo == lat $f; cut -c1-200 gpurun_out/lat_$f.json; done
