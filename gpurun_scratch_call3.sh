mkdir -p gpurun_out
Q="python tools/quick_bench.py"
$Q 2 3 4 5 --steps 4 > gpurun_out/c3_flat.json 2> gpurun_out/c3_err.txt
DMB_EXTRA_NVCC_FLAGS="-DDMB_LU_FLAT=0 -DDMB_BS_FLAT=0" $Q 2 3 4 5 --steps 4 > gpurun_out/c3_base.json 2>> gpurun_out/c3_err.txt
DMB_EXTRA_NVCC_FLAGS="-DDMB_BS_FLAT=0" $Q 3 4 --steps 4 > gpurun_out/c3_lu_only.json 2>> gpurun_out/c3_err.txt
$Q 2 --steps 4 >> gpurun_out/c3_flat.json 2>> gpurun_out/c3_err.txt
for f in flat base lu_only; do echo == $f; cut -c1-175 gpurun_out/c3_$f.json; done
timeout 700 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/c3_tests.log 2>&1
echo tests rc=$? >> gpurun_out/c3_tests.log
tail -12 gpurun_out/c3_tests.log
