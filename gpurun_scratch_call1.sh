mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_gpu.txt 2>&1
python tools/quick_bench.py 2 --steps 5 > gpurun_out/c1_reglu1.json 2> gpurun_out/c1_reglu1.err
DMB_EXTRA_NVCC_FLAGS=-DDMB_REG_LU=0 python tools/quick_bench.py 2 --steps 5 > gpurun_out/c1_reglu0.json 2> gpurun_out/c1_reglu0.err
python tools/quick_bench.py 2 --steps 5 >> gpurun_out/c1_reglu1.json 2>> gpurun_out/c1_reglu1.err
timeout 800 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/c1_tests.log 2>&1
echo tests rc=$? >> gpurun_out/c1_tests.log
tail -5 gpurun_out/c1_tests.log; cat gpurun_out/c1_reglu1.json gpurun_out/c1_reglu0.json | cut -c1-400
