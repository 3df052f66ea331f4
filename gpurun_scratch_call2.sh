mkdir -p gpurun_out
Q="python tools/quick_bench.py 2 --steps 5"
$Q > gpurun_out/c2_base_a.json 2> gpurun_out/c2_err.txt
DMB_EXTRA_NVCC_FLAGS=-DDMB_NOINLINE_PHASES $Q > gpurun_out/c2_noinline.json 2>> gpurun_out/c2_err.txt
DMB_EXTRA_NVCC_FLAGS=-DDMB_EB_LDCG $Q > gpurun_out/c2_ebldcg.json 2>> gpurun_out/c2_err.txt
DMB_EXTRA_NVCC_FLAGS=-Xptxas=-dlcm=cg $Q > gpurun_out/c2_dlcmcg.json 2>> gpurun_out/c2_err.txt
DMB_FORCE_THREADS=128 $Q > gpurun_out/c2_t128.json 2>> gpurun_out/c2_err.txt
DMB_FORCE_THREADS=192 $Q > gpurun_out/c2_t192.json 2>> gpurun_out/c2_err.txt
DMB_L2_PERSIST=1 $Q > gpurun_out/c2_l2persist.json 2>> gpurun_out/c2_err.txt
$Q > gpurun_out/c2_base_b.json 2>> gpurun_out/c2_err.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dmb_solve -s 1 -c 1 -f -o gpurun_out/r2t_cfg3 python tools/quick_bench.py 3 --steps 1 > gpurun_out/c2_ncu_cfg3.log 2>&1
for f in base_a noinline ebldcg dlcmcg t128 t192 l2persist base_b; do echo $f $(cut -c1-130 gpurun_out/c2_$f.json); done
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/c2_ncu_cfg3.log
