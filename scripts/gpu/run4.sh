mkdir -p gpurun_out
UB_MODES=1 timeout 200 scripts/ubench/ubench > gpurun_out/ubench_modes.log 2>&1; cat gpurun_out/ubench_modes.log
python bench.py --scale 0.2 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 1 -c 1 -f -o gpurun_out/prof_v2 python bench.py --scale 0.2 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/ncu.log | cut -c1-300
