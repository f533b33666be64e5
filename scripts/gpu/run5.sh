mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/plain.log 2>&1 && \
timeout 800 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 1 -c 1 -f -o gpurun_out/prof_v4 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/ncu.log | cut -c1-300
