mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:verify_kernel -s 1 -c 1 -f -o gpurun_out/prof_verify $CMD > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log | cut -c1-200
