mkdir -p gpurun_out
CMD="python scripts/gpu/configs_probe.py cfg5 --scale=0.05"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dense_scan_kernel -s 1 -c 1 -f -o gpurun_out/prof_dense $CMD > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log | cut -c1-200
