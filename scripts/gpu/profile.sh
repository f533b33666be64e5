# Round evidence: launch list of one bench step + full capture of the dominant kernel (profiles/ gets the summaries)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:scan_kernel|verify_kernel|rs_|bsort_|order_ties|order_long_runs|pack_kernel|derive_planes|encode_records|build_buckets|mark_chains" -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 1 -c 1 -f -o gpurun_out/prof_scan $CMD > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log | cut -c1-200
