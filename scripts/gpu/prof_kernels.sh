# Per-kernel ncu table: duration + DRAM bytes of every launch of every kernel of the path (profiles/ gets the summary)
mkdir -p gpurun_out
CMD="python scripts/gpu/all_kernels.py ${1:-512}"
$CMD > gpurun_out/all_kernels_plain.log 2>&1 || { tail -5 gpurun_out/all_kernels_plain.log; exit 1; }
cat gpurun_out/all_kernels_plain.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size,launch__block_size \
    --clock-control none -c 2000 --csv --log-file gpurun_out/kernels.csv $CMD > gpurun_out/ncu_kernels.log 2>&1
tail -2 gpurun_out/ncu_kernels.log | cut -c1-200
python scripts/ncu_kernels_table.py gpurun_out/kernels.csv
