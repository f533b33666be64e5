# round 2, second GPU call
mkdir -p gpurun_out
nvidia-smi -L; nproc; lscpu | grep -E "Model name|Socket|NUMA|Thread|Core" | head -8
timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2b_pytest.log
timeout 300 python scripts/gpu/hostpack_rate.py > gpurun_out/r2b_hostpack.log 2>&1; cat gpurun_out/r2b_hostpack.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"; cat gpurun_out/r2b_bench.json; tail -3 gpurun_out/r2b_bench.err
timeout 600 python bench.py --steps 30 --warmup 5 --as-shard 3/8 --no-cpu-baseline > gpurun_out/r2b_shard8.json 2> gpurun_out/r2b_shard8.err; echo "shard rc=$?"; cat gpurun_out/r2b_shard8.json
timeout 120 scripts/ubench/tma_gather4 1 > gpurun_out/r2b_gather4_box1.log 2>&1; echo "gather4 box1 rc=$?"; cat gpurun_out/r2b_gather4_box1.log
timeout 120 scripts/ubench/tma_gather4 4 > gpurun_out/r2b_gather4_box4.log 2>&1; echo "gather4 box4 rc=$?"; cat gpurun_out/r2b_gather4_box4.log
timeout 600 python bench.py --steps 3 --warmup 2 --as-shard 3/8 --no-cpu-baseline --no-e2e > gpurun_out/r2b_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2b_shard8_launches.csv \
    python bench.py --steps 3 --warmup 2 --as-shard 3/8 --no-cpu-baseline --no-e2e > gpurun_out/r2b_ncu.log 2>&1; echo "ncu rc=$?"
