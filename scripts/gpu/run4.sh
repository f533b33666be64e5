# round 2, fourth GPU call: small-list sort in one CTA, verifier polling, profile of the 1/8-shard step
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2d_pytest.log
timeout 600 python bench.py --steps 30 --warmup 5 --as-shard 3/8 --no-cpu-baseline --no-e2e > gpurun_out/r2d_shard8.json 2> gpurun_out/r2d_shard8.err; echo "shard8 rc=$?"; cat gpurun_out/r2d_shard8.json
timeout 600 python bench.py --steps 30 --warmup 5 --as-shard 1/2 --no-cpu-baseline --no-e2e > gpurun_out/r2d_shard2.json 2> gpurun_out/r2d_shard2.err; echo "shard2 rc=$?"; cat gpurun_out/r2d_shard2.json
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2d_n1.json 2> gpurun_out/r2d_n1.err; echo "n1 rc=$?"; cat gpurun_out/r2d_n1.json
timeout 600 python bench.py --steps 3 --warmup 2 --as-shard 3/8 --no-cpu-baseline --no-e2e > gpurun_out/r2d_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'scan_kernel|verify_kernel|sort_small' -s 9 -c 3 -o gpurun_out/r2d_shard8_prof \
    python bench.py --steps 3 --warmup 2 --as-shard 3/8 --no-cpu-baseline --no-e2e > gpurun_out/r2d_ncu.log 2>&1; echo "ncu rc=$?"; ls -la gpurun_out/*.ncu-rep
