mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "sort or regrows or long_runs or synthetic or append" > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2e_pytest.log
for sh in 3/8 1/2; do
timeout 600 python bench.py --steps 30 --warmup 5 --as-shard $sh --no-cpu-baseline --no-e2e > gpurun_out/r2e_shard.json 2> gpurun_out/r2e_shard.err; echo "shard $sh rc=$?"; python -c "
import json; j=json.loads([l for l in open('gpurun_out/r2e_shard.json') if l.startswith('{')][-1]); print('$sh step ms', j['ms_per_step'], 'scan', j['roofline']['kernel_ms'], 'verify', j['roofline']['verify_kernel_ms'], 'hits', j['config']['hits_per_gpu'], 'launches', j['gpu_launches'])"
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2e_n1.json 2> gpurun_out/r2e_n1.err; echo "n1 rc=$?"; python -c "
import json; j=json.loads([l for l in open('gpurun_out/r2e_n1.json') if l.startswith('{')][-1]); print('N=1 step ms', j['ms_per_step'], 'scan', j['roofline']['kernel_ms'], 'verify', j['roofline']['verify_kernel_ms'], 'found', j['config']['planted_found'], j['config']['sorted'], 'launches', j['gpu_launches'])"
timeout 600 python bench.py --steps 3 --warmup 2 --as-shard 3/8 --no-cpu-baseline --no-e2e > gpurun_out/r2e_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'scan_kernel|verify_kernel|bsort|rs_sort|order_' -c 45 --csv --log-file gpurun_out/r2e_shard8_launches.csv \
    python bench.py --steps 3 --warmup 2 --as-shard 3/8 --no-cpu-baseline --no-e2e > gpurun_out/r2e_ncu.log 2>&1; echo "ncu rc=$?"
