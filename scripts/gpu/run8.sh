mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2h_pytest.log
for sh in 3/8 1/2; do
timeout 600 python bench.py --steps 30 --warmup 5 --as-shard $sh --no-cpu-baseline --no-e2e > gpurun_out/r2h_shard.json 2> gpurun_out/r2h_shard.err; echo "shard $sh rc=$?"; tail -2 gpurun_out/r2h_shard.err; python -c "
import json; j=json.loads([l for l in open('gpurun_out/r2h_shard.json') if l.startswith('{')][-1]); print('$sh step ms', j['ms_per_step'], 'synced', j['config']['ms_per_step_host_synced'], 'scan', j['roofline']['kernel_ms'], 'verify', j['roofline']['verify_kernel_ms'], 'hits', j['config']['hits_per_gpu'], 'launches', j['gpu_launches'])"
done
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2h_n1.json 2> gpurun_out/r2h_n1.err; echo "n1 rc=$?"; cat gpurun_out/r2h_n1.json; tail -2 gpurun_out/r2h_n1.err
