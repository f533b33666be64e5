# round 2, third GPU call: one-call step, verify fix, hybrid wire, file -> text
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2c_pytest.log
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"; cat gpurun_out/r2c_bench.json; tail -5 gpurun_out/r2c_bench.err
timeout 600 python bench.py --steps 30 --warmup 5 --as-shard 3/8 --no-cpu-baseline > gpurun_out/r2c_shard8.json 2> gpurun_out/r2c_shard8.err; echo "shard rc=$?"; cat gpurun_out/r2c_shard8.json
MPCR_HYBRID_WIRE=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e-file > gpurun_out/r2c_nohybrid.json 2> gpurun_out/r2c_nohybrid.err; echo "nohybrid rc=$?"; python -c "
import json; j=json.loads([l for l in open('gpurun_out/r2c_nohybrid.json') if l.startswith('{')][-1]); print('no-hybrid e2e', j['e2e'])"
timeout 600 python bench.py --steps 3 --warmup 2 --as-shard 3/8 --no-cpu-baseline --no-e2e > gpurun_out/r2c_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'scan_kernel|verify_kernel|rank_sort|rs_sort|order_' -c 60 --csv --log-file gpurun_out/r2c_shard8_launches.csv \
    python bench.py --steps 3 --warmup 2 --as-shard 3/8 --no-cpu-baseline --no-e2e > gpurun_out/r2c_ncu.log 2>&1; echo "ncu rc=$?"
