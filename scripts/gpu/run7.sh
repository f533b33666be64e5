mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
for dbg in 1 0; do
MPCR_DEBUG=$dbg timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_dbg$dbg.log 2>&1
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_dbg$dbg.log') if x.startswith('{')]
if l:
    j=json.loads(l[-1]); print('dbg$dbg step ms', round(j['ms_per_step'],3), 'scan ms', round(j['roofline']['kernel_ms'],3), 'count', j['config']['hits_per_gpu'], 'found', j['config']['planted_found'])
else:
    print('dbg$dbg FAILED'); print(open('gpurun_out/bench_dbg$dbg.log').read()[-800:])
PY
done
timeout 800 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 1 -c 1 -f -o gpurun_out/prof_cur python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log | cut -c1-200
