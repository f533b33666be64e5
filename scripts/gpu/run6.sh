mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
for v in "" _t512i2 _t768i2 _t768i3 _t1024i2; do
  for dbg in 1 0; do
    MPCR_B200_LIB=$PWD/merpcr_b200/lib/libmerpcr_b200$v.so MPCR_DEBUG=$dbg timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_v${v}_dbg$dbg.log 2>&1
    python - <<PY
import json
l=[x for x in open('gpurun_out/bench_v${v}_dbg$dbg.log') if x.startswith('{')]
if l:
    j=json.loads(l[-1]); print('variant[$v] dbg$dbg step ms', round(j['ms_per_step'],3), 'scan ms', round(j['roofline']['kernel_ms'],3), 'count', j['config']['hits_per_gpu'], 'found', j['config']['planted_found'])
else:
    print('variant[$v] dbg$dbg FAILED'); print(open('gpurun_out/bench_v${v}_dbg$dbg.log').read()[-600:])
PY
  done
done
