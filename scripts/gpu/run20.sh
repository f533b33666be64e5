# linear (FP-assisted) filter map: pipe-overlap micro-benchmark, parity, A/B against the multiplicative map
mkdir -p gpurun_out
( cd scripts/ubench && ./pipes ) > gpurun_out/ubench_pipes.log 2>&1; cat gpurun_out/ubench_pipes.log
( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q ) > gpurun_out/pytest_parity.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/pytest_parity.log
bash scripts/gpu/matrix2.sh default nolin linfp linfptrim
for v in s1lin s1nolin s1linfp; do
  MPCR_B200_LIB=$PWD/merpcr_b200/lib/libmerpcr_b200_$v.so timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_$v.log 2>&1
  python - <<PY
import json
l=[x for x in open('gpurun_out/bench_$v.log') if x.startswith('{')]
if l:
    j=json.loads(l[-1]); print('variant $v: scan ms', round(j['roofline']['kernel_ms'],4), 'candidates', j['config']['hits_per_gpu'])
else: print(open('gpurun_out/bench_$v.log').read()[-600:])
PY
done
