# paired (FFMA2) stage 1: micro-benchmark with the packed variants, parity of the candidate, A/B
mkdir -p gpurun_out
( cd scripts/ubench && ./pipes ) > gpurun_out/ubench_pipes.log 2>&1; tail -4 gpurun_out/ubench_pipes.log
( MPCR_B200_LIB=$PWD/merpcr_b200/lib/libmerpcr_b200_pairfptrim.so timeout 900 python -m pytest tests/test_gpu_parity.py -x -q ) > gpurun_out/pytest_parity_pairfptrim.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/pytest_parity_pairfptrim.log
bash scripts/gpu/matrix2.sh default pair pairfp pairtrim pairfptrim linfptrim2
for v in s1pair s1pairfp; do
  MPCR_B200_LIB=$PWD/merpcr_b200/lib/libmerpcr_b200_$v.so timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_$v.log 2>&1
  python - <<PY
import json
l=[x for x in open('gpurun_out/bench_$v.log') if x.startswith('{')]
if l:
    j=json.loads(l[-1]); print('variant $v: scan ms', round(j['roofline']['kernel_ms'],4))
else: print(open('gpurun_out/bench_$v.log').read()[-600:])
PY
done
