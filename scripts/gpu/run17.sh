mkdir -p gpurun_out
for v in s1 s1nc; do
  MPCR_B200_LIB=$PWD/merpcr_b200/lib/libmerpcr_b200_$v.so timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_$v.log 2>&1
  python - <<PY
import json
l=[x for x in open('gpurun_out/bench_$v.log') if x.startswith('{')]
if l:
    j=json.loads(l[-1]); print('variant $v: scan ms', round(j['roofline']['kernel_ms'],4))
else: print(open('gpurun_out/bench_$v.log').read()[-600:])
PY
done
