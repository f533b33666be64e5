# usage: matrix.sh variant1 variant2 ...   ("default" = the product build); prints scan-kernel ms per variant
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = default ]; then lib=$PWD/merpcr_b200/lib/libmerpcr_b200.so; else lib=$PWD/merpcr_b200/lib/libmerpcr_b200_$v.so; fi
  MPCR_B200_LIB=$lib timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_m_$v.log 2>&1
  python - <<PY
import json
l=[x for x in open('gpurun_out/bench_m_$v.log') if x.startswith('{')]
if l:
    j=json.loads(l[-1]); print('variant $v: step ms', round(j['ms_per_step'],3), 'scan ms', round(j['roofline']['kernel_ms'],3), 'hits', j['config']['hits_per_gpu'], 'found', j['config']['planted_found'], 'sorted', j['config']['sorted'])
else:
    print('variant $v FAILED'); print(open('gpurun_out/bench_m_$v.log').read()[-800:])
PY
done
