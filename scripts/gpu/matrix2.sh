# usage: matrix2.sh variant1 variant2 ...   ("default" = the product build)
# per variant: step / scan-kernel / verify-kernel ms for the whole cfg3 genome and for rank 3's share of an 8-way split
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = default ]; then lib=$PWD/merpcr_b200/lib/libmerpcr_b200.so; else lib=$PWD/merpcr_b200/lib/libmerpcr_b200_$v.so; fi
  for mode in whole shard; do
    extra=""; [ $mode = shard ] && extra="--as-shard 3/8"
    MPCR_B200_LIB=$lib timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e $extra > gpurun_out/bench_m2_${v}_$mode.log 2>&1
    python - <<PY
import json
l=[x for x in open('gpurun_out/bench_m2_${v}_$mode.log') if x.startswith('{')]
if l:
    j=json.loads(l[-1]); r=j['roofline']
    print('variant $v $mode: step ms', round(j['ms_per_step'],4), 'synced', round(j['config']['ms_per_step_host_synced'],4), 'scan', round(r['kernel_ms'],4), 'verify', round(r['verify_kernel_ms'],4), 'hits', j['config']['hits_per_gpu'], 'found', j['config']['planted_found'], 'sorted', j['config']['sorted'])
else:
    print('variant $v $mode FAILED'); print(open('gpurun_out/bench_m2_${v}_$mode.log').read()[-800:])
PY
  done
done
