# scan-kernel time against shard size (one GPU): fixed cost of a launch vs per-base rate
mkdir -p gpurun_out
for s in 5/1024 5/256 5/64 5/32 3/16 3/8 1/4 1/2; do
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --as-shard $s > gpurun_out/shard_line.log 2>&1
  python - <<PY
import json
l=[x for x in open('gpurun_out/shard_line.log') if x.startswith('{')]
if l:
    j=json.loads(l[-1]); r=j['roofline']
    print('shard $s bp', j['config']['bp_per_gpu'], 'step ms', round(j['ms_per_step'],4), 'scan', round(r['kernel_ms'],4), 'verify', round(r['verify_kernel_ms'],4), 'hits', j['config']['hits_per_gpu'])
else: print('shard $s FAILED', open('gpurun_out/shard_line.log').read()[-300:])
PY
done
