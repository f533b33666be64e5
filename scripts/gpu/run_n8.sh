# 8-GPU call: the sharded bench at N = 8 (and N = 4 with "4" as the first argument too)
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "bench n8 rc=$?"; grep -v "OMP_NUM\|^\*\*\*\|unbatched P2P" gpurun_out/r2_bench_n8.err | tail -5
if [ "$1" = 4 ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err; echo "bench n4 rc=$?"
fi
python - <<'PY'
import json, os
for n in (8,4):
    f=f'gpurun_out/r2_bench_n{n}.json'
    if not os.path.exists(f): continue
    try:
        j=json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(n, 'step', round(j['ms_per_step'],4), 'synced', round(j['config']['ms_per_step_host_synced'],4), 'scan', round(j['roofline']['kernel_ms'],4), 'verify', round(j['roofline']['verify_kernel_ms'],4), 'e2e', (j.get('e2e') or {}).get('ms_per_step'), 'file', (j.get('e2e_file') or {}).get('ms'))
        print('  per rank', [r['ms_per_step'] for r in j['config']['per_rank']])
    except Exception as e: print(n, 'failed', e)
PY
