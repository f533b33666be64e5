# warp-per-block ingest kernels: GPU tier, per-kernel ncu table, bench with the file leg
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
bash scripts/gpu/prof_kernels.sh 512 > gpurun_out/prof_kernels.log 2>&1; grep -E "^kernel|fasta_|pack_kernel|derive" gpurun_out/prof_kernels.log
( timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_n1_ingest.log 2> gpurun_out/bench_n1_ingest.err; echo "bench rc=$?"
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_n1_ingest.log') if x.startswith('{')]
if l:
    j=json.loads(l[-1]); print('step', round(j['ms_per_step'],4), 'e2e', j['e2e']['ms_per_step'], 'file', j['e2e_file'])
PY
