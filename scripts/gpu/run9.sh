# ingest kernels after the fusion: parity tests, the per-kernel ncu table, file -> text timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "fasta or ingest or fixture or goldens_bit_exact or cli" > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2i_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?"; python -c "
import json; j=json.loads([l for l in open('gpurun_out/r2i_bench.json') if l.startswith('{')][-1]); print('step', j['ms_per_step'], 'e2e', j['e2e']['ms_per_step'], 'file', j['e2e_file'])"
timeout 900 bash scripts/gpu/prof_kernels.sh 512 > gpurun_out/r2i_kernels_table.txt 2>&1; echo "prof rc=$?"; cat gpurun_out/r2i_kernels_table.txt
