# round 2, first GPU call: GPU test tier, the bench line, scan-kernel variants, the TMA gather4 micro-benchmark
mkdir -p gpurun_out
nvidia-smi -L; nproc; free -g | head -2
timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2a_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"; cat gpurun_out/r2a_bench.json
timeout 900 bash scripts/gpu/matrix.sh default qp qp_ilp3 > gpurun_out/r2a_matrix.log 2>&1; cat gpurun_out/r2a_matrix.log
timeout 300 scripts/ubench/tma_gather4 > gpurun_out/r2a_gather4.log 2>&1; echo "gather4 rc=$?"; cat gpurun_out/r2a_gather4.log
