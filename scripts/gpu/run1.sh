set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv | tee gpurun_out/gpu.txt
nproc | tee gpurun_out/nproc.txt
python __graft_entry__.py > gpurun_out/build.log 2>&1; tail -2 gpurun_out/build.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -30 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -5 gpurun_out/smoke.log
timeout 600 python bench.py --scale 0.05 --steps 5 --warmup 3 > gpurun_out/bench_small.log 2>&1; echo "bench small exit $?"; tail -5 gpurun_out/bench_small.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.log 2>&1; echo "bench full exit $?"; tail -5 gpurun_out/bench_full.log
timeout 600 compute-sanitizer --tool memcheck --log-file gpurun_out/memcheck.log python __graft_entry__.py --smoke > gpurun_out/smoke_memcheck.log 2>&1; echo "memcheck exit $?"; tail -5 gpurun_out/memcheck.log
