# final-state validation: GPU test tier, default bench (both arms), launch list + full capture of the scan kernel
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_n1.err
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"
bash scripts/gpu/profile.sh
