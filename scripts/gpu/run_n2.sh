# 2-GPU call: the multi-GPU CLI test (merged text == single-GPU text) and the sharded bench at N=2
mkdir -p gpurun_out
nvidia-smi -L; nproc
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 900 -rA > gpurun_out/r2_gpu_multi_pytest.log 2>&1; echo "multi pytest rc=$?"; tail -8 gpurun_out/r2_gpu_multi_pytest.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench n2 rc=$?"; cat gpurun_out/r2_bench_n2.json; grep -v "OMP_NUM\|^\*\*\*" gpurun_out/r2_bench_n2.err | tail -5
