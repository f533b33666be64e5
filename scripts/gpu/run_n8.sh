mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "bench n8 rc=$?"; grep -v "OMP_NUM\|^\*\*\*\|unbatched P2P" gpurun_out/r2_bench_n8.err | tail -5
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 30 --warmup 5 --no-e2e > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err; echo "bench n4 rc=$?"
