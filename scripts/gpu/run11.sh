# block tables: parity on the GPU, the dense-vs-blocks probe, and the verifier variants again (hot-kernel regression check)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2k_pytest.log
timeout 900 python scripts/gpu/blocks_probe.py 256 100000 > gpurun_out/r2k_blocks_probe.log 2>&1; echo "probe rc=$?"; cat gpurun_out/r2k_blocks_probe.log | tail -12
bash scripts/gpu/matrix2.sh default v4 l4 v4l4
