# 10-letter tags: GPU tier + default bench (both shapes)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
bash scripts/gpu/matrix2.sh default
