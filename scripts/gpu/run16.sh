mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2p_pytest.log
bash scripts/gpu/matrix2.sh default
MPCR_SMALL_SORT=0 bash scripts/gpu/matrix2.sh default
