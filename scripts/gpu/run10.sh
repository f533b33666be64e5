# verifier schedule / scanner prologue: parity tests on the product build, then the variant matrix
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2j_pytest.log
bash scripts/gpu/matrix2.sh default v4 v2 dyn
