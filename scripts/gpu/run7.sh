mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2g_pytest.log
timeout 1500 python scripts/gpu/configs_probe.py cfg5 > gpurun_out/r2g_cfg5.jsonl 2> gpurun_out/r2g_cfg5.err; echo "cfg5 rc=$?"; cat gpurun_out/r2g_cfg5.jsonl; tail -3 gpurun_out/r2g_cfg5.err
MPCR_SAMPLING=0 timeout 1500 python scripts/gpu/configs_probe.py cfg5 > gpurun_out/r2g_cfg5_nosamp.jsonl 2> gpurun_out/r2g_cfg5_nosamp.err; echo "cfg5 nosamp rc=$?"; cat gpurun_out/r2g_cfg5_nosamp.jsonl
