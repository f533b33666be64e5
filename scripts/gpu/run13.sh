# verifier: static schedule for every list length (variant sall) on the heavy configs; hybrid wire sweep
mkdir -p gpurun_out
MPCR_B200_LIB=$PWD/merpcr_b200/lib/libmerpcr_b200_sall.so timeout 900 python scripts/gpu/configs_probe.py cfg4 cfg5 --timing-only > gpurun_out/r2m_configs_sall.jsonl 2>&1; cat gpurun_out/r2m_configs_sall.jsonl
timeout 900 python scripts/gpu/configs_probe.py cfg4 --timing-only > gpurun_out/r2m_configs_def.jsonl 2>&1; cat gpurun_out/r2m_configs_def.jsonl
timeout 900 python scripts/gpu/hybrid_probe.py > gpurun_out/r2m_hybrid.log 2>&1; cat gpurun_out/r2m_hybrid.log
