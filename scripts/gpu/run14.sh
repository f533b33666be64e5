mkdir -p gpurun_out
timeout 900 python scripts/gpu/configs_probe.py cfg4 cfg5 --timing-only > gpurun_out/r2n_configs_def.jsonl 2>&1; cat gpurun_out/r2n_configs_def.jsonl
MPCR_B200_LIB=$PWD/merpcr_b200/lib/libmerpcr_b200_ee.so timeout 900 python scripts/gpu/configs_probe.py cfg4 cfg5 --timing-only > gpurun_out/r2n_configs_ee.jsonl 2>&1; cat gpurun_out/r2n_configs_ee.jsonl
timeout 900 python scripts/gpu/configs_probe.py cfg5n1 > gpurun_out/r2n_cfg5n1.jsonl 2>&1; cut -c1-900 gpurun_out/r2n_cfg5n1.jsonl
