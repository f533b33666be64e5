# configs 4, 5 and 5-with-one-mismatch at full size on the new verifier / block tables; three-CTA verifier for comparison
mkdir -p gpurun_out
timeout 1500 python scripts/gpu/configs_probe.py cfg4 cfg5 cfg5n1 > gpurun_out/r2l_configs.jsonl 2> gpurun_out/r2l_configs.err; echo "probe rc=$?"; cut -c1-1200 gpurun_out/r2l_configs.jsonl; tail -3 gpurun_out/r2l_configs.err
MPCR_B200_LIB=$PWD/merpcr_b200/lib/libmerpcr_b200_v3.so timeout 900 python scripts/gpu/configs_probe.py cfg4 cfg5 --timing-only > gpurun_out/r2l_configs_v3.jsonl 2>&1; cat gpurun_out/r2l_configs_v3.jsonl
timeout 900 python scripts/gpu/configs_probe.py cfg4 cfg5 --timing-only > gpurun_out/r2l_configs_v4.jsonl 2>&1; cat gpurun_out/r2l_configs_v4.jsonl
