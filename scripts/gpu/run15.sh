mkdir -p gpurun_out
MPCR_B200_LIB=$PWD/merpcr_b200/lib/libmerpcr_b200_sw.so timeout 900 python scripts/gpu/configs_probe.py cfg4 cfg5 --timing-only > gpurun_out/r2o_configs_sw.jsonl 2>&1; cat gpurun_out/r2o_configs_sw.jsonl
timeout 900 python bench.py > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; echo "bench rc=$?"; cat gpurun_out/r2o_bench.json; tail -3 gpurun_out/r2o_bench.err
