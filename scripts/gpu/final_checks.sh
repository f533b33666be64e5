# after the tag revert: GPU tier, configs at full size, fuzz campaign
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
( time timeout 2000 python scripts/gpu/configs_probe.py cfg2 cfg4 cfg5 cfg5n1 > gpurun_out/r2_configs_final.jsonl 2> gpurun_out/r2_configs_final.err ); echo "probe rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2_configs_final.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); print(d['config'][:40], 'step', d['step_ms'], 'scan', d['scan_kernel_ms'], 'verify', d['verify_kernel_ms'], 'hits', d['hits'], 'found', d['planted_found'], 'exact', d['oracle_bit_exact'])
PY
( time timeout 1700 python scripts/gpu/fuzz_campaign.py --small 3000 --medium 100 --seed0 300000 --out gpurun_out/r2_fuzz_campaign_final.json ) > gpurun_out/fuzz.log 2>&1; echo "fuzz rc=$?"; tail -2 gpurun_out/fuzz.log
