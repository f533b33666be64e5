# fuzz campaign on the final kernels (linear filter map, ten-letter tags, warp-per-block ingest)
mkdir -p gpurun_out
( time timeout 1700 python scripts/gpu/fuzz_campaign.py --small 3000 --medium 100 --seed0 300000 --out gpurun_out/r2_fuzz_campaign_final.json ) > gpurun_out/fuzz.log 2>&1; echo "fuzz rc=$?"; tail -6 gpurun_out/fuzz.log
