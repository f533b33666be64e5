mkdir -p gpurun_out
bash scripts/gpu/matrix2.sh default tagb q8 q6tagb
