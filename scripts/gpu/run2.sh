set -x
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/exp_phases.log 2>&1
import os, sys, time
sys.path[:0] = ['.', 'tests']
import numpy as np, torch
import synth
from merpcr_b200 import MerPCR
dev = torch.device('cuda', 0)
L = 400_000_000
genome = synth.dna_torch(7, 0, L, dev)
for n_sts in (1000, 10000, 100000):
    sts = synth.make_sts_set(8, n_sts)
    open('/tmp/x.sts', 'wb').write(synth.sts_lines(sts))
    for dbg in ('0', '1'):
        os.environ['MPCR_DEBUG'] = dbg
        eng = MerPCR(mismatches=1, device=0)
        eng.load_sts_file('/tmp/x.sts')
        lay = eng.make_layout([L])
        sh = eng.upload(lay, [genome])
        for _ in range(3):
            eng.scan_device(lay, sh, sort=False)
        ms = []
        for _ in range(5):
            _, n = eng.scan_device(lay, sh, sort=False)
            ms.append(eng._be.lib.mpcr_last_scan_ms(eng._ctx))
        print(f"n_sts={n_sts} debug={dbg} count={n} scan_ms={np.mean(ms):.3f} Gbp/s={L/np.mean(ms)/1e6:.1f}", flush=True)
        eng.close()
PY
cat gpurun_out/exp_phases.log
python bench.py --scale 0.05 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 1 -c 1 -o gpurun_out/prof_v1 python bench.py --scale 0.05 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/ncu.log
