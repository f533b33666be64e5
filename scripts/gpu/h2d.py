import time, torch
dev = torch.device('cuda', 0)
for mb in (64, 256, 1024):
    h = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    d = torch.empty(mb << 20, dtype=torch.uint8, device=dev)
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"H2D pinned {mb} MiB: {ms:.2f} ms  {(mb << 20) / ms / 1e6:.1f} GB/s")
    e0.record()
    for _ in range(5):
        h.copy_(d, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"D2H pinned {mb} MiB: {ms:.2f} ms  {(mb << 20) / ms / 1e6:.1f} GB/s")
# two streams concurrently
h1 = torch.empty(512 << 20, dtype=torch.uint8).pin_memory(); h2 = torch.empty(512 << 20, dtype=torch.uint8).pin_memory()
d1 = torch.empty(512 << 20, dtype=torch.uint8, device=dev); d2 = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(3):
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"H2D two streams: {3 * 2 * (512 << 20) / dt / 1e9:.1f} GB/s")
import subprocess
print(subprocess.run(['nvidia-smi', '--query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max', '--format=csv'], capture_output=True, text=True).stdout)
