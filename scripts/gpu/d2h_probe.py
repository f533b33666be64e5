import time, torch, numpy as np
dev = torch.device('cuda', 0)
hits = torch.randint(0, 255, (81071 * 24,), dtype=torch.uint8, device=dev)
pinned = torch.empty(1 << 22, dtype=torch.uint8).pin_memory()
torch.cuda.synchronize()
for it in range(4):
    t0 = time.perf_counter()
    st = pinned[: hits.numel()]
    st.copy_(hits, non_blocking=True)
    t1 = time.perf_counter()
    torch.cuda.current_stream().synchronize()
    t2 = time.perf_counter()
    a = st.numpy().view(np.dtype([('a', '<u4', 6)])).copy()
    t3 = time.perf_counter()
    b = hits.cpu().numpy()
    t4 = time.perf_counter()
    print(f"copy_ enqueue {1e3*(t1-t0):.3f} ms, sync {1e3*(t2-t1):.3f} ms, numpy copy {1e3*(t3-t2):.3f} ms | pageable .cpu() {1e3*(t4-t3):.3f} ms")
