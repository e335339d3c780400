import torch, time
dev = torch.device("cuda", 0)
for mb in (53, 106, 212):
    n = mb * 1024 * 1024 // 8
    h = torch.empty(n, dtype=torch.int64).pin_memory()
    d = torch.empty(n, dtype=torch.int64, device=dev)
    for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
        for _ in range(2): fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): fn()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        print(f"{name} {mb} MB: {ms:.3f} ms  {mb * 1.048576 / ms:.1f} GB/s")
# both directions at once on two streams
n = 212 * 1024 * 1024 // 8
h1 = torch.empty(n, dtype=torch.int64).pin_memory(); d1 = torch.empty(n, dtype=torch.int64, device=dev)
h2 = torch.empty(n // 4, dtype=torch.int64).pin_memory(); d2 = torch.empty(n // 4, dtype=torch.int64, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize()
print("duplex 212 MB H2D + 53 MB D2H:", (time.perf_counter() - t0) / 5 * 1e3, "ms")
