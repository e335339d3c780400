import torch
dev = torch.device("cuda", 0)
for mb in (357, 1024):
    x = torch.empty(mb << 20, dtype=torch.uint8, device=dev).view(torch.int64)
    for _ in range(3):
        x.fill_(-100)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        x.fill_(-100)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"fill_ {mb} MB: {ms:.4f} ms = {mb * 1.048576 / ms:.0f} GB/s")
