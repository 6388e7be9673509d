import torch
x = torch.empty(2**30, dtype=torch.float32, device='cuda')  # 4 GB
y = torch.empty_like(x)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
ms = t(lambda: x.fill_(1.0)); print("fill (pure write) GB/s", 4.295/ms*1e3)
ms = t(lambda: x.zero_()); print("memset (pure write) GB/s", 4.295/ms*1e3)
ms = t(lambda: torch.sum(x)); print("sum (pure read) GB/s", 4.295/ms*1e3)
ms = t(lambda: y.copy_(x)); print("copy (r+w) GB/s", 2*4.295/ms*1e3)
