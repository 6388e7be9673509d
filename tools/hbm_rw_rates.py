import torch
x=torch.empty(1<<30, dtype=torch.float32, device='cuda')   # 4 GiB
y=torch.empty_like(x)
def t(f,n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n*1e-3
b=x.numel()*4
print("write-only fill_   GB/s", b/t(lambda: x.fill_(1.0))/1e9)
print("write-only zero_   GB/s", b/t(lambda: x.zero_())/1e9)
print("read-only sum      GB/s", b/t(lambda: x.sum())/1e9)
print("copy (r+w counted) GB/s", 2*b/t(lambda: y.copy_(x))/1e9)
print("memset cudaMemsetAsync GB/s", b/t(lambda: torch.cuda.memory._cudart if False else x.view(torch.uint8).zero_())/1e9)
