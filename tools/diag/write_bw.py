import torch, statistics
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts=[]
    for _ in range(n):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return statistics.median(ts)
x=torch.empty(600*1024*1024, dtype=torch.float32, device="cuda")   # 2.4 GB
y=torch.empty_like(x)
ms=t(lambda: x.zero_()); print(f"memset 2.4GB (torch fill): {ms*1e3:.0f} us -> {x.numel()*4/ms/1e6:.0f} GB/s")
ms=t(lambda: torch.cuda.current_stream().synchronize() or x.fill_(1.0)); print(f"fill: {ms*1e3:.0f} us -> {x.numel()*4/ms/1e6:.0f} GB/s")
ms=t(lambda: y.copy_(x)); print(f"copy 2.4GB->2.4GB: {ms*1e3:.0f} us -> {2*x.numel()*4/ms/1e6:.0f} GB/s (r+w)")
import ctypes
ms=t(lambda: torch.cuda.memset if False else x.zero_())
z=torch.empty(150*1024*1024, dtype=torch.float32, device="cuda")
ms=t(lambda: z.zero_()); print(f"memset 0.6GB: {ms*1e3:.0f} us -> {z.numel()*4/ms/1e6:.0f} GB/s")
ms=t(lambda: x.sum()); print(f"read-only reduce 2.4GB: {ms*1e3:.0f} us -> {x.numel()*4/ms/1e6:.0f} GB/s")
