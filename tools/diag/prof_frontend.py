"""Where a frontend keyframe step spends its GPU time through the host mirror (CorrBlock / PooledCorrBlock):
8 new edges built (offset heads + Gaussian head + tcgen05 build), then lookups over the 48-edge window."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch, torch.nn as nn
from torch.profiler import profile, ProfilerActivity
import inputs, lgu_slam_b200
from importlib import import_module
corr = import_module("lgu-slam_b200.corr")
dev = "cuda"
H, W, C = 48, 64, 128
g = inputs.gen(3)
torch.manual_seed(0)
ofsMap = nn.Conv2d(256, 98, 3, padding=1).to(dev); ofs_res = nn.Conv2d(256, 98, 3, padding=1).to(dev)
GA = corr.GaussianMask(H, W).to(dev)
f1n = torch.randn(1, 8, C, H, W, generator=g).half().to(dev); f2n = torch.randn(1, 8, C, H, W, generator=g).half().to(dev)
f1w = torch.randn(1, 48, C, H, W, generator=g).half().to(dev); f2w = torch.randn(1, 48, C, H, W, generator=g).half().to(dev)
coords = inputs.make_coords(48, H, W, H, W, g).permute(0, 2, 3, 1).contiguous().view(1, 48, H, W, 2).to(dev)

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    return statistics.median(ts)

with torch.no_grad(), torch.autocast("cuda", enabled=True):
    win = corr.CorrBlock(ofsMap, ofs_res, GA, f1w, f2w)
    print("CorrBlock.__init__ 8 new edges  : %8.1f us" % timeit(lambda: corr.CorrBlock(ofsMap, ofs_res, GA, f1n, f2n)))
    print("CorrBlock.__init__ 48 edges     : %8.1f us" % timeit(lambda: corr.CorrBlock(ofsMap, ofs_res, GA, f1w, f2w)))
    print("CorrBlock.__call__ 48 edges     : %8.1f us" % timeit(lambda: win(coords)))
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            corr.CorrBlock(ofsMap, ofs_res, GA, f1n, f2n)
        torch.cuda.synchronize()
    rows = sorted(((e.self_device_time_total / 5, e.count // 5, e.key) for e in prof.key_averages() if e.self_device_time_total > 0), reverse=True)
    print("per CorrBlock.__init__ (8 edges): %.1f us of device time in %d launches" % (sum(r[0] for r in rows), sum(r[1] for r in rows)))
    for t, n, k in rows[:16]:
        print(f"{t:9.1f} us  {n:4d}  {k[:120]}")
