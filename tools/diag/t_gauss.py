import os, sys, statistics, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import inputs, lgu_slam_b200
ops = lgu_slam_b200.ops
E, H, W = 48, 48, 64; dev = "cuda"
g = torch.Generator(device=dev); g.manual_seed(1)
vol = torch.randn(E, H, W, H, W, device=dev, generator=g)
c = inputs.gaussian_case(1, H, W, H, W, 4, seed=3)
means = c["means"].to(dev).expand(E, -1, -1, -1).contiguous(); covs = c["covs"].to(dev).expand(E, -1, -1, -1).contiguous()
def t(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    return statistics.median(ts)
print("gaussianMask forward  E=48: %.1f us" % t(lambda: ops.gaussianMask(means, covs, vol, 4)))
print("gaussianMask backward E=48: %.1f us" % t(lambda: ops.gaussianMask_backward(means, covs, vol, vol, 4)))
