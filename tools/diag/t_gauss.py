"""Time the Gaussian backward legs at E = 48: the drop-in gaussianMask_backward and lgu_build_backward_gauss."""
import os, sys, statistics, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import inputs, lgu_slam_b200
ops = lgu_slam_b200.ops
E, H, W = 48, 48, 64; dev = "cuda"
g = torch.Generator(device=dev); g.manual_seed(1)
fc = inputs.frontend_case(E=E, T=20, seed=5, half_fmaps=True)
means, covs = fc["means"].to(dev), fc["covs"].to(dev)
den = (6.28 * torch.sqrt(covs[..., 0] * covs[..., 1])).contiguous()
vol = torch.randn(E, H, W, H, W, device=dev, generator=g)
grads = [torch.randn(E, H, W, H >> l, W >> l, device=dev, generator=g) for l in range(4)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t(fn, name):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(20):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    print(os.path.basename(os.environ.get("LGU_CORR_LIB", "default")), name, "median us", round(statistics.median(ts), 1), "min", round(min(ts), 1))
t(lambda: ops.gaussianMask_backward(means, covs, vol, grads[0], 4), "gaussianMask_backward")
t(lambda: ops.build_backward_gauss(means, covs, den, vol, grads, 4), "build_backward_gauss")
