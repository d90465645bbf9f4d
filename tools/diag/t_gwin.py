"""Fused lookup backward with / without the Gaussian window record, and the Gaussian-head backward from the four level
gradients vs from the record (E = 48)."""
import os, sys, statistics, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import inputs, lgu_slam_b200
ops = lgu_slam_b200.ops
E, H, W = 48, 48, 64; dev = "cuda"
g = torch.Generator(device=dev); g.manual_seed(1)
fc = inputs.frontend_case(E=E, T=20, seed=5, half_fmaps=True)
pyr = [torch.randn(E, H, W, H >> l, W >> l, device=dev, generator=g) for l in range(4)]
o1 = fc["offsets"][1].to(dev); o0 = fc["offsets"][0].to(dev); co = fc["coords"].to(dev)
means, covs = fc["means"].to(dev).contiguous(), fc["covs"].to(dev).contiguous()
den = (6.28 * torch.sqrt(covs[..., 0] * covs[..., 1])).contiguous()
cum = torch.ones(E, H, W, device=dev)
_, mask = ops.corr_lookup_fused(pyr, co, o0, o1, 3, return_mask=True, cum_mask=cum)
gc = torch.randn(E, 196, H, W, device=dev, generator=g)

def t(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    return round(statistics.median(ts), 1)

print("bwd dense               us", t(lambda: ops.corr_lookup_fused_backward(pyr, co, o0, o1, mask, gc, cum_mask=cum)))
print("bwd dense + window      us", t(lambda: ops.corr_lookup_fused_backward(pyr, co, o0, o1, mask, gc, cum_mask=cum, gauss_window_means=means)))
print("bwd dense + gauss head  us", t(lambda: ops.corr_lookup_fused_backward(pyr, co, o0, o1, mask, gc, cum_mask=cum, gauss_head=(means, covs, den))))
out = ops.corr_lookup_fused_backward(pyr, co, o0, o1, mask, gc, cum_mask=cum, gauss_window_means=means)
print("gauss from levels       us", t(lambda: ops.build_backward_gauss(means, covs, den, pyr[0], list(out[:4]), 4)))
print("gauss from window       us", t(lambda: ops.build_backward_gauss(means, covs, den, pyr[0], None, 4, window=out[6])))
a = ops.build_backward_gauss(means, covs, den, pyr[0], list(out[:4]), 4)
b = ops.build_backward_gauss(means, covs, den, pyr[0], None, 4, window=out[6])
print("identical:", [bool(torch.equal(u, v)) for u, v in zip(a, b)])
f = ops.corr_lookup_fused_backward(pyr, co, o0, o1, mask, gc, cum_mask=cum, gauss_head=(means, covs, den))
print("fused identical:", [bool(torch.equal(u, v)) for u, v in zip(a, f[6:])])
