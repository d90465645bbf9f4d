"""Does cudaLimitMaxL2FetchGranularity (32 / 64 / 128 B) change the gather kernels?  Times the fused lookup forward, the fused
backward and both Gaussian backward legs at E = 48 under each setting (the limit is a per-process device hint)."""
import ctypes, os, sys, statistics, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import inputs, lgu_slam_b200
ops = lgu_slam_b200.ops
rt = ctypes.CDLL("libcudart.so.12")
E, H, W = 48, 48, 64; dev = "cuda"
g = torch.Generator(device=dev); g.manual_seed(1)
fc = inputs.frontend_case(E=E, T=20, seed=5, half_fmaps=True)
pyr = [torch.randn(E, H, W, H >> l, W >> l, device=dev, generator=g) for l in range(4)]
o1 = fc["offsets"][1].to(dev); o0 = fc["offsets"][0].to(dev); co = fc["coords"].to(dev)
means, covs = fc["means"].to(dev), fc["covs"].to(dev)
den = (6.28 * torch.sqrt(covs[..., 0] * covs[..., 1])).contiguous()
gc = torch.randn(E, 196, H, W, device=dev, generator=g)
cum = torch.ones(E, H, W, device=dev)
_, mask = ops.corr_lookup_fused(pyr, co, o0, o1, 3, return_mask=True, cum_mask=cum)
hi, _ = ops.pack_fmaps(fc["fmaps"].half().to(dev)); ii, jj = fc["ii"].to(dev), fc["jj"].to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    return round(statistics.median(ts), 1)
for gran in (64, 32, 128, 64):
    rc = rt.cudaDeviceSetLimit(ctypes.c_int(5), ctypes.c_size_t(gran))
    val = ctypes.c_size_t(0); rt.cudaDeviceGetLimit(ctypes.byref(val), ctypes.c_int(5))
    grads = ops.corr_lookup_fused_backward(pyr, co, o0, o1, mask, gc, cum_mask=cum)
    print(f"L2 fetch granularity {gran} (rc {rc}, now {val.value}):",
          "lookup_fwd", t(lambda: ops.corr_lookup_fused(pyr, co, o0, o1, 3, cum_mask=cum)),
          "| lookup_bwd", t(lambda: ops.corr_lookup_fused_backward(pyr, co, o0, o1, mask, gc, cum_mask=cum)),
          "| gauss_bwd(fused)", t(lambda: ops.build_backward_gauss(means, covs, den, pyr[0], list(grads[:4]), 4)),
          "| gauss_bwd(dropin)", t(lambda: ops.gaussianMask_backward(means, covs, pyr[0], grads[0], 4)),
          "| build", t(lambda: ops.build_pyramid(hi, None, ii, jj, H, W, means=means, covs=covs, den=den, gauss_radius=4)), flush=True)
