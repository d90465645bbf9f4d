"""Time the fused build (pack excluded) at E = 48, fp16 maps, Gaussian on."""
import os, sys, statistics, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import inputs, lgu_slam_b200
ops = lgu_slam_b200.ops
E, H, W = 48, 48, 64; dev = "cuda"
fc = inputs.frontend_case(E=E, T=20, seed=5, half_fmaps=True)
hi, _ = ops.pack_fmaps(fc["fmaps"].half().to(dev))
ii, jj, means, covs = (fc[k].to(dev) for k in ("ii", "jj", "means", "covs"))
f = lambda: ops.build_pyramid(hi, None, ii, jj, H, W, means=means, covs=covs, den=None, gauss_radius=4)
for _ in range(3): f()
torch.cuda.synchronize(); ts = []
for _ in range(10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
print(os.path.basename(os.environ.get("LGU_CORR_LIB", "default")), "build median us", round(statistics.median(ts), 1))
