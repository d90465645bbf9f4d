"""Time the fused lookup backward at E = 48 (dense contract), cumulative-mask form."""
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
cum = torch.ones(E, H, W, device=dev)
_, mask = ops.corr_lookup_fused(pyr, co, o0, o1, 3, return_mask=True, cum_mask=cum)
gc = torch.randn(E, 196, H, W, device=dev, generator=g)
acc = [torch.zeros_like(p) for p in pyr]
for name, kw in (("dense", {}), ("accumulate", {"accumulate_into": acc})):
    f = lambda: ops.corr_lookup_fused_backward(pyr, co, o0, o1, mask, gc, cum_mask=cum, **kw)
    for _ in range(3): f()
    torch.cuda.synchronize(); ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    print(os.path.basename(os.environ.get("LGU_CORR_LIB", "default")), name, "fused lookup backward median us", round(statistics.median(ts), 1))
