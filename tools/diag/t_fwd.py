import os, sys, statistics, torch
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests")]
import inputs, lgu_slam_b200
ops = lgu_slam_b200.ops
E, H, W = 48, 48, 64; dev = "cuda"
g = torch.Generator(device=dev); g.manual_seed(1)
fc = inputs.frontend_case(E=E, T=20, seed=5, half_fmaps=True)
pyr = [torch.randn(E, H, W, H >> l, W >> l, device=dev, generator=g) for l in range(4)]
o1 = fc["offsets"][1].to(dev); o0 = fc["offsets"][0].to(dev); co = fc["coords"].to(dev)
def f(): ops.corr_lookup_fused(pyr, co, o0, o1.clone(), 3)
for _ in range(3): f()
torch.cuda.synchronize(); ts = []
for _ in range(20):
    o = o1.clone()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ops.corr_lookup_fused(pyr, co, o0, o, 3); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
print(os.environ.get("LGU_CORR_LIB", "default"), "fused lookup forward median us", statistics.median(ts))
