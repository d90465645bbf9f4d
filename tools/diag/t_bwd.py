import os, sys, statistics, torch
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests")]
import inputs, lgu_slam_b200
ops = lgu_slam_b200.ops
E,H,W=48,48,64; dev="cuda"
g = torch.Generator(device=dev); g.manual_seed(1)
fc = inputs.frontend_case(E=E, T=20, seed=5, half_fmaps=True)
pyr = [torch.randn(E, H, W, H >> l, W >> l, device=dev, generator=g) for l in range(4)]
o1 = fc["offsets"][1].to(dev); o0 = fc["offsets"][0].to(dev); co = fc["coords"].to(dev)
corr_, mask_ = ops.corr_lookup_fused(pyr, co, o0, o1, 3, return_mask=True)
gout = torch.randn(E, 196, H, W, device=dev, generator=g)
acc = [torch.zeros_like(p) for p in pyr] if os.environ.get("LGU_BWD_ACC") else None
def f(): ops.corr_lookup_fused_backward(pyr, co, o0, o1, mask_, gout, accumulate_into=acc)
for _ in range(3): f()
torch.cuda.synchronize(); ts=[]
for _ in range(10):
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b)*1e3)
print("fused backward", "accumulate" if acc else "dense", "median us", statistics.median(ts))
