import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import inputs, lgu_slam_b200
from oracle import oracle
ops = lgu_slam_b200.ops
c = inputs.frontend_case(E=1, T=2, seed=34, half_fmaps=True)
dev = "cuda"
hi, lo = ops.pack_fmaps(c["fmaps"].to(dev))
E, H, W = c["means"].shape[:3]
got = ops.build_pyramid(hi, None, c["ii"].to(dev), c["jj"].to(dev), H, W, num_levels=2, gauss_radius=0, precision=1, round_half=True)
raw = ops.build_pyramid(hi, None, c["ii"].to(dev), c["jj"].to(dev), H, W, num_levels=2, gauss_radius=0, precision=1, round_half=False)
f1 = c["fmaps"][c["ii"].long()].contiguous(); f2 = c["fmaps"][c["jj"].long()].contiguous()
want = oracle.corr_volume(f1, f2, True)
wraw = oracle.corr_volume(f1, f2, False)
g = got[0].cpu(); d = (g - want).abs()
ulp = torch.maximum(want.abs(), torch.tensor(2.0 ** -14)) * 2.0 ** -10
bad = (d > ulp).nonzero()
print("n bad", bad.shape[0], "of", d.numel(), "n diff", (d > 0).sum().item())
for b in bad[:10]:
    b = tuple(b.tolist())
    print(b, "got", g[b].item(), "want", want[b].item(), "raw gpu", raw[0].cpu()[b].item(), "raw cpu", wraw[b].item())
print("max raw diff", (raw[0].cpu() - wraw).abs().max().item())
