import os, sys, torch, torch.nn as nn
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import inputs, lgu_slam_b200
from importlib import import_module
corr = import_module("lgu-slam_b200.corr")
ops = lgu_slam_b200.ops
dev = "cuda"
torch.manual_seed(4)
ofsMap = nn.Conv2d(256, 98, 3, padding=1).to(dev); ofs_res = nn.Conv2d(256, 98, 3, padding=1).to(dev)
GA = corr.GaussianMask(48, 64).to(dev)
g = inputs.gen(11)
T, E = 6, 7
fmaps = torch.randn(1, T, 128, 48, 64, generator=g).to(dev)
ii = torch.tensor([0, 1, 2, 3, 4, 5, 2], device=dev); jj = torch.tensor([1, 2, 3, 4, 5, 4, 0], device=dev)
coords = inputs.make_coords(E, 48, 64, 48, 64, g).permute(0, 2, 3, 1).contiguous().view(1, E, 48, 64, 2).to(dev)
torch.backends.cudnn.allow_tf32 = False
with torch.no_grad():
    for strict in (True, False):
        a = corr.AltCorrBlock(ofsMap, ofs_res, GA, fmaps, strict_ref=strict, materialize=True)
        b = corr.AltCorrBlock(ofsMap, ofs_res, GA, fmaps, strict_ref=strict, materialize=False)
        oa = a(coords, ii, jj); ob = b(coords, ii, jj)
        d = (a.offset[1] - b.offset[1]).reshape(E, 48, 64, 49, 2)
        print("strict", strict, "out err", (oa - ob).abs().max().item())
        bad = (d.abs() > 1e-5).nonzero()
        print("n bad", bad.shape[0], "first", bad[:5].tolist())
        print("per-edge max", d.abs().amax(dim=(1, 2, 3, 4)).tolist())
        print("per-tap max", d.abs().amax(dim=(0, 1, 2, 4)).tolist()[:49])
