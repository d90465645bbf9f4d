"""Minimal single-op driver for ncu captures:  python tools/prof_one.py <op> [--edges E] [--lvl L] [--iters N]
ops: fwd | fwd1 | bwd | gauss | gauss_bwd | build | fused | fusedbwd | fusedbwd_acc | bbwd"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import inputs  # noqa: E402
import lgu_slam_b200  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("op")
    ap.add_argument("--edges", type=int, default=48)
    ap.add_argument("--lvl", type=int, default=0)
    ap.add_argument("--iters", type=int, default=3)
    a = ap.parse_args()
    ops = lgu_slam_b200.ops
    E, H, W, r = a.edges, 48, 64, 3
    H2, W2 = H >> a.lvl, W >> a.lvl
    dev = "cuda"
    g = torch.Generator(device=dev); g.manual_seed(1)
    vol = torch.randn(E, H, W, H2, W2, device=dev, generator=g)
    coords = inputs.make_coords(E, H, W, H2, W2, inputs.gen(2)).to(dev)
    off = inputs.make_offset(E, H, W, r, inputs.gen(7), zero=(a.lvl >= 2)).to(dev)
    grad = torch.randn(E, 7, 7, H, W, device=dev, generator=g)
    c = inputs.gaussian_case(1, H, W, H2, W2, 4, seed=3)
    means = c["means"].to(dev).expand(E, -1, -1, -1).contiguous()
    covs = c["covs"].to(dev).expand(E, -1, -1, -1).contiguous()
    for _ in range(a.iters):
        if a.op == "fwd":
            ops.defCorr_index_forward(vol, coords, off, r)
        elif a.op == "bwd":
            ops.defCorr_index_backward(vol, coords, off, grad, r)
        elif a.op == "gauss":
            ops.gaussianMask(means, covs, vol, 4)
        elif a.op == "gauss_bwd":
            ops.gaussianMask_backward(means, covs, vol, vol, 4)
        elif a.op == "build":
            if "bp" not in globals():
                global bp
                fc = inputs.frontend_case(E=E, T=20, seed=5, half_fmaps=True)
                hi, _ = ops.pack_fmaps(fc["fmaps"].half().to(dev))
                den = (6.28 * torch.sqrt(fc["covs"][..., 0] * fc["covs"][..., 1])).to(dev).contiguous()
                bp = (hi, fc["ii"].to(dev), fc["jj"].to(dev), fc["means"].to(dev), fc["covs"].to(dev), den)
            ops.build_pyramid(bp[0], None, bp[1], bp[2], H, W, means=bp[3], covs=bp[4], den=bp[5])
        elif a.op == "fused":
            if "fz" not in globals():
                global fz
                fc = inputs.frontend_case(E=E, T=20, seed=5, half_fmaps=True)
                pyr = [torch.randn(E, H, W, H >> l, W >> l, device=dev, generator=g) for l in range(4)]
                fz = (pyr, fc["coords"].to(dev), fc["offsets"][0].to(dev), fc["offsets"][1].to(dev))
            ops.corr_lookup_fused(fz[0], fz[1], fz[2], fz[3].clone(), 3)
        elif a.op == "fusedbwd":
            if "fzb" not in globals():
                global fzb
                fc = inputs.frontend_case(E=E, T=20, seed=5, half_fmaps=True)
                pyr = [torch.randn(E, H, W, H >> l, W >> l, device=dev, generator=g) for l in range(4)]
                o1 = fc["offsets"][1].to(dev)
                corr_, mask_ = ops.corr_lookup_fused(pyr, fc["coords"].to(dev), fc["offsets"][0].to(dev), o1, 3, return_mask=True)
                fzb = (pyr, fc["coords"].to(dev), fc["offsets"][0].to(dev), o1, mask_, torch.randn(E, 196, H, W, device=dev, generator=g))
            ops.corr_lookup_fused_backward(fzb[0], fzb[1], fzb[2], fzb[3], fzb[4], fzb[5])
        elif a.op == "fusedbwd_acc":
            if "fza" not in globals():
                global fza
                fc = inputs.frontend_case(E=E, T=20, seed=5, half_fmaps=True)
                pyr = [torch.randn(E, H, W, H >> l, W >> l, device=dev, generator=g) for l in range(4)]
                o1 = fc["offsets"][1].to(dev)
                corr_, mask_ = ops.corr_lookup_fused(pyr, fc["coords"].to(dev), fc["offsets"][0].to(dev), o1, 3, return_mask=True)
                fza = (pyr, fc["coords"].to(dev), fc["offsets"][0].to(dev), o1, mask_, torch.randn(E, 196, H, W, device=dev, generator=g),
                       [torch.zeros_like(p) for p in pyr])
            ops.corr_lookup_fused_backward(fza[0], fza[1], fza[2], fza[3], fza[4], fza[5], accumulate_into=fza[6])
        elif a.op == "bbwd":
            if "bbw" not in globals():
                global bbw
                bbw = ([torch.randn(E, H, W, H >> l, W >> l, device=dev, generator=g) for l in range(4)],
                       torch.randn(E, 128, H, W, device=dev, generator=g), torch.randn(E, 128, H, W, device=dev, generator=g))
            ops.build_backward_fmaps(bbw[0], bbw[1], bbw[2])
        elif a.op == "fwd1":
            ops.corr_index_forward(vol, coords, 1)
        else:
            raise SystemExit("unknown op")
    torch.cuda.synchronize()
    print("done", a.op)


if __name__ == "__main__":
    main()
