"""CUDA-event timing of single operators at the frontend size:  python tools/time_ops.py [--edges 48] [ops...]"""
import argparse, os, statistics, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import inputs, lgu_slam_b200
ops = lgu_slam_b200.ops

def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    return statistics.median(ts), min(ts)

def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--edges", type=int, default=48); ap.add_argument("which", nargs="*")
    a = ap.parse_args(); E, H, W = a.edges, 48, 64; dev = "cuda"; P = H * W
    c = inputs.frontend_case(E=E, T=20, seed=5, half_fmaps=True)
    fm = c["fmaps"].half().to(dev)
    ii, jj = c["ii"].to(dev), c["jj"].to(dev)
    means, covs = c["means"].to(dev), c["covs"].to(dev)
    den = (6.28 * torch.sqrt(covs[..., 0] * covs[..., 1])).contiguous()
    hi, _ = ops.pack_fmaps(fm)
    pyr = ops.build_pyramid(hi, None, ii, jj, H, W, means=means, covs=covs, den=den)
    coords = c["coords"].to(dev)
    off0, off1 = c["offsets"][0].to(dev), c["offsets"][1].to(dev)
    cc = coords.permute(0, 3, 1, 2).contiguous()
    cl = [(cc / 2 ** l).contiguous() for l in range(4)]
    offs = [off0.view(E, H, W, 7, 7, 2), off1.view(E, H, W, 7, 7, 2), torch.zeros(E, H, W, 7, 7, 2, device=dev), torch.zeros(E, H, W, 7, 7, 2, device=dev)]
    grad = torch.randn(E, 7, 7, H, W, device=dev)
    GB = lambda b, us: b / us / 1e3
    res = {}
    def rep(name, us, bytes_):
        print(f"{name:28s} median {us[0]:8.1f} us  min {us[1]:8.1f} us   alg {bytes_/1e6:8.1f} MB  -> {GB(bytes_, us[0]):7.1f} GB/s ({GB(bytes_, us[0])/6552*100:4.1f}% of 6552)")
    w = set(a.which)
    def on(k): return (not w and k != 'buildx') or k in w
    if on("build"):
        rep("build_pyramid", timeit(lambda: ops.build_pyramid(hi, None, ii, jj, H, W, means=means, covs=covs, den=den)), E * 51.757e6)
    if on("buildx"):
        for lv in (1, 2, 3, 4):
            for gr in (0, 4):
                rep(f"build levels={lv} gauss={gr}", timeit(lambda: ops.build_pyramid(hi, None, ii, jj, H, W, means=means, covs=covs, den=den, num_levels=lv, gauss_radius=gr)), E * (2 * P * 128 * 2 + 4 * P * sum(P >> (2 * l) for l in range(lv))))
    if on("fused"):
        o1 = off1.clone()
        rep("corr_lookup_fused", timeit(lambda: ops.corr_lookup_fused(pyr, coords, off0, o1, 3)), E * P * 3656)
    if on("fwd"):
        gath = [49 * 16, 49 * 16, 256, 256]
        for l in range(4):
            rep(f"defCorr_index_forward l{l}", timeit(lambda: ops.defCorr_index_forward(pyr[l], cl[l], offs[l], 3)), E * P * (8 + 392 + gath[l] + 196))
        rep("corr_index_forward r1 l1", timeit(lambda: ops.corr_index_forward(pyr[1], cl[1], 1)), E * P * (8 + 64 + 36))
    if on("bwd"):
        gath = [49 * 16, 49 * 16, 256, 256]
        for l in range(4):
            Q = (H >> l) * (W >> l)
            rep(f"defCorr_index_backward l{l}", timeit(lambda: ops.defCorr_index_backward(pyr[l], cl[l], offs[l], grad, 3)), E * P * (8 + 392 + 196 + 392 + gath[l] + 4 * Q))
        g1 = torch.randn(E, 3, 3, H, W, device=dev)
        rep("corr_index_backward r1 l1", timeit(lambda: ops.corr_index_backward(pyr[1], cl[1], g1, 1)), E * P * (8 + 36 + 4 * 768))
    if on("gauss"):
        vol = pyr[0]
        rep("gaussianMask", timeit(lambda: ops.gaussianMask(means, covs, vol, 4)), E * P * (81 * 4 + 4 * P + 16))
        rep("gaussianMask_backward", timeit(lambda: ops.gaussianMask_backward(means, covs, vol, vol, 4)), E * P * (2 * 81 * 4 + 32))
    if on("lowmem"):
        B = E
        lc = inputs.lowmem_case(B=B, N=1, H1=H, W1=W, H2=H, W2=W, C=128, r=3, seed=3)
        f1 = lc["fmap1"].to(dev); co = lc["coords"].to(dev); of = lc["offset"].to(dev)
        f2s = [lc["fmap2"].to(dev)]
        for l in range(1, 4):
            f2s.append(torch.nn.functional.avg_pool2d(f2s[-1].permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1).contiguous())
        for l in range(4):
            Ql = (H >> l) * (W >> l)
            cl_ = (co / 2 ** l).contiguous()
            rep(f"lowMem_defSample l{l}", timeit(lambda: ops.lowMem_defSample(f1, f2s[l], cl_, of, 3), iters=5), B * (4 * P * 128 + 4 * Ql * 128 + P * (8 + 392 + 196)))
        cl_ = (co / 2).contiguous()
        rep("altcorr_forward r1 l1", timeit(lambda: ops.altcorr_forward(f1, f2s[1], cl_, 1), iters=5), B * (4 * P * 128 + 4 * 768 * 128 + P * (8 + 36)))

def backend():
    """Backend chunk (AltCorrBlock.__call__): materialised tcgen05 path vs the lowMem operator sequence."""
    import torch.nn as nn
    from importlib import import_module
    corr = import_module("lgu-slam_b200.corr")
    dev = "cuda"
    torch.manual_seed(0)
    ofsMap = nn.Conv2d(256, 98, 3, padding=1).to(dev); ofs_res = nn.Conv2d(256, 98, 3, padding=1).to(dev)
    GA = corr.GaussianMask(48, 64).to(dev)
    T, E = 64, 48
    g = inputs.gen(3)
    fmaps = torch.randn(1, T, 128, 48, 64, generator=g).half().to(dev)
    ii, jj = inputs.edge_list(T, E, g)
    ii, jj = ii.long().to(dev), jj.long().to(dev)
    coords = inputs.make_coords(E, 48, 64, 48, 64, g).permute(0, 2, 3, 1).contiguous().view(1, E, 48, 64, 2).to(dev)
    with torch.no_grad():
        for mat in (True, False):
            for strict in (True, False):
                blk = corr.AltCorrBlock(ofsMap, ofs_res, GA, fmaps, strict_ref=strict, materialize=mat)
                us = timeit(lambda: blk(coords, ii, jj), iters=5, warm=2)
                print(f"AltCorrBlock.__call__ E={E} materialize={mat} strict_ref={strict}: median {us[0]:9.1f} us "
                      f"-> {E / us[0] * 1e6:9.0f} edges/s (incl. offset convs)")
        blk = corr.AltCorrBlock(ofsMap, ofs_res, GA, fmaps, strict_ref=False, materialize=True)
        planes = blk._level_planes()
        ii32, jj32 = ii.int(), jj.int()
        for l in range(4):
            us = timeit(lambda: ops.build_volume(planes[0][0], None, planes[l][0], None, ii32, jj32))
            Q = (48 >> l) * (64 >> l)
            rep(f"build_volume l{l}", us, E * (3072 * Q * 4 + 3072 * 256 + Q * 256))


def rep(name, us, bytes_):
    print(f"{name:28s} median {us[0]:8.1f} us  min {us[1]:8.1f} us   alg {bytes_/1e6:8.1f} MB  -> {bytes_ / us[0] / 1e3:7.1f} GB/s ({bytes_ / us[0] / 1e3 / 6552 * 100:4.1f}% of 6552)")


def reference_cuda():
    """The reference's own kernels (offersample_LGS/*.cu, src/altcorr_kernel.cu) recompiled unmodified for sm_100
    (oracle/_ref), timed on the same tensors as ours: the bar SURVEY section 2.1 sets."""
    from oracle import build_ref
    ref = build_ref.load_ref("defCorrSample_ref"); alt = build_ref.load_ref("altcorr_ref")
    if ref is None:
        print("oracle/_ref not built"); return
    E, H, W, dev = 48, 48, 64, "cuda"; P = H * W
    c = inputs.frontend_case(E=E, T=20, seed=5, half_fmaps=True)
    fm = c["fmaps"].half().to(dev)
    ii, jj = c["ii"].to(dev), c["jj"].to(dev)
    means, covs = c["means"].to(dev), c["covs"].to(dev)
    den = (6.28 * torch.sqrt(covs[..., 0] * covs[..., 1])).contiguous()
    hi, _ = ops.pack_fmaps(fm)
    pyr = ops.build_pyramid(hi, None, ii, jj, H, W, means=means, covs=covs, den=den)
    coords = c["coords"].to(dev)
    cc = coords.permute(0, 3, 1, 2).contiguous()
    cl = [(cc / 2 ** l).contiguous() for l in range(4)]
    off = [c["offsets"][0].to(dev).view(E, H, W, 7, 7, 2), c["offsets"][1].to(dev).view(E, H, W, 7, 7, 2),
           torch.zeros(E, H, W, 7, 7, 2, device=dev), torch.zeros(E, H, W, 7, 7, 2, device=dev)]
    grad = torch.randn(E, 7, 7, H, W, device=dev); g1 = torch.randn(E, 3, 3, H, W, device=dev)
    rows = []
    def both(name, f_ref, f_ours, iters=10):
        tr = timeit(f_ref, iters=iters)[0]; to = timeit(f_ours, iters=iters)[0]
        rows.append((name, tr, to)); print(f"{name:34s} reference {tr:9.1f} us   ours {to:9.1f} us   x{tr / to:5.2f}")
    both("corr_index_forward r1 l1", lambda: ref.corr_index_forward(pyr[1], cl[1], 1), lambda: ops.corr_index_forward(pyr[1], cl[1], 1))
    for l in range(4):
        both(f"defCorr_index_forward l{l}", lambda: ref.defCorr_index_forward(pyr[l], cl[l], off[l], 3), lambda: ops.defCorr_index_forward(pyr[l], cl[l], off[l], 3))
    both("corr_index_backward r1 l1", lambda: ref.corr_index_backward(pyr[1], cl[1], g1, 1), lambda: ops.corr_index_backward(pyr[1], cl[1], g1, 1))
    for l in range(4):
        both(f"defCorr_index_backward l{l}", lambda: ref.defCorr_index_backward(pyr[l], cl[l], off[l], grad, 3), lambda: ops.defCorr_index_backward(pyr[l], cl[l], off[l], grad, 3))
    both("gaussianMask r4", lambda: ref.gaussianMask(means, covs, pyr[0], 4), lambda: ops.gaussianMask(means, covs, pyr[0], 4))
    both("gaussianMask_backward r4", lambda: ref.gaussianMask_backward(means, covs, pyr[0], pyr[0], 4), lambda: ops.gaussianMask_backward(means, covs, pyr[0], pyr[0], 4))
    lc = inputs.lowmem_case(B=E, N=1, H1=H, W1=W, H2=H, W2=W, C=128, r=3, seed=3)
    f1 = lc["fmap1"].to(dev); co = lc["coords"].to(dev); of = lc["offset"].to(dev)
    f2s = [lc["fmap2"].to(dev)]
    for l in range(1, 4):
        f2s.append(torch.nn.functional.avg_pool2d(f2s[-1].permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1).contiguous())
    for l in range(4):
        cl_ = (co / 2 ** l).contiguous()
        both(f"lowMem_defSample l{l}", lambda: ref.lowMem_defSample(f1, f2s[l], cl_, of, 3), lambda: ops.lowMem_defSample(f1, f2s[l], cl_, of, 3), iters=3)
    if alt is not None:
        cl_ = (co / 2).contiguous()
        both("altcorr_forward r1 l1", lambda: alt.altcorr_forward(f1, f2s[1], cl_, 1), lambda: ops.altcorr_forward(f1, f2s[1], cl_, 1), iters=3)
    # composites: the reference's CorrBlock data paths (its own ops + the torch glue it uses) vs the fused launches
    def ref_call():
        m, = ref.corr_index_forward(pyr[1], cl[1], 1)
        mask = torch.sigmoid(torch.var(m.permute(0, 3, 4, 1, 2), dim=[3, 4])).view(E, H, W, 1)
        o = [off[0], (off[1].view(E, H, W, 98) * mask).view(E, H, W, 7, 7, 2), off[2], off[3]]
        return torch.cat([ref.defCorr_index_forward(pyr[l], cl[l], o[l], 3)[0].view(E, 49, H, W) for l in range(4)], 1)
    o1 = off[1].reshape(E, H, W, 98).clone()
    both("CorrBlock.__call__ data path", ref_call, lambda: ops.corr_lookup_fused(pyr, coords, off[0].view(E, H, W, 98), o1, 3))
    cum = torch.ones(E, H, W, device=dev)
    both("CorrBlock.__call__ data path (cumulative-mask form)", ref_call,
         lambda: ops.corr_lookup_fused(pyr, coords, off[0].view(E, H, W, 98), o1, 3, cum_mask=cum))
    def ref_bwd():
        for l in range(4):
            ref.defCorr_index_backward(pyr[l], cl[l], off[l], grad, 3)
        ref.corr_index_backward(pyr[1], cl[1], g1, 1)
    corr_, mask_ = ops.corr_lookup_fused(pyr, coords, off[0].view(E, H, W, 98), o1, 3, return_mask=True)
    gout = torch.randn(E, 196, H, W, device=dev)
    both("its backward (5 ops vs 1 launch)", ref_bwd, lambda: ops.corr_lookup_fused_backward(pyr, coords, off[0].view(E, H, W, 98), o1, mask_, gout))
    f = fm.float()
    def ref_build():
        f1_, f2_ = f[ii.long()].reshape(E, 128, P) / 4, f[jj.long()].reshape(E, 128, P) / 4
        v = torch.matmul(f1_.transpose(1, 2), f2_).view(E, H, W, H, W)
        v1, = ref.gaussianMask(means, covs, v, 4)
        v = v1 / den.view(E, H, W, 1, 1) + v
        cur = v.reshape(E * P, 1, H, W); out = []
        for i in range(4):
            out.append(cur); cur = torch.nn.functional.avg_pool2d(cur, 2, stride=2)
        return out
    both("CorrBlock.__init__ data path", ref_build, lambda: ops.build_pyramid(ops.pack_fmaps(fm)[0], None, ii, jj, H, W, means=means, covs=covs, den=den), iters=5)
    print("\n| operator (E = 48, 48x64x128, B200) | reference CUDA for sm_100, us | this repo, us | speed-up |\n|---|---:|---:|---:|")
    for n_, tr, to in rows:
        print(f"| {n_} | {tr:.1f} | {to:.1f} | {tr / to:.2f}x |")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "refcuda":
        reference_cuda()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "backend":
        backend()
        sys.exit(0)
    main()
