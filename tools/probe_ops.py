"""Per-op timing probe (development tool, not the bench): times each drop-in op at the BASELINE frontend
shape next to the reference CUDA extension (oracle/_ref) on the same tensors, CUDA events, median.

    gpurun -- 'python tools/probe_ops.py --edges 48 > gpurun_out/probe.txt'
"""
import argparse
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import inputs  # noqa: E402
import lgu_slam_b200  # noqa: E402

HBM = 6552.0  # GB/s measured (MEASURED_PEAKS.json)


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--edges", type=int, default=48)
    ap.add_argument("--no-ref", action="store_true")
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    ops = lgu_slam_b200.ops
    ref = alt = None
    if not a.no_ref:
        from oracle import build_ref
        ref, alt = build_ref.load_ref("defCorrSample_ref"), build_ref.load_ref("altcorr_ref")
    E, H, W, r = a.edges, 48, 64, 3
    P = H * W
    dev = "cuda"
    g = torch.Generator(device=dev); g.manual_seed(1)
    rows = []

    def report(name, ours_ms, ref_ms, bytes_per_edge):
        gbs = bytes_per_edge * E / (ours_ms * 1e-3) / 1e9
        rows.append((name, ours_ms, ref_ms, gbs, gbs / HBM))
        print(f"{name:34s} ours {ours_ms:8.3f} ms   ref {ref_ms if ref_ms else float('nan'):8.3f} ms   "
              f"{gbs:7.0f} GB/s algorithmic = {100 * gbs / HBM:5.1f}% of measured HBM peak", flush=True)

    for lvl in range(4):
        if a.only and f"lvl{lvl}" not in a.only and "lookup" not in a.only:
            continue
        H2, W2 = H >> lvl, W >> lvl
        Q = H2 * W2
        vol = torch.randn(E, H, W, H2, W2, device=dev, generator=g)
        coords = inputs.make_coords(E, H, W, H2, W2, inputs.gen(2 + lvl)).to(dev)
        off = inputs.make_offset(E, H, W, r, inputs.gen(7 + lvl), zero=(lvl >= 2)).to(dev)
        grad = torch.randn(E, 7, 7, H, W, device=dev, generator=g)
        gather = 49 * 16 if lvl < 2 else 64 * 4
        fwd_bytes = P * (8 + 392 + gather + 196)
        t = timeit(lambda: ops.defCorr_index_forward(vol, coords, off, r))
        tr = timeit(lambda: ref.defCorr_index_forward(vol, coords, off, r)) if ref else None
        report(f"defCorr_index_forward  lvl{lvl}", t, tr, fwd_bytes)
        bwd_bytes = P * (8 + 392 + 196 + 392 + gather + 4 * Q)
        t = timeit(lambda: ops.defCorr_index_backward(vol, coords, off, grad, r))
        tr = timeit(lambda: ref.defCorr_index_backward(vol, coords, off, grad, r)) if ref else None
        report(f"defCorr_index_backward lvl{lvl}", t, tr, bwd_bytes)
        if lvl == 1:
            c2 = coords
            g1 = torch.randn(E, 3, 3, H, W, device=dev, generator=g)
            t = timeit(lambda: ops.corr_index_forward(vol, c2, 1))
            tr = timeit(lambda: ref.corr_index_forward(vol, c2, 1)) if ref else None
            report("corr_index_forward r=1 lvl1", t, tr, P * (8 + 64 + 36))
            t = timeit(lambda: ops.corr_index_backward(vol, c2, g1, 1))
            tr = timeit(lambda: ref.corr_index_backward(vol, c2, g1, 1)) if ref else None
            report("corr_index_backward r=1 lvl1", t, tr, P * (8 + 36 + 4 * Q))
        if lvl == 0 and (not a.only or "gauss" in a.only or "lvl0" in a.only):
            c = inputs.gaussian_case(1, H, W, H, W, 4, seed=3)
            means = c["means"].to(dev).expand(E, -1, -1, -1).contiguous()
            covs = c["covs"].to(dev).expand(E, -1, -1, -1).contiguous()
            t = timeit(lambda: ops.gaussianMask(means, covs, vol, 4))
            tr = timeit(lambda: ref.gaussianMask(means, covs, vol, 4)) if ref else None
            report("gaussianMask r=4", t, tr, P * (81 * 4 + 4 * P + 16))
            gout = torch.randn(E, H, W, H, W, device=dev, generator=g)
            t = timeit(lambda: ops.gaussianMask_backward(means, covs, vol, gout, 4))
            tr = timeit(lambda: ref.gaussianMask_backward(means, covs, vol, gout, 4)) if ref else None
            report("gaussianMask_backward r=4", t, tr, P * (2 * 81 * 4 + 32))
            del gout
        del vol, coords, off, grad
        torch.cuda.empty_cache()

    if not a.only or "lowmem" in a.only:
        B = min(E, 64)
        for lvl in range(4):
            c = inputs.lowmem_case(B, 1, H, W, H >> lvl, W >> lvl, 128, 3, seed=20 + lvl)
            f1, f2, coords, off = (c[k].to(dev) for k in ("fmap1", "fmap2", "coords", "offset"))
            t = timeit(lambda: ops.lowMem_defSample(f1, f2, coords, off, 3), iters=5, warm=1)
            tr = timeit(lambda: ref.lowMem_defSample(f1, f2, coords, off, 3), iters=5, warm=1) if ref else None
            Ql = (H >> lvl) * (W >> lvl)
            by = 4 * P * 128 + 4 * Ql * 128 + 8 * P + 392 * P + 196 * P
            gbs = by * B / (t * 1e-3) / 1e9
            print(f"lowMem_defSample lvl{lvl} (B={B})        ours {t:8.3f} ms   ref {tr if tr else float('nan'):8.3f} ms   "
                  f"{gbs:7.0f} GB/s algorithmic = {100 * gbs / HBM:5.1f}%", flush=True)
            if lvl == 1 and alt is not None:
                t = timeit(lambda: ops.altcorr_forward(f1, f2, coords, 1), iters=5, warm=1)
                tr = timeit(lambda: alt.altcorr_forward(f1, f2, coords, 1), iters=5, warm=1)
                print(f"altcorr_forward r=1 lvl1 (B={B})       ours {t:8.3f} ms   ref {tr:8.3f} ms", flush=True)


if __name__ == "__main__":
    main()
