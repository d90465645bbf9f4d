#!/usr/bin/env python
"""BASELINE.json configs[4]: edge-count sweep of the backend correlation (AltCorrBlock.__call__ over reference chunks)
on one GPU -- this repo's tensor-core path, this repo's drop-in lowMem operators, and the reference's own CUDA kernels
recompiled for sm_100 (oracle/_ref) inside the same Python glue.  Multi-GPU points: tools/bench_backend.py.
  python tools/bench_sweep.py [--max-edges 16384] [--ref-max 1024]"""
import argparse, json, os, sys, types
import torch, torch.nn as nn
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import inputs, lgu_slam_b200  # noqa: E402
from importlib import import_module  # noqa: E402
corr = import_module("lgu-slam_b200.corr"); sharded = import_module("lgu-slam_b200.sharded")
sys.path.insert(0, os.path.join(ROOT, "tools"))
from bench_backend import backend_edges  # noqa: E402


def run(blk, coords, ii, jj, plan, dev, steps=1):
    def step():
        return [blk(coords[:, v.to(dev)], ii[v.to(dev)], jj[v.to(dev)]) for v in plan.chunk_edges]
    step(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = step()
        del out
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-edges", type=int, default=16384)
    ap.add_argument("--ref-max", type=int, default=1024)
    a = ap.parse_args()
    dev = torch.device("cuda")
    H, W, C = 48, 64, 128
    torch.manual_seed(0)
    ofsMap = nn.Conv2d(256, 98, 3, padding=1).to(dev); ofs_res = nn.Conv2d(256, 98, 3, padding=1).to(dev)
    GA = corr.GaussianMask(H, W).to(dev)
    refops = None
    try:
        from oracle import build_ref
        r, al = build_ref.load_ref("defCorrSample_ref"), build_ref.load_ref("altcorr_ref")
        if r is not None and al is not None:
            refops = types.SimpleNamespace(lowMem_defSample=r.lowMem_defSample, altcorr_forward=al.altcorr_forward)
    except Exception:
        pass
    E = 64
    while E <= a.max_edges:
        T = max(16, E // 16)
        g = inputs.gen(100 + E)
        ii, jj = backend_edges(T, E, g)
        coords = inputs.make_coords(E, H, W, H, W, g).permute(0, 2, 3, 1).contiguous().view(1, E, H, W, 2).to(dev)
        fmaps = torch.randn(1, T, C, H, W, generator=g).half().to(dev)
        plan = sharded.partition_edges(ii, jj, 1)
        ii_d, jj_d = ii.to(dev), jj.to(dev)
        row = {"edges": E, "frames": T, "chunks": len(plan.chunk_edges), "visited": plan.num_edges}
        with torch.no_grad():
            blk = corr.AltCorrBlock(ofsMap, ofs_res, GA, fmaps, materialize=True)
            ms = run(blk, coords, ii_d, jj_d, plan, dev, steps=2)
            row["tcgen05_volumes_fused_lookup"] = {"ms": ms, "edges_per_s": plan.num_edges / ms * 1e3}
            if E <= 4096:
                blk = corr.AltCorrBlock(ofsMap, ofs_res, GA, fmaps, materialize=False)
                ms = run(blk, coords, ii_d, jj_d, plan, dev)
                row["lowmem_operators"] = {"ms": ms, "edges_per_s": plan.num_edges / ms * 1e3}
            if refops is not None and E <= a.ref_max:
                blk = corr.AltCorrBlock(ofsMap, ofs_res, GA, fmaps, materialize=False, sampler_ops=refops)
                ms = run(blk, coords, ii_d, jj_d, plan, dev)
                row["reference_cuda_sm100"] = {"ms": ms, "edges_per_s": plan.num_edges / ms * 1e3}
        print(json.dumps(row), flush=True)
        E *= 4


if __name__ == "__main__":
    main()
