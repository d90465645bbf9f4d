#!/usr/bin/env python
"""Backend (global-BA) correlation throughput, BASELINE.json configs[3]/[4]: ~4k edges over 256 keyframes,
edge-sharded over the GPUs of one box.

One STEP = one BA iteration's correlation work (factor_graph.py:265-279): for every reference chunk (edges whose
source frame lies in a block of 8), AltCorrBlock.__call__ (offset convs + mask + 4-level deformable sampling).
Ranks start with T/N keyframes each, all-gather the fp16 feature maps once (outside the step loop, as the
reference builds its AltCorrBlock once per backend call, factor_graph.py:263), then every rank runs its chunks.
Outputs stay sharded (`--gather none`, the GRU update can run data-parallel on them) or are gathered on rank 0.

  python tools/bench_backend.py                                     # 1 GPU
  torchrun --nproc-per-node N tools/bench_backend.py [--gather dst] # N GPUs

Prints one JSON line per configuration (rank 0).  Timing: CUDA events, max over ranks."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import inputs  # noqa: E402
import lgu_slam_b200  # noqa: E402
from importlib import import_module  # noqa: E402

corr = import_module("lgu-slam_b200.corr")
sharded = import_module("lgu-slam_b200.sharded")


def backend_edges(T, E, g):
    """Proximity-style edge set: every frame linked to neighbours within +-8 (factor_graph.py:319-383 in spirit)."""
    ii = torch.randint(0, T, (E,), generator=g)
    jj = (ii + torch.randint(-8, 9, (E,), generator=g)).clamp(0, T - 1)
    jj = torch.where(jj == ii, (ii + 1).clamp(max=T - 1), jj)
    return ii, jj


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--edges", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--gather", default="none", choices=["none", "dst", "stream"])
    ap.add_argument("--lowmem-ops", action="store_true", help="reference op sequence (altcorr + 4 x lowMem_defSample)")
    ap.add_argument("--strict-ref", type=int, default=1)
    ap.add_argument("--cache", type=int, default=0, help="1: keep per-chunk offsets and volumes across BA steps")
    ap.add_argument("--volume-cache-gb", type=float, default=None, help="HBM budget of the volume cache (default: half of free)")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    H, W, C = 48, 64, 128
    g = inputs.gen(4242)
    T, E = a.frames, a.edges
    ii, jj = backend_edges(T, E, g)
    coords = inputs.make_coords(E, H, W, H, W, g).permute(0, 2, 3, 1).contiguous().view(1, E, H, W, 2)
    torch.manual_seed(0)
    ofsMap = nn.Conv2d(256, 98, 3, padding=1).to(dev); ofs_res = nn.Conv2d(256, 98, 3, padding=1).to(dev)
    GA = corr.GaussianMask(H, W).to(dev)
    # each rank owns a contiguous block of keyframes (as if it had encoded them) -> all-gather (collective 1)
    per = (T + world - 1) // world
    lo, hi = rank * per, min(T, (rank + 1) * per)
    gl = torch.Generator().manual_seed(7)
    all_maps = torch.randn(T, C, H, W, generator=gl).half()
    mine = all_maps[lo:hi].to(dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        sharded.all_gather_frames(mine)            # first collective: NCCL connection set-up, not steady state
    torch.cuda.synchronize()
    ev0.record()
    fmaps = sharded.all_gather_frames(mine) if world > 1 else mine
    ev1.record()
    torch.cuda.synchronize()
    gather_ms = ev0.elapsed_time(ev1)
    assert fmaps.shape[0] == T
    coords_d, ii_d, jj_d = coords.to(dev), ii.to(dev), jj.to(dev)
    with torch.no_grad():
        blk = corr.AltCorrBlock(ofsMap, ofs_res, GA, fmaps.view(1, T, C, H, W), strict_ref=bool(a.strict_ref),
                                materialize=not a.lowmem_ops, cache=bool(a.cache), volume_cache_gb=a.volume_cache_gb)
        eng = sharded.ShardedBackendCorr(lambda c, i, j: blk(c, i, j)) if world > 1 else None
        if eng is not None:
            plan = eng.set_edges(ii, jj)
        else:
            plan = sharded.partition_edges(ii, jj, 1)

        def step():
            if eng is not None:
                if a.gather == "stream":
                    return eng.lookup_streamed_to(coords_d, ii_d, jj_d, dst=0)
                return eng(coords_d, ii_d, jj_d, gather=None if a.gather == "none" else "dst")
            outs = [blk(coords_d[:, v.to(dev)], ii_d[v.to(dev)], jj_d[v.to(dev)]) for v in plan.chunk_edges]
            return torch.cat(outs, dim=1)

        for _ in range(a.warmup):
            out = step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(a.steps):
            out = step()
        ev1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / a.steps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    visited = plan.num_edges
    if rank == 0:
        print(json.dumps({"workload": "backend_altcorr", "frames": T, "edges": E, "edges_visited": visited,
                          "chunks": len(plan.chunk_edges), "n_gpus": world, "ms_per_step": ms,
                          "edges_per_s": visited / ms * 1e3, "gather": a.gather,
                          "path": "lowMem operator sequence" if a.lowmem_ops else "tcgen05 volumes + fused lookup",
                          "strict_ref": bool(a.strict_ref), "cache": bool(a.cache),
                          "volume_cache_GB": round(blk._vol_bytes / 2**30, 1), "fmap_allgather_ms": gather_ms,
                          "edges_per_rank": plan.counts()}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
