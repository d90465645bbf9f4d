"""2-GPU check of sharded.PeerOutput: every rank's AltCorrBlock stores its edges into rank 0's buffer through NVLink
peer memory; rank 0 compares with its own single-process result.  torchrun --nproc-per-node 2 tools/diag/peer_check.py"""
import os, sys, torch, torch.distributed as dist, torch.nn as nn
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import inputs, lgu_slam_b200
from importlib import import_module
corr = import_module("lgu-slam_b200.corr"); sh = import_module("lgu-slam_b200.sharded")
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
H, W, C, T, E = 48, 64, 128, 24, 60
g = inputs.gen(5)
fm = torch.randn(T, C, H, W, generator=g).half()
ii = torch.randint(0, T, (E,), generator=g); jj = (ii + torch.randint(1, 4, (E,), generator=g)).clamp(max=T - 1)
coords = inputs.make_coords(E, H, W, H, W, g).permute(0, 2, 3, 1).contiguous().view(1, E, H, W, 2).to(dev)
torch.manual_seed(0)
ofs, ofr = nn.Conv2d(256, 98, 3, padding=1).to(dev), nn.Conv2d(256, 98, 3, padding=1).to(dev)
GA = corr.GaussianMask(H, W).to(dev)
per = T // world
full = sh.all_gather_frames(fm[rank * per:(rank + 1) * per].to(dev))
assert torch.equal(full.cpu(), fm)
with torch.no_grad():
    blk = corr.AltCorrBlock(ofs, ofr, GA, full.view(1, T, C, H, W), materialize=True)
    def compute(c, i, j, out=None, out_index=None, pass_edges=None, pass_hook=None):
        return blk(c, i, j, out=out, out_index=out_index, pass_edges=pass_edges, pass_hook=pass_hook)
    eng = sh.ShardedBackendCorr(compute); plan = eng.set_edges(ii, jj)
    for dtype, via, chunked in ((torch.float32, "store", False), (torch.float16, "store", False), (torch.float16, "copy", True),
                                (torch.float16, "copy", False)):
        eng.SHIP_EDGES = 8
        peer = sh.PeerOutput(E, (196, H, W), dtype, dev, dst=0, plan=plan if chunked else None)
        res = eng.lookup_into_peer(coords, ii.to(dev), jj.to(dev), peer, via=via)
        if rank == 0 and chunked:
            res = res[:, peer.row_of_edge.to(dev)]
        if rank == 0:
            one = sh.ShardedBackendCorr(compute, single_process=True); one.set_edges(ii, jj)
            own = sh.PeerOutput(E, (196, H, W), dtype, dev, single_process=True)
            want = one.lookup_into_peer(coords, ii.to(dev), jj.to(dev), own)
            print("peer", dtype, via, chunked, "counts", plan.counts(), "equal:", torch.equal(res, want), "nonzero frac", (res != 0).float().mean().item(), flush=True)
            assert torch.equal(res, want)
            own.close()
        peer.close()
dist.barrier(); dist.destroy_process_group()
if rank == 0: print("peer_check ok")
