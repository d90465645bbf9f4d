"""CPU tests of the edge-sharded backend plumbing (lgu-slam_b200/sharded.py): the partitioner alone, and the
world_size-2 path over `gloo` with the CPU oracle standing in for the GPU sampler (tests may use the oracle; the
product path on a GPU box passes AltCorrBlock).  Sharded == single-process, bit for bit, in the original edge order."""
import importlib
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def _sharded():
    # import the module file directly: the package __init__ would load the CUDA library, which is irrelevant here
    import importlib.util
    spec = importlib.util.spec_from_file_location("lgu_sharded", os.path.join(ROOT, "lgu-slam_b200", "sharded.py"))
    m = importlib.util.module_from_spec(spec)
    sys.modules["lgu_sharded"] = m
    spec.loader.exec_module(m)
    return m


def _edges(T, E, seed):
    g = torch.Generator().manual_seed(seed)
    ii = torch.randint(0, T, (E,), generator=g)
    jj = (ii + torch.randint(-3, 4, (E,), generator=g)).clamp(0, T - 1)
    return ii, jj


def test_reference_chunking_rule():
    sh = _sharded()
    ii = torch.tensor([0, 9, 3, 17, 8, 7, 16, 2])
    jj = torch.tensor([1, 8, 4, 16, 9, 6, 17, 3])
    chunks = sh.reference_chunks(ii, jj)
    # factor_graph.py:272-279: i = 0, 8, 16 (range(0, jj.max()+1, 8)); v = (ii >= i) & (ii < i+8)
    assert [c.tolist() for c in chunks] == [[0, 2, 5, 7], [1, 4], [3, 6]]
    # edges whose source frame is beyond the last loop start are never visited by the reference loop either
    assert sh.reference_chunks(torch.tensor([20]), torch.tensor([3])) == []
    assert sh.reference_chunks(torch.zeros(0, dtype=torch.long), torch.zeros(0, dtype=torch.long)) == []


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_partition_covers_every_edge_once_and_balances(world):
    sh = _sharded()
    ii, jj = _edges(T=256, E=4096, seed=world)
    plan = sh.partition_edges(ii, jj, world)
    allpos = torch.cat(plan.rank_edges)
    visited = torch.cat(sh.reference_chunks(ii, jj))
    assert sorted(allpos.tolist()) == sorted(visited.tolist())
    assert len(set(allpos.tolist())) == allpos.numel()
    # chunks are never split
    for c, own in enumerate(plan.chunk_owner):
        assert set(plan.chunk_edges[c].tolist()) <= set(plan.rank_edges[own].tolist())
    counts = plan.counts()
    assert max(counts) - min(counts) <= max(e.numel() for e in plan.chunk_edges)     # LPT bound
    assert max(counts) <= 1.25 * (sum(counts) / world) + 1
    # deterministic: same plan from the same edge list
    plan2 = sh.partition_edges(ii.clone(), jj.clone(), world)
    assert plan2.chunk_owner == plan.chunk_owner


def test_partition_with_a_destination_rank_gives_it_the_lighter_share():
    """Outputs returned to one rank: from DST_HANDICAP_FROM ranks on, the destination starts the LPT assignment with a
    handicap (the other ranks' rows arrive in its HBM while its own kernels run), so it ends with no more edges than any
    other rank; every edge is still assigned exactly once, chunks are never split, and the plan is deterministic."""
    sh = _sharded()
    ii, jj = _edges(T=256, E=4096, seed=3)
    world = sh.DST_HANDICAP_FROM
    plain = sh.partition_edges(ii, jj, world)
    plan = sh.partition_edges(ii, jj, world, dst=0)
    assert sorted(torch.cat(plan.rank_edges).tolist()) == sorted(torch.cat(plain.rank_edges).tolist())
    for c, own in enumerate(plan.chunk_owner):
        assert set(plan.chunk_edges[c].tolist()) <= set(plan.rank_edges[own].tolist())
    counts = plan.counts()
    assert counts[0] == min(counts) and counts[0] <= plain.counts()[0]
    assert max(counts) <= 1.25 * (sum(counts) / world) + 1
    assert sh.partition_edges(ii.clone(), jj.clone(), world, dst=0).chunk_owner == plan.chunk_owner
    # below that rank count the destination is an ordinary rank
    assert sh.partition_edges(ii, jj, 2, dst=0).chunk_owner == sh.partition_edges(ii, jj, 2).chunk_owner


@pytest.mark.parametrize("n", [1, 5, 36, 37, 40, 48, 73, 100, 128, 154, 256, 300])
def test_ship_schedule_covers_the_chunk(n):
    """Pass sizes of a chunk whose rows are shipped pass by pass (sharded.ShardedBackendCorr._ship_schedule): they cover
    the chunk (the consumer repeats the last size and clips the last pass), a rank's last chunk ends in the tapering
    tail, and no pass is tiny."""
    sh = _sharded()
    S = sh.ShardedBackendCorr
    for last in (False, True):
        sizes = S._ship_schedule(n, last=last)
        assert all(x > 0 for x in sizes)
        covered, k = 0, 0
        while covered < n:                                        # how AltCorrBlock._corr_materialized walks the schedule
            covered += sizes[min(k, len(sizes) - 1)]
            k += 1
        if len(sizes) > 1:
            assert sum(sizes) == n
            assert all(x >= 6 for x in sizes)
        if last and n > sum(S.SHIP_TAIL) and n > S.SHIP_EDGES:
            assert tuple(sizes[-len(S.SHIP_TAIL):])[1:] == S.SHIP_TAIL[1:]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_compute(pyr_levels):
    """AltCorrBlock.__call__-shaped stand-in on the CPU oracle: zero offsets, 2 levels, r=1 (small and fast)."""
    from oracle import oracle as orc

    def compute(coords, ii, jj):
        e, H, W = coords.shape[1:4]
        outs = []
        for l, f in enumerate(pyr_levels):
            f1 = pyr_levels[0][ii].contiguous()
            f2 = f[jj].contiguous()
            c = (coords[0] / 2 ** l).reshape(e, 1, H, W, 2).contiguous()
            o, = orc.lowMem_defSample(f1, f2, c, torch.zeros(e, H, W, 3, 3, 2), 1)
            outs.append(o.view(e, 9, H, W))
        return torch.cat(outs, dim=1)[None]
    return compute


def _case(T=20, E=37, H=6, W=8, C=32, seed=5):
    g = torch.Generator().manual_seed(seed)
    fmaps = torch.randn(T, H, W, C, generator=g)
    ii, jj = _edges(T, E, seed)
    coords = torch.rand(1, E, H, W, 2, generator=g) * torch.tensor([W - 1.0, H - 1.0])
    return fmaps, ii, jj, coords


def _levels(fmaps):
    l1 = torch.nn.functional.avg_pool2d(fmaps.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1).contiguous()
    return [fmaps.contiguous(), l1]


def _worker(rank, world, port, gather, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sh = _sharded()
        fmaps, ii, jj, coords = _case()
        T = fmaps.shape[0]
        # each rank starts with its own contiguous block of keyframes (uneven on purpose)
        cut = [0, 7, T] if world == 2 else [0, T]
        full = sh.all_gather_frames(fmaps[cut[rank]:cut[rank + 1]].contiguous())
        assert torch.equal(full, fmaps)
        eng = sh.ShardedBackendCorr(_oracle_compute(_levels(full)))
        plan = eng.set_edges(ii, jj)
        if gather in ("peer", "peer_chunk"):
            peer = sh.PeerOutput(ii.numel(), (18, 6, 8), torch.float32, "cpu", dst=0,
                                 plan=plan if gather == "peer_chunk" else None)
            out = eng.lookup_into_peer(coords, ii, jj, peer)
            if out is not None and gather == "peer_chunk":          # chunk-major rows -> original edge order
                assert sorted(peer.row_of_edge.tolist()) == list(range(ii.numel()))
                out = out[:, peer.row_of_edge]
            out = out.clone() if out is not None else None
            peer.close()
        elif gather == "stream":
            out = eng.lookup_streamed_to(coords, ii, jj, dst=0)
        else:
            out = eng(coords, ii, jj, gather=gather)
        if gather is None:
            local, pos = out
            q.put((rank, "local", local, pos))
        elif out is not None:
            q.put((rank, "full", out, None))
        else:
            q.put((rank, "none", None, None))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("gather", ["all", "dst", "stream", "peer", "peer_chunk", None])
def test_world_size_2_gloo_equals_single_process(gather):
    fmaps, ii, jj, coords = _case()
    sh = _sharded()
    want_plan = sh.partition_edges(ii, jj, 1)
    compute = _oracle_compute(_levels(fmaps))
    visited = want_plan.rank_edges[0]
    want = torch.zeros(1, visited.numel(), 18, 6, 8)
    # single process, chunk by chunk in reference order
    ref = torch.cat([compute(coords[:, v], ii[v], jj[v]) for v in want_plan.chunk_edges], dim=1)
    want[:, visited] = ref if visited.numel() == ii.numel() else want
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, gather, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got.sort(key=lambda t: t[0])
    full = torch.empty_like(ref)
    full[:, visited] = ref                      # expected output in ORIGINAL edge order
    if gather == "all":
        for _, kind, out, _ in got:
            assert kind == "full" and torch.equal(out, full)
    elif gather in ("dst", "stream", "peer", "peer_chunk"):
        assert got[0][1] == "full" and torch.equal(got[0][2], full) and got[1][1] == "none"
    else:
        seen = torch.zeros(ii.numel(), dtype=torch.bool)
        for _, kind, local, pos in got:
            assert kind == "local"
            assert torch.equal(local, full[:, pos])
            seen[pos] = True
        assert seen.all()


def test_edges_the_reference_loop_never_visits_stay_zero_in_the_gathered_output():
    """factor_graph.py:272-279 stops at ceil((jj.max()+1)/8)*8: an edge whose source frame lies beyond that is skipped by
    the reference loop.  The gathered output keeps the FULL edge count (such rows are zero), in every single-process
    gather mode, and the streamed / peer variants work without a process group (world = 1)."""
    sh = _sharded()
    ii = torch.tensor([0, 1, 9, 2])
    jj = torch.tensor([1, 0, 3, 3])                       # jj.max() = 3 -> one loop start (0): edge 2 (ii = 9) is skipped
    H, W = 6, 8
    coords = torch.rand(1, 4, H, W, 2)

    def compute(c, a, b):
        return (a.float() * 10 + b.float()).view(1, -1, 1, 1, 1).expand(1, a.numel(), 3, H, W).contiguous()

    eng = sh.ShardedBackendCorr(compute)
    plan = eng.set_edges(ii, jj)
    assert plan.total_edges == 4 and plan.num_edges == 3
    want = compute(coords, ii, jj).clone()
    want[:, 2] = 0
    assert torch.equal(eng(coords, ii, jj, gather="all"), want)
    assert torch.equal(eng.lookup_streamed_to(coords, ii, jj), want)
    peer = sh.PeerOutput(4, (3, H, W), torch.float32, "cpu")
    assert torch.equal(eng.lookup_into_peer(coords, ii, jj, peer), want)
    peer.close()
    local, pos = eng(coords, ii, jj, gather=None)
    assert pos.tolist() == [0, 1, 3] and torch.equal(local, want[:, pos])
