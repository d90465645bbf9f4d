"""Backend level 0 three ways at 128 edges (one pass): dense volume, sparse volume rows, compact per-pixel boxes -- build time
and the fused per-corner-gated lookup on each."""
import os, sys, statistics, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import inputs, lgu_slam_b200
ops = lgu_slam_b200.ops
E, H, W, T = 128, 48, 64, 32; dev = "cuda"
g = inputs.gen(5)
fm = (torch.randn(T, 128, H, W, generator=g) / 4).half().to(dev)
planes, cur = [], fm.float()
for l in range(4):
    planes.append(cur.permute(0, 2, 3, 1).reshape(T, -1, 128).half().contiguous())
    cur = torch.nn.functional.avg_pool2d(cur, 2, stride=2)
ii = torch.randint(0, T, (E,), generator=g).to(torch.int32).to(dev); jj = torch.randint(0, T, (E,), generator=g).to(torch.int32).to(dev)
coords = inputs.make_coords(E, H, W, H, W, g).permute(0, 2, 3, 1).contiguous().to(dev)
off = [(4 * torch.tanh(torch.randn(1, H, W, 98, generator=g))).to(dev).contiguous() for _ in range(2)]
hm = ops.volume_half_mask(coords, 0)
out = torch.empty(E, 196, H, W, dtype=torch.float16, device=dev)

def t(f, n=7):
    for _ in range(2): f()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    return round(statistics.median(ts), 1)

print("level 0 dense volume        us", t(lambda: ops.build_volume(planes[0], None, planes[0], None, ii, jj)))
print("level 0 sparse volume rows  us", t(lambda: ops.build_volume(planes[0], None, planes[0], None, ii, jj, half_mask=hm)))
print("level 0 compact boxes       us", t(lambda: ops.build_boxes(planes[0], planes[0], ii, jj, coords, half_mask=hm)))
print("levels 1-3 volumes          us", t(lambda: [ops.build_volume(planes[0], None, planes[l], None, ii, jj) for l in (1, 2, 3)]))
vols = [ops.build_volume(planes[0], None, planes[l], None, ii, jj).view(E, H, W, H >> l, W >> l) for l in range(4)]
boxes = ops.build_boxes(planes[0], planes[0], ii, jj, coords, half_mask=hm)
print("lookup on 4 volumes         us", t(lambda: ops.altcorr_lookup_fused(vols, coords, off[0], off[1], 3, shared_offsets=True, apply_mask=False, out=out)))
print("lookup on boxes + 3 volumes us", t(lambda: ops.altcorr_lookup_fused([None] + vols[1:], coords, off[0], off[1], 3, shared_offsets=True, apply_mask=False, out=out, boxes0=boxes)))
