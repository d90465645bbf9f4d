import os, sys, statistics, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import inputs, lgu_slam_b200
ops = lgu_slam_b200.ops
E, H, W, r = 48, 48, 64, 3
for lvl in (0, 3):
    H2, W2 = H >> lvl, W >> lvl
    vol = torch.randn(E, H, W, H2, W2, device="cuda")
    coords = inputs.make_coords(E, H, W, H2, W2, inputs.gen(2)).cuda()
    off = inputs.make_offset(E, H, W, r, inputs.gen(7), zero=(lvl >= 2)).cuda()
    for _ in range(3): ops.defCorr_index_forward(vol, coords, off, r)
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.defCorr_index_forward(vol, coords, off, r); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print("LGU_EXP", os.environ.get("LGU_EXP", "0"), "lvl", lvl, "median ms", statistics.median(ts), "min", min(ts))
