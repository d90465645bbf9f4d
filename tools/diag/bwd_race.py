import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import inputs, lgu_slam_b200
ops = lgu_slam_b200.ops
dev = "cuda"
for E, big in ((2, True), (2, False), (8, True)):
    g = inputs.gen(71)
    pyr = [torch.randn(E, 48, 64, 48 >> l, 64 >> l, generator=g).to(dev) for l in range(4)]
    c = inputs.frontend_case(E=E, T=max(3, E // 2), seed=72, half_fmaps=True)
    if big:
        g2 = inputs.gen(73)
        c["offsets"][0] = (9.0 * torch.randn(E, 48, 64, 98, generator=g2)).contiguous()
        c["offsets"][1] = (7.0 * torch.randn(E, 48, 64, 98, generator=g2)).contiguous()
    coords, off0, off1 = c["coords"].to(dev), c["offsets"][0].to(dev), c["offsets"][1].to(dev)
    _, mask = ops.corr_lookup_fused(pyr, coords, off0, off1, 3, return_mask=True)
    g_corr = torch.randn(E, 196, 48, 64, generator=g).to(dev)
    g_up = (0.1 * torch.randn(E, 48, 64, 98, generator=g)).to(dev)
    runs = []
    for mode in ("dense", "dense", "acc", "acc"):
        acc = [torch.zeros_like(p) for p in pyr] if mode == "acc" else None
        r = ops.corr_lookup_fused_backward(pyr, coords, off0, off1, mask, g_corr, g_up, accumulate_into=acc)
        torch.cuda.synchronize()
        runs.append([t.clone() for t in r])
    names = ["gv0", "gv1", "gv2", "gv3", "g_off0", "g_off1"]
    for a, b, lab in ((0, 1, "dense vs dense"), (2, 3, "acc vs acc"), (0, 2, "dense vs acc")):
        for k in range(6):
            x, y = torch.nan_to_num(runs[a][k]), torch.nan_to_num(runs[b][k])
            if not torch.equal(x, y):
                d = (x - y).abs()
                idx = torch.nonzero(d.view(E, 48 * 64, -1).amax(2) > 0)
                print(f"E={E} big={big} {lab}: {names[k]} differs, max {d.max().item():.3e}, {idx.shape[0]} pixels, first {idx[:6].tolist()}, pixel%4 hist {torch.bincount(idx[:,1] % 4, minlength=4).tolist()}")
print("done")
