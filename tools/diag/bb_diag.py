"""Error anatomy of lgu_build_backward_fmaps: which operand split / level / product is off."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import lgu_slam_b200
ops = lgu_slam_b200.ops
dev = "cuda"
E, C, H, W = 1, 128, 48, 64
g = torch.Generator().manual_seed(1)
def trunc(x): return (x.view(torch.int32) & -8192).view(torch.float32)
def ref(grads, f1, f2):
    gd = torch.zeros(E, H, W, H, W, dtype=torch.float64, device=dev)
    for l, gl in enumerate(grads):
        if gl is None: continue
        k = 1 << l
        gd += gl.double().repeat_interleave(k, 3).repeat_interleave(k, 4) / (k * k)
    gm = gd.view(E, H * W, H * W)
    return (torch.bmm(f2.double().reshape(E, C, -1), gm.transpose(1, 2)) / 16).view_as(f1), (torch.bmm(f1.double().reshape(E, C, -1), gm) / 16).view_as(f2)
for lv in range(4):
    for gex, fex in ((1, 1), (0, 1), (1, 0), (0, 0)):
        f1 = torch.randn(E, C, H, W, generator=g); f2 = torch.randn(E, C, H, W, generator=g)
        G = torch.randn(E, H, W, H >> lv, W >> lv, generator=g)
        if gex: G = trunc(G)
        if fex: f1, f2 = trunc(f1), trunc(f2)     # /16 and 2x2 averaging of tf32-exact values may still need > 10 bits at l > 0
        f1, f2, G = f1.to(dev), f2.to(dev), G.to(dev)
        grads = [None] * 4; grads[lv] = G
        a1, a2 = ops.build_backward_fmaps(grads, f1, f2)
        w1, w2 = ref(grads, f1, f2)
        r1 = ((a1.double() - w1).abs().max() / w1.square().mean().sqrt()).item()
        r2 = ((a2.double() - w2).abs().max() / w2.square().mean().sqrt()).item()
        print(f"level {lv} G_exact={gex} f_exact={fex}: g_f1 err/rms {r1:.2e}   g_f2 err/rms {r2:.2e}")
