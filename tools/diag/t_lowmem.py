"""lowMem_defSample / altcorr_forward per level at B = 48 edges (48x64x128 maps): tensor-core path vs the on-the-fly SIMT
kernel vs the reference's kernel recompiled for sm_100 (oracle/_ref), CUDA events."""
import os, sys, statistics, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import inputs, lgu_slam_b200
from oracle import build_ref
ops = lgu_slam_b200.ops
ref = build_ref.load_ref("defCorrSample_ref"); ref_alt = build_ref.load_ref("altcorr_ref")
B, H, W, C = int(os.environ.get("B", 48)), 48, 64, 128; dev = "cuda"
def t(fn, n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    return statistics.median(ts)
tot = {"tc": 0.0, "simt": 0.0, "ref": 0.0}
for l in range(4):
    c = inputs.lowmem_case(B, 1, H, W, H >> l, W >> l, C, 3, seed=40 + l)
    f1, f2, co, off = (c[k].to(dev) for k in ("fmap1", "fmap2", "coords", "offset"))
    a = t(lambda: ops.lowMem_defSample(f1, f2, co, off, 3))
    b = t(lambda: ops.lowMem_defSample(f1, f2, co, off, 3, tensor_cores=False), 3)
    r = t(lambda: ref.lowMem_defSample(f1, f2, co, off, 3), 3) if ref is not None else float("nan")
    x, = ops.lowMem_defSample(f1, f2, co, off.clone(), 3); y, = ops.lowMem_defSample(f1, f2, co, off.clone(), 3, tensor_cores=False)
    print(f"level {l}: tensor-core {a:8.1f} us | simt {b:8.1f} us | reference kernel {r:8.1f} us | max |tc - simt| {(x - y).abs().max().item():.2e}")
    tot["tc"] += a; tot["simt"] += b; tot["ref"] += r
c = inputs.lowmem_case(B, 1, H, W, H >> 1, W >> 1, C, 1, seed=50)
f1, f2, co = (c[k].to(dev) for k in ("fmap1", "fmap2", "coords"))
a = t(lambda: ops.altcorr_forward(f1, f2, co, 1)); b = t(lambda: ops.altcorr_forward(f1, f2, co, 1, tensor_cores=False), 3)
r = t(lambda: ref_alt.altcorr_forward(f1, f2, co, 1), 3) if ref_alt is not None else float("nan")
print(f"altcorr r=1 level 1: tensor-core {a:8.1f} us | simt {b:8.1f} us | reference kernel {r:8.1f} us")
tot["tc"] += a; tot["simt"] += b; tot["ref"] += r
alg = 17.81e6 * B
print(f"B={B}: 4 levels + altcorr: tensor-core {tot['tc']:.0f} us ({alg / tot['tc'] / 1e3 / 6552:.3f} of the HBM roofline on 17.81 MB/edge) | simt {tot['simt']:.0f} us | reference kernels {tot['ref']:.0f} us")
