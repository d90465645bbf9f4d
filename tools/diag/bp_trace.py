"""Where the build kernel's epilogue warps wait (diagnostic build with -DLGU_BP_TRACE, see build_pyramid.cu):
   LGU_CORR_LIB=.../lib_bptrace.so python tools/diag/bp_trace.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
trace = torch.zeros(4, dtype=torch.int64, device="cuda")
os.environ["LGU_BP_TRACE_PTR"] = str(trace.data_ptr())
import inputs, lgu_slam_b200
ops = lgu_slam_b200.ops
E, H, W = 48, 48, 64
fc = inputs.frontend_case(E=E, T=20, seed=5, half_fmaps=True)
dev = "cuda"
hi, _ = ops.pack_fmaps(fc["fmaps"].half().to(dev))
den = (6.28 * torch.sqrt(fc["covs"][..., 0] * fc["covs"][..., 1])).to(dev).contiguous()
args = (hi, None, fc["ii"].to(dev), fc["jj"].to(dev), H, W)
kw = dict(means=fc["means"].to(dev), covs=fc["covs"].to(dev), den=den)
for gauss in (True, False):
    k = kw if gauss else {}
    for _ in range(2): ops.build_pyramid(*args, **k)
    torch.cuda.synchronize(); trace.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ops.build_pyramid(*args, **k); b.record(); torch.cuda.synchronize()
    t = trace.tolist(); tot = t[3]
    if tot == 0:
        print(f"gauss={gauss}: {a.elapsed_time(b)*1e3:.0f} us (library built without -DLGU_BP_TRACE: no counters)")
        continue
    if os.environ.get("LGU_BP_TRACE_MODE") == "2":
        print(f"gauss={gauss}: {a.elapsed_time(b)*1e3:.0f} us; tcgen05.ld pairs {100*t[0]/tot:.1f} %, staging writes + Gaussian patch "
              f"{100*t[1]/tot:.1f} %, fence.proxy.async + syncwarp + store issue {100*t[2]/tot:.1f} %, everything else {100*(tot-t[0]-t[1]-t[2])/tot:.1f} %")
        continue
    if os.environ.get("LGU_BP_TRACE_MODE") == "3":
        print(f"gauss={gauss}: {a.elapsed_time(b)*1e3:.0f} us; 2x2 pooling {100*t[0]/tot:.1f} %, level-1 path (pair barriers, staging, store) "
              f"{100*t[1]/tot:.1f} %, levels 2/3 {100*t[2]/tot:.1f} %, everything else {100*(tot-t[0]-t[1]-t[2])/tot:.1f} %")
        continue
    print(f"gauss={gauss}: {a.elapsed_time(b)*1e3:.0f} us; epilogue warps (148 x 8): total {tot/1184/1.965e3:.0f} us each; "
          f"waiting for accumulator {100*t[0]/tot:.1f} %, for a free staging buffer (TMA store engine) {100*t[1]/tot:.1f} %, "
          f"pair barriers {100*t[2]/tot:.1f} %, everything else {100*(tot-t[0]-t[1]-t[2])/tot:.1f} %")
