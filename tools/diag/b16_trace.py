"""Where the 16-warp build kernel's epilogue warps spend their cycles (diagnostic build with -DLGU_B16_TRACE):
   cd lgu-slam_b200/csrc && nvcc <flags of the Makefile> -DLGU_B16_TRACE *.cu -shared -o ../lib_b16trace.so
   LGU_CORR_LIB=lgu-slam_b200/lib_b16trace.so python tools/diag/b16_trace.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
trace = torch.zeros(8, dtype=torch.int64, device="cuda")
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
names = ["wait accumulator", "tcgen05.ld x2", "wait staging tile free", "staging writes + patch", "fence+syncwarp+issue",
         "barrier A", "L1 staging + barrier B"]
for gauss in (True, False):
    k = kw if gauss else {}
    for _ in range(2): ops.build_pyramid(*args, **k)
    torch.cuda.synchronize(); trace.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ops.build_pyramid(*args, **k); b.record(); torch.cuda.synchronize()
    t = trace.tolist(); tot = t[7]
    if tot == 0:
        print(f"gauss={gauss}: {a.elapsed_time(b)*1e3:.0f} us (library built without -DLGU_B16_TRACE)"); continue
    rest = tot - sum(t[:7])
    print(f"gauss={gauss} dbg={os.environ.get('LGU_BUILD_DBG','0')}: {a.elapsed_time(b)*1e3:.0f} us; per warp {tot/(148*16)/1.965e3:.0f} us; " +
          ", ".join(f"{n} {100*v/tot:.1f} %" for n, v in zip(names, t[:7])) + f", rest {100*rest/tot:.1f} %")
