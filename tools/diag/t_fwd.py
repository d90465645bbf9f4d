"""Time the fused lookup forward at E = 48 (frontend_w20_e48 shapes).  LGU_CORR_LIB selects the library variant;
modes: writeback (lgu_corr_lookup_fused; includes no clone), cum (lgu_corr_lookup_fused_cum)."""
import os, sys, statistics, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import inputs, lgu_slam_b200
ops = lgu_slam_b200.ops
E, H, W = 48, 48, 64; dev = "cuda"
g = torch.Generator(device=dev); g.manual_seed(1)
fc = inputs.frontend_case(E=E, T=20, seed=5, half_fmaps=True)
pyr = [torch.randn(E, H, W, H >> l, W >> l, device=dev, generator=g) for l in range(4)]
o1 = fc["offsets"][1].to(dev); o0 = fc["offsets"][0].to(dev); co = fc["coords"].to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
modes = ["writeback"] + (["cum"] if hasattr(lgu_slam_b200._lib.lib(), "lgu_corr_lookup_fused_cum") else [])
if hasattr(lgu_slam_b200._lib.lib(), "lgu_corr_lookup_fused_enc"):
    modes += ["enc+corr", "enc", "enc_fp16", "cum+cudnn_conv"]
    conv = torch.nn.Conv2d(196, 128, 1).to(dev)
    frag = ops.pack_conv1x1(conv.weight); bias = conv.bias.detach().contiguous()
for mode in modes:
    cum = torch.ones(E, H, W, device=dev)
    def f(o):
        if mode == "cum": ops.corr_lookup_fused(pyr, co, o0, o, 3, cum_mask=cum)
        elif mode == "enc+corr": ops.corr_lookup_fused_enc(pyr, co, o0, o, cum, frag, bias, keep_corr=True)
        elif mode == "enc": ops.corr_lookup_fused_enc(pyr, co, o0, o, cum, frag, bias)
        elif mode == "enc_fp16": ops.corr_lookup_fused_enc(pyr, co, o0, o, cum, frag, bias, enc_half=True)
        elif mode == "cum+cudnn_conv":
            with torch.no_grad(): torch.relu_(conv(ops.corr_lookup_fused(pyr, co, o0, o, 3, cum_mask=cum)))
        else: ops.corr_lookup_fused(pyr, co, o0, o, 3)
    for _ in range(3): f(o1.clone())
    torch.cuda.synchronize(); ts = []
    for _ in range(20):
        o = o1.clone(); flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(o); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    print(os.path.basename(os.environ.get("LGU_CORR_LIB", "default")), mode, "fused lookup forward median us",
          round(statistics.median(ts), 1), "min", round(min(ts), 1))
