"""Generates tests/golden/ref_golden.pt by running the REFERENCE kernels (oracle/_ref: /root/reference's CUDA
sources recompiled unmodified for sm_100) on small seeded inputs.  Must run on the GPU box:

    gpurun -- 'python tests/golden/make_golden.py'   ->  gpurun_out/ref_golden.pt  (copy into tests/golden/)

The fixture pins the CPU oracle in the `-m "not gpu"` suite (tests/test_golden_cpu.py)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import inputs  # noqa: E402
from oracle import build_ref  # noqa: E402


def main():
    ref = build_ref.load_ref("defCorrSample_ref")
    alt = build_ref.load_ref("altcorr_ref")
    assert ref is not None and alt is not None, "prebuilt reference modules missing"
    out = {"meta": dict(torch=torch.__version__, device=torch.cuda.get_device_name(0),
                        source="/root/reference offersample_LGS/*.cu + src/altcorr_kernel.cu @ sm_100")}
    cu = lambda t: t.cuda().contiguous()  # noqa: E731
    vol_cases = {"lvl_a": dict(E=2, H1=6, W1=8, H2=6, W2=8, r=3, seed=41, probes=True),
                 "lvl_b": dict(E=1, H1=5, W1=7, H2=9, W2=11, r=1, seed=42, probes=True),
                 "lvl_c": dict(E=1, H1=8, W1=8, H2=4, W2=4, r=3, seed=43, probes=False)}
    for name, kw in vol_cases.items():
        c = inputs.volume_case(**kw)
        r = kw["r"]
        vol, coords, grad = cu(c["volume"]), cu(c["coords"]), cu(c["corr_grad"])
        o1, o2 = cu(c["offset"]), cu(c["offset"])
        out[name] = dict(
            kw=kw,
            corr_index_forward=ref.corr_index_forward(vol, coords, r)[0].cpu(),
            corr_index_backward=ref.corr_index_backward(vol, coords, grad, r)[0].cpu(),
            defCorr_index_forward=ref.defCorr_index_forward(vol, coords, o1, r)[0].cpu(),
            defCorr_index_backward=[t.cpu() for t in ref.defCorr_index_backward(vol, coords, o2, grad, r)],
            offset_after=o1.cpu())
    for name, kw in {"gauss_a": dict(E=2, H1=6, W1=8, H2=6, W2=8, r=4, seed=51, probes=True),
                     "gauss_b": dict(E=1, H1=4, W1=4, H2=10, W2=12, r=2, seed=52, probes=False)}.items():
        c = inputs.gaussian_case(**kw)
        r = kw["r"]
        m, cv, vol, g = cu(c["means"]), cu(c["covs"]), cu(c["volume"]), cu(c["out_grad"])
        out[name] = dict(kw=kw, gaussianMask=ref.gaussianMask(m, cv, vol, r)[0].cpu(),
                         gaussianMask_backward=[t.cpu() for t in ref.gaussianMask_backward(m, cv, vol, g, r)])
    for name, kw in {"lowmem_a": dict(B=3, N=1, H1=6, W1=8, H2=6, W2=8, C=64, r=3, seed=61, probes=True),
                     "lowmem_b": dict(B=2, N=2, H1=5, W1=7, H2=4, W2=6, C=32, r=1, seed=62, probes=True)}.items():
        c = inputs.lowmem_case(**kw)
        r = kw["r"]
        f1, f2, coords, off = cu(c["fmap1"]), cu(c["fmap2"]), cu(c["coords"]), cu(c["offset"])
        out[name] = dict(kw=kw, lowMem_defSample=ref.lowMem_defSample(f1, f2, coords, off, r)[0].cpu(),
                         offset_after=off.cpu(), altcorr_forward=alt.altcorr_forward(f1, f2, coords, r)[0].cpu())
    dst = os.path.join(ROOT, "gpurun_out")
    os.makedirs(dst, exist_ok=True)
    torch.save(out, os.path.join(dst, "ref_golden.pt"))
    print("wrote", os.path.join(dst, "ref_golden.pt"), {k: (list(v) if isinstance(v, dict) else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
