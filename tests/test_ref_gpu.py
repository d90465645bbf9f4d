"""GPU tests against the REFERENCE ITSELF: /root/reference's CUDA kernels recompiled, unmodified, for sm_100
(oracle/_ref, built by oracle/build_ref.py in the dev container and shipped prebuilt to the GPU box).

Two things are pinned here:
  1. the CPU oracle (oracle/lgu_oracle.c) == the reference kernels          -> "the oracle is right";
  2. the sm_100a product kernels (through the C ABI) == the reference kernels -> the drop-in claim.
Skipped (not failed) if the prebuilt reference modules are absent."""
import pytest
import torch

import inputs

pytestmark = pytest.mark.gpu
ATOL = 1e-5


@pytest.fixture(scope="module")
def ref():
    from oracle import build_ref
    m = build_ref.load_ref("defCorrSample_ref")
    if m is None:
        pytest.skip("oracle/_ref/defCorrSample_ref*.so not built (run python oracle/build_ref.py in the dev container)")
    return m


@pytest.fixture(scope="module")
def ref_alt():
    from oracle import build_ref
    m = build_ref.load_ref("altcorr_ref")
    if m is None:
        pytest.skip("oracle/_ref/altcorr_ref*.so not built")
    return m


def cu(t):
    return t.cuda().contiguous()


def eq_nan(a, b):
    return torch.equal(torch.isnan(a), torch.isnan(b)) and torch.equal(torch.nan_to_num(a), torch.nan_to_num(b))


SHAPES = [(2, 6, 8, 6, 8, 3, True), (1, 48, 64, 48, 64, 3, True), (2, 48, 64, 24, 32, 3, True),
          (2, 48, 64, 12, 16, 3, False), (2, 48, 64, 6, 8, 3, True), (2, 48, 64, 24, 32, 1, True),
          (1, 5, 7, 9, 11, 2, True), (1, 20, 20, 6, 6, 4, False)]


@pytest.mark.parametrize("E,H1,W1,H2,W2,r,probes", SHAPES)
def test_forward_lookups_bit_exact_vs_reference(ops, oracle, ref, E, H1, W1, H2, W2, r, probes):
    c = inputs.volume_case(E, H1, W1, H2, W2, r, seed=1000 + r, probes=probes)
    vol, coords = cu(c["volume"]), cu(c["coords"])
    # plain lookup
    want, = ref.corr_index_forward(vol, coords, r)
    got, = ops.corr_index_forward(vol, coords, r)
    orc, = oracle.corr_index_forward(c["volume"], c["coords"], r)
    assert eq_nan(got, want), "product corr_index_forward != reference (bit-exact expected)"
    assert eq_nan(orc, want.cpu()), "oracle corr_index_forward != reference (bit-exact expected)"
    # deformable lookup + in-place offset mutation
    o_ref, o_got, o_orc = cu(c["offset"]), cu(c["offset"]), c["offset"].clone()
    want, = ref.defCorr_index_forward(vol, coords, o_ref, r)
    got, = ops.defCorr_index_forward(vol, coords, o_got, r)
    orc, = oracle.defCorr_index_forward(c["volume"], c["coords"], o_orc, r)
    assert eq_nan(got, want) and torch.equal(o_got, o_ref)
    assert eq_nan(orc, want.cpu()) and torch.equal(o_orc, o_ref.cpu())


@pytest.mark.parametrize("E,H1,W1,H2,W2,r,probes", SHAPES)
def test_backward_lookups_vs_reference(ops, oracle, ref, E, H1, W1, H2, W2, r, probes):
    c = inputs.volume_case(E, H1, W1, H2, W2, r, seed=1100 + r, probes=probes)
    vol, coords, grad = cu(c["volume"]), cu(c["coords"]), cu(c["corr_grad"])
    want, = ref.corr_index_backward(vol, coords, grad, r)
    got, = ops.corr_index_backward(vol, coords, grad, r)
    orc, = oracle.corr_index_backward(c["volume"], c["coords"], c["corr_grad"], r)
    assert torch.allclose(got, want, atol=ATOL, rtol=0, equal_nan=True)
    assert eq_nan(orc, want.cpu()), "oracle follows the reference's accumulation order: bit-exact expected"
    o_ref, o_got, o_orc = cu(c["offset"]), cu(c["offset"]), c["offset"].clone()
    wv, wo = ref.defCorr_index_backward(vol, coords, o_ref, grad, r)
    gv, go = ops.defCorr_index_backward(vol, coords, o_got, grad, r)
    ov, oo = oracle.defCorr_index_backward(c["volume"], c["coords"], o_orc, c["corr_grad"], r)
    assert torch.allclose(gv, wv, atol=ATOL, rtol=0, equal_nan=True)
    assert eq_nan(go, wo), "offset_grad: same operation order as the reference -> bit-exact"
    assert torch.equal(o_got, o_ref)
    assert eq_nan(ov, wv.cpu()) and eq_nan(oo, wo.cpu()) and torch.equal(o_orc, o_ref.cpu())


@pytest.mark.parametrize("E,H1,W1,H2,W2,r,probes", [(2, 6, 8, 6, 8, 4, True), (1, 48, 64, 48, 64, 4, True),
                                                    (2, 5, 7, 9, 11, 4, True), (1, 4, 4, 10, 12, 2, False)])
def test_gaussian_vs_reference(ops, oracle, ref, E, H1, W1, H2, W2, r, probes):
    c = inputs.gaussian_case(E, H1, W1, H2, W2, r, seed=1200 + r, probes=probes)
    m, cv, vol, g = cu(c["means"]), cu(c["covs"]), cu(c["volume"]), cu(c["out_grad"])
    want, = ref.gaussianMask(m, cv, vol, r)
    got, = ops.gaussianMask(m, cv, vol, r)
    orc, = oracle.gaussianMask(c["means"], c["covs"], c["volume"], r)
    assert torch.equal(got, want), "same arithmetic + same CUDA expf -> bit-exact"
    assert torch.equal(orc == 0, want.cpu() == 0)
    assert torch.allclose(orc, want.cpu(), atol=ATOL, rtol=1e-6)          # glibc expf vs CUDA expf
    wm, wc = ref.gaussianMask_backward(m, cv, vol, g, r)
    gm, gc = ops.gaussianMask_backward(m, cv, vol, g, r)
    om, oc = oracle.gaussianMask_backward(c["means"], c["covs"], c["volume"], c["out_grad"], r)
    for a, b in ((gm, wm), (gc, wc), (om.cuda(), wm), (oc.cuda(), wc)):
        assert torch.allclose(a, b, atol=ATOL, rtol=1e-5), (a - b).abs().max()


@pytest.mark.parametrize("B,N,H1,W1,H2,W2,C,r,probes", [(3, 1, 6, 8, 6, 8, 128, 3, True),
                                                        (2, 1, 48, 64, 48, 64, 128, 3, True),
                                                        (2, 1, 48, 64, 12, 16, 128, 3, False),
                                                        (2, 2, 5, 7, 4, 6, 64, 2, True)])
def test_lowmem_vs_reference(ops, oracle, ref, B, N, H1, W1, H2, W2, C, r, probes):
    c = inputs.lowmem_case(B, N, H1, W1, H2, W2, C, r, seed=1300 + r, probes=probes)
    f1, f2, coords = cu(c["fmap1"]), cu(c["fmap2"]), cu(c["coords"])
    o_ref, o_got, o_orc = cu(c["offset"]), cu(c["offset"]), c["offset"].clone()
    want, = ref.lowMem_defSample(f1, f2, coords, o_ref, r)                 # the reference IS strict (quirk Q2)
    got, = ops.lowMem_defSample(f1, f2, coords, o_got, r, strict_ref=True)
    orc, = oracle.lowMem_defSample(c["fmap1"], c["fmap2"], c["coords"], o_orc, r, strict_ref=True)
    assert torch.allclose(got, want, atol=ATOL, rtol=0, equal_nan=True), (got - want).abs().nan_to_num().max()
    assert torch.allclose(orc, want.cpu(), atol=ATOL, rtol=0, equal_nan=True)
    assert torch.equal(o_got, o_ref) and torch.equal(o_orc, o_ref.cpu())


@pytest.mark.parametrize("B,N,H1,W1,H2,W2,C,r,probes", [(3, 1, 6, 8, 6, 8, 128, 1, True),
                                                        (2, 1, 48, 64, 24, 32, 128, 1, True),
                                                        (2, 2, 5, 7, 4, 6, 64, 3, True)])
def test_altcorr_vs_reference(ops, oracle, ref_alt, B, N, H1, W1, H2, W2, C, r, probes):
    c = inputs.lowmem_case(B, N, H1, W1, H2, W2, C, r, seed=1400 + r, probes=probes)
    f1, f2, coords = cu(c["fmap1"]), cu(c["fmap2"]), cu(c["coords"])
    want, = ref_alt.altcorr_forward(f1, f2, coords, r)
    got, = ops.altcorr_forward(f1, f2, coords, r)
    orc, = oracle.altcorr_forward(c["fmap1"], c["fmap2"], c["coords"], r)
    assert torch.allclose(got, want, atol=ATOL, rtol=0, equal_nan=True), (got - want).abs().nan_to_num().max()
    assert torch.allclose(orc, want.cpu(), atol=ATOL, rtol=0, equal_nan=True)
