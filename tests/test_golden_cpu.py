"""Pins the CPU oracle against golden vectors produced by the REFERENCE kernels themselves
(tests/golden/ref_golden.pt: /root/reference's CUDA sources recompiled unmodified for sm_100 and run on a B200
by tests/golden/make_golden.py; inputs are regenerated here from the recorded seeds).  Runs without a GPU."""
import os

import pytest
import torch

import inputs

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_golden.pt")
ATOL = 1e-5


@pytest.fixture(scope="module")
def gold():
    if not os.path.exists(GOLD):
        pytest.fail("tests/golden/ref_golden.pt is missing (generate with tests/golden/make_golden.py on the GPU box)")
    return torch.load(GOLD, weights_only=False)


def eq_nan(a, b):
    return torch.equal(torch.isnan(a), torch.isnan(b)) and torch.equal(torch.nan_to_num(a), torch.nan_to_num(b))


@pytest.mark.parametrize("name", ["lvl_a", "lvl_b", "lvl_c"])
def test_lookup_ops_bit_exact_vs_reference_golden(oracle, gold, name):
    G = gold[name]
    c = inputs.volume_case(**G["kw"])
    r = G["kw"]["r"]
    out, = oracle.corr_index_forward(c["volume"], c["coords"], r)
    assert eq_nan(out, G["corr_index_forward"])
    gv, = oracle.corr_index_backward(c["volume"], c["coords"], c["corr_grad"], r)
    assert eq_nan(gv, G["corr_index_backward"])
    off = c["offset"].clone()
    out, = oracle.defCorr_index_forward(c["volume"], c["coords"], off, r)
    assert eq_nan(out, G["defCorr_index_forward"])
    assert torch.equal(off, G["offset_after"]), "in-place centre-tap zeroing"
    gv, go = oracle.defCorr_index_backward(c["volume"], c["coords"], c["offset"].clone(), c["corr_grad"], r)
    assert eq_nan(gv, G["defCorr_index_backward"][0])
    assert eq_nan(go, G["defCorr_index_backward"][1])


@pytest.mark.parametrize("name", ["gauss_a", "gauss_b"])
def test_gaussian_vs_reference_golden(oracle, gold, name):
    G = gold[name]
    c = inputs.gaussian_case(**G["kw"])
    r = G["kw"]["r"]
    out, = oracle.gaussianMask(c["means"], c["covs"], c["volume"], r)
    assert torch.equal(out == 0, G["gaussianMask"] == 0), "window / bounds logic"
    assert torch.allclose(out, G["gaussianMask"], atol=ATOL, rtol=1e-6)        # glibc vs CUDA expf: <= 2 ulp
    gm, gc = oracle.gaussianMask_backward(c["means"], c["covs"], c["volume"], c["out_grad"], r)
    assert torch.allclose(gm, G["gaussianMask_backward"][0], atol=ATOL, rtol=1e-5)
    assert torch.allclose(gc, G["gaussianMask_backward"][1], atol=ATOL, rtol=1e-5)


@pytest.mark.parametrize("name", ["lowmem_a", "lowmem_b"])
def test_lowmem_altcorr_vs_reference_golden(oracle, gold, name):
    G = gold[name]
    c = inputs.lowmem_case(**G["kw"])
    r = G["kw"]["r"]
    off = c["offset"].clone()
    out, = oracle.lowMem_defSample(c["fmap1"], c["fmap2"], c["coords"], off, r, strict_ref=True)
    assert torch.allclose(out, G["lowMem_defSample"], atol=ATOL, rtol=0, equal_nan=True)
    assert torch.equal(off, G["offset_after"])
    out, = oracle.altcorr_forward(c["fmap1"], c["fmap2"], c["coords"], r)
    assert torch.allclose(out, G["altcorr_forward"], atol=ATOL, rtol=0, equal_nan=True)
