"""CPU tests of the oracle (no GPU): oracle/lgu_oracle.c against an independent naive restatement
(tests/pyref.py) on tiny seeded cases, its edge-case behaviour (NaN / inf / huge / border coordinates,
top-left gating, in-place centre-tap zeroing), and analytic properties."""
import numpy as np
import pytest
import torch

import inputs
import pyref

TOL = 2e-6


def _np(t):
    return t.detach().clone().numpy()


@pytest.mark.parametrize("r,deform,probes", [(1, False, True), (3, True, True), (3, True, False), (2, False, False)])
def test_lookup_forward_matches_naive(oracle, r, deform, probes):
    c = inputs.volume_case(E=2, H1=5, W1=7, H2=6, W2=9, r=r, seed=10 + r, probes=probes)
    off = c["offset"].clone()
    if deform:
        out, = oracle.defCorr_index_forward(c["volume"], c["coords"], off, r)
    else:
        out, = oracle.corr_index_forward(c["volume"], c["coords"], r)
    off_np = _np(c["offset"])
    ref = pyref.lookup_forward(_np(c["volume"]), _np(c["coords"]), off_np, r, deform)
    assert np.array_equal(np.isnan(out.numpy()), np.isnan(ref))
    assert np.array_equal(out.numpy() == 0, ref == 0), "zero (gated) pattern must be identical"
    np.testing.assert_allclose(out.numpy(), ref, atol=TOL, rtol=0, equal_nan=True)
    if deform:
        assert np.array_equal(off.numpy(), off_np), "centre tap must be zeroed in place, nothing else touched"
        assert (off[:, :, :, r, r, :] == 0).all()


@pytest.mark.parametrize("r,deform", [(1, False), (3, True)])
def test_lookup_backward_matches_naive(oracle, r, deform):
    c = inputs.volume_case(E=2, H1=4, W1=6, H2=5, W2=8, r=r, seed=20 + r, probes=False)
    off = c["offset"].clone()
    if deform:
        gv, go = oracle.defCorr_index_backward(c["volume"], c["coords"], off, c["corr_grad"], r)
    else:
        gv, = oracle.corr_index_backward(c["volume"], c["coords"], c["corr_grad"], r)
    rgv, rgo = pyref.lookup_backward(_np(c["volume"]), _np(c["coords"]), _np(c["offset"]), _np(c["corr_grad"]), r,
                                     deform)
    np.testing.assert_allclose(gv.numpy(), rgv, atol=1e-5, rtol=0)
    if deform:
        np.testing.assert_allclose(go.numpy(), rgo, atol=1e-5, rtol=0)
        assert np.array_equal(go.numpy() == 0, rgo == 0)


def test_lookup_backward_is_adjoint_of_forward(oracle):
    """<J v, g> == <v, J^T g> for the volume argument (the lookup is linear in the volume)."""
    c = inputs.volume_case(E=1, H1=6, W1=8, H2=6, W2=8, r=3, seed=5)
    out, = oracle.defCorr_index_forward(c["volume"], c["coords"], c["offset"].clone(), 3)
    gv, _ = oracle.defCorr_index_backward(c["volume"], c["coords"], c["offset"].clone(), c["corr_grad"], 3)
    lhs = (out.double() * c["corr_grad"].double()).sum()
    rhs = (c["volume"].double() * gv.double()).sum()
    assert abs(lhs - rhs) < 1e-3 * max(1.0, abs(lhs))


def test_offset_grad_matches_finite_difference(oracle):
    c = inputs.volume_case(E=1, H1=3, W1=4, H2=8, W2=10, r=3, seed=7)
    # keep samples away from integer boundaries so the lookup is differentiable at the probe points
    off = c["offset"].clone()
    coords = c["coords"].clone()
    pos = off[..., 0] + coords[:, 0][..., None, None]
    frac = pos - pos.floor()
    ok = (frac > 0.05) & (frac < 0.95)
    posy = off[..., 1] + coords[:, 1][..., None, None]
    fracy = posy - posy.floor()
    ok &= (fracy > 0.05) & (fracy < 0.95)
    _, go = oracle.defCorr_index_backward(c["volume"], coords, off.clone(), torch.ones_like(c["corr_grad"]), 3)
    eps = 1e-2
    for ch in (0, 1):
        d = torch.zeros_like(off)
        d[..., ch] = eps
        d[:, :, :, 3, 3, :] = 0
        hi, = oracle.defCorr_index_forward(c["volume"], coords, (off + d).contiguous(), 3)
        lo, = oracle.defCorr_index_forward(c["volume"], coords, (off - d).contiguous(), 3)
        fd = ((hi - lo) / (2 * eps)).permute(0, 3, 4, 1, 2)                      # [E,H1,W1,i,j]
        an = go[..., ch]
        sel = ok.clone()
        sel[:, :, :, 3, 3] = False
        assert torch.allclose(fd[sel], an[sel], atol=2e-3), (fd[sel] - an[sel]).abs().max()


@pytest.mark.parametrize("probes", [False, True])
def test_gaussian_matches_naive(oracle, probes):
    c = inputs.gaussian_case(E=2, H1=4, W1=5, H2=7, W2=9, r=4, seed=3, probes=probes)
    out, = oracle.gaussianMask(c["means"], c["covs"], c["volume"], 4)
    ref = pyref.gaussian_forward(_np(c["means"]), _np(c["covs"]), _np(c["volume"]), 4)
    assert np.array_equal(out.numpy() == 0, ref == 0)
    np.testing.assert_allclose(out.numpy(), ref, atol=TOL, rtol=1e-6)
    gm, gc = oracle.gaussianMask_backward(c["means"], c["covs"], c["volume"], c["out_grad"], 4)
    rgm, rgc = pyref.gaussian_backward(_np(c["means"]), _np(c["covs"]), _np(c["volume"]), _np(c["out_grad"]), 4)
    np.testing.assert_allclose(gm.numpy(), rgm, atol=2e-5, rtol=1e-5)
    np.testing.assert_allclose(gc.numpy(), rgc, atol=2e-5, rtol=1e-5)


def test_gaussian_grads_match_finite_difference(oracle):
    c = inputs.gaussian_case(E=1, H1=3, W1=3, H2=12, W2=12, r=4, seed=4)
    means = c["means"].double().add(0.013).float()       # keep floor(mean) stable under +-eps
    frac = means - means.floor()
    means = torch.where((frac < 0.1) | (frac > 0.9), means.floor() + 0.5, means).contiguous()
    g = c["out_grad"]
    gm, gc = oracle.gaussianMask_backward(means, c["covs"], c["volume"], g, 4)
    eps = 1e-3

    def f(m, cv):
        o, = oracle.gaussianMask(m.contiguous(), cv.contiguous(), c["volume"], 4)
        return (o.double() * g.double()).sum(dim=(3, 4))

    for ch in (0, 1):
        d = torch.zeros_like(means); d[..., ch] = eps
        fd = (f(means + d, c["covs"]) - f(means - d, c["covs"])) / (2 * eps)
        assert torch.allclose(fd.float(), gm[..., ch], atol=5e-2, rtol=2e-2)
        fd = (f(means, c["covs"] + d) - f(means, c["covs"] - d)) / (2 * eps)
        assert torch.allclose(fd.float(), gc[..., ch], atol=5e-2, rtol=2e-2)


@pytest.mark.parametrize("strict,N", [(True, 1), (False, 2), (True, 2)])
def test_lowmem_matches_naive(oracle, strict, N):
    c = inputs.lowmem_case(B=2, N=N, H1=3, W1=5, H2=4, W2=6, C=64, r=2, seed=11, probes=True)
    off = c["offset"].clone()
    out, = oracle.lowMem_defSample(c["fmap1"], c["fmap2"], c["coords"], off, 2, strict_ref=strict)
    off_np = _np(c["offset"])
    ref = pyref.lowmem_forward(_np(c["fmap1"]), _np(c["fmap2"]), _np(c["coords"]), off_np, 2, strict)
    np.testing.assert_allclose(out.numpy(), ref, atol=1e-5, rtol=0, equal_nan=True)
    assert np.array_equal(off.numpy(), off_np)


def test_altcorr_matches_naive(oracle):
    c = inputs.lowmem_case(B=2, N=2, H1=3, W1=5, H2=4, W2=6, C=64, r=1, seed=12, probes=True)
    out, = oracle.altcorr_forward(c["fmap1"], c["fmap2"], c["coords"], 1)
    ref = pyref.altcorr_forward(_np(c["fmap1"]), _np(c["fmap2"]), _np(c["coords"]), 1)
    np.testing.assert_allclose(out.numpy(), ref, atol=1e-5, rtol=0, equal_nan=True)


def test_lowmem_equals_volume_lookup_in_the_interior(oracle):
    """lowMem (on the fly) == defCorr lookup of the explicit volume wherever no corner is clipped
    (the two ops differ only in border gating, quirks Q3/Q4)."""
    B, H, W, C, r = 1, 8, 10, 32, 3
    g = inputs.gen(21)
    f1 = torch.randn(B, H, W, C, generator=g)
    f2 = torch.randn(B, H, W, C, generator=g)
    vol = torch.einsum("bhwc,byxc->bhwyx", f1.double(), f2.double()).float().contiguous()
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    coords = torch.stack([xs * 0 + 4.3, ys * 0 + 3.6], 0)[None].contiguous()       # taps stay inside
    off = torch.zeros(B, H, W, 7, 7, 2)
    a, = oracle.defCorr_index_forward(vol, coords, off.clone(), r)
    b, = oracle.lowMem_defSample(f1, f2, coords.permute(0, 2, 3, 1)[:, None].contiguous(), off.clone(), r)
    assert torch.allclose(a, b[:, 0], atol=1e-4)


def test_volume_pool_and_residual(oracle):
    g = inputs.gen(5)
    f1 = torch.randn(2, 16, 4, 6, generator=g)
    f2 = torch.randn(2, 16, 4, 6, generator=g)
    vol = oracle.corr_volume(f1, f2)
    ref = torch.einsum("ecp,ecq->epq", (f1 / 4).flatten(2).double(), (f2 / 4).flatten(2).double()).float()
    assert torch.allclose(vol.view(2, 24, 24), ref, atol=1e-6)
    pooled = oracle.avg_pool2x2(vol)
    assert torch.allclose(pooled, torch.nn.functional.avg_pool2d(vol.view(-1, 1, 4, 6), 2, 2).view(2, 4, 6, 2, 3),
                          atol=1e-7)
    covs = torch.rand(2, 4, 6, 2, generator=g) + 0.1
    masked = torch.randn(2, 4, 6, 4, 6, generator=g)
    den = oracle.gaussian_den(covs)
    res = oracle.gaussian_residual(masked, den, vol)
    assert torch.equal(res, masked / den.view(2, 4, 6, 1, 1) + vol)
    half = oracle.corr_volume(f1, f2, round_half=True)
    assert torch.equal(half, vol.half().float())
