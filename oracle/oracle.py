"""TEST INFRASTRUCTURE ONLY -- Python face of the CPU oracle (oracle/lgu_oracle.c).

Mirrors the reference's operator API one to one (module `defCorrSample`,
/root/reference/offersample_LGS/droid.cpp:138-147, plus `droid_backends.altcorr_forward`,
/root/reference/src/droid.cpp:193-203) on CPU torch tensors, and composes them the way
/root/reference/droid_slam/modules/corr.py:53-109 (CorrBlock) does.  Every op follows the
reference file:line cited in lgu_oracle.c.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may
import this module; the product package never does.
"""
import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liblgu_oracle.so")
_lib = None


def build(force=False):
    if force or not os.path.exists(_LIB_PATH) or \
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "lgu_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"] if force else ["make", "-C", _HERE, "-s"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.orc_num_threads.restype = ctypes.c_int
    return _lib


def num_threads():
    return int(lib().orc_num_threads())


def set_num_threads(n):
    lib().orc_set_num_threads(ctypes.c_int(int(n)))


def _p(t):
    assert t.device.type == "cpu" and t.dtype == torch.float32 and t.is_contiguous(), \
        "oracle works on contiguous fp32 CPU tensors"
    return ctypes.c_void_p(t.data_ptr())


def _i(v):
    return ctypes.c_int(int(v))


def _l(v):
    return ctypes.c_int64(int(v))


# ---------------------------------------------------------------- the 7 reference ops (+ altcorr)
def corr_index_forward(volume, coords, radius):
    E, H1, W1, H2, W2 = volume.shape
    rd = 2 * radius + 1
    out = torch.empty(E, rd, rd, H1, W1, dtype=torch.float32)
    lib().orc_corr_index_forward(_p(volume), _p(coords), _p(out), _i(E), _i(H1), _i(W1), _i(H2), _i(W2), _i(radius))
    return [out]


def corr_index_backward(volume, coords, corr_grad, radius):
    E, H1, W1, H2, W2 = volume.shape
    g = torch.empty_like(volume)
    lib().orc_corr_index_backward(_p(coords), _p(corr_grad), _p(g), _i(E), _i(H1), _i(W1), _i(H2), _i(W2), _i(radius))
    return [g]


def defCorr_index_forward(volume, coords, offset, radius):
    E, H1, W1, H2, W2 = volume.shape
    rd = 2 * radius + 1
    out = torch.empty(E, rd, rd, H1, W1, dtype=torch.float32)
    lib().orc_defcorr_index_forward(_p(volume), _p(coords), _p(offset), _p(out),
                                    _i(E), _i(H1), _i(W1), _i(H2), _i(W2), _i(radius))
    return [out]


def defCorr_index_backward(volume, coords, offset, corr_grad, radius):
    E, H1, W1, H2, W2 = volume.shape
    gv = torch.empty_like(volume)
    go = torch.empty_like(offset)
    lib().orc_defcorr_index_backward(_p(volume), _p(coords), _p(offset), _p(corr_grad), _p(gv), _p(go),
                                     _i(E), _i(H1), _i(W1), _i(H2), _i(W2), _i(radius))
    return [gv, go]


def gaussianMask(means, covs, volume, radius):
    E, H1, W1, H2, W2 = volume.shape
    out = torch.empty_like(volume)
    lib().orc_gaussian_mask_forward(_p(means), _p(covs), _p(volume), _p(out),
                                    _i(E), _i(H1), _i(W1), _i(H2), _i(W2), _i(radius))
    return [out]


def gaussianMask_backward(means, covs, volume, volume1_grad, radius):
    E, H1, W1, H2, W2 = volume.shape
    gm = torch.empty_like(means)
    gc = torch.empty_like(covs)
    lib().orc_gaussian_mask_backward(_p(means), _p(covs), _p(volume), _p(volume1_grad), _p(gm), _p(gc),
                                     _i(E), _i(H1), _i(W1), _i(H2), _i(W2), _i(radius))
    return [gm, gc]


def lowMem_defSample(fmap1, fmap2, coords, offset, radius, strict_ref=True):
    B, H1, W1, C = fmap1.shape
    _, H2, W2, _ = fmap2.shape
    N = coords.shape[1]
    rd = 2 * radius + 1
    out = torch.empty(B, N, rd, rd, H1, W1, dtype=torch.float32)
    lib().orc_lowmem_defsample_forward(_p(fmap1), _p(fmap2), _p(coords), _p(offset), _p(out),
                                       _i(B), _i(N), _i(H1), _i(W1), _i(H2), _i(W2), _i(C), _i(radius),
                                       _i(1 if strict_ref else 0))
    return [out]


def altcorr_forward(fmap1, fmap2, coords, radius):
    B, H1, W1, C = fmap1.shape
    _, H2, W2, _ = fmap2.shape
    N = coords.shape[1]
    rd = 2 * radius + 1
    out = torch.empty(B, N, rd * rd, H1, W1, dtype=torch.float32)
    lib().orc_altcorr_forward(_p(fmap1), _p(fmap2), _p(coords), _p(out),
                              _i(B), _i(N), _i(H1), _i(W1), _i(H2), _i(W2), _i(C), _i(radius))
    return [out]


# ---------------------------------------------------------------- compositions (corr.py / gaussianMask_cuda.py)
def corr_volume(f1, f2, round_half=False):
    """CorrBlock.corr (corr.py:144-152): f1 [E,C,H,W], f2 [E,C,H,W] -> [E,H,W,H,W] fp32."""
    E, C, H, W = f1.shape
    _, _, Hb, Wb = f2.shape
    out = torch.empty(E, H, W, Hb, Wb, dtype=torch.float32)
    lib().orc_corr_volume(_p(f1.float().contiguous()), _p(f2.float().contiguous()), _p(out),
                          _i(E), _i(C), _i(H * W), _i(Hb * Wb), _i(1 if round_half else 0))
    return out


def avg_pool2x2(vol):
    """F.avg_pool2d(corr, 2, stride=2) on the trailing two axes (corr.py:86)."""
    *lead, H2, W2 = vol.shape
    R = 1
    for d in lead:
        R *= d
    out = torch.empty(*lead, H2 // 2, W2 // 2, dtype=torch.float32)
    lib().orc_avg_pool2x2(_p(vol), _p(out), _l(R), _i(H2), _i(W2))
    return out


def gaussian_residual(masked, den, volume):
    """corr1/denominator + corr (gaussianMask_cuda.py:85-86); den [E,H1,W1] fp32."""
    E, H1, W1, H2, W2 = volume.shape
    out = torch.empty_like(volume)
    lib().orc_gaussian_residual(_p(masked), _p(den.contiguous()), _p(volume), _p(out), _l(E * H1 * W1), _l(H2 * W2))
    return out


def gaussian_den(covs):
    """6.28 * sqrt(cov_x * cov_y) with fp32 tensor ops (gaussianMask_cuda.py:77,85)."""
    det = covs[..., 0] * covs[..., 1]
    return (6.28 * torch.sqrt(det)).contiguous()


def build_pyramid(f1, f2, means, covs, num_levels=4, gauss_radius=4, round_half=False):
    """CorrBlock.__init__ minus the offset convs (corr.py:61-86):
    volume -> Gaussian mask + residual -> num_levels-level average pyramid."""
    vol = corr_volume(f1, f2, round_half)
    masked, = gaussianMask(means, covs, vol, gauss_radius)
    cur = gaussian_residual(masked, gaussian_den(covs), vol)
    pyr = []
    for _ in range(num_levels):
        pyr.append(cur)
        cur = avg_pool2x2(cur)
    return pyr


def corr_block_lookup(pyramid, coords, offsets, radius=3):
    """CorrBlock.__call__ (corr.py:88-109).  coords [E,H,W,2]; offsets: list of 4 tensors
    [E,H,W,98]; offsets[1] is replaced (re-multiplied by the mask, quirk Q7) and the centre
    taps of every level are zeroed in place (Q5).  Returns corr [E, 4*49, H, W]."""
    E, H, W, _ = coords.shape
    rd = 2 * radius + 1
    c = coords.permute(0, 3, 1, 2).contiguous()
    m, = corr_index_forward(pyramid[1], (c / 2).contiguous(), 1)
    m = m.permute(0, 3, 4, 1, 2)
    mask = torch.sigmoid(torch.var(m, dim=[3, 4])).view(E, H, W, 1)
    offsets[1] = offsets[1] * mask
    outs = []
    for i in range(len(pyramid)):
        o = offsets[i].contiguous().view(E, H, W, rd, rd, 2)
        out, = defCorr_index_forward(pyramid[i], (c / 2 ** i).contiguous(), o, radius)
        offsets[i] = o.view(E, H, W, rd * rd * 2)
        outs.append(out.view(E, rd * rd, H, W))
    return torch.cat(outs, dim=1)
