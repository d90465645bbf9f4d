"""Operator layer: the reference's `defCorrSample` operators (and `droid_backends.altcorr_forward`)
on torch CUDA tensors, implemented by the sm_100a C-ABI library.

Signatures, argument meaning, in-place side effects and error behaviour follow
/root/reference/offersample_LGS/droid.cpp:53-136 (each returns a list of tensors that callers
unpack with `x, = op(...)`); the only checks the reference performs are `is_contiguous`
(droid.cpp:48-49) and the accessor's dtype/ndim checks, which surface as RuntimeError -- same here,
plus explicit device/shape checks.  torch is used only for allocation and the current stream.
"""
import ctypes

import torch

from . import _lib


def _chk(t, name, ndim, dtype=torch.float32):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (this build has no CPU path)")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")          # droid.cpp:48
    if t.dtype != dtype:
        raise RuntimeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if t.dim() != ndim:
        raise RuntimeError(f"{name}: expected {ndim} dimensions, got {t.dim()}")


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _i(v):
    return ctypes.c_int(int(v))


def corr_index_forward(volume, coords, radius):
    """droid.cpp:79-87.  volume [E,H1,W1,H2,W2], coords [E,2,H1,W1] -> [corr [E,rd,rd,H1,W1]]."""
    _chk(volume, "volume", 5); _chk(coords, "coords", 4)
    E, H1, W1, H2, W2 = volume.shape
    if tuple(coords.shape) != (E, 2, H1, W1):
        raise RuntimeError(f"coords shape {tuple(coords.shape)} != {(E, 2, H1, W1)}")
    rd = 2 * int(radius) + 1
    corr = torch.empty(E, rd, rd, H1, W1, dtype=volume.dtype, device=volume.device)
    with torch.cuda.device(volume.device):
        st = _lib.lib().lgu_corr_index_forward(_p(volume), _p(coords), _p(corr), _i(E), _i(H1), _i(W1), _i(H2),
                                               _i(W2), _i(radius), _stream(volume))
    _lib.check(st, "corr_index_forward")
    return [corr]


def corr_index_backward(volume, coords, corr_grad, radius):
    """droid.cpp:89-99 -> [volume_grad] (dense)."""
    _chk(volume, "volume", 5); _chk(coords, "coords", 4); _chk(corr_grad, "corr_grad", 5)
    E, H1, W1, H2, W2 = volume.shape
    rd = 2 * int(radius) + 1
    if tuple(corr_grad.shape) != (E, rd, rd, H1, W1):
        raise RuntimeError(f"corr_grad shape {tuple(corr_grad.shape)} != {(E, rd, rd, H1, W1)}")
    volume_grad = torch.empty_like(volume)
    with torch.cuda.device(volume.device):
        st = _lib.lib().lgu_corr_index_backward(_p(coords), _p(corr_grad), _p(volume_grad), _i(E), _i(H1), _i(W1),
                                                _i(H2), _i(W2), _i(radius), _stream(volume))
    _lib.check(st, "corr_index_backward")
    return [volume_grad]


def defCorr_index_forward(volume, coords, offset, radius):
    """droid.cpp:53-63.  offset [E,H1,W1,rd,rd,2] is mutated (centre tap zeroed)."""
    _chk(volume, "volume", 5); _chk(coords, "coords", 4); _chk(offset, "offset", 6)
    E, H1, W1, H2, W2 = volume.shape
    rd = 2 * int(radius) + 1
    if tuple(offset.shape) != (E, H1, W1, rd, rd, 2):
        raise RuntimeError(f"offset shape {tuple(offset.shape)} != {(E, H1, W1, rd, rd, 2)}")
    if tuple(coords.shape) != (E, 2, H1, W1):
        raise RuntimeError(f"coords shape {tuple(coords.shape)} != {(E, 2, H1, W1)}")
    corr = torch.empty(E, rd, rd, H1, W1, dtype=volume.dtype, device=volume.device)
    with torch.cuda.device(volume.device):
        st = _lib.lib().lgu_defcorr_index_forward(_p(volume), _p(coords), _p(offset), _p(corr), _i(E), _i(H1),
                                                  _i(W1), _i(H2), _i(W2), _i(radius), _stream(volume))
    _lib.check(st, "defCorr_index_forward")
    return [corr]


def defCorr_index_backward(volume, coords, offset, corr_grad, radius):
    """droid.cpp:65-77 -> [volume_grad, offset_grad]."""
    _chk(volume, "volume", 5); _chk(coords, "coords", 4); _chk(offset, "offset", 6); _chk(corr_grad, "corr_grad", 5)
    E, H1, W1, H2, W2 = volume.shape
    rd = 2 * int(radius) + 1
    if tuple(offset.shape) != (E, H1, W1, rd, rd, 2):
        raise RuntimeError(f"offset shape {tuple(offset.shape)} != {(E, H1, W1, rd, rd, 2)}")
    if tuple(corr_grad.shape) != (E, rd, rd, H1, W1):
        raise RuntimeError(f"corr_grad shape {tuple(corr_grad.shape)} != {(E, rd, rd, H1, W1)}")
    volume_grad = torch.empty_like(volume)
    offset_grad = torch.empty_like(offset)
    with torch.cuda.device(volume.device):
        st = _lib.lib().lgu_defcorr_index_backward(_p(volume), _p(coords), _p(offset), _p(corr_grad),
                                                   _p(volume_grad), _p(offset_grad), _i(E), _i(H1), _i(W1), _i(H2),
                                                   _i(W2), _i(radius), _stream(volume))
    _lib.check(st, "defCorr_index_backward")
    return [volume_grad, offset_grad]


def gaussianMask(means, covs, volume, radius):
    """droid.cpp:100-110.  means, covs [E,H1,W1,2] -> [volume1] (dense, zero outside the window)."""
    _chk(means, "means", 4); _chk(covs, "covs", 4); _chk(volume, "volume", 5)
    E, H1, W1, H2, W2 = volume.shape
    if tuple(means.shape) != (E, H1, W1, 2) or tuple(covs.shape) != (E, H1, W1, 2):
        raise RuntimeError("means/covs must be [E,H1,W1,2]")
    out = torch.empty_like(volume)
    with torch.cuda.device(volume.device):
        st = _lib.lib().lgu_gaussian_mask_forward(_p(means), _p(covs), _p(volume), _p(out), _i(E), _i(H1), _i(W1),
                                                  _i(H2), _i(W2), _i(radius), _stream(volume))
    _lib.check(st, "gaussianMask")
    return [out]


def gaussianMask_backward(means, covs, volume, volume1_grad, radius):
    """droid.cpp:112-123 -> [means_grad, covs_grad]."""
    _chk(means, "means", 4); _chk(covs, "covs", 4); _chk(volume, "volume", 5); _chk(volume1_grad, "volume_grad", 5)
    E, H1, W1, H2, W2 = volume.shape
    if tuple(means.shape) != (E, H1, W1, 2) or tuple(covs.shape) != (E, H1, W1, 2):
        raise RuntimeError("means/covs must be [E,H1,W1,2]")
    if volume1_grad.shape != volume.shape:
        raise RuntimeError("volume_grad must have the volume's shape")
    gm = torch.empty_like(means)
    gc = torch.empty_like(covs)
    with torch.cuda.device(volume.device):
        st = _lib.lib().lgu_gaussian_mask_backward(_p(means), _p(covs), _p(volume), _p(volume1_grad), _p(gm), _p(gc),
                                                   _i(E), _i(H1), _i(W1), _i(H2), _i(W2), _i(radius),
                                                   _stream(volume))
    _lib.check(st, "gaussianMask_backward")
    return [gm, gc]


def lowMem_defSample(fmap1, fmap2, coords, offset, radius, strict_ref=True):
    """droid.cpp:124-136.  fmap1 [B,H1,W1,C], fmap2 [B,H2,W2,C], coords [B,N,H1,W1,2],
    offset [slabs,H1,W1,rd,rd,2] (mutated) -> [corr [B,N,rd,rd,H1,W1]].
    strict_ref=True keeps the reference's offset[b*n] slab indexing (quirk Q2)."""
    _chk(fmap1, "fmap1", 4); _chk(fmap2, "fmap2", 4); _chk(coords, "coords", 5); _chk(offset, "offset", 6)
    B, H1, W1, C = fmap1.shape
    B2, H2, W2, C2 = fmap2.shape
    Bc, N, Hc, Wc, two = coords.shape
    rd = 2 * int(radius) + 1
    if B2 != B or C2 != C or Bc != B or (Hc, Wc, two) != (H1, W1, 2):
        raise RuntimeError("lowMem_defSample: inconsistent fmap/coords shapes")
    need = ((B - 1) * (N - 1) + 1) if strict_ref else B * N
    if tuple(offset.shape[1:]) != (H1, W1, rd, rd, 2) or offset.shape[0] < need:
        raise RuntimeError(f"offset shape {tuple(offset.shape)} incompatible with {(need, H1, W1, rd, rd, 2)}")
    corr = torch.empty(B, N, rd, rd, H1, W1, dtype=fmap1.dtype, device=fmap1.device)
    with torch.cuda.device(fmap1.device):
        st = _lib.lib().lgu_lowmem_defsample_forward(_p(fmap1), _p(fmap2), _p(coords), _p(offset), _p(corr), _i(B),
                                                     _i(N), _i(H1), _i(W1), _i(H2), _i(W2), _i(C), _i(radius),
                                                     _i(1 if strict_ref else 0), _stream(fmap1))
    _lib.check(st, "lowMem_defSample")
    return [corr]


def altcorr_forward(fmap1, fmap2, coords, radius):
    """droid_backends.altcorr_forward (src/droid.cpp:193-203) -> [corr [B,N,rd*rd,H1,W1]]."""
    _chk(fmap1, "fmap1", 4); _chk(fmap2, "fmap2", 4); _chk(coords, "coords", 5)
    B, H1, W1, C = fmap1.shape
    B2, H2, W2, C2 = fmap2.shape
    Bc, N, Hc, Wc, two = coords.shape
    rd = 2 * int(radius) + 1
    if B2 != B or C2 != C or Bc != B or (Hc, Wc, two) != (H1, W1, 2):
        raise RuntimeError("altcorr_forward: inconsistent fmap/coords shapes")
    corr = torch.empty(B, N, rd * rd, H1, W1, dtype=fmap1.dtype, device=fmap1.device)
    with torch.cuda.device(fmap1.device):
        st = _lib.lib().lgu_altcorr_forward(_p(fmap1), _p(fmap2), _p(coords), _p(corr), _i(B), _i(N), _i(H1), _i(W1),
                                            _i(H2), _i(W2), _i(C), _i(radius), _stream(fmap1))
    _lib.check(st, "altcorr_forward")
    return [corr]
