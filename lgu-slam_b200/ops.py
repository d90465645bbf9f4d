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


# Workspace of the tensor-core lowMem / altcorr path: one grow-only byte buffer per device, reused by every call (the
# C ABI allocates nothing).  A call whose volume would not fit LOWMEM_WORKSPACE_CAP is processed in sub-batches of edges.
LOWMEM_WORKSPACE_CAP = 32 << 30
_workspaces = {}


def _workspace(device, nbytes):
    buf = _workspaces.get(device)
    if buf is None or buf.numel() < nbytes:
        _workspaces[device] = None                      # release the old buffer before growing
        buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _workspaces[device] = buf
    return buf


def release_workspace():
    """Drop the cached tensor-core workspaces (they are re-created on demand)."""
    _workspaces.clear()


def _ws_bytes(B, N, H1, W1, H2, W2, C, radius):
    L = _lib.lib()
    L.lgu_lowmem_workspace_bytes.restype = ctypes.c_longlong
    return int(L.lgu_lowmem_workspace_bytes(_i(B), _i(N), _i(H1), _i(W1), _i(H2), _i(W2), _i(C), _i(radius)))


def _sub_batches(B, per_call_bytes_fn):
    """Split B edges into runs whose workspace fits LOWMEM_WORKSPACE_CAP (usually one run)."""
    step = B
    while step > 1 and per_call_bytes_fn(step) > LOWMEM_WORKSPACE_CAP:
        step = (step + 1) // 2
    return [(s0, min(B, s0 + step)) for s0 in range(0, B, step)]


def lowMem_defSample(fmap1, fmap2, coords, offset, radius, strict_ref=True, tensor_cores=True):
    """droid.cpp:124-136.  fmap1 [B,H1,W1,C], fmap2 [B,H2,W2,C], coords [B,N,H1,W1,2],
    offset [slabs,H1,W1,rd,rd,2] (mutated) -> [corr [B,N,rd,rd,H1,W1]].
    strict_ref=True keeps the reference's offset[b*n] slab indexing (quirk Q2).
    tensor_cores (default): for the reference's call shape (N = 1, C = 128, r = 3, H1*W1 % 128 == 0, W2 % 4 == 0) this
    level's [B,P,Q] volume is built on tcgen05 into a cached workspace and sampled with the TMA-staged lookup
    (lgu_lowmem_defsample_forward_ws); other shapes, or tensor_cores=False, run the on-the-fly SIMT kernel."""
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
    if tensor_cores and B > 0 and _ws_bytes(B, N, H1, W1, H2, W2, C, radius) > 0:
        runs = _sub_batches(B, lambda b: _ws_bytes(b, N, H1, W1, H2, W2, C, radius))
        with torch.cuda.device(fmap1.device):
            for b0, b1 in runs:
                nb = _ws_bytes(b1 - b0, N, H1, W1, H2, W2, C, radius)
                ws = _workspace(fmap1.device, nb)
                off = offset if strict_ref else offset[b0:]          # strict: every run reads slab 0 of the call (Q2)
                st = _lib.lib().lgu_lowmem_defsample_forward_ws(
                    _p(fmap1[b0:b1]), _p(fmap2[b0:b1]), _p(coords[b0:b1]), _p(off), _p(corr[b0:b1]), _i(b1 - b0), _i(N),
                    _i(H1), _i(W1), _i(H2), _i(W2), _i(C), _i(radius), _i(1 if strict_ref else 0), _p(ws),
                    ctypes.c_longlong(ws.numel()), _stream(fmap1))
                _lib.check(st, "lowMem_defSample (tensor cores)")
        return [corr]
    with torch.cuda.device(fmap1.device):
        st = _lib.lib().lgu_lowmem_defsample_forward(_p(fmap1), _p(fmap2), _p(coords), _p(offset), _p(corr), _i(B),
                                                     _i(N), _i(H1), _i(W1), _i(H2), _i(W2), _i(C), _i(radius),
                                                     _i(1 if strict_ref else 0), _stream(fmap1))
    _lib.check(st, "lowMem_defSample")
    return [corr]


def altcorr_forward(fmap1, fmap2, coords, radius, tensor_cores=True):
    """droid_backends.altcorr_forward (src/droid.cpp:193-203) -> [corr [B,N,rd*rd,H1,W1]].
    tensor_cores (default): for N = 1, C = 128, r = 1 (the reference's only call, corr.py:202) the volume is built on
    tcgen05 and sampled with per-corner gating (lgu_altcorr_forward_ws); otherwise the on-the-fly SIMT kernel."""
    _chk(fmap1, "fmap1", 4); _chk(fmap2, "fmap2", 4); _chk(coords, "coords", 5)
    B, H1, W1, C = fmap1.shape
    B2, H2, W2, C2 = fmap2.shape
    Bc, N, Hc, Wc, two = coords.shape
    rd = 2 * int(radius) + 1
    if B2 != B or C2 != C or Bc != B or (Hc, Wc, two) != (H1, W1, 2):
        raise RuntimeError("altcorr_forward: inconsistent fmap/coords shapes")
    corr = torch.empty(B, N, rd * rd, H1, W1, dtype=fmap1.dtype, device=fmap1.device)
    if tensor_cores and B > 0 and _ws_bytes(B, N, H1, W1, H2, W2, C, radius) > 0:
        runs = _sub_batches(B, lambda b: _ws_bytes(b, N, H1, W1, H2, W2, C, radius))
        with torch.cuda.device(fmap1.device):
            for b0, b1 in runs:
                ws = _workspace(fmap1.device, _ws_bytes(b1 - b0, N, H1, W1, H2, W2, C, radius))
                st = _lib.lib().lgu_altcorr_forward_ws(
                    _p(fmap1[b0:b1]), _p(fmap2[b0:b1]), _p(coords[b0:b1]), _p(corr[b0:b1]), _i(b1 - b0), _i(N), _i(H1),
                    _i(W1), _i(H2), _i(W2), _i(C), _i(radius), _p(ws), ctypes.c_longlong(ws.numel()), _stream(fmap1))
                _lib.check(st, "altcorr_forward (tensor cores)")
        return [corr]
    with torch.cuda.device(fmap1.device):
        st = _lib.lib().lgu_altcorr_forward(_p(fmap1), _p(fmap2), _p(coords), _p(corr), _i(B), _i(N), _i(H1), _i(W1),
                                            _i(H2), _i(W2), _i(C), _i(radius), _stream(fmap1))
    _lib.check(st, "altcorr_forward")
    return [corr]


# ----------------------------------------------------------------------------------------------
# Fused ops (no single reference operator; they replace sequences of the reference's Python glue)
# ----------------------------------------------------------------------------------------------
def offset_heads(c0, c1, eps=1e-5):
    """The offset heads' glue (corr.py:117-135 / 217-235) in two launches: c0 [E,CH,H,W] = ofsMap(t), c1 [E,CH,H/2,W/2] =
    ofs_residual(avg_pool2d(t, 2)) -> (off0, off1) [E,H,W,CH], channels-last and contiguous:
    off0 = 4 tanh(norm(c0)), off1 = (4 tanh(norm(upsample(c1))) + off0) / 2, norm = per-edge standardisation over
    (CH,H,W) with biased variance + eps (per_Corr_Normalization, corr.py:44-51).  Forward only."""
    _chk(c0, "c0", 4); _chk(c1, "c1", 4)
    E, CH, H, W = c0.shape
    if tuple(c1.shape) != (E, CH, H // 2, W // 2):
        raise RuntimeError(f"c1 shape {tuple(c1.shape)} != {(E, CH, H // 2, W // 2)}")
    off0 = torch.empty(E, H, W, CH, dtype=torch.float32, device=c0.device)
    off1 = torch.empty_like(off0)
    stats = torch.empty(E, 4, dtype=torch.float32, device=c0.device)
    with torch.cuda.device(c0.device):
        st = _lib.lib().lgu_offset_heads(_p(c0), _p(c1), _p(off0), _p(off1), _p(stats), _i(E), _i(CH), _i(H), _i(W),
                                         ctypes.c_float(eps), _stream(c0))
    _lib.check(st, "offset_heads")
    return off0, off1


def pack_fmaps(fmaps, split=False):
    """fmaps [T,C,H,W] (fp32 or fp16 CUDA, as the encoders emit) -> channels-last fp16 planes for the
    tcgen05 build: hi [T,H*W,C] = fp16(x/4) and, if split, lo = fp16(x/4 - hi) (else None).
    The /4 is CorrBlock.corr's pre-scale (corr.py:148-149)."""
    if not (isinstance(fmaps, torch.Tensor) and fmaps.is_cuda and fmaps.dim() == 4 and fmaps.is_contiguous()):
        raise RuntimeError("fmaps must be a contiguous CUDA tensor [T,C,H,W]")
    if fmaps.dtype not in (torch.float32, torch.float16):
        raise RuntimeError(f"fmaps: expected float32 or float16, got {fmaps.dtype}")
    T, C, H, W = fmaps.shape
    hi = torch.empty(T, H * W, C, dtype=torch.float16, device=fmaps.device)
    lo = torch.empty_like(hi) if split else None
    with torch.cuda.device(fmaps.device):
        st = _lib.lib().lgu_pack_fmaps(_p(fmaps), _i(fmaps.dtype == torch.float16), _p(hi),
                                       _p(lo) if split else ctypes.c_void_p(0), _i(T), _i(C), _i(H * W),
                                       _stream(fmaps))
    _lib.check(st, "pack_fmaps")
    return hi, lo


def build_pyramid(hi, lo, ii, jj, H, W, means=None, covs=None, den=None, num_levels=4, gauss_radius=4,
                  precision=None, round_half=False, out=None, out_slots=None):
    """CorrBlock.__init__'s data path (corr.py:61-86) in one tcgen05/TMA kernel: all-pairs volume of edge
    (ii[e] -> jj[e]) + Gaussian residual (gaussianMask_cuda.py:84-86) + num_levels-level average pyramid.
    hi/lo from pack_fmaps; ii, jj int32 [E]; means, covs [E,H,W,2]; den [E,H,W] = 6.28*sqrt(cov_x*cov_y).
    Returns [lvl0 [E,H,W,H,W], lvl1 [E,H,W,H/2,W/2], ...].
    Edge-slot pool: out = the pool's level storages [S,H,W,H>>l,W>>l] and out_slots = int32 [E] slot of each edge;
    the levels are then written in place (nothing is allocated) and `out` is returned."""
    if hi.dtype != torch.float16 or not hi.is_cuda or hi.dim() != 3 or not hi.is_contiguous():
        raise RuntimeError("hi must be a contiguous CUDA fp16 tensor [T,H*W,C] (see pack_fmaps)")
    T, P, C = hi.shape
    if P != H * W:
        raise RuntimeError(f"hi has {P} pixels per map, expected H*W = {H * W}")
    if precision is None:
        precision = 2 if lo is not None else 1
    if precision == 2 and (lo is None or lo.shape != hi.shape or lo.dtype != torch.float16):
        raise RuntimeError("precision 2 needs the lo plane from pack_fmaps(split=True)")
    for t, name in ((ii, "ii"), (jj, "jj")):
        if not (t.is_cuda and t.dtype == torch.int32 and t.dim() == 1 and t.is_contiguous()):
            raise RuntimeError(f"{name} must be a contiguous CUDA int32 vector")
    E = ii.numel()
    if jj.numel() != E:
        raise RuntimeError("ii and jj must have the same length")
    use_gauss = gauss_radius > 0 and means is not None
    if use_gauss:
        _chk(means, "means", 4); _chk(covs, "covs", 4)
        if tuple(means.shape) != (E, H, W, 2) or tuple(covs.shape) != (E, H, W, 2):
            raise RuntimeError("means/covs must be [E,H,W,2]")
        if den is not None:                      # None: the kernel forms 6.28*sqrt(cov_x*cov_y) in its prologue
            _chk(den, "den", 3)
            if tuple(den.shape) != (E, H, W):
                raise RuntimeError("den must be [E,H,W]")
    if not 1 <= num_levels <= 4:
        raise RuntimeError("num_levels must be in 1..4")
    null = ctypes.c_void_p(0)
    if out is not None:
        if out_slots is None or not (out_slots.is_cuda and out_slots.dtype == torch.int32 and out_slots.numel() == E
                                     and out_slots.is_contiguous()):
            raise RuntimeError("out_slots must be a contiguous CUDA int32 vector with one slot per edge")
        S = out[0].shape[0]
        for l, t in enumerate(out[:num_levels]):
            _chk(t, f"out[{l}]", 5)
            if tuple(t.shape) != (S, H, W, H >> l, W >> l):
                raise RuntimeError(f"out[{l}] shape {tuple(t.shape)} != {(S, H, W, H >> l, W >> l)}")
        lv = list(out[:num_levels])
        ptr = [_p(t) for t in lv] + [null] * (4 - num_levels)
        with torch.cuda.device(hi.device):
            st = _lib.lib().lgu_build_pyramid_slots(
                _p(hi), _p(lo) if precision == 2 else null, _p(ii), _p(jj), _p(means) if use_gauss else null,
                _p(covs) if use_gauss else null, _p(den) if (use_gauss and den is not None) else null, ptr[0], ptr[1], ptr[2],
                ptr[3], _p(out_slots), _i(S), _i(T), _i(E), _i(H), _i(W), _i(C), _i(gauss_radius if use_gauss else 0),
                _i(precision), _i(1 if round_half else 0), _stream(hi))
        _lib.check(st, "build_pyramid (slots)")
        return lv
    lv = [torch.empty(E, H, W, H >> l, W >> l, dtype=torch.float32, device=hi.device) for l in range(num_levels)]
    ptr = [_p(t) for t in lv] + [null] * (4 - num_levels)
    with torch.cuda.device(hi.device):
        st = _lib.lib().lgu_build_pyramid(_p(hi), _p(lo) if precision == 2 else null, _p(ii), _p(jj),
                                          _p(means) if use_gauss else null, _p(covs) if use_gauss else null,
                                          _p(den) if (use_gauss and den is not None) else null, ptr[0], ptr[1], ptr[2], ptr[3],
                                          _i(T), _i(E), _i(H), _i(W), _i(C), _i(gauss_radius if use_gauss else 0),
                                          _i(precision), _i(1 if round_half else 0), _stream(hi))
    _lib.check(st, "build_pyramid")
    return lv


def corr_lookup_fused(pyramid, coords, off0, off1, radius=3, return_mask=False, slots=None, cum_mask=None):
    """CorrBlock.__call__'s data path (corr.py:88-109) in one TMA-staged kernel.
    pyramid: 4 tensors [E,H,W,H>>l,W>>l]; coords [E,H,W,2] (x,y; level-0 units); off0, off1 [E,H,W,98] or
    [E,H,W,7,7,2]; off1 is mutated in place (multiplied by this call's uncertainty mask, Q7); the centre taps are
    read as 0 (Q5) but left untouched in memory.  Returns corr [E,196,H,W] (and the mask [E,H,W] if return_mask).
    cum_mask [S,H,W] (fp32, initialised to 1 by the caller): off1 is then READ ONLY and stays pristine; the running
    product of all masks so far lives in cum_mask (updated in place) and is applied in registers -- the same
    cumulative semantics without writing 392 B per pixel back (lgu_corr_lookup_fused_cum)."""
    if len(pyramid) != 4:
        raise RuntimeError("corr_lookup_fused needs a 4-level pyramid")
    S, H, W = pyramid[0].shape[:3]                    # S = storage slots (== edges without a pool)
    for l, t in enumerate(pyramid):
        _chk(t, f"pyramid[{l}]", 5)
        if tuple(t.shape) != (S, H, W, H >> l, W >> l):
            raise RuntimeError(f"pyramid[{l}] shape {tuple(t.shape)} != {(S, H, W, H >> l, W >> l)}")
    _chk(coords, "coords", 4)
    E = coords.shape[0] if slots is not None else S
    if slots is not None and not (slots.is_cuda and slots.dtype == torch.int32 and slots.is_contiguous()
                                  and slots.numel() == E):
        raise RuntimeError("slots must be a contiguous CUDA int32 vector with one slot per edge of coords")
    if tuple(coords.shape) != (E, H, W, 2):
        raise RuntimeError(f"coords shape {tuple(coords.shape)} != {(E, H, W, 2)}")
    rd = 2 * int(radius) + 1
    for name, o in (("off0", off0), ("off1", off1)):
        if not (isinstance(o, torch.Tensor) and o.is_cuda and o.is_contiguous() and o.dtype == torch.float32):
            raise RuntimeError(f"{name} must be a contiguous fp32 CUDA tensor")
        if o.numel() != S * H * W * rd * rd * 2:
            raise RuntimeError(f"{name} has {o.numel()} elements, expected {S * H * W * rd * rd * 2}")
    corr = torch.empty(E, 4 * rd * rd, H, W, dtype=torch.float32, device=coords.device)
    mask = torch.empty(E, H, W, dtype=torch.float32, device=coords.device) if return_mask else None
    if cum_mask is not None:
        _chk(cum_mask, "cum_mask", 3)
        if tuple(cum_mask.shape) != (S, H, W):
            raise RuntimeError(f"cum_mask shape {tuple(cum_mask.shape)} != {(S, H, W)}")
    with torch.cuda.device(coords.device):
        if cum_mask is not None:
            st = _lib.lib().lgu_corr_lookup_fused_cum(_p(pyramid[0]), _p(pyramid[1]), _p(pyramid[2]), _p(pyramid[3]),
                                                      _p(coords), _p(off0), _p(off1), _p(cum_mask), _p(corr),
                                                      _p(mask) if return_mask else ctypes.c_void_p(0),
                                                      _p(slots) if slots is not None else ctypes.c_void_p(0),
                                                      _i(S), _i(E), _i(H), _i(W), _i(4), _i(radius), _stream(coords))
        elif slots is not None:
            st = _lib.lib().lgu_corr_lookup_fused_slots(_p(pyramid[0]), _p(pyramid[1]), _p(pyramid[2]), _p(pyramid[3]),
                                                        _p(coords), _p(off0), _p(off1), _p(corr),
                                                        _p(mask) if return_mask else ctypes.c_void_p(0), _p(slots),
                                                        _i(S), _i(E), _i(H), _i(W), _i(4), _i(radius), _stream(coords))
        else:
            st = _lib.lib().lgu_corr_lookup_fused(_p(pyramid[0]), _p(pyramid[1]), _p(pyramid[2]), _p(pyramid[3]),
                                                  _p(coords), _p(off0), _p(off1), _p(corr),
                                                  _p(mask) if return_mask else ctypes.c_void_p(0),
                                                  _i(E), _i(H), _i(W), _i(4), _i(radius), _stream(coords))
    _lib.check(st, "corr_lookup_fused")
    return (corr, mask) if return_mask else corr


def pack_conv1x1(weight):
    """corr_encoder[0].weight [128,196] (or [128,196,1,1]) -> the tensor-core operand fragments of the fused encoder
    epilogue (TF32 hi/lo split, m16n8k8 A-fragment order): a float32 tensor of 8*25*32*8 elements."""
    if not (isinstance(weight, torch.Tensor) and weight.is_cuda and weight.dtype == torch.float32):
        raise RuntimeError("weight must be a CUDA fp32 tensor")
    w = weight.detach().reshape(weight.shape[0], -1).contiguous()
    if tuple(w.shape) != (128, 196):
        raise RuntimeError(f"only the 196 -> 128 corr_encoder is implemented, got weight {tuple(weight.shape)}")
    frag = torch.empty(8 * 25 * 32 * 8, dtype=torch.float32, device=w.device)
    with torch.cuda.device(w.device):
        st = _lib.lib().lgu_pack_conv1x1(_p(w), _p(frag), _i(128), _i(196), _stream(w))
    _lib.check(st, "pack_conv1x1")
    return frag


def corr_lookup_fused_enc(pyramid, coords, off0, off1, cum_mask, wfrag, bias=None, relu=True, keep_corr=False,
                          enc_half=False, slots=None):
    """corr_lookup_fused (cumulative-mask form) with the consumer's first layer folded in: UpdateModule.corr_encoder[0:2] =
    Conv2d(196,128,1) + ReLU (droid_net.py:74-76,115) evaluated on the staged tile with TF32 tensor-core MMAs (3-term
    split, <= 1e-5 of F.conv2d in fp32).  Returns (corr [E,196,H,W] or None, enc [E,128,H,W] fp32 / fp16).
    keep_corr=False: the 196-channel tensor is never written.  wfrag from pack_conv1x1; bias [128] or None."""
    if len(pyramid) != 4:
        raise RuntimeError("corr_lookup_fused_enc needs a 4-level pyramid")
    S, H, W = pyramid[0].shape[:3]
    for l, t in enumerate(pyramid):
        _chk(t, f"pyramid[{l}]", 5)
        if tuple(t.shape) != (S, H, W, H >> l, W >> l):
            raise RuntimeError(f"pyramid[{l}] shape {tuple(t.shape)} != {(S, H, W, H >> l, W >> l)}")
    _chk(coords, "coords", 4); _chk(cum_mask, "cum_mask", 3)
    E = coords.shape[0] if slots is not None else S
    if tuple(coords.shape) != (E, H, W, 2) or tuple(cum_mask.shape) != (S, H, W):
        raise RuntimeError("corr_lookup_fused_enc: inconsistent coords / cum_mask shapes")
    for name, o in (("off0", off0), ("off1", off1)):
        if not (isinstance(o, torch.Tensor) and o.is_cuda and o.is_contiguous() and o.dtype == torch.float32
                and o.numel() == S * H * W * 98):
            raise RuntimeError(f"{name} must be a contiguous fp32 CUDA tensor with {S * H * W * 98} elements")
    if not (wfrag.is_cuda and wfrag.dtype == torch.float32 and wfrag.numel() == 8 * 25 * 32 * 8 and wfrag.is_contiguous()):
        raise RuntimeError("wfrag must come from pack_conv1x1")
    if bias is not None and not (bias.is_cuda and bias.dtype == torch.float32 and bias.numel() == 128 and bias.is_contiguous()):
        raise RuntimeError("bias must be a contiguous CUDA fp32 vector of 128")
    if keep_corr and enc_half:
        raise RuntimeError("enc_half needs keep_corr=False")
    corr = torch.empty(E, 196, H, W, dtype=torch.float32, device=coords.device) if keep_corr else None
    enc = torch.empty(E, 128, H, W, dtype=torch.float16 if enc_half else torch.float32, device=coords.device)
    null = ctypes.c_void_p(0)
    with torch.cuda.device(coords.device):
        st = _lib.lib().lgu_corr_lookup_fused_enc(
            _p(pyramid[0]), _p(pyramid[1]), _p(pyramid[2]), _p(pyramid[3]), _p(coords), _p(off0), _p(off1), _p(cum_mask),
            _p(corr) if keep_corr else null, _p(enc), _p(wfrag), _p(bias) if bias is not None else null,
            _i(1 if relu else 0), _i(1 if enc_half else 0), null, _p(slots) if slots is not None else null, _i(S), _i(E),
            _i(H), _i(W), _i(4), _i(3), _stream(coords))
    _lib.check(st, "corr_lookup_fused_enc")
    return corr, enc


def corr_lookup_fused_backward(pyramid, coords, off0, off1_out, mask, corr_grad, off1_out_grad=None,
                               accumulate_into=None, cum_mask=None, gauss_window_means=None, gauss_head=None):
    """Backward of corr_lookup_fused in one launch (what autograd runs for corr.py:88-109 in training).
    pyramid: the 4 levels (only levels 0-1 are read); off1_out: off1 after the forward; mask [E,H,W] from the
    forward; corr_grad [E,196,H,W]; off1_out_grad: upstream gradient on the post-mask offsets (later calls) or None.
    Returns (gv0, gv1, gv2, gv3, off0_grad, off1_grad) -- dense level gradients, off1_grad w.r.t. the PRE-mask off1.
    accumulate_into: 4 persistent level-gradient buffers (same shapes as the pyramid, zeroed once by the caller); the
    launch ADDS this call's gradient into them, touching only the per-pixel footprints, and returns them as gv0..gv3.
    cum_mask [E,H,W]: the forward ran in the cumulative-mask form -- `off1_out` is then the PRISTINE offset[1] and the
    post-mask offsets are off1_out * cum_mask (cum_mask as left by that forward), formed in registers.
    gauss_window_means [E,H,W,2] (dense cum form only): also return, as a 7th value, the [E,H,W,81] window record of the
    merged level-0 gradient around floor(means) that build_backward_gauss(..., window=) consumes
    (lgu_corr_lookup_fused_backward_win in include/lgu_corr.h).
    gauss_head = (means [E,H,W,2], covs [E,H,W,2], den [E,H,W]) (dense cum form only): the Gaussian head's backward of
    the build is folded into the launch; returns 9 values, the last three being (means_grad, covs_grad, den_grad) =
    build_backward_gauss(means, covs, den, pyramid[0], this call's level gradients, 4)
    (lgu_corr_lookup_fused_backward_gauss)."""
    E, H, W = pyramid[0].shape[:3]
    for l, t in enumerate(pyramid):
        _chk(t, f"pyramid[{l}]", 5)
    _chk(coords, "coords", 4); _chk(mask, "mask", 3); _chk(corr_grad, "corr_grad", 4)
    if tuple(corr_grad.shape) != (E, 196, H, W) or tuple(mask.shape) != (E, H, W) or tuple(coords.shape) != (E, H, W, 2):
        raise RuntimeError("corr_lookup_fused_backward: inconsistent shapes")
    for name, o in (("off0", off0), ("off1_out", off1_out), ("off1_out_grad", off1_out_grad)):
        if o is None and name == "off1_out_grad":
            continue
        if not (isinstance(o, torch.Tensor) and o.is_cuda and o.is_contiguous() and o.dtype == torch.float32
                and o.numel() == E * H * W * 98):
            raise RuntimeError(f"{name} must be a contiguous fp32 CUDA tensor with {E * H * W * 98} elements")
    if accumulate_into is not None:
        gv = list(accumulate_into)
        for l, (a, t) in enumerate(zip(gv, pyramid)):
            _chk(a, f"accumulate_into[{l}]", 5)
            if a.shape != t.shape or a.device != t.device:
                raise RuntimeError(f"accumulate_into[{l}] must match pyramid[{l}]")
        fn = _lib.lib().lgu_corr_lookup_fused_backward_accumulate
    else:
        gv = [torch.empty_like(t) for t in pyramid]
        fn = _lib.lib().lgu_corr_lookup_fused_backward
    g0 = torch.empty(E, H, W, 98, dtype=torch.float32, device=coords.device)
    g1 = torch.empty_like(g0)
    if cum_mask is not None:
        _chk(cum_mask, "cum_mask", 3)
        if tuple(cum_mask.shape) != (E, H, W):
            raise RuntimeError("cum_mask must be [E,H,W]")
        if gauss_head is not None:
            if accumulate_into is not None or gauss_window_means is not None:
                raise RuntimeError("corr_lookup_fused_backward: gauss_head needs the dense form (and excludes the window record)")
            means, covs, den = gauss_head
            _chk(means, "means", 4); _chk(covs, "covs", 4); _chk(den, "den", 3)
            if tuple(means.shape) != (E, H, W, 2) or tuple(covs.shape) != (E, H, W, 2) or tuple(den.shape) != (E, H, W):
                raise RuntimeError("gauss_head must be (means [E,H,W,2], covs [E,H,W,2], den [E,H,W])")
            gm, gc, gd = torch.empty_like(means), torch.empty_like(covs), torch.empty_like(den)
            with torch.cuda.device(coords.device):
                st = _lib.lib().lgu_corr_lookup_fused_backward_gauss(
                    _p(pyramid[0]), _p(pyramid[1]), _p(coords), _p(off0), _p(off1_out), _p(cum_mask), _p(mask),
                    _p(corr_grad), _p(off1_out_grad) if off1_out_grad is not None else ctypes.c_void_p(0),
                    _p(means), _p(covs), _p(den), _p(gv[0]), _p(gv[1]), _p(gv[2]), _p(gv[3]), _p(g0), _p(g1),
                    _p(gm), _p(gc), _p(gd), _i(E), _i(H), _i(W), _i(4), _i(3), _i(4), _stream(coords))
            _lib.check(st, "corr_lookup_fused_backward (gauss)")
            return gv[0], gv[1], gv[2], gv[3], g0, g1, gm, gc, gd
        if gauss_window_means is not None:
            if accumulate_into is not None:
                raise RuntimeError("corr_lookup_fused_backward: the Gaussian window record needs the dense form")
            _chk(gauss_window_means, "gauss_window_means", 4)
            if tuple(gauss_window_means.shape) != (E, H, W, 2):
                raise RuntimeError("gauss_window_means must be [E,H,W,2]")
            gwin = torch.empty(E, H, W, 81, dtype=torch.float32, device=coords.device)
            with torch.cuda.device(coords.device):
                st = _lib.lib().lgu_corr_lookup_fused_backward_win(
                    _p(pyramid[0]), _p(pyramid[1]), _p(coords), _p(off0), _p(off1_out), _p(cum_mask), _p(mask),
                    _p(corr_grad), _p(off1_out_grad) if off1_out_grad is not None else ctypes.c_void_p(0),
                    _p(gauss_window_means), _p(gv[0]), _p(gv[1]), _p(gv[2]), _p(gv[3]), _p(g0), _p(g1), _p(gwin),
                    _i(E), _i(H), _i(W), _i(4), _i(3), _stream(coords))
            _lib.check(st, "corr_lookup_fused_backward (window)")
            return gv[0], gv[1], gv[2], gv[3], g0, g1, gwin
        with torch.cuda.device(coords.device):
            st = _lib.lib().lgu_corr_lookup_fused_backward_cum(
                _p(pyramid[0]), _p(pyramid[1]), _p(coords), _p(off0), _p(off1_out), _p(cum_mask), _p(mask), _p(corr_grad),
                _p(off1_out_grad) if off1_out_grad is not None else ctypes.c_void_p(0),
                _p(gv[0]), _p(gv[1]), _p(gv[2]), _p(gv[3]), _p(g0), _p(g1), _i(E), _i(H), _i(W), _i(4), _i(3),
                _i(1 if accumulate_into is not None else 0), _stream(coords))
        _lib.check(st, "corr_lookup_fused_backward (cum)")
        return gv[0], gv[1], gv[2], gv[3], g0, g1
    if gauss_window_means is not None or gauss_head is not None:
        raise RuntimeError("corr_lookup_fused_backward: gauss_window_means / gauss_head need cum_mask (the cumulative-mask form)")
    with torch.cuda.device(coords.device):
        st = fn(
            _p(pyramid[0]), _p(pyramid[1]), _p(coords), _p(off0), _p(off1_out), _p(mask), _p(corr_grad),
            _p(off1_out_grad) if off1_out_grad is not None else ctypes.c_void_p(0),
            _p(gv[0]), _p(gv[1]), _p(gv[2]), _p(gv[3]), _p(g0), _p(g1), _i(E), _i(H), _i(W), _i(4), _i(3),
            _stream(coords))
    _lib.check(st, "corr_lookup_fused_backward")
    return gv[0], gv[1], gv[2], gv[3], g0, g1


def build_backward_gauss(means, covs, den, lvl0, level_grads, radius, window=None):
    """Gaussian-head gradients of the fused build straight from the four level gradients (no dense pass):
    means, covs [E,H,W,2], den [E,H,W], lvl0 [E,H,W,H,W], level_grads = 4 tensors [E,H,W,H>>l,W>>l] or None.
    window [E,H,W,81]: the record corr_lookup_fused_backward(gauss_window_means=means) returned -- used INSTEAD of
    level_grads (radius 4; same result bit for bit, one contiguous read per pixel).
    Returns (means_grad, covs_grad, den_grad); see lgu_build_backward_gauss in include/lgu_corr.h."""
    _chk(means, "means", 4); _chk(covs, "covs", 4); _chk(den, "den", 3); _chk(lvl0, "lvl0", 5)
    E, H, W = lvl0.shape[:3]
    if tuple(lvl0.shape) != (E, H, W, H, W) or tuple(means.shape) != (E, H, W, 2) or tuple(covs.shape) != (E, H, W, 2) \
            or tuple(den.shape) != (E, H, W):
        raise RuntimeError("build_backward_gauss: inconsistent shapes")
    if window is not None:
        _chk(window, "window", 4)
        if tuple(window.shape) != (E, H, W, 81) or radius != 4:
            raise RuntimeError("build_backward_gauss: window must be [E,H,W,81] and radius 4")
        gm, gc, gd = torch.empty_like(means), torch.empty_like(covs), torch.empty_like(den)
        with torch.cuda.device(lvl0.device):
            st = _lib.lib().lgu_build_backward_gauss_window(_p(means), _p(covs), _p(den), _p(lvl0), _p(window), _p(gm),
                                                            _p(gc), _p(gd), _i(E), _i(H), _i(W), _stream(lvl0))
        _lib.check(st, "build_backward_gauss (window)")
        return gm, gc, gd
    ptrs = []
    for l, g in enumerate(level_grads):
        if g is None:
            ptrs.append(ctypes.c_void_p(0))
            continue
        _chk(g, f"level_grads[{l}]", 5)
        if tuple(g.shape) != (E, H, W, H >> l, W >> l):
            raise RuntimeError(f"level_grads[{l}] must be [E,H,W,{H >> l},{W >> l}]")
        ptrs.append(_p(g))
    gm, gc, gd = torch.empty_like(means), torch.empty_like(covs), torch.empty_like(den)
    with torch.cuda.device(lvl0.device):
        st = _lib.lib().lgu_build_backward_gauss(_p(means), _p(covs), _p(den), _p(lvl0), ptrs[0], ptrs[1], ptrs[2],
                                                 ptrs[3], _p(gm), _p(gc), _p(gd), _i(E), _i(H), _i(W), _i(radius),
                                                 _stream(lvl0))
    _lib.check(st, "build_backward_gauss")
    return gm, gc, gd


def tf32_split(x, scale=1.0):
    """x * scale = hi + lo with hi carrying 10 mantissa bits (exact split; the operand form of build_backward_fmaps)."""
    if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32):
        raise RuntimeError("tf32_split: fp32 CUDA tensor expected")
    x = x.contiguous()
    hi, lo = torch.empty_like(x), torch.empty_like(x)
    with torch.cuda.device(x.device):
        st = _lib.lib().lgu_tf32_split(_p(x), ctypes.c_float(scale), _p(hi), _p(lo), ctypes.c_longlong(x.numel()),
                                       _stream(x))
    _lib.check(st, "tf32_split")
    return hi, lo


def build_backward_fmaps(level_grads, f1, f2):
    """Feature-map gradients of the fused build from the level gradients, on tcgen05 (kind::tf32, 3-term split).
    level_grads: up to 4 tensors [E,H,W,H>>l,W>>l] or None; f1, f2 [E,C,H,W] fp32 (the maps the volume was built from).
    Returns (g_f1, g_f2) [E,C,H,W] fp32: g_f1 = g_V f2 / 16, g_f2 = g_V^T f1 / 16 with g_V = sum_l avg_pool^T(level_grads[l])."""
    import torch.nn.functional as F
    E, C, H, W = f1.shape
    L = len(level_grads)
    f1h, f1l = tf32_split(f1.float().reshape(E, C, H * W), 1.0 / 16.0)
    planes, cur = [], f2.float()
    for l in range(L):
        if l > 0:
            cur = F.avg_pool2d(cur, 2, stride=2)
        planes.append(tf32_split(cur.reshape(E, C, -1), 1.0 / 16.0) if level_grads[l] is not None else (None, None))
    g_f1 = torch.empty(E, C, H * W, dtype=torch.float32, device=f1.device)
    g_f2l = [torch.empty(E, C, (H >> l) * (W >> l), dtype=torch.float32, device=f1.device) if level_grads[l] is not None
             else None for l in range(L)]
    arr = ctypes.c_void_p * L

    def ptrs(ts):
        return arr(*[t.data_ptr() if t is not None else None for t in ts])

    for l, g in enumerate(level_grads):
        if g is not None:
            _chk(g, f"level_grads[{l}]", 5)
            if tuple(g.shape) != (E, H, W, H >> l, W >> l):
                raise RuntimeError(f"level_grads[{l}] must be [E,H,W,{H >> l},{W >> l}]")
    with torch.cuda.device(f1.device):
        st = _lib.lib().lgu_build_backward_fmaps(ptrs(level_grads), _p(f1h), _p(f1l), ptrs([p[0] for p in planes]),
                                                 ptrs([p[1] for p in planes]), _p(g_f1), ptrs(g_f2l), _i(L), _i(E), _i(H),
                                                 _i(W), _i(C), _stream(f1))
    _lib.check(st, "build_backward_fmaps")
    g_f2 = None
    for l, t in enumerate(g_f2l):
        if t is None:
            continue
        t = t.view(E, C, H >> l, W >> l)
        if l > 0:
            k = 1 << l
            t = (t / float(k * k)).repeat_interleave(k, dim=2).repeat_interleave(k, dim=3)
        g_f2 = t if g_f2 is None else g_f2 + t
    return g_f1.view(E, C, H, W), g_f2


def volume_half_mask(coords, level=0):
    """Which 256-column halves of a level's [E,P,Q] volume the fused backend lookup at `coords` [E,H,W,2] can touch
    (offsets bounded by 4): uint32 [E*H*W/128] as int32 storage, for build_volume(..., half_mask=)."""
    _chk(coords, "coords", 4)
    E, H, W, _ = coords.shape
    mask = torch.empty(E * H * W // 128, dtype=torch.int32, device=coords.device)
    with torch.cuda.device(coords.device):
        st = _lib.lib().lgu_volume_half_mask(_p(coords), _p(mask), _i(E), _i(H), _i(W), _i(level), _stream(coords))
    _lib.check(st, "volume_half_mask")
    return mask


def build_boxes(f1_hi, f2_hi, ii, jj, coords, half_mask=None, level=0):
    """Level 0 of the backend path as compact per-pixel boxes (lgu_build_boxes): f1_hi [T1,P,C], f2_hi [T2,P,C] fp16 planes,
    ii, jj int32 [E], coords [E,H,W,2] (W = 64) -> [E,P,16,20] fp32: for every source pixel the 16 x 20 window of its
    correlation slice that altcorr_lookup_fused(..., boxes0=) stages around coords, zeros outside the grid.
    level = 1: f2_hi holds the level-1 pooled maps [T2,P/4,C]; the result goes to altcorr_lookup_fused(..., boxes1=)."""
    _chk(coords, "coords", 4)
    E, H, W, _ = coords.shape
    for t, name in ((f1_hi, "f1_hi"), (f2_hi, "f2_hi")):
        if not (t.is_cuda and t.dtype == torch.float16 and t.dim() == 3 and t.is_contiguous()):
            raise RuntimeError(f"{name} must be a contiguous CUDA fp16 tensor [T,P,C]")
    T1, P, C = f1_hi.shape
    T2, Q, C2 = f2_hi.shape
    if C2 != C or P != H * W or Q != (H >> level) * (W >> level) or level not in (0, 1):
        raise RuntimeError("build_boxes: f1 [T,H*W,C] and f2 [T,(H>>level)*(W>>level),C] expected, level 0 or 1")
    for t, name in ((ii, "ii"), (jj, "jj")):
        if not (t.is_cuda and t.dtype == torch.int32 and t.dim() == 1 and t.is_contiguous() and t.numel() == E):
            raise RuntimeError(f"{name} must be a contiguous CUDA int32 vector with one entry per edge")
    if half_mask is None:
        half_mask = volume_half_mask(coords, level)
    boxes = torch.empty(E, P, 16, 20, dtype=torch.float32, device=coords.device)
    with torch.cuda.device(coords.device):
        st = _lib.lib().lgu_build_boxes(_p(f1_hi), _p(f2_hi), _p(ii), _p(jj), _p(coords), _p(half_mask), _p(boxes), _i(T1),
                                        _i(T2), _i(E), _i(H), _i(W), _i(C), _i(level), _stream(coords))
    _lib.check(st, "build_boxes")
    return boxes


def build_volume(f1_hi, f1_lo, f2_hi, f2_lo, ii, jj, half_mask=None):
    """volume[e,p,q] = sum_c f1[ii[e],p,c] * f2[jj[e],q,c] on tcgen05 (fp32 accumulate).  f1_* [T1,P,C],
    f2_* [T2,Q,C] contiguous CUDA fp16 planes (lo planes None for single-product precision); ii, jj int32 [E].
    Returns [E,P,Q] fp32.  The backend path's per-level volume (source level 0 x pooled target level l).
    half_mask (volume_half_mask): sparse form -- only the named halves are computed; the rest of the result is
    UNINITIALISED memory that the fused backend lookup at those coords never reads."""
    for t, name in ((f1_hi, "f1_hi"), (f2_hi, "f2_hi")):
        if not (t.is_cuda and t.dtype == torch.float16 and t.dim() == 3 and t.is_contiguous()):
            raise RuntimeError(f"{name} must be a contiguous CUDA fp16 tensor [T,P,C]")
    split = f1_lo is not None
    if split and (f2_lo is None or f1_lo.shape != f1_hi.shape or f2_lo.shape != f2_hi.shape):
        raise RuntimeError("precision 2 needs lo planes for both map sets")
    T1, P, C = f1_hi.shape
    T2, Q, C2 = f2_hi.shape
    if C2 != C:
        raise RuntimeError("channel mismatch")
    for t, name in ((ii, "ii"), (jj, "jj")):
        if not (t.is_cuda and t.dtype == torch.int32 and t.dim() == 1 and t.is_contiguous()):
            raise RuntimeError(f"{name} must be a contiguous CUDA int32 vector")
    E = ii.numel()
    vol = torch.empty(E, P, Q, dtype=torch.float32, device=f1_hi.device)
    null = ctypes.c_void_p(0)
    if half_mask is not None:
        if not (half_mask.is_cuda and half_mask.dtype == torch.int32 and half_mask.is_contiguous()
                and half_mask.numel() == E * (P // 128)):
            raise RuntimeError(f"half_mask must be a contiguous CUDA int32 vector of {E * (P // 128)} words")
        with torch.cuda.device(f1_hi.device):
            st = _lib.lib().lgu_build_volume_sparse(_p(f1_hi), _p(f1_lo) if split else null, _p(f2_hi),
                                                    _p(f2_lo) if split else null, _p(ii), _p(jj), _p(half_mask), _p(vol),
                                                    _i(T1), _i(T2), _i(E), _i(P), _i(Q), _i(C), _i(2 if split else 1),
                                                    _stream(f1_hi))
        _lib.check(st, "build_volume (sparse)")
        return vol
    with torch.cuda.device(f1_hi.device):
        st = _lib.lib().lgu_build_volume(_p(f1_hi), _p(f1_lo) if split else null, _p(f2_hi), _p(f2_lo) if split else null,
                                         _p(ii), _p(jj), _p(vol), _i(T1), _i(T2), _i(E), _i(P), _i(Q), _i(C),
                                         _i(2 if split else 1), _stream(f1_hi))
    _lib.check(st, "build_volume")
    return vol


def altcorr_lookup_fused(volumes, coords, off0, off1, radius=3, shared_offsets=False, apply_mask=True,
                         return_mask=False, out=None, out_index=None, boxes0=None, boxes1=None):
    """corr_lookup_fused with the backend samplers' semantics (per-corner gating, quirk Q4; lowMem_defSample.cu /
    altcorr_kernel.cu).  volumes: 4 tensors [E,H,W,H>>l,W>>l] from build_volume.  shared_offsets: every edge reads
    offset slab 0 (quirk Q2; off0/off1 then hold >= 1 slab); apply_mask=False: off1 is used as given.
    out: destination [E_out,196,H,W], fp32 or fp16 (rounded to nearest), possibly peer memory of another GPU mapped
    into this process; out_index int32 [E]: row of `out` that receives edge e (default: e).  Returns `out` then.
    boxes0 [E,H*W,16,20] (build_boxes): level 0 as compact per-pixel boxes; volumes[0] is then ignored (may be None);
    boxes1: the same for level 1 (needs boxes0)."""
    E, H, W = volumes[2].shape[:3]
    if boxes1 is not None and boxes0 is None:
        raise RuntimeError("boxes1 needs boxes0")
    for name, bx in (("boxes0", boxes0), ("boxes1", boxes1)):
        if bx is not None:
            _chk(bx, name, 4)
            if tuple(bx.shape) != (E, H * W, 16, 20):
                raise RuntimeError(f"{name} must be [E,H*W,16,20], got {tuple(bx.shape)}")
    for l, t in enumerate(volumes):
        if (l == 0 and boxes0 is not None) or (l == 1 and boxes1 is not None):
            continue
        _chk(t, f"volumes[{l}]", 5)
        if tuple(t.shape) != (E, H, W, H >> l, W >> l):
            raise RuntimeError(f"volumes[{l}] shape {tuple(t.shape)} != {(E, H, W, H >> l, W >> l)}")
    _chk(coords, "coords", 4)
    need = H * W * 98 * (1 if shared_offsets else E)
    for name, o in (("off0", off0), ("off1", off1)):
        if not (isinstance(o, torch.Tensor) and o.is_cuda and o.is_contiguous() and o.dtype == torch.float32
                and o.numel() >= need):
            raise RuntimeError(f"{name} must be a contiguous fp32 CUDA tensor with >= {need} elements")
    mask = torch.empty(E, H, W, dtype=torch.float32, device=coords.device) if return_mask else None
    if out is not None:
        if not (isinstance(out, torch.Tensor) and out.is_cuda and out.is_contiguous() and out.dim() == 4
                and tuple(out.shape[1:]) == (196, H, W) and out.dtype in (torch.float32, torch.float16)):
            raise RuntimeError("out must be a contiguous CUDA tensor [E_out,196,H,W], fp32 or fp16")
        if out_index is not None:
            if not (out_index.is_cuda and out_index.dtype == torch.int32 and out_index.is_contiguous()
                    and out_index.numel() == E and out_index.device == coords.device):
                raise RuntimeError("out_index must be a contiguous CUDA int32 vector with one row per edge")
        elif out.shape[0] < E:
            raise RuntimeError(f"out has {out.shape[0]} rows, {E} edges")
        null = ctypes.c_void_p(0)
        if boxes0 is not None:
            head = (_p(boxes0), _p(boxes1) if boxes1 is not None else null, _p(volumes[1]) if boxes1 is None else null)
            fn = _lib.lib().lgu_altcorr_lookup_boxes_into
        else:
            head = (_p(volumes[0]), _p(volumes[1]))
            fn = _lib.lib().lgu_altcorr_lookup_fused_into
        with torch.cuda.device(coords.device):
            st = fn(
                *head, _p(volumes[2]), _p(volumes[3]), _p(coords), _p(off0), _p(off1), _p(out),
                _p(out_index) if out_index is not None else ctypes.c_void_p(0), _i(out.dtype == torch.float16),
                _p(mask) if return_mask else ctypes.c_void_p(0), _i(E), _i(H), _i(W), _i(4), _i(radius),
                _i(1 if shared_offsets else 0), _i(1 if apply_mask else 0), _stream(coords))
        _lib.check(st, "altcorr_lookup_fused (into)")
        return (out, mask) if return_mask else out
    corr = torch.empty(E, 196, H, W, dtype=torch.float32, device=coords.device)
    if boxes0 is not None:
        with torch.cuda.device(coords.device):
            null = ctypes.c_void_p(0)
            st = _lib.lib().lgu_altcorr_lookup_boxes_into(
                _p(boxes0), _p(boxes1) if boxes1 is not None else null, _p(volumes[1]) if boxes1 is None else null,
                _p(volumes[2]), _p(volumes[3]), _p(coords), _p(off0), _p(off1), _p(corr),
                ctypes.c_void_p(0), _i(0), _p(mask) if return_mask else ctypes.c_void_p(0), _i(E), _i(H), _i(W), _i(4),
                _i(radius), _i(1 if shared_offsets else 0), _i(1 if apply_mask else 0), _stream(coords))
        _lib.check(st, "altcorr_lookup_fused (boxes)")
        return (corr, mask) if return_mask else corr
    with torch.cuda.device(coords.device):
        st = _lib.lib().lgu_altcorr_lookup_fused(_p(volumes[0]), _p(volumes[1]), _p(volumes[2]), _p(volumes[3]),
                                                 _p(coords), _p(off0), _p(off1), _p(corr),
                                                 _p(mask) if return_mask else ctypes.c_void_p(0), _i(E), _i(H), _i(W),
                                                 _i(4), _i(radius), _i(1 if shared_offsets else 0),
                                                 _i(1 if apply_mask else 0), _stream(coords))
    _lib.check(st, "altcorr_lookup_fused")
    return (corr, mask) if return_mask else corr
