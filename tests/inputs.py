"""Seeded synthetic inputs for the correlation hot path.

Everything is generated on the CPU with a fixed torch.Generator so that the oracle, the
compiled reference and the sm_100a kernels see identical bits.
"""
import math

import torch
import torch.nn.functional as F


def gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


def make_coords(E, H1, W1, H2, W2, g, flow_sigma=4.0, oob_frac=0.02, probes=False):
    """[E,2,H1,W1] (ch0 = x, ch1 = y) in the units of an (H2,W2) target grid: identity grid scaled to the
    level + smooth flow + a fraction of pixels pushed far out of bounds (+ optional NaN/inf/huge probes)."""
    ys, xs = torch.meshgrid(torch.arange(H1, dtype=torch.float32), torch.arange(W1, dtype=torch.float32),
                            indexing="ij")
    sx, sy = W2 / W1, H2 / H1
    lo = torch.randn(E, 2, max(H1 // 8, 1), max(W1 // 8, 1), generator=g) * flow_sigma
    flow = F.interpolate(lo, size=(H1, W1), mode="bilinear", align_corners=False)
    c = torch.stack([xs, ys], 0)[None] + flow
    c[:, 0] *= sx
    c[:, 1] *= sy
    oob = torch.rand(E, 1, H1, W1, generator=g) < oob_frac
    push = (torch.rand(E, 2, H1, W1, generator=g) - 0.5) * 200.0
    c = torch.where(oob, c + push, c)
    if probes and E * H1 * W1 >= 16:
        flat = c.permute(0, 2, 3, 1).reshape(-1, 2)
        vals = [float("nan"), float("inf"), -float("inf"), 1e10, -1e10, 2147483648.0, -2147483904.0,
                -0.0, 0.0, -1e-30, float(W2), float(W2) - 1e-4, -1.0, -0.5]
        for k, v in enumerate(vals):
            flat[k, k % 2] = v
        flat[len(vals)] = torch.tensor([float("nan"), float("nan")])
        c = flat.reshape(E, H1, W1, 2).permute(0, 3, 1, 2)
    return c.contiguous()


def make_offset(E, H1, W1, r, g, scale=4.0, zero=False):
    rd = 2 * r + 1
    if zero:
        return torch.zeros(E, H1, W1, rd, rd, 2)
    return (scale * torch.tanh(torch.randn(E, H1, W1, rd, rd, 2, generator=g))).contiguous()


def volume_case(E=2, H1=6, W1=8, H2=6, W2=8, r=3, seed=0, probes=False, zero_offset=False):
    g = gen(seed)
    rd = 2 * r + 1
    return dict(
        volume=torch.randn(E, H1, W1, H2, W2, generator=g),
        coords=make_coords(E, H1, W1, H2, W2, g, probes=probes),
        offset=make_offset(E, H1, W1, r, g, zero=zero_offset),
        corr_grad=torch.randn(E, rd, rd, H1, W1, generator=g),
        radius=r,
    )


def gaussian_case(E=2, H1=6, W1=8, H2=6, W2=8, r=4, seed=0, probes=False):
    g = gen(seed)
    ys, xs = torch.meshgrid(torch.arange(H1, dtype=torch.float32), torch.arange(W1, dtype=torch.float32),
                            indexing="ij")
    grid = torch.stack([xs * (W2 / W1), ys * (H2 / H1)], -1)[None]
    means = (grid + 0.5 * torch.randn(E, H1, W1, 2, generator=g)).contiguous()
    if probes:
        m = means.view(-1, 2)
        m[0, 0] = -3.5; m[1, 1] = H2 + 2.25; m[2, 0] = W2 - 0.5; m[3] = torch.tensor([-10.0, -10.0])
        m[4, 0] = 1e9; m[5, 1] = -1e9
    covs = (torch.sigmoid(torch.randn(E, H1, W1, 2, generator=g)) * 5 + 0.05).contiguous()
    return dict(means=means, covs=covs, volume=torch.randn(E, H1, W1, H2, W2, generator=g),
                out_grad=torch.randn(E, H1, W1, H2, W2, generator=g), radius=r)


def lowmem_case(B=3, N=1, H1=6, W1=8, H2=6, W2=8, C=128, r=3, seed=0, probes=False, half_exact=True):
    """Channels-last fmaps [B,H,W,C] (values fp16-representable when half_exact, like the backend's
    fp16 frame buffer, depth_video.py:36), coords [B,N,H1,W1,2], offsets [B*N,H1,W1,rd,rd,2]."""
    g = gen(seed)
    f1 = torch.randn(B, H1, W1, C, generator=g) / 4
    f2 = torch.randn(B, H2, W2, C, generator=g) / 4
    if half_exact:
        f1, f2 = f1.half().float(), f2.half().float()
    c = make_coords(B * N, H1, W1, H2, W2, g, probes=probes)          # [B*N,2,H1,W1]
    coords = c.view(B, N, 2, H1, W1).permute(0, 1, 3, 4, 2).contiguous()
    return dict(fmap1=f1.contiguous(), fmap2=f2.contiguous(), coords=coords,
                offset=make_offset(B * N, H1, W1, r, g), radius=r)


def edge_list(T, E, g, max_gap=3):
    """E (i,j) frame pairs over T keyframes: neighbours |i-j| <= max_gap first, then random proximity pairs."""
    pairs = [(i, j) for i in range(T) for j in range(T) if i != j and abs(i - j) <= max_gap]
    perm = torch.randperm(len(pairs), generator=g).tolist()
    pairs = [pairs[k] for k in perm]
    while len(pairs) < E:
        i, j = torch.randint(0, T, (2,), generator=g).tolist()
        if i != j:
            pairs.append((i, j))
    pairs = pairs[:E]
    ii = torch.tensor([p[0] for p in pairs], dtype=torch.int32)
    jj = torch.tensor([p[1] for p in pairs], dtype=torch.int32)
    return ii, jj


def frontend_case(E=8, T=6, H=48, W=64, C=128, seed=1235, half_fmaps=True):
    """BASELINE configs[0]/[1] shape: fmaps [T,C,H,W], E edges, Gaussian params, level-0 coords, offsets."""
    g = gen(seed)
    fmaps = torch.randn(T, C, H, W, generator=g)
    if half_fmaps:
        fmaps = fmaps.half().float()
    ii, jj = edge_list(T, E, g)
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32),
                            indexing="ij")
    grid = torch.stack([xs, ys], -1)[None]
    means = (grid + 0.5 * torch.randn(E, H, W, 2, generator=g)).contiguous()
    covs = (torch.sigmoid(torch.randn(E, H, W, 2, generator=g)) * 5 + 0.05).contiguous()
    coords = make_coords(E, H, W, H, W, g).permute(0, 2, 3, 1).contiguous()      # [E,H,W,2]
    off0 = 4 * torch.tanh(torch.randn(E, H, W, 98, generator=g))
    off1 = (4 * torch.tanh(torch.randn(E, H, W, 98, generator=g)) + off0) / 2
    return dict(fmaps=fmaps, ii=ii, jj=jj, means=means, covs=covs, coords=coords,
                offsets=[off0.contiguous(), off1.contiguous(), torch.zeros(E, H, W, 98), torch.zeros(E, H, W, 98)])
