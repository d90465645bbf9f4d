"""Second, deliberately naive restatement of the reference ops in pure Python/numpy loops (tiny sizes
only).  Used by the CPU tests to cross-check oracle/lgu_oracle.c before either is trusted; written from
SURVEY.md Appendix A, independently of the C file (no FMA emulation -> compare with 1e-6 tolerance)."""
import math

import numpy as np

f32 = np.float32


def _f2i(v):
    if v != v:
        return 0
    if v >= 2147483648.0:
        return 2**31 - 1
    if v <= -2147483648.0:
        return -2**31
    return int(v)


def _wrap(v):
    return (v + 2**31) % 2**32 - 2**31


def _bil(V, y1, x1, dx, dy):
    H2, W2 = V.shape
    x2, y2 = _wrap(x1 + 1), _wrap(y1 + 1)
    xo, yo = 0 <= x2 < W2, 0 <= y2 < H2
    q11 = V[y1, x1]
    q21 = V[y1, x2] if xo else f32(0)
    q12 = V[y2, x1] if yo else f32(0)
    q22 = V[y2, x2] if (xo and yo) else f32(0)
    one = f32(1)
    return (q11 * ((one - dy) * (one - dx)) + q21 * ((one - dy) * dx) + q12 * (dy * (one - dx)) + q22 * (dy * dx)), \
        (q11, q21, q12, q22, xo, yo, x2, y2)


def lookup_forward(volume, coords, offset, r, deform):
    E, H1, W1, H2, W2 = volume.shape
    rd = 2 * r + 1
    out = np.zeros((E, rd, rd, H1, W1), f32)
    if deform:
        offset[:, :, :, r, r, :] = 0
    with np.errstate(all="ignore"):
        for n in range(E):
            for y in range(H1):
                for x in range(W1):
                    x0, y0 = coords[n, 0, y, x], coords[n, 1, y, x]
                    for i in range(rd):
                        for j in range(rd):
                            if deform:
                                px, py = f32(offset[n, y, x, i, j, 0] + x0), f32(offset[n, y, x, i, j, 1] + y0)
                                fx, fy = _f2i(np.floor(px)), _f2i(np.floor(py))
                                dx, dy = f32(px - f32(fx)), f32(py - f32(fy))
                            else:
                                fx, fy = _f2i(np.floor(x0)), _f2i(np.floor(y0))
                                dx, dy = f32(x0 - np.floor(x0)), f32(y0 - np.floor(y0))
                            x1, y1 = _wrap(fx - r + i), _wrap(fy - r + j)
                            if 0 <= y1 < H2 and 0 <= x1 < W2:
                                out[n, i, j, y, x], _ = _bil(volume[n, y, x], y1, x1, dx, dy)
    return out


def lookup_backward(volume, coords, offset, corr_grad, r, deform):
    E, H1, W1, H2, W2 = volume.shape
    rd = 2 * r + 1
    gv = np.zeros_like(volume)
    go = np.zeros((E, H1, W1, rd, rd, 2), f32)
    if deform:
        offset[:, :, :, r, r, :] = 0
    one = f32(1)
    with np.errstate(all="ignore"):
        for n in range(E):
            for y in range(H1):
                for x in range(W1):
                    x0, y0 = coords[n, 0, y, x], coords[n, 1, y, x]
                    for i in range(rd):
                        for j in range(rd):
                            if deform:
                                px, py = f32(offset[n, y, x, i, j, 0] + x0), f32(offset[n, y, x, i, j, 1] + y0)
                                fx, fy = _f2i(np.floor(px)), _f2i(np.floor(py))
                                dx, dy = f32(px - f32(fx)), f32(py - f32(fy))
                            else:
                                fx, fy = _f2i(np.floor(x0)), _f2i(np.floor(y0))
                                dx, dy = f32(x0 - np.floor(x0)), f32(y0 - np.floor(y0))
                            x1, y1 = _wrap(fx - r + i), _wrap(fy - r + j)
                            if not (0 <= y1 < H2 and 0 <= x1 < W2):
                                continue
                            g = corr_grad[n, i, j, y, x]
                            _, (q11, q21, q12, q22, xo, yo, x2, y2) = _bil(volume[n, y, x], y1, x1, dx, dy)
                            G = gv[n, y, x]
                            G[y1, x1] += (one - dy) * (one - dx) * g
                            if xo:
                                G[y1, x2] += (one - dy) * dx * g
                            if yo:
                                G[y2, x1] += dy * (one - dx) * g
                            if xo and yo:
                                G[y2, x2] += dy * dx * g
                            go[n, y, x, i, j, 1] = (-q11 * (one - dx) - q21 * dx + q12 * (one - dx) + q22 * dx) * g
                            go[n, y, x, i, j, 0] = (-q11 * (one - dy) + q21 * (one - dy) - q12 * dy + q22 * dy) * g
    return gv, go


def gaussian_forward(means, covs, volume, r):
    E, H1, W1, H2, W2 = volume.shape
    rd = 2 * r + 1
    out = np.zeros_like(volume)
    for n in range(E):
        for y in range(H1):
            for x in range(W1):
                mx, my = means[n, y, x]
                c1, c2 = covs[n, y, x]
                cx, cy = _f2i(np.floor(mx)), _f2i(np.floor(my))
                for i in range(rd):
                    for j in range(rd):
                        x1, y1 = _wrap(cx - r + i), _wrap(cy - r + j)
                        if 0 <= y1 < H2 and 0 <= x1 < W2:
                            ddx, ddy = f32(f32(x1) - mx), f32(f32(y1) - my)
                            f = f32(-0.5) * (ddx / c1 * ddx + ddy / c2 * ddy)
                            out[n, y, x, y1, x1] = volume[n, y, x, y1, x1] * f32(3) * f32(math.exp(f))
    return out


def gaussian_backward(means, covs, volume, gout, r):
    E, H1, W1, H2, W2 = volume.shape
    rd = 2 * r + 1
    gm = np.zeros_like(means)
    gc = np.zeros_like(covs)
    for n in range(E):
        for y in range(H1):
            for x in range(W1):
                mx, my = means[n, y, x]
                c1, c2 = covs[n, y, x]
                cx, cy = _f2i(np.floor(mx)), _f2i(np.floor(my))
                for i in range(rd):
                    for j in range(rd):
                        x1, y1 = _wrap(cx - r + i), _wrap(cy - r + j)
                        if 0 <= y1 < H2 and 0 <= x1 < W2:
                            ddx, ddy = float(f32(x1) - mx), float(f32(y1) - my)
                            e = math.exp(-0.5 * (ddx * ddx / c1 + ddy * ddy / c2))
                            v, g = float(volume[n, y, x, y1, x1]), float(gout[n, y, x, y1, x1])
                            gm[n, y, x, 0] += 3 * v * e * ddx / c1 * g
                            gm[n, y, x, 1] += 3 * v * e * ddy / c2 * g
                            gc[n, y, x, 0] += 3 * v * e * 0.5 * ddx * ddx / (c1 * c1) * g
                            gc[n, y, x, 1] += 3 * v * e * 0.5 * ddy * ddy / (c2 * c2) * g
    return gm, gc


def lowmem_forward(f1, f2, coords, offset, r, strict_ref=True):
    B, H1, W1, C = f1.shape
    _, H2, W2, _ = f2.shape
    N = coords.shape[1]
    rd = 2 * r + 1
    out = np.zeros((B, N, rd, rd, H1, W1), f32)
    one = f32(1)
    with np.errstate(all="ignore"):
        for b in range(B):
            for n in range(N):
                slab = b * n if strict_ref else b * N + n
                offset[slab, :, :, r, r, :] = 0
        for b in range(B):
            for n in range(N):
                slab = b * n if strict_ref else b * N + n
                for h in range(H1):
                    for w in range(W1):
                        for iy in range(rd):
                            for ix in range(rd):
                                px = f32(coords[b, n, h, w, 0] + offset[slab, h, w, ix, iy, 0])
                                py = f32(coords[b, n, h, w, 1] + offset[slab, h, w, ix, iy, 1])
                                dx, dy = f32(px - np.floor(px)), f32(py - np.floor(py))
                                h2, w2 = _wrap(_f2i(np.floor(py)) - r + iy), _wrap(_f2i(np.floor(px)) - r + ix)

                                def at(hh, ww):
                                    return f2[b, hh, ww] if (0 <= hh < H2 and 0 <= ww < W2) else np.zeros(C, f32)
                                q = at(h2, w2) * ((one - dy) * (one - dx)) + at(h2, _wrap(w2 + 1)) * ((one - dy) * dx) \
                                    + at(_wrap(h2 + 1), w2) * (dy * (one - dx)) + at(_wrap(h2 + 1), _wrap(w2 + 1)) * (dy * dx)
                                out[b, n, ix, iy, h, w] = np.dot(f1[b, h, w].astype(np.float64), q.astype(np.float64))
    return out


def altcorr_forward(f1, f2, coords, r):
    B, H1, W1, C = f1.shape
    _, H2, W2, _ = f2.shape
    N = coords.shape[1]
    rd = 2 * r + 1
    out = np.zeros((B, N, rd * rd, H1, W1), np.float64)
    with np.errstate(all="ignore"):
        for b in range(B):
            for n in range(N):
                for h in range(H1):
                    for w in range(W1):
                        x, y = coords[b, n, h, w]
                        dx, dy = float(f32(x - np.floor(x))), float(f32(y - np.floor(y)))
                        bx, by = _f2i(np.floor(x)), _f2i(np.floor(y))
                        for iy in range(rd + 1):
                            for ix in range(rd + 1):
                                h2, w2 = _wrap(by - r + iy), _wrap(bx - r + ix)
                                s = 0.0
                                if 0 <= h2 < H2 and 0 <= w2 < W2:
                                    s = float(np.dot(f1[b, h, w].astype(np.float64), f2[b, h2, w2].astype(np.float64)))
                                if iy > 0 and ix > 0:
                                    out[b, n, (iy - 1) + rd * (ix - 1), h, w] += s * dy * dx
                                if iy > 0 and ix < rd:
                                    out[b, n, (iy - 1) + rd * ix, h, w] += s * dy * (1 - dx)
                                if iy < rd and ix > 0:
                                    out[b, n, iy + rd * (ix - 1), h, w] += s * (1 - dy) * dx
                                if iy < rd and ix < rd:
                                    out[b, n, iy + rd * ix, h, w] += s * (1 - dy) * (1 - dx)
    return out.astype(f32)
