// lookup_bwd.cu -- corr_index_backward / defCorr_index_backward for sm_100a.
//
// Replaces /root/reference/offersample_LGS/corrSample_kernel.cu:84-136,170-199 and
// defCorrSample_kernel.cu:93-162,198-231, which zero-fill a dense volume_grad
// (torch::zeros_like) and then read-modify-write it from one thread per pixel.
//
// The API contract keeps volume_grad DENSE ([E,H1,W1,H2,W2], 4*H2*W2 bytes per source pixel),
// so this op is bound by that write.  Design: one pass, every byte written exactly once.
//   * a CTA owns 32 consecutive source pixels; corr_grad's rd^2 x 32 tile is loaded with
//     128-byte rows and transposed through shared memory;
//   * ONE WARP PER PIXEL: lane t owns tap t; the <=4 corner contributions of all taps are
//     accumulated into a per-warp H2xW2 slice in SHARED memory (collisions between lanes of
//     one instruction are resolved with __match_any_sync and applied in lane order, so the
//     result is deterministic);
//   * the finished slice is streamed out with 16-byte stores (rows no tap touched are
//     written as zeros without reading shared memory) and the slice is re-zeroed on the fly;
//   * offset_grad records (8*rd^2 B per pixel) are written as one coalesced float2 row.
// Index logic identical to the forward (reference quirks Q1, Q3, Q5).
#include "common.cuh"

namespace lgu {

constexpr int kBwGroup = 4;  // source pixels per warp (all their loads are issued before any store)

// Scatter-add of one value per lane into the warp's shared slice; `idx` < 0: nothing to add.
// DISTINCT = the caller guarantees that no two lanes of this call share an index (all taps of the pixel
// carry the same offset, so tap (i,j) -> cell is injective): plain read-modify-write.  Otherwise lanes may
// collide (learned offsets can map two taps onto one cell) and a shared-memory atomic resolves it.
// (__match_any_sync-based conflict resolution was measured at ~1100 cycles per call with 32 distinct keys.)
template <bool DISTINCT>
__device__ __forceinline__ void warp_scatter_add(float* __restrict__ slice, int idx, float val) {
  if (DISTINCT) {
    if (idx >= 0) slice[idx] += val;
  } else {
    if (idx >= 0) atomicAdd(slice + idx, val);
  }
  __syncwarp();
}

template <int R, bool DEFORM, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
lookup_bwd_kernel(const float* __restrict__ volume, const float* __restrict__ coords, float* __restrict__ offset,
                  const float* __restrict__ corr_grad, float* __restrict__ volume_grad,
                  float* __restrict__ offset_grad, int P, int H2, int W2, int tiles_per_edge) {
  constexpr int RD = 2 * R + 1, TAPS = RD * RD, PASSES = (TAPS + 31) / 32;
  constexpr int PIX_PER_WARP = kBwGroup;
  constexpr int kBwTile = WARPS * kBwGroup;          // source pixels per CTA
  extern __shared__ __align__(16) float smem[];
  float(*s_g)[kBwTile + 1] = reinterpret_cast<float(*)[kBwTile + 1]>(smem);   // [TAPS][33]
  const int Q = H2 * W2;
  const int Qpad = (Q + 3) & ~3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* slice = smem + ((TAPS * (kBwTile + 1) + 3) & ~3) + warp * Qpad;

  const int n = blockIdx.x / tiles_per_edge;
  const int p0 = (blockIdx.x - n * tiles_per_edge) * kBwTile;

  // stage the upstream gradient tile (coalesced rows) and clear the slices
  {
    // rows of kBwTile floats (64 or 128 B), several tap rows per warp instruction
    for (int q = threadIdx.x; q < TAPS * kBwTile; q += WARPS * 32) {
      const int t = q / kBwTile, c = q - t * kBwTile;
      s_g[t][c] = (p0 + c) < P ? __ldg(corr_grad + ((size_t)n * TAPS + t) * P + p0 + c) : 0.0f;
    }
    for (int q = lane; q < Qpad; q += 32) slice[q] = 0.0f;
  }
  __syncthreads();

  const float* cx = coords + (size_t)n * 2 * P;
  const float* cy = cx + P;

  // Pixels are walked in groups of GRP: all coords / offset records of a group are requested first
  // (phase A), then all corner gathers (phase B, deformable only), then each pixel is scattered and
  // streamed out (phase C).  One phase per pixel left the kernel latency-bound (ncu: serialized
  // ~5k-cycle long-scoreboard waits).
  constexpr int GRP = PIX_PER_WARP < 4 ? PIX_PER_WARP : 4;
#pragma unroll 1
  for (int k0 = 0; k0 < PIX_PER_WARP; k0 += GRP) {
    if (p0 + warp * PIX_PER_WARP + k0 >= P) break;   // warp-uniform
    float dxs[GRP][PASSES], dys[GRP][PASSES], q[GRP][PASSES][4];
    int i11s[GRP][PASSES];
    unsigned gates[GRP][PASSES];                     // bit0 gate, bit1 xo, bit2 yo
    // ---- phase A
    float x0[GRP], y0[GRP];
    float2 o[GRP][PASSES];
    bool uniform_off[GRP];
#pragma unroll
    for (int g = 0; g < GRP; ++g) {
      const int p = min(p0 + warp * PIX_PER_WARP + k0 + g, P - 1);
      const size_t pix = (size_t)n * P + p;
      x0[g] = __ldg(cx + p);
      y0[g] = __ldg(cy + p);
#pragma unroll
      for (int ps = 0; ps < PASSES; ++ps) {
        o[g][ps] = make_float2(0.0f, 0.0f);
        if (DEFORM) {
          float2* O = reinterpret_cast<float2*>(offset) + pix * TAPS;
          const int t = ps * 32 + lane;
          o[g][ps] = O[min(t, TAPS - 1)];
          if (t == R * RD + R) o[g][ps] = make_float2(0.0f, 0.0f);   // defCorrSample_kernel.cu:122-123 (Q5), stored below
        }
      }
    }
    // all taps of a pixel share one offset (always true for the plain lookup, and for the zero-offset
    // pyramid levels)?  then tap -> cell is injective and the scatter needs no atomics
#pragma unroll
    for (int g = 0; g < GRP; ++g) {
      bool same = true;
      if (DEFORM) {
        const float ox0 = __shfl_sync(0xffffffffu, o[g][0].x, 0), oy0 = __shfl_sync(0xffffffffu, o[g][0].y, 0);
#pragma unroll
        for (int ps = 0; ps < PASSES; ++ps)
          same = same && (ps * 32 + lane >= TAPS || (o[g][ps].x == ox0 && o[g][ps].y == oy0));
        same = __all_sync(0xffffffffu, same);
      }
      uniform_off[g] = same;
    }
    // ---- phase B
#pragma unroll
    for (int g = 0; g < GRP; ++g) {
      const int p = min(p0 + warp * PIX_PER_WARP + k0 + g, P - 1);
      const float* V = DEFORM ? volume + ((size_t)n * P + p) * (size_t)Q : nullptr;
#pragma unroll
      for (int ps = 0; ps < PASSES; ++ps) {
        const int t = ps * 32 + lane;
        const bool has_tap = t < TAPS;
        const int tt = has_tap ? t : 0;
        const int i = tt / RD, j = tt - i * RD;
        float dx, dy;
        int fx, fy;
        if (DEFORM) {
          const float px = __fadd_rn(o[g][ps].x, x0[g]), py = __fadd_rn(o[g][ps].y, y0[g]);
          fx = floor_to_int(px);
          fy = floor_to_int(py);
          dx = __fsub_rn(px, (float)fx);
          dy = __fsub_rn(py, (float)fy);
        } else {
          dx = __fsub_rn(x0[g], floorf(x0[g]));
          dy = __fsub_rn(y0[g], floorf(y0[g]));
          fx = floor_to_int(x0[g]);
          fy = floor_to_int(y0[g]);
        }
        const int x1 = tap_coord(fx, R, i), y1 = tap_coord(fy, R, j);
        const int x2 = wrap_inc(x1), y2 = wrap_inc(y1);
        const bool gate = has_tap && in_bounds(y1, x1, H2, W2);
        const bool xo = gate && (x2 >= 0 && x2 < W2), yo = gate && (y2 >= 0 && y2 < H2);
        const int i11 = gate ? y1 * W2 + x1 : 0;
        dxs[g][ps] = dx;
        dys[g][ps] = dy;
        i11s[g][ps] = i11;
        gates[g][ps] = (gate ? 1u : 0u) | (xo ? 2u : 0u) | (yo ? 4u : 0u);
        if (DEFORM) {   // unconditional loads from always-valid addresses; masked in phase C
          q[g][ps][0] = __ldg(V + i11);
          q[g][ps][1] = __ldg(V + (xo ? i11 + 1 : i11));
          q[g][ps][2] = __ldg(V + (yo ? i11 + W2 : i11));
          q[g][ps][3] = __ldg(V + ((xo && yo) ? i11 + W2 + 1 : i11));
        }
      }
    }
    // ---- phase C
#pragma unroll
    for (int g = 0; g < GRP; ++g) {
      const int pl = warp * PIX_PER_WARP + k0 + g;
      const int p = p0 + pl;
      if (p >= P) break;                               // warp-uniform
      const size_t pix = (size_t)n * P + p;
      float2* GO = DEFORM ? reinterpret_cast<float2*>(offset_grad) + pix * TAPS : nullptr;
      if (DEFORM && lane == 0)   // in-place zeroing of the centre tap, after every load of this warp has been issued
        reinterpret_cast<float2*>(offset)[pix * TAPS + R * RD + R] = make_float2(0.0f, 0.0f);
      int ymin = 0x7fffffff, ymax = -1;
#pragma unroll
      for (int ps = 0; ps < PASSES; ++ps) {
        const int t = ps * 32 + lane;
        const bool has_tap = t < TAPS;
        const bool gate = gates[g][ps] & 1u, xo = gates[g][ps] & 2u, yo = gates[g][ps] & 4u;
        const float dx = dxs[g][ps], dy = dys[g][ps];
        const float omdx = __fsub_rn(1.0f, dx), omdy = __fsub_rn(1.0f, dy);
        const int i11 = i11s[g][ps];
        const float gq = s_g[has_tap ? t : 0][pl];
        if (DEFORM && has_tap) {
          float2 go = make_float2(0.0f, 0.0f);
          if (gate) {
            const float q11 = q[g][ps][0], q21 = xo ? q[g][ps][1] : 0.0f, q12 = yo ? q[g][ps][2] : 0.0f,
                        q22 = (xo && yo) ? q[g][ps][3] : 0.0f;
            // defCorrSample_kernel.cu:156-157 in the reference's SASS operation order
            float ty = __fmaf_rn(-q11, omdx, -__fmul_rn(dx, q21));
            ty = __fmaf_rn(omdx, q12, ty);
            ty = __fmaf_rn(dx, q22, ty);
            float tx = __fmaf_rn(omdy, q21, -__fmul_rn(q11, omdy));
            tx = __fmaf_rn(-dy, q12, tx);
            tx = __fmaf_rn(dy, q22, tx);
            go.x = __fmul_rn(tx, gq);
            go.y = __fmul_rn(ty, gq);
          }
          GO[t] = go;
        }
        if (gate) {
          const int y1 = i11 / W2;
          ymin = min(ymin, y1);
          ymax = max(ymax, yo ? y1 + 1 : y1);
        }
        const float w11 = __fmul_rn(__fmul_rn(omdy, omdx), gq), w21 = __fmul_rn(__fmul_rn(omdy, dx), gq);
        const float w12 = __fmul_rn(__fmul_rn(dy, omdx), gq), w22 = __fmul_rn(__fmul_rn(dy, dx), gq);
        if (uniform_off[g]) {
          warp_scatter_add<true>(slice, gate ? i11 : -1, w11);
          warp_scatter_add<true>(slice, xo ? i11 + 1 : -1, w21);
          warp_scatter_add<true>(slice, yo ? i11 + W2 : -1, w12);
          warp_scatter_add<true>(slice, (xo && yo) ? i11 + W2 + 1 : -1, w22);
        } else {
          warp_scatter_add<false>(slice, gate ? i11 : -1, w11);
          warp_scatter_add<false>(slice, xo ? i11 + 1 : -1, w21);
          warp_scatter_add<false>(slice, yo ? i11 + W2 : -1, w12);
          warp_scatter_add<false>(slice, (xo && yo) ? i11 + W2 + 1 : -1, w22);
        }
      }

      // stream the slice out (and re-zero it); rows outside [ymin,ymax] are known zeros
      ymin = __reduce_min_sync(0xffffffffu, ymin);
      ymax = __reduce_max_sync(0xffffffffu, ymax);
      const int lo = (ymax >= 0) ? ymin * W2 : 0x7fffffff;    // first touched element
      const int hi = (ymax >= 0) ? (ymax + 1) * W2 : 0;       // one past the last
      float* G = volume_grad + pix * (size_t)Q;
      if ((Q & 3) == 0 && ((reinterpret_cast<uintptr_t>(G) & 15) == 0)) {
        float4* G4 = reinterpret_cast<float4*>(G);
        float4* S4 = reinterpret_cast<float4*>(slice);
        const float4 z = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        for (int q4 = lane; q4 < (Q >> 2); q4 += 32) {
          const int e0 = q4 << 2;
          float4 v = z;
          if (e0 + 3 >= lo && e0 < hi) {
            v = S4[q4];
            S4[q4] = z;
          }
          __stcs(G4 + q4, v);
        }
      } else {
        for (int qq = lane; qq < Q; qq += 32) {
          float v = 0.0f;
          if (qq >= lo && qq < hi) {
            v = slice[qq];
            slice[qq] = 0.0f;
          }
          __stcs(G + qq, v);
        }
      }
      __syncwarp();
    }
  }
}

// Fallback for slices too large for shared memory or unusual radii: memset + one thread per
// source pixel applying the taps in the reference's order (race-free: a thread owns its slice).
template <bool DEFORM>
__global__ void lookup_bwd_generic_kernel(const float* __restrict__ volume, const float* __restrict__ coords,
                                          float* __restrict__ offset, const float* __restrict__ corr_grad,
                                          float* __restrict__ volume_grad, float* __restrict__ offset_grad, int r,
                                          int P, int H2, int W2, long long npix) {
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= npix) return;
  const int rd = 2 * r + 1, taps = rd * rd;
  const int n = (int)(pix / P), p = (int)(pix - (long long)n * P);
  const size_t Q = (size_t)H2 * W2;
  const float x0 = coords[(size_t)n * 2 * P + p], y0 = coords[(size_t)n * 2 * P + P + p];
  const float* V = DEFORM ? volume + (size_t)pix * Q : nullptr;
  float* G = volume_grad + (size_t)pix * Q;
  float2* O = DEFORM ? reinterpret_cast<float2*>(offset) + (size_t)pix * taps : nullptr;
  float2* GO = DEFORM ? reinterpret_cast<float2*>(offset_grad) + (size_t)pix * taps : nullptr;
  if (DEFORM) O[r * rd + r] = make_float2(0.0f, 0.0f);
  for (int t = 0; t < taps; ++t) {
    const int i = t / rd, j = t - i * rd;
    float dx, dy;
    int fx, fy;
    if (DEFORM) {
      const float2 o = O[t];
      const float px = __fadd_rn(o.x, x0), py = __fadd_rn(o.y, y0);
      fx = floor_to_int(px);
      fy = floor_to_int(py);
      dx = __fsub_rn(px, (float)fx);
      dy = __fsub_rn(py, (float)fy);
    } else {
      dx = __fsub_rn(x0, floorf(x0));
      dy = __fsub_rn(y0, floorf(y0));
      fx = floor_to_int(x0);
      fy = floor_to_int(y0);
    }
    const int x1 = tap_coord(fx, r, i), y1 = tap_coord(fy, r, j);
    const int x2 = wrap_inc(x1), y2 = wrap_inc(y1);
    float2 go = make_float2(0.0f, 0.0f);
    if (in_bounds(y1, x1, H2, W2)) {
      const bool xo = (x2 >= 0 && x2 < W2), yo = (y2 >= 0 && y2 < H2);
      const float g = corr_grad[((size_t)n * taps + t) * P + p];
      const float omdx = __fsub_rn(1.0f, dx), omdy = __fsub_rn(1.0f, dy);
      const size_t i11 = (size_t)y1 * W2 + x1;
      G[i11] = __fmaf_rn(__fmul_rn(omdy, omdx), g, G[i11]);
      if (xo) G[i11 + 1] = __fmaf_rn(__fmul_rn(omdy, dx), g, G[i11 + 1]);
      if (yo) G[i11 + W2] = __fmaf_rn(__fmul_rn(dy, omdx), g, G[i11 + W2]);
      if (xo && yo) G[i11 + W2 + 1] = __fmaf_rn(__fmul_rn(dy, dx), g, G[i11 + W2 + 1]);
      if (DEFORM) {
        const float q11 = V[i11], q21 = xo ? V[i11 + 1] : 0.0f, q12 = yo ? V[i11 + W2] : 0.0f,
                    q22 = (xo && yo) ? V[i11 + W2 + 1] : 0.0f;
        float ty = __fmaf_rn(-q11, omdx, -__fmul_rn(dx, q21));
        ty = __fmaf_rn(omdx, q12, ty);
        ty = __fmaf_rn(dx, q22, ty);
        float tx = __fmaf_rn(omdy, q21, -__fmul_rn(q11, omdy));
        tx = __fmaf_rn(-dy, q12, tx);
        tx = __fmaf_rn(dy, q22, tx);
        go.x = __fmul_rn(tx, g);
        go.y = __fmul_rn(ty, g);
      }
    }
    if (DEFORM) GO[t] = go;
  }
}

template <int R, bool DEFORM, int WARPS>
static int launch_bwd_cfg(const float* volume, const float* coords, float* offset, const float* corr_grad,
                          float* volume_grad, float* offset_grad, int E, int P, int H2, int W2, cudaStream_t st) {
  constexpr int TAPS = (2 * R + 1) * (2 * R + 1);
  const int Q = H2 * W2, Qpad = (Q + 3) & ~3;
  constexpr int kBwTile = WARPS * kBwGroup;
  const size_t smem = (size_t)(((TAPS * (kBwTile + 1) + 3) & ~3) + WARPS * Qpad) * sizeof(float);
  auto kern = lookup_bwd_kernel<R, DEFORM, WARPS>;
  if (smem > 48 * 1024) {
    if (int rc = optin_smem(reinterpret_cast<const void*>(kern), (int)smem, "lookup backward")) return rc;
  }
  const int tiles = (P + kBwTile - 1) / kBwTile;
  kern<<<(unsigned)((long long)E * tiles), WARPS * 32, smem, st>>>(volume, coords, offset, corr_grad, volume_grad,
                                                                   offset_grad, P, H2, W2, tiles);
  return check_launch(DEFORM ? "lgu_defcorr_index_backward" : "lgu_corr_index_backward");
}

template <int R, bool DEFORM>
static int launch_bwd_r(const float* volume, const float* coords, float* offset, const float* corr_grad,
                        float* volume_grad, float* offset_grad, int E, int P, int H2, int W2, cudaStream_t st) {
  // Per-warp slice in shared memory: pick the warp count so that >= 2 CTAs fit per SM.
  const size_t slice = (size_t)((H2 * W2 + 3) & ~3) * sizeof(float);
  if (slice * 8 <= 40 * 1024)
    return launch_bwd_cfg<R, DEFORM, 8>(volume, coords, offset, corr_grad, volume_grad, offset_grad, E, P, H2, W2, st);
  if (slice * 4 <= 100 * 1024)
    return launch_bwd_cfg<R, DEFORM, 4>(volume, coords, offset, corr_grad, volume_grad, offset_grad, E, P, H2, W2, st);
  if (slice * 2 <= 200 * 1024)
    return launch_bwd_cfg<R, DEFORM, 2>(volume, coords, offset, corr_grad, volume_grad, offset_grad, E, P, H2, W2, st);
  return -1;  // caller falls back to the generic path
}

template <bool DEFORM>
static int launch_lookup_bwd(const float* volume, const float* coords, float* offset, const float* corr_grad,
                             float* volume_grad, float* offset_grad, int E, int H1, int W1, int H2, int W2, int r,
                             cudaStream_t st) {
  const int P = H1 * W1;
  const long long nblk = (long long)E * ((P + 7) / 8);   // smallest tile is 2 warps x 4 pixels
  LGU_REQUIRE(nblk < 2147483647LL, "lookup backward: grid too large (%lld CTAs)", nblk);
  int rc = -1;
  switch (r) {
    case 1: rc = launch_bwd_r<1, DEFORM>(volume, coords, offset, corr_grad, volume_grad, offset_grad, E, P, H2, W2, st); break;
    case 2: rc = launch_bwd_r<2, DEFORM>(volume, coords, offset, corr_grad, volume_grad, offset_grad, E, P, H2, W2, st); break;
    case 3: rc = launch_bwd_r<3, DEFORM>(volume, coords, offset, corr_grad, volume_grad, offset_grad, E, P, H2, W2, st); break;
    case 4: rc = launch_bwd_r<4, DEFORM>(volume, coords, offset, corr_grad, volume_grad, offset_grad, E, P, H2, W2, st); break;
    default: break;
  }
  if (rc >= 0) return rc;
  // generic path
  const long long npix = (long long)E * P;
  cudaError_t e = cudaMemsetAsync(volume_grad, 0, sizeof(float) * (size_t)npix * H2 * W2, st);
  if (e != cudaSuccess) {
    set_error("lookup backward: memset failed: %s", cudaGetErrorString(e));
    return LGU_ERR_LAUNCH;
  }
  const long long nb = (npix + 127) / 128;
  LGU_REQUIRE(nb < 2147483647LL, "lookup backward: too many pixels (%lld)", npix);
  lookup_bwd_generic_kernel<DEFORM><<<(unsigned)nb, 128, 0, st>>>(volume, coords, offset, corr_grad, volume_grad,
                                                                  offset_grad, r, P, H2, W2, npix);
  return check_launch(DEFORM ? "lgu_defcorr_index_backward(generic)" : "lgu_corr_index_backward(generic)");
}

}  // namespace lgu

namespace lgu {
int launch_level_tma(bool bwd, const float* volume, const float* coords, float* offset, float* corr,
                     const float* corr_grad, float* volume_grad, float* offset_grad, int E, int H1, int W1, int H2,
                     int W2, cudaStream_t st);   // lookup_level_tma.cu
}
extern "C" int lgu_corr_index_backward(const float* coords, const float* corr_grad, float* volume_grad, int E, int H1,
                                       int W1, int H2, int W2, int radius, void* stream) {
  if (E == 0) return LGU_OK;   // empty edge set: nothing to do (pointers may be null)
  LGU_REQUIRE(coords && corr_grad && volume_grad, "lgu_corr_index_backward: null pointer");
  LGU_REQUIRE(E >= 0 && H1 > 0 && W1 > 0 && H2 > 0 && W2 > 0 && radius >= 0,
              "lgu_corr_index_backward: bad sizes E=%d H1=%d W1=%d H2=%d W2=%d r=%d", E, H1, W1, H2, W2, radius);
  LGU_REQUIRE((long long)H2 * W2 < (1LL << 30), "lgu_corr_index_backward: H2*W2 too large");
  return lgu::launch_lookup_bwd<false>(nullptr, coords, nullptr, corr_grad, volume_grad, nullptr, E, H1, W1, H2, W2,
                                       radius, (cudaStream_t)stream);
}

extern "C" int lgu_defcorr_index_backward(const float* volume, const float* coords, float* offset,
                                          const float* corr_grad, float* volume_grad, float* offset_grad, int E,
                                          int H1, int W1, int H2, int W2, int radius, void* stream) {
  if (E == 0) return LGU_OK;   // empty edge set: nothing to do (pointers may be null)
  LGU_REQUIRE(volume && coords && offset && corr_grad && volume_grad && offset_grad,
              "lgu_defcorr_index_backward: null pointer");
  LGU_REQUIRE(E >= 0 && H1 > 0 && W1 > 0 && H2 > 0 && W2 > 0 && radius >= 0,
              "lgu_defcorr_index_backward: bad sizes E=%d H1=%d W1=%d H2=%d W2=%d r=%d", E, H1, W1, H2, W2, radius);
  LGU_REQUIRE((long long)H2 * W2 < (1LL << 30), "lgu_defcorr_index_backward: H2*W2 too large");
  if (radius == 3) {   // TMA-staged footprints + box accumulator (lookup_level_tma.cu)
    const int rc = lgu::launch_level_tma(true, volume, coords, offset, nullptr, corr_grad, volume_grad, offset_grad, E,
                                         H1, W1, H2, W2, (cudaStream_t)stream);
    if (rc >= 0) return rc;
  }
  return lgu::launch_lookup_bwd<true>(volume, coords, offset, corr_grad, volume_grad, offset_grad, E, H1, W1, H2, W2,
                                      radius, (cudaStream_t)stream);
}
