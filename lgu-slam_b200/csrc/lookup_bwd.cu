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

constexpr int kBwTile = 32;  // source pixels per CTA

// Deterministic scatter-add of one value per lane into the warp's shared slice.
// `idx` < 0 means "this lane has nothing to add".
__device__ __forceinline__ void warp_scatter_add(float* __restrict__ slice, int idx, float val, int lane) {
  const unsigned key = idx >= 0 ? (unsigned)idx : (0x80000000u | (unsigned)lane);
  const unsigned peers = __match_any_sync(0xffffffffu, key);
  const int rank = __popc(peers & ((1u << lane) - 1u));
  const int rounds = __reduce_max_sync(0xffffffffu, (unsigned)rank);
  for (int it = 0; it <= rounds; ++it) {
    if (idx >= 0 && rank == it) slice[idx] += val;
    __syncwarp();
  }
}

template <int R, bool DEFORM, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
lookup_bwd_kernel(const float* __restrict__ volume, const float* __restrict__ coords, float* __restrict__ offset,
                  const float* __restrict__ corr_grad, float* __restrict__ volume_grad,
                  float* __restrict__ offset_grad, int P, int H2, int W2, int tiles_per_edge) {
  constexpr int RD = 2 * R + 1, TAPS = RD * RD, PASSES = (TAPS + 31) / 32;
  constexpr int PIX_PER_WARP = kBwTile / WARPS;
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
    const bool live = (p0 + lane) < P;
    const float* g = corr_grad + (size_t)n * TAPS * P + p0 + lane;
    for (int t = warp; t < TAPS; t += WARPS) s_g[t][lane] = live ? __ldg(g + (size_t)t * P) : 0.0f;
    for (int q = lane; q < Qpad; q += 32) slice[q] = 0.0f;
  }
  __syncthreads();

  const float* cx = coords + (size_t)n * 2 * P;
  const float* cy = cx + P;

#pragma unroll 1
  for (int k = 0; k < PIX_PER_WARP; ++k) {
    const int pl = warp * PIX_PER_WARP + k;
    const int p = p0 + pl;
    if (p >= P) break;                               // warp-uniform
    const size_t pix = (size_t)n * P + p;
    const float x0 = __ldg(cx + p), y0 = __ldg(cy + p);
    const float* V = DEFORM ? volume + pix * (size_t)Q : nullptr;
    float2* O = DEFORM ? reinterpret_cast<float2*>(offset) + pix * TAPS : nullptr;
    float2* GO = DEFORM ? reinterpret_cast<float2*>(offset_grad) + pix * TAPS : nullptr;
    int ymin = 0x7fffffff, ymax = -1;

#pragma unroll
    for (int ps = 0; ps < PASSES; ++ps) {
      const int t = ps * 32 + lane;
      const bool has_tap = t < TAPS;
      const int tt = has_tap ? t : 0;
      const int i = tt / RD, j = tt - i * RD;
      float dx, dy;
      int fx, fy;
      if (DEFORM) {
        float2 o = make_float2(0.0f, 0.0f);
        if (has_tap) {
          if (t == R * RD + R) O[t] = o;             // defCorrSample_kernel.cu:122-123 (Q5)
          else o = O[t];
        }
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
      const int x1 = tap_coord(fx, R, i), y1 = tap_coord(fy, R, j);
      const int x2 = wrap_inc(x1), y2 = wrap_inc(y1);
      const bool gate = has_tap && in_bounds(y1, x1, H2, W2);
      const bool xo = gate && (x2 >= 0 && x2 < W2), yo = gate && (y2 >= 0 && y2 < H2);
      const float g = s_g[tt][pl];
      const float omdx = __fsub_rn(1.0f, dx), omdy = __fsub_rn(1.0f, dy);
      const int i11 = y1 * W2 + x1;

      if (DEFORM) {
        float q11 = 0.0f, q21 = 0.0f, q12 = 0.0f, q22 = 0.0f;
        if (gate) q11 = __ldg(V + i11);
        if (xo) q21 = __ldg(V + i11 + 1);
        if (yo) q12 = __ldg(V + i11 + W2);
        if (xo && yo) q22 = __ldg(V + i11 + W2 + 1);
        if (has_tap) {
          float2 go = make_float2(0.0f, 0.0f);
          if (gate) {
            // defCorrSample_kernel.cu:156-157 in the reference's SASS operation order
            float ty = __fmaf_rn(-q11, omdx, -__fmul_rn(dx, q21));
            ty = __fmaf_rn(omdx, q12, ty);
            ty = __fmaf_rn(dx, q22, ty);
            float tx = __fmaf_rn(omdy, q21, -__fmul_rn(q11, omdy));
            tx = __fmaf_rn(-dy, q12, tx);
            tx = __fmaf_rn(dy, q22, tx);
            go.x = __fmul_rn(tx, g);
            go.y = __fmul_rn(ty, g);
          }
          GO[t] = go;
        }
      }

      if (gate) {
        ymin = min(ymin, y1);
        ymax = max(ymax, yo ? y2 : y1);
      }
      warp_scatter_add(slice, gate ? i11 : -1, __fmul_rn(__fmul_rn(omdy, omdx), g), lane);
      warp_scatter_add(slice, xo ? i11 + 1 : -1, __fmul_rn(__fmul_rn(omdy, dx), g), lane);
      warp_scatter_add(slice, yo ? i11 + W2 : -1, __fmul_rn(__fmul_rn(dy, omdx), g), lane);
      warp_scatter_add(slice, (xo && yo) ? i11 + W2 + 1 : -1, __fmul_rn(__fmul_rn(dy, dx), g), lane);
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
      for (int q = lane; q < Q; q += 32) {
        float v = 0.0f;
        if (q >= lo && q < hi) {
          v = slice[q];
          slice[q] = 0.0f;
        }
        __stcs(G + q, v);
      }
    }
    __syncwarp();
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
  const size_t smem = (size_t)(((TAPS * (kBwTile + 1) + 3) & ~3) + WARPS * Qpad) * sizeof(float);
  auto kern = lookup_bwd_kernel<R, DEFORM, WARPS>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("lookup backward: cannot opt in to %zu B of shared memory: %s", smem, cudaGetErrorString(e));
      return LGU_ERR_LAUNCH;
    }
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
  const long long nblk = (long long)E * ((P + kBwTile - 1) / kBwTile);
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
  return lgu::launch_lookup_bwd<true>(volume, coords, offset, corr_grad, volume_grad, offset_grad, E, H1, W1, H2, W2,
                                      radius, (cudaStream_t)stream);
}
