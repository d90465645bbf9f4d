// gaussian.cu -- gaussianMask forward / backward (stand-alone drop-in variants) for sm_100a.
//
// Replaces /root/reference/offersample_LGS/gaussianAttn.cu:19-68,134-163 (forward: dense
// zeros_like + 81 scattered 4-byte writes per pixel) and :72-131,165-200 (backward: 81x2
// scalar gathers per thread, accumulators RMW'd in global memory).
//
// Forward is bound by its dense output (4*H2*W2 B per source pixel): ONE WARP PER PIXEL streams
// the pixel's whole H2xW2 slice with 16-byte stores; only the chunks that intersect the
// (2r+1)^2 window read the volume and evaluate the Gaussian, everything else is a zero store.
// Backward: lane t owns window tap t, four partial sums are warp-reduced with shuffles.
// Arithmetic follows the reference's SASS: IEEE divides, s = fma(t2, dy, t1*dx), f = -0.5*s,
// full-precision expf, ((V*3)*e); fp64 only for the covariance-gradient factor (:120,122).
#include "common.cuh"
#include "gauss_window.cuh"

namespace lgu {

constexpr int kGaWarps = 8;

__device__ __forceinline__ float gauss_weight(int x1, int y1, float mx, float my, float c1, float c2) {
  const float ddx = __fsub_rn((float)x1, mx), ddy = __fsub_rn((float)y1, my);
  const float t1 = __fdiv_rn(ddx, c1), t2 = __fdiv_rn(ddy, c2);
  const float s = __fmaf_rn(ddy, t2, __fmul_rn(t1, ddx));
  return expf(__fmul_rn(s, -0.5f));
}

constexpr int kGaZeroBytes = 8192;

template <bool BULK>
__global__ void __launch_bounds__(kGaWarps * 32)
gaussian_fwd_kernel(const float* __restrict__ means, const float* __restrict__ covs,
                    const float* __restrict__ volume, float* __restrict__ out, long long npix, int H2, int W2,
                    int r) {
  __shared__ __align__(16) uint8_t zero_smem[BULK ? kGaZeroBytes : 16];
  if (BULK) {
    for (int q = threadIdx.x; q < kGaZeroBytes / 16; q += blockDim.x)
      reinterpret_cast<float4*>(zero_smem)[q] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int Q = H2 * W2;
  const unsigned rd = 2u * (unsigned)r + 1u;
  for (long long pix = wid; pix < npix; pix += nwarps) {
    const float2 m = __ldg(reinterpret_cast<const float2*>(means) + pix);
    const float2 c = __ldg(reinterpret_cast<const float2*>(covs) + pix);
    const int cx = floor_to_int(m.x), cy = floor_to_int(m.y);
    // window membership with the reference's wrapping arithmetic: x1 == cx - r + i for some 0 <= i < rd
    const unsigned bx = (unsigned)cx - (unsigned)r, by = (unsigned)cy - (unsigned)r;
    const float* V = volume + (size_t)pix * Q;
    float* O = out + (size_t)pix * Q;
    if ((W2 & 3) == 0 && ((reinterpret_cast<uintptr_t>(O) | reinterpret_cast<uintptr_t>(V)) & 15) == 0) {
      const int W4 = W2 >> 2;
      int q4_begin = 0, q4_end = Q >> 2;
      if (BULK) {
        // rows that miss the window are pure zeros: they leave as cp.async.bulk copies from the CTA's zero buffer
        // (one elected lane); only the band of window rows is evaluated and stored by the lanes
        long long yb0 = (long long)cy - r, yb1 = yb0 + rd;       // cy may be a saturated conversion: 64-bit
        int y0 = (int)(yb0 < 0 ? 0 : (yb0 > H2 ? H2 : yb0)), y1 = (int)(yb1 < 0 ? 0 : (yb1 > H2 ? H2 : yb1));
        if (y1 <= y0) y0 = y1 = H2;
        if (lane == 0) {
          float* d = O;
          for (int bytes = y0 * W2 * 4; bytes > 0;) {
            const int nb = bytes < kGaZeroBytes ? bytes : kGaZeroBytes;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(d),
                         "r"((uint32_t)__cvta_generic_to_shared(zero_smem)), "r"(nb) : "memory");
            d += nb >> 2; bytes -= nb;
          }
          d = O + (size_t)y1 * W2;
          for (int bytes = (H2 - y1) * W2 * 4; bytes > 0;) {
            const int nb = bytes < kGaZeroBytes ? bytes : kGaZeroBytes;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(d),
                         "r"((uint32_t)__cvta_generic_to_shared(zero_smem)), "r"(nb) : "memory");
            d += nb >> 2; bytes -= nb;
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        q4_begin = y0 * W4;
        q4_end = y1 * W4;
      }
      for (int q4 = q4_begin + lane; q4 < q4_end; q4 += 32) {
        const int y1 = q4 / W4, x1 = (q4 - y1 * W4) << 2;
        float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        const bool row_in = ((unsigned)y1 - by) < rd;
        // any of x1..x1+3 inside [bx, bx+rd) (mod 2^32)?
        const bool col_in = (((unsigned)x1 - bx) < rd) || (((unsigned)(x1 + 3) - bx) < rd) ||
                            ((bx - (unsigned)x1) < 4u);
        if (row_in && col_in) {
          const float4 in = __ldg(reinterpret_cast<const float4*>(V) + q4);
          if (((unsigned)(x1 + 0) - bx) < rd) v.x = __fmul_rn(__fmul_rn(in.x, 3.0f), gauss_weight(x1 + 0, y1, m.x, m.y, c.x, c.y));
          if (((unsigned)(x1 + 1) - bx) < rd) v.y = __fmul_rn(__fmul_rn(in.y, 3.0f), gauss_weight(x1 + 1, y1, m.x, m.y, c.x, c.y));
          if (((unsigned)(x1 + 2) - bx) < rd) v.z = __fmul_rn(__fmul_rn(in.z, 3.0f), gauss_weight(x1 + 2, y1, m.x, m.y, c.x, c.y));
          if (((unsigned)(x1 + 3) - bx) < rd) v.w = __fmul_rn(__fmul_rn(in.w, 3.0f), gauss_weight(x1 + 3, y1, m.x, m.y, c.x, c.y));
        }
        __stcs(reinterpret_cast<float4*>(O) + q4, v);
      }
    } else {
      for (int q = lane; q < Q; q += 32) {
        const int y1 = q / W2, x1 = q - y1 * W2;
        float v = 0.0f;
        if (((unsigned)y1 - by) < rd && ((unsigned)x1 - bx) < rd)
          v = __fmul_rn(__fmul_rn(__ldg(V + q), 3.0f), gauss_weight(x1, y1, m.x, m.y, c.x, c.y));
        __stcs(O + q, v);
      }
    }
  }
  if (BULK) {                                                    // the zero buffer must outlive the engine's reads
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kGaWarps * 32)
gaussian_bwd_kernel(const float* __restrict__ means, const float* __restrict__ covs,
                    const float* __restrict__ volume, const float* __restrict__ out_grad,
                    float* __restrict__ means_grad, float* __restrict__ covs_grad, long long npix, int H2, int W2,
                    int r) {
  const int lane = threadIdx.x & 31;
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int Q = H2 * W2;
  const int rd = 2 * r + 1, taps = rd * rd;
  for (long long pix = wid; pix < npix; pix += nwarps) {
    const float2 m = __ldg(reinterpret_cast<const float2*>(means) + pix);
    const float2 c = __ldg(reinterpret_cast<const float2*>(covs) + pix);
    const int cx = floor_to_int(m.x), cy = floor_to_int(m.y);
    const float* V = volume + (size_t)pix * Q;
    const float* G = out_grad + (size_t)pix * Q;
    float gm0 = 0.0f, gm1 = 0.0f, gc0 = 0.0f, gc1 = 0.0f;
    // lanes sweep the window row-major in memory (y outer, x inner) so a warp instruction touches ~3.5 consecutive
    // 36-byte row segments.  All passes' loads are issued before any is consumed (the kernel was stalling once per
    // pass on the volume / gradient gathers); radius <= 4 -> at most 3 passes, larger windows take the tail loop.
    constexpr int kHoist = 3;
    float v3s[kHoist], gs[kHoist];
    int xs_[kHoist], ys_[kHoist];
    bool ok[kHoist];
#pragma unroll
    for (int ps = 0; ps < kHoist; ++ps) {
      const int t = ps * 32 + lane;
      const int tt = t < taps ? t : 0;
      const int j = tt / rd, i = tt - j * rd;          // j: y tap, i: x tap
      xs_[ps] = tap_coord(cx, r, i);
      ys_[ps] = tap_coord(cy, r, j);
      ok[ps] = t < taps && in_bounds(ys_[ps], xs_[ps], H2, W2);
      const int q = ok[ps] ? ys_[ps] * W2 + xs_[ps] : 0;
      v3s[ps] = __ldg(V + q);
      gs[ps] = __ldg(G + q);
    }
    // per-pixel reciprocals (the reference divides per tap; multiplying by the correctly rounded reciprocal differs
    // by <= 1 ulp per factor, far inside the 1e-5 gradient tolerance, and removes 4 fp32 + 2 fp64 divisions per tap)
    const float rcx = __fdiv_rn(1.0f, c.x), rcy = __fdiv_rn(1.0f, c.y);
    const double rccx = 1.0 / (double)__fmul_rn(c.x, c.x), rccy = 1.0 / (double)__fmul_rn(c.y, c.y);
    auto accumulate = [&](int x1, int y1, float vraw, float g) {
      const float ddx = __fsub_rn((float)x1, m.x), ddy = __fsub_rn((float)y1, m.y);
      const float t1 = __fmul_rn(ddx, rcx), t2 = __fmul_rn(ddy, rcy);
      const float s = __fmaf_rn(ddy, t2, __fmul_rn(t1, ddx));
      const float e = expf(__fmul_rn(s, -0.5f));
      const float v3 = __fmul_rn(vraw, 3.0f);
      gm0 = __fmaf_rn(__fmul_rn(v3, __fmul_rn(__fmul_rn(ddx, e), rcx)), g, gm0);   // gaussianAttn.cu:117
      gm1 = __fmaf_rn(__fmul_rn(v3, __fmul_rn(__fmul_rn(ddy, e), rcy)), g, gm1);   // :118
      const double eh = (double)e * 0.5;                                           // :120,122 (fp64)
      const float dE1 = (float)(((eh * (double)ddx) * (double)ddx) * rccx);
      const float dE2 = (float)(((eh * (double)ddy) * (double)ddy) * rccy);
      gc0 = __fmaf_rn(__fmul_rn(dE1, v3), g, gc0);                                 // :125
      gc1 = __fmaf_rn(__fmul_rn(dE2, v3), g, gc1);                                 // :126
    };
#pragma unroll
    for (int ps = 0; ps < kHoist; ++ps)
      if (ok[ps]) accumulate(xs_[ps], ys_[ps], v3s[ps], gs[ps]);
    for (int t = kHoist * 32 + lane; t < taps; t += 32) {
      const int j = t / rd, i = t - j * rd;
      const int x1 = tap_coord(cx, r, i), y1 = tap_coord(cy, r, j);
      if (in_bounds(y1, x1, H2, W2)) {
        const int q = y1 * W2 + x1;
        accumulate(x1, y1, __ldg(V + q), __ldg(G + q));
      }
    }
    gm0 = warp_sum(gm0); gm1 = warp_sum(gm1); gc0 = warp_sum(gc0); gc1 = warp_sum(gc1);
    if (lane == 0) {
      reinterpret_cast<float2*>(means_grad)[pix] = make_float2(gm0, gm1);
      reinterpret_cast<float2*>(covs_grad)[pix] = make_float2(gc0, gc1);
    }
  }
}

// Gaussian-head gradients of the fused build, from the four level gradients (see lgu_build_backward_gauss in
// include/lgu_corr.h).  Same tap arithmetic as gaussian_bwd_kernel; the merged upstream gradient and the raw volume
// are formed per tap from lvl0 and the level gradients (5 gathers per tap instead of two dense passes per edge).
__global__ void __launch_bounds__(kGaWarps * 32)
build_bwd_gauss_kernel(const float* __restrict__ means, const float* __restrict__ covs, const float* __restrict__ den,
                       const float* __restrict__ lvl0, const float* __restrict__ g0, const float* __restrict__ g1,
                       const float* __restrict__ g2, const float* __restrict__ g3, float* __restrict__ means_grad,
                       float* __restrict__ covs_grad, float* __restrict__ den_grad, long long npix, int H2, int W2,
                       int r) {
  const int lane = threadIdx.x & 31;
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int Q = H2 * W2;
  const int rd = 2 * r + 1, taps = rd * rd;
  for (long long pix = wid; pix < npix; pix += nwarps) {
    const float2 m = __ldg(reinterpret_cast<const float2*>(means) + pix);
    const float2 c = __ldg(reinterpret_cast<const float2*>(covs) + pix);
    const float dn = __ldg(den + pix);
    const int cx = floor_to_int(m.x), cy = floor_to_int(m.y);
    const float* L0 = lvl0 + (size_t)pix * Q;
    const float* G0 = g0 != nullptr ? g0 + (size_t)pix * Q : nullptr;
    const float* G1 = g1 != nullptr ? g1 + (size_t)pix * (Q >> 2) : nullptr;
    const float* G2 = g2 != nullptr ? g2 + (size_t)pix * (Q >> 4) : nullptr;
    const float* G3 = g3 != nullptr ? g3 + (size_t)pix * (Q >> 6) : nullptr;
    const float rcx = __fdiv_rn(1.0f, c.x), rcy = __fdiv_rn(1.0f, c.y);
    const double rccx = 1.0 / (double)__fmul_rn(c.x, c.x), rccy = 1.0 / (double)__fmul_rn(c.y, c.y);
    float gm0 = 0.0f, gm1 = 0.0f, gc0 = 0.0f, gc1 = 0.0f, gd = 0.0f;
    for (int t = lane; t < taps; t += 32) {
      const int j = t / rd, i = t - j * rd;
      const int x1 = tap_coord(cx, r, i), y1 = tap_coord(cy, r, j);
      if (!in_bounds(y1, x1, H2, W2)) continue;
      const float l0 = __ldg(L0 + y1 * W2 + x1);
      float g = G0 != nullptr ? __ldg(G0 + y1 * W2 + x1) : 0.0f;                 // avg_pool2d^T, level by level
      if (G1 != nullptr) g = __fadd_rn(g, __fmul_rn(__ldg(G1 + (y1 >> 1) * (W2 >> 1) + (x1 >> 1)), 0.25f));
      if (G2 != nullptr) g = __fadd_rn(g, __fmul_rn(__ldg(G2 + (y1 >> 2) * (W2 >> 2) + (x1 >> 2)), 0.0625f));
      if (G3 != nullptr) g = __fadd_rn(g, __fmul_rn(__ldg(G3 + (y1 >> 3) * (W2 >> 3) + (x1 >> 3)), 0.015625f));
      const float ddx = __fsub_rn((float)x1, m.x), ddy = __fsub_rn((float)y1, m.y);
      const float t1 = __fmul_rn(ddx, rcx), t2 = __fmul_rn(ddy, rcy);
      const float s = __fmaf_rn(ddy, t2, __fmul_rn(t1, ddx));
      const float e = expf(__fmul_rn(s, -0.5f));
      const float vraw = __fdiv_rn(l0, __fadd_rn(1.0f, __fdiv_rn(__fmul_rn(3.0f, e), dn)));   // lvl0 = V (1 + 3 e / den)
      const float gq = __fdiv_rn(g, dn);
      const float v3 = __fmul_rn(vraw, 3.0f);
      gm0 = __fmaf_rn(__fmul_rn(v3, __fmul_rn(__fmul_rn(ddx, e), rcx)), gq, gm0);
      gm1 = __fmaf_rn(__fmul_rn(v3, __fmul_rn(__fmul_rn(ddy, e), rcy)), gq, gm1);
      const double eh = (double)e * 0.5;
      const float dE1 = (float)(((eh * (double)ddx) * (double)ddx) * rccx);
      const float dE2 = (float)(((eh * (double)ddy) * (double)ddy) * rccy);
      gc0 = __fmaf_rn(__fmul_rn(dE1, v3), gq, gc0);
      gc1 = __fmaf_rn(__fmul_rn(dE2, v3), gq, gc1);
      gd = __fmaf_rn(g, __fsub_rn(l0, vraw), gd);
    }
    gm0 = warp_sum(gm0); gm1 = warp_sum(gm1); gc0 = warp_sum(gc0); gc1 = warp_sum(gc1); gd = warp_sum(gd);
    if (lane == 0) {
      reinterpret_cast<float2*>(means_grad)[pix] = make_float2(gm0, gm1);
      reinterpret_cast<float2*>(covs_grad)[pix] = make_float2(gc0, gc1);
      den_grad[pix] = -__fdiv_rn(gd, dn);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Round-2 backward for the reference's only radius (r = 4, a 9 x 9 window).  The round-1 kernels evaluated a
// full-precision expf, two IEEE divides and ~10 fp64 operations PER TAP and were issue-bound (68 % issue-active at 16 %
// of the HBM roofline).  The Gaussian is separable, e(i,j) = ex(i) * ey(j), and so is every factor the reference forms
// from it (gaussianAttn.cu:112-126):
//     d/dmean_x : 3 V g * [ex(i) ddx(i) / cx] * ey(j)          d/dcov_x : 3 V g * [0.5 ex(i) ddx(i)^2 / cx^2] * ey(j)
//     d/dmean_y : 3 V g * ex(i) * [ey(j) ddy(j) / cy]          d/dcov_y : 3 V g * ex(i) * [0.5 ey(j) ddy(j)^2 / cy^2]
// Lanes 0-8 evaluate the 9 COLUMN records, lanes 9-17 the 9 ROW records (one expf each; the fp64 factor of the reference,
// :120,122, 18 times per pixel instead of 162); records of out-of-range columns / rows are zero, which is the reference's
// bounds gate.  Lane = tap as before (coalesced 36-byte row segments: a lane = row mapping was measured and is bound by
// L1 wavefronts, 27 lines per load), the two records of a tap arrive by shuffle, a tap costs 2 multiplies and 4 FMAs,
// and the four sums are reduced with a value-splitting butterfly (6 shuffles instead of 20).  Differences from the
// reference's per-tap evaluation are products of correctly rounded factors in another order (a few 1e-7 relative),
// inside the 1e-5 gradient bar.  FUSED: the upstream gradient and the raw volume are formed per tap from lvl0 and the
// four level gradients (lgu_build_backward_gauss).
// The kernel is LATENCY-bound otherwise (two dependent DRAM round trips per pixel -- parameters, then the window -- with
// ~40 resident warps per SM: measured 86 us at E = 48 whatever the instruction count), so a warp takes NP = 4
// consecutive pixels per trip: lanes 0..3 fetch the four parameter records together and the 4 x 3 gather passes are all
// in flight before the first tap is consumed.
// MODE 0: drop-in gaussianMask backward (g0 = upstream gradient of the masked volume, vol = raw volume);
// MODE 1: from the four level gradients (FUSED above); MODE 2: as 1, but g0 is the [npix, 81] WINDOW RECORD of the merged
// gradient that lgu_corr_lookup_fused_backward_win emits -- one contiguous 324-byte read per pixel instead of four strided
// window gathers (a 64-byte DRAM granule per 36-byte row: 513 MB per launch at E = 48 against 124 MB of algorithmic bytes).
template <int MODE, int NP, int MINB = 2>
__global__ void __launch_bounds__(kGaWarps * 32, MINB)
gaussian_bwd_sep_kernel(const float* __restrict__ means, const float* __restrict__ covs, const float* __restrict__ den,
                        const float* __restrict__ vol, const float* __restrict__ g0, const float* __restrict__ g1,
                        const float* __restrict__ g2, const float* __restrict__ g3, float* __restrict__ means_grad,
                        float* __restrict__ covs_grad, float* __restrict__ den_grad, long long npix, int H2, int W2) {
  constexpr int r = 4, rd = 9, taps = 81, kPasses = 3;
  constexpr bool FUSED = MODE != 0;
  const int lane = threadIdx.x & 31;
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int Q = H2 * W2;
  // per-lane tap constants of the three passes
  int ti[kPasses], tj[kPasses];
#pragma unroll
  for (int ps = 0; ps < kPasses; ++ps) {
    const int t = min(ps * 32 + lane, taps - 1);
    tj[ps] = t / rd;
    ti[ps] = t - tj[ps] * rd;
  }
  for (long long pb = wid * NP; pb < npix; pb += nwarps * NP) {
    // ---- parameters of the NP pixels: lane q loads pixel pb + q (the tail repeats the last pixel; its store is skipped)
    const long long pmine = min(pb + (lane < NP ? lane : 0), npix - 1);
    const float2 m_l = __ldg(reinterpret_cast<const float2*>(means) + pmine);
    const float2 c_l = __ldg(reinterpret_cast<const float2*>(covs) + pmine);
    const float d_l = FUSED ? __ldg(den + pmine) : 1.0f;
    float2 m[NP], c[NP];
    float dn[NP];
    int x0[NP], y0[NP];
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      m[q] = make_float2(__shfl_sync(0xffffffffu, m_l.x, q), __shfl_sync(0xffffffffu, m_l.y, q));
      c[q] = make_float2(__shfl_sync(0xffffffffu, c_l.x, q), __shfl_sync(0xffffffffu, c_l.y, q));
      dn[q] = __shfl_sync(0xffffffffu, d_l, q);
      x0[q] = tap_coord(floor_to_int(m[q].x), r, 0);
      y0[q] = tap_coord(floor_to_int(m[q].y), r, 0);
    }
    // ---- all gathers of all NP pixels (clamped addresses: the records below zero what the reference's bounds test drops)
    float v[NP][kPasses], g[NP][kPasses];
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      const long long pix = min(pb + q, npix - 1);
#pragma unroll
      for (int ps = 0; ps < kPasses; ++ps) {
        const int xc = min(max(tap_coord(x0[q], 0, ti[ps]), 0), W2 - 1), yc = min(max(tap_coord(y0[q], 0, tj[ps]), 0), H2 - 1);
        const size_t at = (size_t)pix * Q + (size_t)yc * W2 + xc;
        v[q][ps] = __ldg(vol + at);
        float gg;
        if (MODE == 2) gg = __ldg(g0 + (size_t)pix * taps + min(ps * 32 + lane, taps - 1));
        else gg = g0 != nullptr ? __ldg(g0 + at) : 0.0f;
        if (MODE == 1) {                                        // avg_pool2d^T, level by level
          if (g1 != nullptr) gg = __fadd_rn(gg, __fmul_rn(__ldg(g1 + (size_t)pix * (Q >> 2) + (yc >> 1) * (W2 >> 1) + (xc >> 1)), 0.25f));
          if (g2 != nullptr) gg = __fadd_rn(gg, __fmul_rn(__ldg(g2 + (size_t)pix * (Q >> 4) + (yc >> 2) * (W2 >> 2) + (xc >> 2)), 0.0625f));
          if (g3 != nullptr) gg = __fadd_rn(gg, __fmul_rn(__ldg(g3 + (size_t)pix * (Q >> 6) + (yc >> 3) * (W2 >> 3) + (xc >> 3)), 0.015625f));
        }
        g[q][ps] = gg;
      }
    }
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      float o0, o1, o2, o3, od;
      gauss_window_grads<FUSED>(v[q], g[q], m[q], c[q], dn[q], x0[q], y0[q], H2, W2, lane, o0, o1, o2, o3, od);
      if (lane == 0 && pb + q < npix) {
        if (FUSED) den_grad[pb + q] = od;
        reinterpret_cast<float2*>(means_grad)[pb + q] = make_float2(o0, o1);
        reinterpret_cast<float2*>(covs_grad)[pb + q] = make_float2(o2, o3);
      }
    }
  }
}

static inline unsigned grid_for_warps(long long nwarps_needed, int warps_per_cta, int ctas_per_sm) {
  long long want = (nwarps_needed + warps_per_cta - 1) / warps_per_cta;
  long long cap = (long long)kNumSMs * ctas_per_sm;
  return (unsigned)(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace lgu

extern "C" int lgu_gaussian_mask_forward(const float* means, const float* covs, const float* volume, float* volume1,
                                         int E, int H1, int W1, int H2, int W2, int radius, void* stream) {
  if (E == 0) return LGU_OK;   // empty edge set: nothing to do (pointers may be null)
  LGU_REQUIRE(means && covs && volume && volume1, "lgu_gaussian_mask_forward: null pointer");
  LGU_REQUIRE(E >= 0 && H1 > 0 && W1 > 0 && H2 > 0 && W2 > 0 && radius >= 0,
              "lgu_gaussian_mask_forward: bad sizes E=%d H1=%d W1=%d H2=%d W2=%d r=%d", E, H1, W1, H2, W2, radius);
  LGU_REQUIRE((long long)H2 * W2 < (1LL << 30), "lgu_gaussian_mask_forward: H2*W2 too large");
  const long long npix = (long long)E * H1 * W1;
  const unsigned grid = lgu::grid_for_warps(npix, lgu::kGaWarps, 8);
  // rows outside the window leave through the TMA engine when the slice rows are 16-byte multiples and aligned
  const bool bulk = (W2 & 3) == 0 && ((reinterpret_cast<uintptr_t>(volume1) | reinterpret_cast<uintptr_t>(volume)) & 15) == 0 &&
                    (((long long)H2 * W2) & 3) == 0;
  if (bulk)
    lgu::gaussian_fwd_kernel<true><<<grid, lgu::kGaWarps * 32, 0, (cudaStream_t)stream>>>(means, covs, volume, volume1,
                                                                                         npix, H2, W2, radius);
  else
    lgu::gaussian_fwd_kernel<false><<<grid, lgu::kGaWarps * 32, 0, (cudaStream_t)stream>>>(means, covs, volume, volume1,
                                                                                          npix, H2, W2, radius);
  return lgu::check_launch("lgu_gaussian_mask_forward");
}

extern "C" int lgu_gaussian_mask_backward(const float* means, const float* covs, const float* volume,
                                          const float* volume1_grad, float* means_grad, float* covs_grad, int E,
                                          int H1, int W1, int H2, int W2, int radius, void* stream) {
  if (E == 0) return LGU_OK;   // empty edge set: nothing to do (pointers may be null)
  LGU_REQUIRE(means && covs && volume && volume1_grad && means_grad && covs_grad,
              "lgu_gaussian_mask_backward: null pointer");
  LGU_REQUIRE(E >= 0 && H1 > 0 && W1 > 0 && H2 > 0 && W2 > 0 && radius >= 0,
              "lgu_gaussian_mask_backward: bad sizes E=%d H1=%d W1=%d H2=%d W2=%d r=%d", E, H1, W1, H2, W2, radius);
  LGU_REQUIRE((long long)H2 * W2 < (1LL << 30), "lgu_gaussian_mask_backward: H2*W2 too large");
  const long long npix = (long long)E * H1 * W1;
  if (radius == 4) {                                             // the reference's only radius: lane = (pixel, window row)
    const unsigned grid3 = lgu::grid_for_warps((npix + 3) / 4, lgu::kGaWarps, 8);
    lgu::gaussian_bwd_sep_kernel<0, 4><<<grid3, lgu::kGaWarps * 32, 0, (cudaStream_t)stream>>>(
        means, covs, nullptr, volume, volume1_grad, nullptr, nullptr, nullptr, means_grad, covs_grad, nullptr, npix, H2, W2);
    return lgu::check_launch("lgu_gaussian_mask_backward");
  }
  const unsigned grid = lgu::grid_for_warps(npix, lgu::kGaWarps, 8);
  lgu::gaussian_bwd_kernel<<<grid, lgu::kGaWarps * 32, 0, (cudaStream_t)stream>>>(
      means, covs, volume, volume1_grad, means_grad, covs_grad, npix, H2, W2, radius);
  return lgu::check_launch("lgu_gaussian_mask_backward");
}

extern "C" int lgu_build_backward_gauss(const float* means, const float* covs, const float* den, const float* lvl0,
                                        const float* g0, const float* g1, const float* g2, const float* g3,
                                        float* means_grad, float* covs_grad, float* den_grad, int E, int H, int W,
                                        int radius, void* stream) {
  if (E == 0) return LGU_OK;
  LGU_REQUIRE(means && covs && den && lvl0 && means_grad && covs_grad && den_grad,
              "lgu_build_backward_gauss: null pointer");
  LGU_REQUIRE(E >= 0 && H > 0 && W > 0 && radius >= 0 && (H % 8) == 0 && (W % 8) == 0,
              "lgu_build_backward_gauss: bad sizes E=%d H=%d W=%d r=%d (H, W must be multiples of 8)", E, H, W, radius);
  LGU_REQUIRE((long long)H * W < (1LL << 30), "lgu_build_backward_gauss: H*W too large");
  const long long npix = (long long)E * H * W;
  if (radius == 4) {
    const unsigned grid3 = lgu::grid_for_warps((npix + 3) / 4, lgu::kGaWarps, 8);
    lgu::gaussian_bwd_sep_kernel<1, 4><<<grid3, lgu::kGaWarps * 32, 0, (cudaStream_t)stream>>>(
        means, covs, den, lvl0, g0, g1, g2, g3, means_grad, covs_grad, den_grad, npix, H, W);
    return lgu::check_launch("lgu_build_backward_gauss");
  }
  const unsigned grid = lgu::grid_for_warps(npix, lgu::kGaWarps, 8);
  lgu::build_bwd_gauss_kernel<<<grid, lgu::kGaWarps * 32, 0, (cudaStream_t)stream>>>(
      means, covs, den, lvl0, g0, g1, g2, g3, means_grad, covs_grad, den_grad, npix, H, W, radius);
  return lgu::check_launch("lgu_build_backward_gauss");
}

// Gaussian-head gradients from the window record of the merged level-0 gradient (lgu_corr_lookup_fused_backward_win):
// the same sums as lgu_build_backward_gauss, bit for bit, with one contiguous 324-byte read per pixel where the
// from-levels form gathers four strided windows.  Radius 4 (the reference's only one, gaussianMask_cuda.py:77).
extern "C" int lgu_build_backward_gauss_window(const float* means, const float* covs, const float* den, const float* lvl0,
                                               const float* gwin, float* means_grad, float* covs_grad, float* den_grad,
                                               int E, int H, int W, void* stream) {
  if (E == 0) return LGU_OK;
  LGU_REQUIRE(means && covs && den && lvl0 && gwin && means_grad && covs_grad && den_grad,
              "lgu_build_backward_gauss_window: null pointer");
  LGU_REQUIRE(E >= 0 && H > 0 && W > 0, "lgu_build_backward_gauss_window: bad sizes E=%d H=%d W=%d", E, H, W);
  LGU_REQUIRE((long long)H * W < (1LL << 30), "lgu_build_backward_gauss_window: H*W too large");
  const long long npix = (long long)E * H * W;
  // (occupancy does not matter here: 2 / 3 / 4 / 5 CTAs per SM with 4 or 2 pixels per trip all ran within 7 % of each other)
  const unsigned grid3 = lgu::grid_for_warps((npix + 3) / 4, lgu::kGaWarps, 8);
  lgu::gaussian_bwd_sep_kernel<2, 4><<<grid3, lgu::kGaWarps * 32, 0, (cudaStream_t)stream>>>(
      means, covs, den, lvl0, gwin, nullptr, nullptr, nullptr, means_grad, covs_grad, den_grad, npix, H, W);
  return lgu::check_launch("lgu_build_backward_gauss_window");
}
