// lookup_fwd.cu -- corr_index_forward / defCorr_index_forward for sm_100a.
//
// Replaces /root/reference/offersample_LGS/corrSample_kernel.cu:24-82 and
// defCorrSample_kernel.cu:25-91 (one thread per source pixel, 49x4 scalar gathers with a
// 12 KB lane stride, offsets read as 98 strided floats).
//
// Design (gather-bound, HBM roofline; algorithmic bytes per source pixel and level =
// 8 coords + 8*rd^2 offsets + <=16*rd^2 gather + 4*rd^2 out):
//   * a CTA owns a tile of 32 consecutive source pixels of one edge; each of its 8 warps
//     walks 4 of them, ONE WARP PER PIXEL: lane t handles tap t (and t+32), so
//       - the pixel's 8*rd^2-byte offset record is one coalesced float2 load,
//       - all gathers of an instruction fall into the pixel's own H2xW2 slice (<= 16 rows
//         of 64 B instead of 32 slices 12 KB apart) and the four corner loads of
//         neighbouring taps share L1 sectors,
//   * the rd^2 x 32 result tile is transposed through shared memory ([tap][33] padding,
//     conflict-free both ways) and written as 128-byte rows, one per tap plane.
//   * index logic is the reference's, bit for bit: F2I.FLOOR saturating conversion,
//     wrapping tap arithmetic, whole tap gated on the top-left corner (quirk Q3),
//     x2/y2 corners gated individually, centre offset tap zeroed in place (Q5).
#include <cstdlib>
#include "common.cuh"

namespace lgu {

constexpr int kLkWarps = 8;
constexpr int kLkThreads = kLkWarps * 32;
constexpr int kLkTile = 32;                       // source pixels per CTA
constexpr int kLkPixPerWarp = kLkTile / kLkWarps; // 4

// Address / weight set-up of one tap (index logic of defCorrSample_kernel.cu:56-67 and
// corrSample_kernel.cu:52-60, bit for bit): returns the gate and the four corner indices, which are
// always valid addresses inside the pixel's slice (gated-off corners alias the top-left / slice start),
// so that the loads can be issued unconditionally and early.
struct TapAddr {
  float dx, dy;
  int i11, i21, i12, i22;
  bool gate, xo, yo;
};

template <bool DEFORM>
__device__ __forceinline__ TapAddr tap_setup(float x0, float y0, float ox, float oy, int i, int j, int r, int H2,
                                             int W2) {
  TapAddr a;
  int fx, fy;
  if (DEFORM) {
    const float px = __fadd_rn(ox, x0), py = __fadd_rn(oy, y0);   // defCorrSample_kernel.cu:56-57
    fx = floor_to_int(px);
    fy = floor_to_int(py);
    a.dx = __fsub_rn(px, (float)fx);                               // :60-61 (via the int)
    a.dy = __fsub_rn(py, (float)fy);
  } else {
    a.dx = __fsub_rn(x0, floorf(x0));                              // corrSample_kernel.cu:52-53
    a.dy = __fsub_rn(y0, floorf(y0));
    fx = floor_to_int(x0);
    fy = floor_to_int(y0);
  }
  const int x1 = tap_coord(fx, r, i), y1 = tap_coord(fy, r, j);
  const int x2 = wrap_inc(x1), y2 = wrap_inc(y1);
  a.gate = in_bounds(y1, x1, H2, W2);                              // whole tap gated on top-left (Q3)
  a.xo = a.gate && (x2 >= 0 && x2 < W2);
  a.yo = a.gate && (y2 >= 0 && y2 < H2);
  a.i11 = a.gate ? y1 * W2 + x1 : 0;
  a.i21 = a.xo ? a.i11 + 1 : a.i11;
  a.i12 = a.yo ? a.i11 + W2 : a.i11;
  a.i22 = (a.xo && a.yo) ? a.i11 + W2 + 1 : a.i11;
  return a;
}

// One tap of one pixel (generic-radius path): returns the blended value (0 when gated off).
template <bool DEFORM>
__device__ __forceinline__ float sample_tap(const float* __restrict__ V, float x0, float y0, float ox, float oy,
                                            int i, int j, int r, int H2, int W2) {
  const TapAddr a = tap_setup<DEFORM>(x0, y0, ox, oy, i, j, r, H2, W2);
  const float q11 = __ldg(V + a.i11), q21 = __ldg(V + a.i21), q12 = __ldg(V + a.i12), q22 = __ldg(V + a.i22);
  return a.gate ? blend4(q11, a.xo ? q21 : 0.0f, a.yo ? q12 : 0.0f, (a.xo && a.yo) ? q22 : 0.0f, a.dx, a.dy) : 0.0f;
}

// R > 0: compile-time radius.  The warp's 4 pixels x PASSES tap groups are processed in three
// branch-free phases so that every global load of a phase is in flight at once (the kernel is
// latency-bound otherwise: ncu showed 16 serialized ~5k-cycle waits per warp with one phase per tap group):
//   A: coords + offset records (coalesced float2 rows)   B: 4 corner gathers per tap   C: blend + smem transpose
template <int R, bool DEFORM>
__global__ void __launch_bounds__(kLkThreads, 2)
lookup_fwd_kernel(const float* __restrict__ volume, const float* __restrict__ coords, float* __restrict__ offset,
                  float* __restrict__ corr, int P, int W1, int H2, int W2, int tiles_per_edge) {
  constexpr int RD = 2 * R + 1, TAPS = RD * RD, PASSES = (TAPS + 31) / 32;
  constexpr int CENTER = R * RD + R;
  __shared__ float s_out[TAPS][kLkTile + 1];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.x / tiles_per_edge;
  const int p0 = (blockIdx.x - n * tiles_per_edge) * kLkTile;
  const size_t Q = (size_t)H2 * W2;
  const float* cx = coords + (size_t)n * 2 * P;
  const float* cy = cx + P;

  // ---- phase A: coords and offsets of all 4 pixels
  float x0[kLkPixPerWarp], y0[kLkPixPerWarp];
  float2 o[kLkPixPerWarp][PASSES];
  const float* V[kLkPixPerWarp];
#pragma unroll
  for (int k = 0; k < kLkPixPerWarp; ++k) {
    const int p = min(p0 + warp * kLkPixPerWarp + k, P - 1);   // tail pixels recompute the last one (never stored)
    const size_t pix = (size_t)n * P + p;
    x0[k] = __ldg(cx + p);
    y0[k] = __ldg(cy + p);
    V[k] = volume + pix * Q;
    if (DEFORM) {
      float2* O = reinterpret_cast<float2*>(offset) + pix * TAPS;
#pragma unroll
      for (int ps = 0; ps < PASSES; ++ps) {
        const int t = ps * 32 + lane;
        o[k][ps] = O[min(t, TAPS - 1)];
        if (t == CENTER) o[k][ps] = make_float2(0.0f, 0.0f);   // the centre tap reads as zero (Q5) ...
      }
    } else {
#pragma unroll
      for (int ps = 0; ps < PASSES; ++ps) o[k][ps] = make_float2(0.0f, 0.0f);
    }
  }
  // ---- phase B: set up every tap and issue all gathers
  TapAddr ta[kLkPixPerWarp][PASSES];
  float q[kLkPixPerWarp][PASSES][4];
#pragma unroll
  for (int k = 0; k < kLkPixPerWarp; ++k) {
#pragma unroll
    for (int ps = 0; ps < PASSES; ++ps) {
      const int t = min(ps * 32 + lane, TAPS - 1);
      const int i = t / RD, j = t - i * RD;                       // i: x tap, j: y tap (quirk Q1)
      ta[k][ps] = tap_setup<DEFORM>(x0[k], y0[k], o[k][ps].x, o[k][ps].y, i, j, R, H2, W2);
      q[k][ps][0] = __ldg(V[k] + ta[k][ps].i11);
      q[k][ps][1] = __ldg(V[k] + ta[k][ps].i21);
      q[k][ps][2] = __ldg(V[k] + ta[k][ps].i12);
      q[k][ps][3] = __ldg(V[k] + ta[k][ps].i22);
    }
  }
  // ---- phase C: blend, transpose through shared memory
#pragma unroll
  for (int k = 0; k < kLkPixPerWarp; ++k) {
    const int pl = warp * kLkPixPerWarp + k;
#pragma unroll
    for (int ps = 0; ps < PASSES; ++ps) {
      const int t = ps * 32 + lane;
      const TapAddr& a = ta[k][ps];
      const float val = a.gate ? blend4(q[k][ps][0], a.xo ? q[k][ps][1] : 0.0f, a.yo ? q[k][ps][2] : 0.0f,
                                        (a.xo && a.yo) ? q[k][ps][3] : 0.0f, a.dx, a.dy)
                               : 0.0f;
      if (t < TAPS) s_out[t][pl] = val;
    }
  }
  __syncthreads();
  const bool live = (p0 + lane) < P;
  float* out = corr + (size_t)n * TAPS * P + p0 + lane;
#pragma unroll 1
  for (int t = warp; t < TAPS; t += kLkWarps)
    if (live) out[(size_t)t * P] = s_out[t][lane];
  // ... and is zeroed in the caller's tensor LAST: a global store ahead of the gathers of the same warp
  // stalls them (measured 2.7x on the whole kernel), and no other warp reads this pixel's record.
  if (DEFORM && lane < kLkPixPerWarp) {
    const int p = p0 + warp * kLkPixPerWarp + lane;
    if (p < P) reinterpret_cast<float2*>(offset)[((size_t)n * P + p) * TAPS + CENTER] = make_float2(0.0f, 0.0f);
  }
}

// Plain lookup with r = 1 (the uncertainty-mask lookup, corr.py:94): only 9 taps, so warp-per-pixel leaves 23 lanes
// idle.  One THREAD per pixel instead: its 4 x 4 footprint is loaded once (16 bounds-checked loads, zero outside),
// the 9 taps are blended from registers, and consecutive threads store consecutive pixels of each tap plane
// (coalesced).  Index logic as corrSample_kernel.cu:52-77 (top-left gating, Q3).
// PC: altcorr_forward's semantics on a materialised volume (src/altcorr_kernel.cu:27-149: dot products at integer
// positions, each gated on its own, bilinearly splatted = a per-corner-gated bilinear tap, quirk Q4), coords interleaved
// [E,P,2] as that operator receives them.
template <bool PC>
__global__ void __launch_bounds__(128)
lookup_fwd_r1_kernel(const float* __restrict__ volume, const float* __restrict__ coords, float* __restrict__ corr,
                     int P, int H2, int W2, long long npix) {
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= npix) return;
  const int n = (int)(pix / P), p = (int)(pix - (long long)n * P);
  float x0, y0;
  if (PC) {
    const float2 c = __ldg(reinterpret_cast<const float2*>(coords) + pix);
    x0 = c.x; y0 = c.y;
  } else {
    x0 = __ldg(coords + (size_t)n * 2 * P + p); y0 = __ldg(coords + (size_t)n * 2 * P + P + p);
  }
  const float dx = __fsub_rn(x0, floorf(x0)), dy = __fsub_rn(y0, floorf(y0));
  const int fx = floor_to_int(x0), fy = floor_to_int(y0);
  const float* V = volume + (size_t)pix * H2 * W2;
  int xs[4], ys[4];
  xs[0] = tap_coord(fx, 1, 0); ys[0] = tap_coord(fy, 1, 0);
#pragma unroll
  for (int k = 1; k < 4; ++k) { xs[k] = wrap_inc(xs[k - 1]); ys[k] = wrap_inc(ys[k - 1]); }
  float q[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
      q[a][b] = in_bounds(ys[a], xs[b], H2, W2) ? __ldg(V + (long long)ys[a] * W2 + xs[b]) : 0.0f;
  float* out = corr + (size_t)n * 9 * P + p;
#pragma unroll
  for (int i = 0; i < 3; ++i)          // x tap
#pragma unroll
    for (int j = 0; j < 3; ++j) {      // y tap
      const bool gate = PC ? true : in_bounds(ys[j], xs[i], H2, W2);      // out-of-bounds corners already read as 0
      out[(size_t)(i * 3 + j) * P] = gate ? blend4(q[j][i], q[j][i + 1], q[j + 1][i], q[j + 1][i + 1], dx, dy) : 0.0f;
    }
}

int launch_r1_pc(const float* volume, const float* coords, float* corr, int B, int P, int H2, int W2, cudaStream_t st) {
  const long long npix = (long long)B * P;
  lookup_fwd_r1_kernel<true><<<(unsigned)((npix + 127) / 128), 128, 0, st>>>(volume, coords, corr, P, H2, W2, npix);
  return check_launch("lgu_altcorr_forward(volume)");
}

// Any radius (slow path, rarely used): same mapping, direct strided stores.
template <bool DEFORM>
__global__ void __launch_bounds__(kLkThreads)
lookup_fwd_generic_kernel(const float* __restrict__ volume, const float* __restrict__ coords,
                          float* __restrict__ offset, float* __restrict__ corr, int r, int P, int W1, int H2,
                          int W2, long long npix) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = ((long long)blockIdx.x * kLkThreads + threadIdx.x) >> 5;
  if (warp_global >= npix) return;
  const int rd = 2 * r + 1, taps = rd * rd;
  const long long pix = warp_global;
  const int n = (int)(pix / P), p = (int)(pix - (long long)n * P);
  const float x0 = __ldg(coords + (size_t)n * 2 * P + p), y0 = __ldg(coords + (size_t)n * 2 * P + P + p);
  const float* V = volume + (size_t)pix * H2 * W2;
  float2* O = DEFORM ? reinterpret_cast<float2*>(offset) + (size_t)pix * taps : nullptr;
  for (int t = lane; t < taps; t += 32) {
    const int i = t / rd, j = t - i * rd;
    float2 o = make_float2(0.0f, 0.0f);
    if (DEFORM) {
      if (t == r * rd + r) O[t] = o;
      else o = O[t];
    }
    corr[((size_t)n * taps + t) * P + p] = sample_tap<DEFORM>(V, x0, y0, o.x, o.y, i, j, r, H2, W2);
  }
}

template <bool DEFORM>
static int launch_lookup_fwd(const float* volume, const float* coords, float* offset, float* corr, int E, int H1,
                             int W1, int H2, int W2, int r, cudaStream_t st) {
  const int P = H1 * W1;
  const int tiles = (P + kLkTile - 1) / kLkTile;
  const long long nblk = (long long)E * tiles;
  LGU_REQUIRE(nblk < 2147483647LL, "lookup forward: E*ceil(H1*W1/32) = %lld exceeds the grid limit", nblk);
  const dim3 grid((unsigned)nblk), block(kLkThreads);
  switch (r) {
    case 1:
      if (!DEFORM) {
        const long long npix = (long long)E * P;
        lookup_fwd_r1_kernel<false><<<(unsigned)((npix + 127) / 128), 128, 0, st>>>(volume, coords, corr, P, H2, W2, npix);
        break;
      }
      lookup_fwd_kernel<1, DEFORM><<<grid, block, 0, st>>>(volume, coords, offset, corr, P, W1, H2, W2, tiles);
      break;
    case 2: lookup_fwd_kernel<2, DEFORM><<<grid, block, 0, st>>>(volume, coords, offset, corr, P, W1, H2, W2, tiles); break;
    case 3: lookup_fwd_kernel<3, DEFORM><<<grid, block, 0, st>>>(volume, coords, offset, corr, P, W1, H2, W2, tiles); break;
    case 4: lookup_fwd_kernel<4, DEFORM><<<grid, block, 0, st>>>(volume, coords, offset, corr, P, W1, H2, W2, tiles); break;
    default: {
      const long long npix = (long long)E * P;
      const long long nb = (npix * 32 + kLkThreads - 1) / kLkThreads;
      LGU_REQUIRE(nb < 2147483647LL, "lookup forward: too many pixels for one launch (%lld)", npix);
      lookup_fwd_generic_kernel<DEFORM><<<(unsigned)nb, kLkThreads, 0, st>>>(volume, coords, offset, corr, r, P, W1,
                                                                            H2, W2, npix);
    }
  }
  return check_launch(DEFORM ? "lgu_defcorr_index_forward" : "lgu_corr_index_forward");
}

}  // namespace lgu

extern "C" int lgu_corr_index_forward(const float* volume, const float* coords, float* corr, int E, int H1, int W1,
                                      int H2, int W2, int radius, void* stream) {
  if (E == 0) return LGU_OK;   // empty edge set: nothing to do (pointers may be null)
  LGU_REQUIRE(volume && coords && corr, "lgu_corr_index_forward: null pointer");
  LGU_REQUIRE(E >= 0 && H1 > 0 && W1 > 0 && H2 > 0 && W2 > 0 && radius >= 0,
              "lgu_corr_index_forward: bad sizes E=%d H1=%d W1=%d H2=%d W2=%d r=%d", E, H1, W1, H2, W2, radius);
  LGU_REQUIRE((long long)H2 * W2 < (1LL << 30), "lgu_corr_index_forward: H2*W2 too large for 32-bit slice offsets");
  return lgu::launch_lookup_fwd<false>(volume, coords, nullptr, corr, E, H1, W1, H2, W2, radius,
                                       (cudaStream_t)stream);
}

namespace lgu {
int launch_level_tma(bool bwd, const float* volume, const float* coords, float* offset, float* corr,
                     const float* corr_grad, float* volume_grad, float* offset_grad, int E, int H1, int W1, int H2,
                     int W2, cudaStream_t st);   // lookup_level_tma.cu
}
extern "C" int lgu_defcorr_index_forward(const float* volume, const float* coords, float* offset, float* corr, int E,
                                         int H1, int W1, int H2, int W2, int radius, void* stream) {
  if (E == 0) return LGU_OK;   // empty edge set: nothing to do (pointers may be null)
  LGU_REQUIRE(volume && coords && offset && corr, "lgu_defcorr_index_forward: null pointer");
  LGU_REQUIRE(E >= 0 && H1 > 0 && W1 > 0 && H2 > 0 && W2 > 0 && radius >= 0,
              "lgu_defcorr_index_forward: bad sizes E=%d H1=%d W1=%d H2=%d W2=%d r=%d", E, H1, W1, H2, W2, radius);
  LGU_REQUIRE((long long)H2 * W2 < (1LL << 30), "lgu_defcorr_index_forward: H2*W2 too large for 32-bit slice offsets");
  if (radius == 3) {   // the reference's only deformable radius: TMA-staged footprints (lookup_level_tma.cu)
    const int rc = lgu::launch_level_tma(false, volume, coords, offset, corr, nullptr, nullptr, nullptr, E, H1, W1, H2,
                                         W2, (cudaStream_t)stream);
    if (rc >= 0) return rc;
  }
  return lgu::launch_lookup_fwd<true>(volume, coords, offset, corr, E, H1, W1, H2, W2, radius, (cudaStream_t)stream);
}
