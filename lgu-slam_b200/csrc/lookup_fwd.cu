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
#include "common.cuh"

namespace lgu {

constexpr int kLkWarps = 8;
constexpr int kLkThreads = kLkWarps * 32;
constexpr int kLkTile = 32;                       // source pixels per CTA
constexpr int kLkPixPerWarp = kLkTile / kLkWarps; // 4

template <bool DEFORM>
struct TapSample {
  float dx, dy;
  int x1, y1;
};

// One tap of one pixel: returns the blended value (0 when gated off).
template <bool DEFORM>
__device__ __forceinline__ float sample_tap(const float* __restrict__ V, float x0, float y0, float ox, float oy,
                                            int i, int j, int r, int H2, int W2) {
  float dx, dy;
  int fx, fy;
  if (DEFORM) {
    const float px = __fadd_rn(ox, x0), py = __fadd_rn(oy, y0);   // defCorrSample_kernel.cu:56-57
    fx = floor_to_int(px);
    fy = floor_to_int(py);
    dx = __fsub_rn(px, (float)fx);                                 // :60-61 (via the int)
    dy = __fsub_rn(py, (float)fy);
  } else {
    dx = __fsub_rn(x0, floorf(x0));                                // corrSample_kernel.cu:52-53
    dy = __fsub_rn(y0, floorf(y0));
    fx = floor_to_int(x0);
    fy = floor_to_int(y0);
  }
  const int x1 = tap_coord(fx, r, i), y1 = tap_coord(fy, r, j);
  const int x2 = wrap_inc(x1), y2 = wrap_inc(y1);
  if (!in_bounds(y1, x1, H2, W2)) return 0.0f;                     // whole tap gated on top-left (Q3)
  const bool xo = (x2 >= 0 && x2 < W2), yo = (y2 >= 0 && y2 < H2);
  const float* row1 = V + y1 * W2;
  const float* row2 = row1 + W2;
  const float q11 = __ldg(row1 + x1);
  const float q21 = xo ? __ldg(row1 + x2) : 0.0f;
  const float q12 = yo ? __ldg(row2 + x1) : 0.0f;
  const float q22 = (xo && yo) ? __ldg(row2 + x2) : 0.0f;
  return blend4(q11, q21, q12, q22, dx, dy);
}

// R > 0: compile-time radius (taps fully unrolled, smem-transposed coalesced stores).
template <int R, bool DEFORM>
__global__ void __launch_bounds__(kLkThreads)
lookup_fwd_kernel(const float* __restrict__ volume, const float* __restrict__ coords, float* __restrict__ offset,
                  float* __restrict__ corr, int P, int W1, int H2, int W2, int tiles_per_edge) {
  constexpr int RD = 2 * R + 1, TAPS = RD * RD, PASSES = (TAPS + 31) / 32;
  __shared__ float s_out[TAPS][kLkTile + 1];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.x / tiles_per_edge;
  const int p0 = (blockIdx.x - n * tiles_per_edge) * kLkTile;
  const size_t Q = (size_t)H2 * W2;
  const float* cx = coords + (size_t)n * 2 * P;
  const float* cy = cx + P;

#pragma unroll
  for (int k = 0; k < kLkPixPerWarp; ++k) {
    const int pl = warp * kLkPixPerWarp + k;
    const int p = min(p0 + pl, P - 1);          // tail pixels recompute the last one (never stored)
    const size_t pix = (size_t)n * P + p;
    const float x0 = __ldg(cx + p), y0 = __ldg(cy + p);
    const float* V = volume + pix * Q;
    float2* O = DEFORM ? reinterpret_cast<float2*>(offset) + pix * TAPS : nullptr;
#pragma unroll
    for (int ps = 0; ps < PASSES; ++ps) {
      const int t = ps * 32 + lane;
      if (t < TAPS) {
        const int i = t / RD, j = t - i * RD;   // i: x tap, j: y tap (quirk Q1)
        float2 o = make_float2(0.0f, 0.0f);
        if (DEFORM) {
          if (t == R * RD + R) O[t] = o;        // in-place zeroing of the centre tap (Q5)
          else o = O[t];
        }
        s_out[t][pl] = sample_tap<DEFORM>(V, x0, y0, o.x, o.y, i, j, R, H2, W2);
      }
    }
  }
  __syncthreads();
  const bool live = (p0 + lane) < P;
  float* out = corr + (size_t)n * TAPS * P + p0 + lane;
#pragma unroll 1
  for (int t = warp; t < TAPS; t += kLkWarps)
    if (live) out[(size_t)t * P] = s_out[t][lane];
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
    case 1: lookup_fwd_kernel<1, DEFORM><<<grid, block, 0, st>>>(volume, coords, offset, corr, P, W1, H2, W2, tiles); break;
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

extern "C" int lgu_defcorr_index_forward(const float* volume, const float* coords, float* offset, float* corr, int E,
                                         int H1, int W1, int H2, int W2, int radius, void* stream) {
  if (E == 0) return LGU_OK;   // empty edge set: nothing to do (pointers may be null)
  LGU_REQUIRE(volume && coords && offset && corr, "lgu_defcorr_index_forward: null pointer");
  LGU_REQUIRE(E >= 0 && H1 > 0 && W1 > 0 && H2 > 0 && W2 > 0 && radius >= 0,
              "lgu_defcorr_index_forward: bad sizes E=%d H1=%d W1=%d H2=%d W2=%d r=%d", E, H1, W1, H2, W2, radius);
  LGU_REQUIRE((long long)H2 * W2 < (1LL << 30), "lgu_defcorr_index_forward: H2*W2 too large for 32-bit slice offsets");
  return lgu::launch_lookup_fwd<true>(volume, coords, offset, corr, E, H1, W1, H2, W2, radius, (cudaStream_t)stream);
}
