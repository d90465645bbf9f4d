// lowmem.cu -- on-the-fly correlation from feature maps (no volume): lowMem_defSample and
// altcorr_forward for sm_100a.
//
// Replaces /root/reference/offersample_LGS/lowMem_defSample.cu:27-134,137-168 (32-thread
// CTAs, per-tap __syncthreads, 4x128 scalar channel loads per tap per pixel) and
// /root/reference/src/altcorr_kernel.cu:27-149,290-319.
//
// Design v1: a CTA owns 32 consecutive source pixels of one edge, ONE WARP PER PIXEL with the
// CHANNEL axis across lanes: f1[pixel] lives in registers (one float4 per lane per 128
// channels), every corner of every tap is one coalesced 16-byte-per-lane row read of the
// channels-last fmap2 (512 B per 128 channels), the per-channel bilinear blend is the
// reference's (blend4), the dot product is a shuffle reduction, and the rd^2 x 32 result
// tile leaves through the same shared-memory transpose as lookup_fwd.cu (128-byte rows).
// Index logic: every corner gated separately (quirk Q4), [ix][iy] output order (Q1),
// offset slab b*n in strict_ref mode (Q2), centre offset tap zeroed in place (Q5).
#include <cuda_fp16.h>
#include "common.cuh"

namespace lgu {

// ---------------------------------------------------------------------------------------------
// Tensor-core path behind the SAME operator names (round 2).  The reference samples on the fly because 12-24 GB GPUs
// cannot hold a chunk's all-pairs volumes; per call that costs 1.2 GB/edge of L2 gathers (4 x 128 scalar channel loads
// per tap per pixel).  With a caller-supplied workspace the operator instead
//   1. splits the fp32 channels-last maps into fp16 hi / lo planes (exact for the backend's fp16 frame buffer; for
//      general fp32 maps the 3-MMA split keeps |error| <= ~2^-21 relative per product),
//   2. builds the [B, P, Q] volume of THIS level on tcgen05 (lgu_build_volume, fp32 accumulation in TMEM),
//   3. samples it with the TMA-staged single-level lookup in per-corner-gating mode (quirk Q4, lookup_level_tma.cu;
//      r = 1 for altcorr_forward).
// "Dot-then-interpolate" is algebraically the reference's "interpolate-then-dot" (both gate every corner on its own);
// the results differ by fp32 summation order only (<= 1e-5, tests/test_ref_gpu.py).  Shapes the path does not cover
// (C != 128, N != 1, H1*W1 % 128 != 0, W2 % 4 != 0, other radii) run the SIMT kernel below.
int launch_level_tma_pc(const float* volume, const float* coords, float* offset, float* corr, int B, int H1, int W1, int H2,
                        int W2, long long off_edge_stride, cudaStream_t st);                       // lookup_level_tma.cu
int launch_r1_pc(const float* volume, const float* coords, float* corr, int B, int P, int H2, int W2, cudaStream_t st);

__global__ void __launch_bounds__(256)
split_planes_kernel(const float* __restrict__ x, __half* __restrict__ hi, __half* __restrict__ lo, long long n4,
                    int32_t* __restrict__ idx, int nidx) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nidx) idx[t] = (int32_t)t;
  for (long long q = t; q < n4; q += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + q);
    const __half h0 = __float2half_rn(v.x), h1 = __float2half_rn(v.y), h2 = __float2half_rn(v.z), h3 = __float2half_rn(v.w);
    const __half l0 = __float2half_rn(v.x - __half2float(h0)), l1 = __float2half_rn(v.y - __half2float(h1));
    const __half l2 = __float2half_rn(v.z - __half2float(h2)), l3 = __float2half_rn(v.w - __half2float(h3));
    reinterpret_cast<__half2*>(hi)[2 * q] = __halves2half2(h0, h1);
    reinterpret_cast<__half2*>(hi)[2 * q + 1] = __halves2half2(h2, h3);
    reinterpret_cast<__half2*>(lo)[2 * q] = __halves2half2(l0, l1);
    reinterpret_cast<__half2*>(lo)[2 * q + 1] = __halves2half2(l2, l3);
  }
}

struct LowMemWs {
  size_t f1_hi, f1_lo, f2_hi, f2_lo, idx, volume, total;
};
static inline size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }
// 0 total: the tensor-core path does not apply to this shape
static LowMemWs lowmem_ws_layout(int B, int N, int H1, int W1, int H2, int W2, int C, int radius) {
  LowMemWs w = {0, 0, 0, 0, 0, 0, 0};
  const long long P = (long long)H1 * W1, Q = (long long)H2 * W2;
  if (N != 1 || C != 128 || (radius != 3 && radius != 1) || (P % 128) != 0 || (W2 % 4) != 0 || B <= 0 ||
      (long long)B * P >= 2147483647LL)
    return w;
  size_t at = 0;
  w.f1_hi = at; at = up256(at + (size_t)B * P * C * 2);
  w.f1_lo = at; at = up256(at + (size_t)B * P * C * 2);
  w.f2_hi = at; at = up256(at + (size_t)B * Q * C * 2);
  w.f2_lo = at; at = up256(at + (size_t)B * Q * C * 2);
  w.idx = at;   at = up256(at + (size_t)B * 4);
  w.volume = at; at = up256(at + (size_t)B * P * Q * 4);
  w.total = at;
  return w;
}

static int lowmem_volume(const float* fmap1, const float* fmap2, int B, int P, int Q, int C, const LowMemWs& w, char* ws,
                         cudaStream_t st) {
  auto hi1 = reinterpret_cast<__half*>(ws + w.f1_hi), lo1 = reinterpret_cast<__half*>(ws + w.f1_lo);
  auto hi2 = reinterpret_cast<__half*>(ws + w.f2_hi), lo2 = reinterpret_cast<__half*>(ws + w.f2_lo);
  auto idx = reinterpret_cast<int32_t*>(ws + w.idx);
  const long long n1 = (long long)B * P * C / 4, n2 = (long long)B * Q * C / 4;
  split_planes_kernel<<<(unsigned)((n1 + 255) / 256 < 4096 ? (n1 + 255) / 256 : 4096), 256, 0, st>>>(fmap1, hi1, lo1, n1, idx, B);
  split_planes_kernel<<<(unsigned)((n2 + 255) / 256 < 4096 ? (n2 + 255) / 256 : 4096), 256, 0, st>>>(fmap2, hi2, lo2, n2, idx, 0);
  if (int rc = check_launch("lowMem: operand split")) return rc;
  return lgu_build_volume(hi1, lo1, hi2, lo2, idx, idx, reinterpret_cast<float*>(ws + w.volume), B, B, B, P, Q, C, 2, st);
}

constexpr int kLmWarps = 8;
constexpr int kLmThreads = kLmWarps * 32;
constexpr int kLmTile = 32;
constexpr int kLmPixPerWarp = kLmTile / kLmWarps;
constexpr int kLmMaxC4 = 4;   // float4 groups per lane: C <= 512 on the vector path

struct LowMemArgs {
  const float* f1; const float* f2; const float* coords; float* offset; float* out;
  int B, N, H1, W1, H2, W2, C, r, strict_ref, tiles_per_map;
};

// VEC: C % 128 == 0, lane owns channels {128*m + 4*lane .. +3}.  Otherwise lane owns {32*m + lane}.
template <bool VEC, bool DEFORM>
__global__ void __launch_bounds__(kLmThreads) lowmem_kernel(const LowMemArgs a) {
  extern __shared__ float s_out_raw[];
  float(*s_out)[kLmTile + 1] = reinterpret_cast<float(*)[kLmTile + 1]>(s_out_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int P = a.H1 * a.W1;
  const int rd = 2 * a.r + 1, taps = DEFORM ? rd * rd : (rd + 1) * (rd + 1);
  const int out_taps = rd * rd;
  const int bn = blockIdx.x / a.tiles_per_map;               // b*N + n
  const int b = bn / a.N, n = bn - b * a.N;
  const int p0 = (blockIdx.x - bn * a.tiles_per_map) * kLmTile;
  const int C = a.C;
  const int groups = VEC ? (C >> 7) : (C >> 5);
  const float* F2 = a.f2 + (size_t)b * a.H2 * a.W2 * C;

  if (!DEFORM) {   // altcorr accumulates splats: start from zero
    for (int q = threadIdx.x; q < out_taps * (kLmTile + 1); q += kLmThreads) s_out_raw[q] = 0.0f;
    __syncthreads();
  }

#pragma unroll 1
  for (int k = 0; k < kLmPixPerWarp; ++k) {
    const int pl = warp * kLmPixPerWarp + k;
    const int p = min(p0 + pl, P - 1);
    const float* F1 = a.f1 + ((size_t)b * P + p) * C;
    float4 f1v[kLmMaxC4];
    float f1s[16];
    if (VEC) {
#pragma unroll
      for (int m = 0; m < kLmMaxC4; ++m)
        if (m < groups) f1v[m] = __ldg(reinterpret_cast<const float4*>(F1) + m * 32 + lane);
    } else {
#pragma unroll
      for (int m = 0; m < 16; ++m)
        if (m < groups) f1s[m] = __ldg(F1 + m * 32 + lane);
    }
    const float2 c = __ldg(reinterpret_cast<const float2*>(a.coords) + (size_t)bn * P + p);

    float2 my_off[3];   // lane t holds the offsets of taps t, t+32, t+64 (rd <= 9)
    if (DEFORM) {
      const long long slab = a.strict_ref ? (long long)b * n : (long long)b * a.N + n;   // quirk Q2
      float2* O = reinterpret_cast<float2*>(a.offset) + ((size_t)slab * P + p) * taps;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int t = s * 32 + lane;
        my_off[s] = make_float2(0.0f, 0.0f);
        if (t < taps) {
          if (t == a.r * rd + a.r) O[t] = my_off[s];         // Q5
          else my_off[s] = O[t];
        }
      }
    }

    for (int t = 0; t < taps; ++t) {
      // DEFORM: t = ix*rd + iy (offset memory order == output order).  altcorr: iy outer, ix inner.
      int ix, iy;
      float px = c.x, py = c.y;
      if (DEFORM) {
        ix = t / rd; iy = t - ix * rd;
        const int s = t >> 5;
        const float2 src = s == 0 ? my_off[0] : (s == 1 ? my_off[1] : my_off[2]);
        const float ox = __shfl_sync(0xffffffffu, src.x, t & 31);
        const float oy = __shfl_sync(0xffffffffu, src.y, t & 31);
        px = __fadd_rn(c.x, ox);                             // lowMem_defSample.cu:82-83
        py = __fadd_rn(c.y, oy);
      } else {
        iy = t / (rd + 1); ix = t - iy * (rd + 1);
      }
      const float flx = floorf(px), fly = floorf(py);
      const float dx = __fsub_rn(px, flx), dy = __fsub_rn(py, fly);
      const int h2 = tap_coord(floor_to_int(py), a.r, iy), w2 = tap_coord(floor_to_int(px), a.r, ix);

      float acc = 0.0f;
      if (DEFORM) {
        const int h2h = wrap_inc(h2), w2h = wrap_inc(w2);
        const bool b11 = in_bounds(h2, w2, a.H2, a.W2), b21 = in_bounds(h2, w2h, a.H2, a.W2);
        const bool b12 = in_bounds(h2h, w2, a.H2, a.W2), b22 = in_bounds(h2h, w2h, a.H2, a.W2);
        const float* r11 = F2 + ((size_t)h2 * a.W2 + w2) * C;     // only dereferenced when gated in
        const float* r21 = r11 + C;
        const float* r12 = r11 + (size_t)a.W2 * C;
        const float* r22 = r12 + C;
        if (VEC) {
          const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int m = 0; m < kLmMaxC4; ++m) {
            if (m < groups) {
              const int o = m * 32 + lane;
              const float4 q11 = b11 ? __ldg(reinterpret_cast<const float4*>(r11) + o) : z;
              const float4 q21 = b21 ? __ldg(reinterpret_cast<const float4*>(r21) + o) : z;
              const float4 q12 = b12 ? __ldg(reinterpret_cast<const float4*>(r12) + o) : z;
              const float4 q22 = b22 ? __ldg(reinterpret_cast<const float4*>(r22) + o) : z;
              acc = __fmaf_rn(f1v[m].x, blend4(q11.x, q21.x, q12.x, q22.x, dx, dy), acc);
              acc = __fmaf_rn(f1v[m].y, blend4(q11.y, q21.y, q12.y, q22.y, dx, dy), acc);
              acc = __fmaf_rn(f1v[m].z, blend4(q11.z, q21.z, q12.z, q22.z, dx, dy), acc);
              acc = __fmaf_rn(f1v[m].w, blend4(q11.w, q21.w, q12.w, q22.w, dx, dy), acc);
            }
          }
        } else {
#pragma unroll
          for (int m = 0; m < 16; ++m) {
            if (m < groups) {
              const int o = m * 32 + lane;
              const float q11 = b11 ? __ldg(r11 + o) : 0.f, q21 = b21 ? __ldg(r21 + o) : 0.f;
              const float q12 = b12 ? __ldg(r12 + o) : 0.f, q22 = b22 ? __ldg(r22 + o) : 0.f;
              acc = __fmaf_rn(f1s[m], blend4(q11, q21, q12, q22, dx, dy), acc);
            }
          }
        }
        acc = warp_sum(acc);
        if (lane == 0) s_out[t][pl] = acc;
      } else {
        // altcorr: dot at the integer position, then splat to <= 4 neighbouring taps (altcorr_kernel.cu:102-142)
        if (in_bounds(h2, w2, a.H2, a.W2)) {
          const float* rr = F2 + ((size_t)h2 * a.W2 + w2) * C;
          if (VEC) {
#pragma unroll
            for (int m = 0; m < kLmMaxC4; ++m) {
              if (m < groups) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(rr) + m * 32 + lane);
                acc = __fmaf_rn(f1v[m].x, q.x, acc); acc = __fmaf_rn(f1v[m].y, q.y, acc);
                acc = __fmaf_rn(f1v[m].z, q.z, acc); acc = __fmaf_rn(f1v[m].w, q.w, acc);
              }
            }
          } else {
#pragma unroll
            for (int m = 0; m < 16; ++m)
              if (m < groups) acc = __fmaf_rn(f1s[m], __ldg(rr + m * 32 + lane), acc);
          }
        }
        const float s = warp_sum(acc);
        // lanes 0..3 each own one splat target (distinct taps -> no conflict)
        if (lane < 4) {
          const int oy = iy - 1 + (lane >> 1), ox = ix - 1 + (lane & 1);       // nw, ne, sw, se
          const float wy = (lane >> 1) ? __fsub_rn(1.0f, dy) : dy;
          const float wx = (lane & 1) ? __fsub_rn(1.0f, dx) : dx;
          if (oy >= 0 && oy < rd && ox >= 0 && ox < rd)
            s_out[oy + rd * ox][pl] += __fmul_rn(s, __fmul_rn(wy, wx));
        }
        __syncwarp();
      }
    }
  }
  __syncthreads();
  const bool live = (p0 + lane) < P;
  float* out = a.out + (size_t)bn * out_taps * P + p0 + lane;
  for (int t = warp; t < out_taps; t += kLmWarps)
    if (live) out[(size_t)t * P] = s_out[t][lane];
}

template <bool DEFORM>
static int launch_lowmem(const LowMemArgs& a0, cudaStream_t st, const char* name) {
  LowMemArgs a = a0;
  const int P = a.H1 * a.W1;
  a.tiles_per_map = (P + kLmTile - 1) / kLmTile;
  const long long nblk = (long long)a.B * a.N * a.tiles_per_map;
  LGU_REQUIRE(nblk < 2147483647LL, "%s: grid too large (%lld CTAs)", name, nblk);
  const int rd = 2 * a.r + 1;
  const size_t smem = (size_t)rd * rd * (kLmTile + 1) * sizeof(float);
  LGU_REQUIRE(smem <= 48 * 1024, "%s: radius %d too large", name, a.r);
  if ((a.C & 127) == 0 && a.C <= 128 * kLmMaxC4)
    lowmem_kernel<true, DEFORM><<<(unsigned)nblk, kLmThreads, smem, st>>>(a);
  else
    lowmem_kernel<false, DEFORM><<<(unsigned)nblk, kLmThreads, smem, st>>>(a);
  return check_launch(name);
}

}  // namespace lgu

extern "C" long long lgu_lowmem_workspace_bytes(int B, int N, int H1, int W1, int H2, int W2, int C, int radius) {
  return (long long)lgu::lowmem_ws_layout(B, N, H1, W1, H2, W2, C, radius).total;
}

extern "C" int lgu_lowmem_defsample_forward_ws(const float* fmap1, const float* fmap2, const float* coords, float* offset,
                                               float* corr, int B, int N, int H1, int W1, int H2, int W2, int C, int radius,
                                               int strict_ref, void* workspace, long long workspace_bytes, void* stream) {
  if (B == 0) return LGU_OK;
  const lgu::LowMemWs w = lgu::lowmem_ws_layout(B, N, H1, W1, H2, W2, C, radius);
  if (w.total == 0 || radius != 3 || workspace == nullptr)       // shape outside the tensor-core path: the SIMT kernel
    return lgu_lowmem_defsample_forward(fmap1, fmap2, coords, offset, corr, B, N, H1, W1, H2, W2, C, radius, strict_ref, stream);
  LGU_REQUIRE(fmap1 && fmap2 && coords && offset && corr, "lgu_lowmem_defsample_forward_ws: null pointer");
  LGU_REQUIRE((size_t)workspace_bytes >= w.total && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
              "lgu_lowmem_defsample_forward_ws: workspace of %lld B (256-byte aligned) needed, got %lld", (long long)w.total,
              workspace_bytes);
  LGU_REQUIRE(((reinterpret_cast<uintptr_t>(fmap1) | reinterpret_cast<uintptr_t>(fmap2)) & 15) == 0,
              "lgu_lowmem_defsample_forward_ws: feature maps must be 16-byte aligned");
  const int P = H1 * W1, Q = H2 * W2;
  char* ws = reinterpret_cast<char*>(workspace);
  if (int rc = lgu::lowmem_volume(fmap1, fmap2, B, P, Q, C, w, ws, (cudaStream_t)stream)) return rc;
  // offset slabs: the reference reads offset[b*n] with n == 0 (quirk Q2: every edge uses slab 0); fixed mode: offset[b]
  const long long stride = strict_ref ? 0 : (long long)P * 49;
  return lgu::launch_level_tma_pc(reinterpret_cast<const float*>(ws + w.volume), coords, offset, corr, B, H1, W1, H2, W2,
                                  stride, (cudaStream_t)stream);
}

extern "C" int lgu_altcorr_forward_ws(const float* fmap1, const float* fmap2, const float* coords, float* corr, int B, int N,
                                      int H1, int W1, int H2, int W2, int C, int radius, void* workspace,
                                      long long workspace_bytes, void* stream) {
  if (B == 0) return LGU_OK;
  const lgu::LowMemWs w = lgu::lowmem_ws_layout(B, N, H1, W1, H2, W2, C, radius);
  if (w.total == 0 || radius != 1 || workspace == nullptr)
    return lgu_altcorr_forward(fmap1, fmap2, coords, corr, B, N, H1, W1, H2, W2, C, radius, stream);
  LGU_REQUIRE(fmap1 && fmap2 && coords && corr, "lgu_altcorr_forward_ws: null pointer");
  LGU_REQUIRE((size_t)workspace_bytes >= w.total && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
              "lgu_altcorr_forward_ws: workspace of %lld B (256-byte aligned) needed, got %lld", (long long)w.total,
              workspace_bytes);
  LGU_REQUIRE(((reinterpret_cast<uintptr_t>(fmap1) | reinterpret_cast<uintptr_t>(fmap2)) & 15) == 0,
              "lgu_altcorr_forward_ws: feature maps must be 16-byte aligned");
  const int P = H1 * W1, Q = H2 * W2;
  char* ws = reinterpret_cast<char*>(workspace);
  if (int rc = lgu::lowmem_volume(fmap1, fmap2, B, P, Q, C, w, ws, (cudaStream_t)stream)) return rc;
  return lgu::launch_r1_pc(reinterpret_cast<const float*>(ws + w.volume), coords, corr, B, P, H2, W2, (cudaStream_t)stream);
}

extern "C" int lgu_lowmem_defsample_forward(const float* fmap1, const float* fmap2, const float* coords,
                                            float* offset, float* corr, int B, int N, int H1, int W1, int H2, int W2,
                                            int C, int radius, int strict_ref, void* stream) {
  if (B == 0) return LGU_OK;   // empty edge set: nothing to do (pointers may be null)
  LGU_REQUIRE(fmap1 && fmap2 && coords && offset && corr, "lgu_lowmem_defsample_forward: null pointer");
  LGU_REQUIRE(B >= 0 && N > 0 && H1 > 0 && W1 > 0 && H2 > 0 && W2 > 0 && C > 0 && radius >= 0,
              "lgu_lowmem_defsample_forward: bad sizes");
  LGU_REQUIRE((C & 31) == 0 && C <= 512, "lgu_lowmem_defsample_forward: C=%d must be a multiple of 32, <= 512", C);
  LGU_REQUIRE(radius <= 4, "lgu_lowmem_defsample_forward: radius %d > 4 unsupported", radius);
  lgu::LowMemArgs a{fmap1, fmap2, coords, offset, corr, B, N, H1, W1, H2, W2, C, radius, strict_ref, 0};
  return lgu::launch_lowmem<true>(a, (cudaStream_t)stream, "lgu_lowmem_defsample_forward");
}

extern "C" int lgu_altcorr_forward(const float* fmap1, const float* fmap2, const float* coords, float* corr, int B,
                                   int N, int H1, int W1, int H2, int W2, int C, int radius, void* stream) {
  if (B == 0) return LGU_OK;   // empty edge set: nothing to do (pointers may be null)
  LGU_REQUIRE(fmap1 && fmap2 && coords && corr, "lgu_altcorr_forward: null pointer");
  LGU_REQUIRE(B >= 0 && N > 0 && H1 > 0 && W1 > 0 && H2 > 0 && W2 > 0 && C > 0 && radius >= 0,
              "lgu_altcorr_forward: bad sizes");
  LGU_REQUIRE((C & 31) == 0 && C <= 512, "lgu_altcorr_forward: C=%d must be a multiple of 32, <= 512", C);
  LGU_REQUIRE(radius <= 4, "lgu_altcorr_forward: radius %d > 4 unsupported", radius);
  lgu::LowMemArgs a{fmap1, fmap2, coords, nullptr, corr, B, N, H1, W1, H2, W2, C, radius, 0, 0};
  return lgu::launch_lowmem<false>(a, (cudaStream_t)stream, "lgu_altcorr_forward");
}
