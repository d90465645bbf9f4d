// offset_heads.cu -- the offset heads' glue of CorrBlock.fpn_offset_generate / AltCorrBlock.offset_generate
// (/root/reference/droid_slam/modules/corr.py:117-135, 217-235; per_Corr_Normalization :44-51) for sm_100a.
//
// The reference runs, on the two conv outputs, mean -> 3 x unsqueeze -> var -> sqrt -> sub -> div -> tanh -> mul (twice),
// F.interpolate, add, div, 2 x permute and, at the first lookup, 2 x .contiguous(): ~14 elementwise / reduction kernels that
// each stream the [E,98,H,W] tensors through HBM (70 % of an uncached backend step was this glue).  Here:
//   offset_stats_kernel : one CTA per (edge, head): mean and biased variance over (CH,H,W), fp32 loads, fp64 accumulation
//   offset_apply_kernel : one CTA per (edge, 32-pixel tile): normalise, 4 tanh, the level-1 average with level 0, nearest
//                         upsampling of the residual head by indexing, and the NCHW -> NHWC transpose through shared memory,
//                         so the [E,H,W,98] records the lookups read are written once, fully coalesced.
// tanhf / IEEE division as torch's fp32 CUDA kernels use them; statistics differ from torch's Welford reduction in the last
// bits only (values agree to a few 1e-7, tests/test_dropin_gpu.py compares against the reference's own Python).
#include "common.cuh"

namespace lgu {

constexpr int kOhThreads = 256;

__global__ void __launch_bounds__(1024)
offset_stats_kernel(const float* __restrict__ c0, const float* __restrict__ c1, float* __restrict__ stats, long long n0,
                    long long n1, float eps) {
  const int e = blockIdx.x >> 1, head = blockIdx.x & 1;
  const long long n = head ? n1 : n0;
  const float* x = (head ? c1 : c0) + (size_t)e * n;
  double s = 0.0, ss = 0.0;
  if ((n & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    for (long long q = threadIdx.x; q < (n >> 2); q += blockDim.x) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x) + q);
      s += (double)v.x + (double)v.y + (double)v.z + (double)v.w;
      ss += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    }
  } else {
    for (long long q = threadIdx.x; q < n; q += blockDim.x) {
      const float v = __ldg(x + q);
      s += v;
      ss += (double)v * v;
    }
  }
  __shared__ double sh[2][32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s; sh[1][threadIdx.x >> 5] = ss; }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int nw = blockDim.x >> 5;
    s = threadIdx.x < nw ? sh[0][threadIdx.x] : 0.0;
    ss = threadIdx.x < nw ? sh[1][threadIdx.x] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    if (threadIdx.x == 0) {
      const double mean = s / (double)n;
      double var = ss / (double)n - mean * mean;                // biased (unbiased=False, corr.py:47)
      if (var < 0.0) var = 0.0;
      stats[(size_t)e * 4 + head * 2 + 0] = (float)mean;
      stats[(size_t)e * 4 + head * 2 + 1] = sqrtf((float)var + eps);       // the divisor: sqrt(var + eps), fp32 like torch
    }
  }
}

// 32 consecutive pixels of one edge x all CH channels.  Thread t: pixel t & 31, channels (t >> 5) + 8 m.
__global__ void __launch_bounds__(kOhThreads)
offset_apply_kernel(const float* __restrict__ c0, const float* __restrict__ c1, const float* __restrict__ stats,
                    float* __restrict__ off0, float* __restrict__ off1, int CH, int H, int W, int tiles_per_edge) {
  extern __shared__ float sm[];                                  // [2][32][CH + 1]
  const int e = blockIdx.x / tiles_per_edge;
  const int p0 = (blockIdx.x - e * tiles_per_edge) * 32;
  const int P = H * W, Wh = W >> 1, Ph = (H >> 1) * Wh;
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const float m0 = __ldg(stats + (size_t)e * 4 + 0), d0 = __ldg(stats + (size_t)e * 4 + 1);
  const float m1 = __ldg(stats + (size_t)e * 4 + 2), d1 = __ldg(stats + (size_t)e * 4 + 3);
  const int p = p0 + lane, y = p / W, x = p - y * W;
  const int ph = (y >> 1) * Wh + (x >> 1);                       // F.interpolate(..., (h, w)) nearest: source = floor(dst / 2)
  const int pitch = CH + 1;
  float* s0 = sm;
  float* s1 = sm + 32 * pitch;
  for (int c = grp; c < CH; c += kOhThreads / 32) {
    const float a = __ldg(c0 + ((size_t)e * CH + c) * P + p);
    const float b = __ldg(c1 + ((size_t)e * CH + c) * Ph + ph);
    const float o0 = __fmul_rn(tanhf(__fdiv_rn(__fsub_rn(a, m0), d0)), 4.0f);
    const float o1 = __fdiv_rn(__fadd_rn(__fmul_rn(tanhf(__fdiv_rn(__fsub_rn(b, m1), d1)), 4.0f), o0), 2.0f);
    s0[lane * pitch + c] = o0;
    s1[lane * pitch + c] = o1;
  }
  __syncthreads();
  // the 32 pixels' records are contiguous in the NHWC output: 32 * CH floats
  float* g0 = off0 + ((size_t)e * P + p0) * CH;
  float* g1 = off1 + ((size_t)e * P + p0) * CH;
  for (int q = threadIdx.x; q < 32 * CH; q += kOhThreads) {
    const int pl = q / CH, c = q - pl * CH;
    g0[q] = s0[pl * pitch + c];
    g1[q] = s1[pl * pitch + c];
  }
}

}  // namespace lgu

extern "C" int lgu_offset_heads(const float* c0, const float* c1, float* off0, float* off1, float* stats, int E, int CH, int H,
                                int W, float eps, void* stream) {
  using namespace lgu;
  if (E == 0) return LGU_OK;
  LGU_REQUIRE(c0 && c1 && off0 && off1 && stats, "lgu_offset_heads: null pointer");
  LGU_REQUIRE(E > 0 && CH > 0 && CH <= 128 && H > 0 && W > 0 && (W % 32) == 0 && (H % 2) == 0,
              "lgu_offset_heads: bad sizes E=%d CH=%d H=%d W=%d (needs CH <= 128, W %% 32 == 0, H even)", E, CH, H, W);
  const long long n0 = (long long)CH * H * W, n1 = (long long)CH * (H / 2) * (W / 2);
  offset_stats_kernel<<<2 * E, 1024, 0, (cudaStream_t)stream>>>(c0, c1, stats, n0, n1, eps);
  if (int rc = check_launch("lgu_offset_heads (stats)")) return rc;
  const int tiles = H * W / 32;
  const size_t smem = (size_t)2 * 32 * (CH + 1) * sizeof(float);
  offset_apply_kernel<<<(unsigned)((long long)E * tiles), kOhThreads, smem, (cudaStream_t)stream>>>(c0, c1, stats, off0, off1,
                                                                                                  CH, H, W, tiles);
  return check_launch("lgu_offset_heads");
}
