// Microbenchmark: small-box TMA gathers vs warp-cooperative LDG.128 row loads, for the per-pixel
// 16x16 patch fetch of the deformable lookup.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tW_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

__device__ __forceinline__ uint32_t hash(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

template <int BW, int BH, int DEPTH, int WARPS, int NBOX>
__global__ void __launch_bounds__(WARPS * 32) tma_kernel(const __grid_constant__ CUtensorMap map, float* out, int npix, int H2, int W2, int mode) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int BOXB = BW * BH * 4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* buf = reinterpret_cast<float*>(smem) + warp * DEPTH * NBOX * BW * BH;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WARPS * DEPTH * NBOX * BOXB) + warp * DEPTH;
  if (lane == 0) for (int d = 0; d < DEPTH; ++d) mbar_init(bars + d, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const int wg = blockIdx.x * WARPS + warp, nw = gridDim.x * WARPS;
  float acc = 0.f;
  int issued = 0, done = 0;
  const int mine = (npix - wg + nw - 1) / nw;
  auto issue = [&](int k) {
    const int pix = wg + k * nw;
    const int slot = k % DEPTH;
    if (lane == 0) {
      mbar_expect_tx(bars + slot, NBOX * BOXB);
      for (int b = 0; b < NBOX; ++b) {
        const uint32_t h = hash(pix * 4 + b);
        int x = (int)(h % (W2 + 8)) - 12, y = (int)((h >> 8) % (H2 + 8)) - 12;
        if (mode == 0) { x = 0; y = 0; } else if (mode == 1) { x &= ~3; } else if (mode == 3) { x = (h % 40) + 1; y = (h >> 8) % 30; }
        tma_load_3d(buf + (slot * NBOX + b) * BW * BH, &map, bars + slot, x, y, pix);
      }
    }
  };
  for (; issued < DEPTH && issued < mine; ++issued) issue(issued);
  for (; done < mine; ++done) {
    const int slot = done % DEPTH;
    mbar_wait(bars + slot, (done / DEPTH) & 1);
    const float* b = buf + slot * NBOX * BW * BH;
#pragma unroll
    for (int k = 0; k < NBOX * BW * BH / 32; ++k) acc += b[k * 32 + lane];
    __syncwarp();
    if (issued < mine) { issue(issued); ++issued; }
  }
  if (acc == 12345.678f) out[wg] = acc;
}

// warp-cooperative: 16 rows x 128 B (the aligned 32-float window containing the box), LDG.128 -> STS.128 -> read
template <int WARPS, int PIXB>
__global__ void __launch_bounds__(WARPS * 32) ldg_kernel(const float* __restrict__ vol, float* out, int npix, int H2, int W2) {
  __shared__ __align__(16) float sbuf[WARPS][PIXB][16 * 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wg = blockIdx.x * WARPS + warp, nw = gridDim.x * WARPS;
  float acc = 0.f;
  for (int pix0 = wg * PIXB; pix0 < npix; pix0 += nw * PIXB) {
    float4 v[PIXB][4];
#pragma unroll
    for (int p = 0; p < PIXB; ++p) {
      const int pix = pix0 + p;
      const uint32_t h = hash(pix * 4);
      int x = (int)(h % (W2 - 16)), y = (int)((h >> 8) % (H2 - 16));
      x &= ~31; if (x + 32 > W2) x = W2 - 32;
      const float* base = vol + (size_t)pix * H2 * W2 + (size_t)y * W2 + x;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int row = k * 4 + (lane >> 3), c4 = lane & 7;
        v[p][k] = __ldg(reinterpret_cast<const float4*>(base + (size_t)row * W2) + c4);
      }
    }
#pragma unroll
    for (int p = 0; p < PIXB; ++p)
#pragma unroll
      for (int k = 0; k < 4; ++k) reinterpret_cast<float4*>(sbuf[warp][p])[k * 32 + lane] = v[p][k];
    __syncwarp();
#pragma unroll
    for (int p = 0; p < PIXB; ++p)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc += sbuf[warp][p][(k * 32 + lane * 7) & 511];
    __syncwarp();
  }
  if (acc == 12345.678f) out[wg] = acc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(void* base, int npix, int H2, int W2, int bw, int bh) {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  EncodeTiledFn fn = (EncodeTiledFn)p;
  CUtensorMap m;
  cuuint64_t dims[3] = {(cuuint64_t)W2, (cuuint64_t)H2, (cuuint64_t)npix};
  cuuint64_t strides[2] = {(cuuint64_t)W2 * 4, (cuuint64_t)W2 * H2 * 4};
  cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  return m;
}

static int g_mode = 2;
template <typename F> static float time_it(F f, int iters = 5) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); CK(cudaDeviceSynchronize());
  float best = 1e9;
  for (int i = 0; i < iters; ++i) { cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
  return best;
}

template <int BW, int BH, int DEPTH, int WARPS, int NBOX>
static void run_tma(float* vol, float* out, int npix, int H2, int W2, int ctas_per_sm, const char* name) {
  CUtensorMap m = make_map(vol, npix, H2, W2, BW, BH);
  const int smem = WARPS * DEPTH * NBOX * BW * BH * 4 + WARPS * DEPTH * 8;
  auto k = tma_kernel<BW, BH, DEPTH, WARPS, NBOX>;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  float ms = time_it([&] { k<<<148 * ctas_per_sm, WARPS * 32, smem>>>(m, out, npix, H2, W2, g_mode); });
  CK(cudaGetLastError());
  printf("%-44s %8.1f us  %7.2f Mbox/s*1e-0  %7.1f GB/s(payload)  smem/CTA %d\n", name, ms * 1e3,
         (double)npix * NBOX / (ms * 1e3), (double)npix * NBOX * BW * BH * 4 / (ms * 1e6), smem);
}

int main(int argc, char** argv) {
  if (argc > 1) g_mode = atoi(argv[1]);
  const int E = 48, P = 3072, npix = E * P;
  float *vol, *out;
  CK(cudaMalloc(&vol, (size_t)npix * 48 * 64 * 4)); CK(cudaMemset(vol, 0, (size_t)npix * 48 * 64 * 4));
  CK(cudaMalloc(&out, 1 << 20));
  printf("npix %d\n", npix);
  // level-0 slices 48x64
  run_tma<16, 16, 2, 8, 1>(vol, out, npix, 48, 64, 2, "tma 16x16 depth2 8w x2cta");
  run_tma<16, 16, 4, 8, 1>(vol, out, npix, 48, 64, 2, "tma 16x16 depth4 8w x2cta");
  run_tma<16, 16, 4, 8, 1>(vol, out, npix, 48, 64, 4, "tma 16x16 depth4 8w x4cta");
  run_tma<16, 16, 8, 8, 1>(vol, out, npix, 48, 64, 2, "tma 16x16 depth8 8w x2cta");
  run_tma<32, 16, 4, 8, 1>(vol, out, npix, 48, 64, 2, "tma 32x16 depth4 8w x2cta");
  run_tma<16, 16, 4, 16, 1>(vol, out, npix, 48, 64, 2, "tma 16x16 depth4 16w x2cta");
  run_tma<16, 16, 2, 8, 4>(vol, out, npix, 48, 64, 2, "tma 4 boxes/pixel 16x16 depth2 8w x2cta");
  run_tma<8, 8, 4, 8, 1>(vol, out, npix, 48, 64, 2, "tma 8x8 depth4 8w x2cta");
  run_tma<8, 8, 8, 8, 4>(vol, out, npix, 48, 64, 2, "tma 4 boxes/pixel 8x8 depth8 8w x2cta");
  {
    float ms = time_it([&] { ldg_kernel<8, 2><<<148 * 4, 256>>>(vol, out, npix, 48, 64); });
    printf("%-44s %8.1f us  %7.2f Mbox/s  %7.1f GB/s(payload 2KB)\n", "ldg128 16rows x128B, 2 px/warp, 8w x4cta", ms * 1e3, npix / (ms * 1e3), (double)npix * 2048 / (ms * 1e6));
    ms = time_it([&] { ldg_kernel<4, 4><<<148 * 6, 128>>>(vol, out, npix, 48, 64); });
    printf("%-44s %8.1f us  %7.2f Mbox/s  %7.1f GB/s(payload 2KB)\n", "ldg128 16rows x128B, 4 px/warp, 4w x6cta", ms * 1e3, npix / (ms * 1e3), (double)npix * 2048 / (ms * 1e6));
  }
  return 0;
}
