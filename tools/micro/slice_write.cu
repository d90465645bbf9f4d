// Microbenchmark: ceiling of the fused backward's write pattern -- every warp streams contiguous slices
// (12 KB + 3 KB + 768 B + 192 B per source pixel) with 16-byte st.cs, 16 warps per SM, nothing else.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
template <int MODE>
__global__ void __launch_bounds__(256, 2) k(float* g0, float* g1, float* g2, float* g3, int npix) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  // CTA = 32 consecutive pixels, warp = 4 of them (as in the backward)
  for (int k4 = 0; k4 < 4; ++k4) {
    const size_t pix = (size_t)blockIdx.x * 32 + warp * 4 + k4;
    if (pix >= (size_t)npix) return;
    float4* a = reinterpret_cast<float4*>(g0 + pix * 3072);
    for (int i = lane; i < 768; i += 32) { if (MODE == 0) __stcs(a + i, z); else a[i] = z; }
    float4* b = reinterpret_cast<float4*>(g1 + pix * 768);
    for (int i = lane; i < 192; i += 32) { if (MODE == 0) __stcs(b + i, z); else b[i] = z; }
    float4* c = reinterpret_cast<float4*>(g2 + pix * 192);
    for (int i = lane; i < 48; i += 32) { if (MODE == 0) __stcs(c + i, z); else c[i] = z; }
    float4* d = reinterpret_cast<float4*>(g3 + pix * 48);
    if (lane < 12) { if (MODE == 0) __stcs(d + lane, z); else d[lane] = z; }
  }
}
template <typename F> static float time_it(F f, int iters = 5) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); CK(cudaDeviceSynchronize());
  float best = 1e9;
  for (int i = 0; i < iters; ++i) { cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
  return best;
}
int main() {
  const int npix = 48 * 3072;
  float *g0, *g1, *g2, *g3;
  CK(cudaMalloc(&g0, (size_t)npix * 3072 * 4)); CK(cudaMalloc(&g1, (size_t)npix * 768 * 4));
  CK(cudaMalloc(&g2, (size_t)npix * 192 * 4)); CK(cudaMalloc(&g3, (size_t)npix * 48 * 4));
  const double GB = (double)npix * 4080 * 4 / 1e9;
  float ms = time_it([&] { k<0><<<npix / 32, 256>>>(g0, g1, g2, g3, npix); });
  printf("st.cs   slices, CTA=32 px, 8 warps x 4 px: %7.1f us  %6.0f GB/s\n", ms * 1e3, GB / ms * 1e3);
  ms = time_it([&] { k<1><<<npix / 32, 256>>>(g0, g1, g2, g3, npix); });
  printf("st.wb   slices, same                      : %7.1f us  %6.0f GB/s\n", ms * 1e3, GB / ms * 1e3);
  ms = time_it([&] { cudaMemsetAsync(g0, 0, (size_t)npix * 3072 * 4); });
  printf("cudaMemset level 0 only (1.81 GB)          : %7.1f us  %6.0f GB/s\n", ms * 1e3, (double)npix * 3072 * 4 / 1e9 / ms * 1e3);
  return 0;
}
