// Microbenchmark: TMA bulk-tensor STORE throughput for the build epilogue's tile shapes.
// Output matrix [R rows][3072 floats] (12 KB rows), each warp owns 32 consecutive rows and sweeps the columns.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

// MODE 0: 2D box 32 cols x 32 rows (4 KB, 128B rows), one store per 32 columns
// MODE 1: 3D box {32 floats, 32 rows, CH chunks} in ONE instruction (CH*4 KB)
// MODE 2: 2D box COLS cols x 32 rows, no swizzle (COLS*4-byte rows)
template <int MODE, int CH, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) store_kernel(const __grid_constant__ CUtensorMap map, int nrowblocks, int ncols) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int BUF = CH * 4096;
  uint8_t* mine = smem + warp * 2 * BUF;
  for (int i = lane; i < 2 * BUF / 4; i += 32) reinterpret_cast<float*>(mine)[i] = 1.0f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  int sb = 0;
  for (int rb = blockIdx.x * WARPS + warp; rb < nrowblocks; rb += gridDim.x * WARPS) {
    const int row0 = rb * 32;
    for (int c = 0; c < ncols; c += 32 * CH) {
      if (lane == 0) {
        tma_wait_read<1>();
        if (MODE == 0) {
          for (int k = 0; k < CH; ++k) tma_store_2d(&map, mine + sb * BUF + k * 4096, c + 32 * k, row0);
        } else if (MODE == 1) {
          tma_store_3d(&map, mine + sb * BUF, 0, row0, c / 32);
        } else {
          tma_store_2d(&map, mine + sb * BUF, c, row0);
        }
        tma_commit();
      }
      sb ^= 1;
      __syncwarp();
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// single staging buffer per warp: the next store waits until the engine has READ the previous tile (wait_group.read 0)
template <int CH, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) store_kernel_1buf(const __grid_constant__ CUtensorMap map, int nrowblocks, int ncols) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int BUF = CH * 4096;
  uint8_t* mine = smem + warp * BUF;
  for (int i = lane; i < BUF / 4; i += 32) reinterpret_cast<float*>(mine)[i] = 1.0f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  for (int rb = blockIdx.x * WARPS + warp; rb < nrowblocks; rb += gridDim.x * WARPS) {
    const int row0 = rb * 32;
    for (int c = 0; c < ncols; c += 32 * CH) {
      if (lane == 0) {
        tma_wait_read<0>();
        tma_store_3d(&map, mine, 0, row0, c / 32);
        tma_commit();
      }
      __syncwarp();
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn enc() { void* p = nullptr; cudaDriverEntryPointQueryResult q; CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q)); return (EncodeTiledFn)p; }

template <typename F> static float time_it(F f, int iters = 5) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); CK(cudaDeviceSynchronize());
  float best = 1e9;
  for (int i = 0; i < iters; ++i) { cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
  return best;
}

int main() {
  const long long R = 48LL * 3072, C = 3072;
  float* out; CK(cudaMalloc(&out, R * C * 4));
  const double GB = (double)R * C * 4 / 1e9;
  {  // MODE 0
    CUtensorMap m; cuuint64_t d[2] = {(cuuint64_t)C, (cuuint64_t)R}; cuuint64_t s[1] = {(cuuint64_t)C * 4}; cuuint32_t b[2] = {32, 32}; cuuint32_t e[2] = {1, 1};
    if (enc()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("enc0 failed\n"); return 1; }
    auto k = store_kernel<0, 1, 8>; const int smem = 8 * 2 * 4096;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    float ms = time_it([&] { k<<<148, 256, smem>>>(m, (int)(R / 32), (int)C); }); CK(cudaGetLastError());
    printf("2D 32x32 swz128 (4 KB/store), 8 warps x 2 bufs : %7.1f us  %6.0f GB/s\n", ms * 1e3, GB / ms * 1e3);
    auto k4 = store_kernel<0, 4, 4>; const int smem4 = 4 * 2 * 4 * 4096;
    CK(cudaFuncSetAttribute(k4, cudaFuncAttributeMaxDynamicSharedMemorySize, smem4));
    ms = time_it([&] { k4<<<148, 128, smem4>>>(m, (int)(R / 32), (int)C); }); CK(cudaGetLastError());
    printf("2D 32x32 swz128, 4 stores per group (16 KB)      : %7.1f us  %6.0f GB/s\n", ms * 1e3, GB / ms * 1e3);
    auto k16 = store_kernel<0, 1, 16>; const int smem16 = 16 * 2 * 4096;
    CK(cudaFuncSetAttribute(k16, cudaFuncAttributeMaxDynamicSharedMemorySize, smem16));
    ms = time_it([&] { k16<<<148, 512, smem16>>>(m, (int)(R / 32), (int)C); }); CK(cudaGetLastError());
    printf("2D 32x32 swz128, 16 warps                        : %7.1f us  %6.0f GB/s\n", ms * 1e3, GB / ms * 1e3);
  }
  {  // MODE 1: 3D view {32 floats, rows, chunks}
    CUtensorMap m; cuuint64_t d[3] = {32, (cuuint64_t)R, (cuuint64_t)C / 32}; cuuint64_t s[2] = {(cuuint64_t)C * 4, 128}; cuuint32_t e[3] = {1, 1, 1};
    cuuint32_t b4[3] = {32, 32, 4};
    CUresult r = enc()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, out, d, s, b4, e, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) printf("enc 3D failed %d\n", (int)r);
    else {
      auto k = store_kernel<1, 4, 4>; const int smem = 4 * 2 * 4 * 4096;
      CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      float ms = time_it([&] { k<<<148, 128, smem>>>(m, (int)(R / 32), (int)C); }); CK(cudaGetLastError());
      printf("3D {32,32 rows,4 chunks} one store (16 KB)       : %7.1f us  %6.0f GB/s\n", ms * 1e3, GB / ms * 1e3);
    }
    if (!r) {
      auto k1 = store_kernel_1buf<4, 4>; const int smem1 = 4 * 4 * 4096;
      CK(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1));
      float ms = time_it([&] { k1<<<148, 128, smem1>>>(m, (int)(R / 32), (int)C); }); CK(cudaGetLastError());
      printf("3D {32,32 rows,4 chunks} one store (16 KB), 4 warps x 1 buf (64 KB smem) : %7.1f us  %6.0f GB/s\n", ms * 1e3, GB / ms * 1e3);
    }
    cuuint32_t b2c[3] = {32, 32, 2};
    r = enc()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, out, d, s, b2c, e, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) printf("enc 3D/2 failed %d\n", (int)r);
    else {
      auto k = store_kernel<1, 2, 8>; const int smem = 8 * 2 * 2 * 4096;
      CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      float ms = time_it([&] { k<<<148, 256, smem>>>(m, (int)(R / 32), (int)C); }); CK(cudaGetLastError());
      printf("3D {32,32 rows,2 chunks} one store (8 KB), 8 warps x 2 bufs (128 KB smem): %7.1f us  %6.0f GB/s\n", ms * 1e3, GB / ms * 1e3);
      auto k4 = store_kernel<1, 2, 4>; const int smem4 = 4 * 2 * 2 * 4096;
      CK(cudaFuncSetAttribute(k4, cudaFuncAttributeMaxDynamicSharedMemorySize, smem4));
      ms = time_it([&] { k4<<<148, 128, smem4>>>(m, (int)(R / 32), (int)C); }); CK(cudaGetLastError());
      printf("3D {32,32 rows,2 chunks} one store (8 KB), 4 warps x 2 bufs: %7.1f us  %6.0f GB/s\n", ms * 1e3, GB / ms * 1e3);
    }
  }
  {  // MODE 2: no swizzle, wide rows
    CUtensorMap m; cuuint64_t d[2] = {(cuuint64_t)C, (cuuint64_t)R}; cuuint64_t s[1] = {(cuuint64_t)C * 4}; cuuint32_t e[2] = {1, 1};
    cuuint32_t b[2] = {128, 32};
    CUresult r = enc()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) printf("enc wide failed %d\n", (int)r);
    else {
      auto k = store_kernel<2, 4, 4>; const int smem = 4 * 2 * 4 * 4096;
      CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      float ms = time_it([&] { k<<<148, 128, smem>>>(m, (int)(R / 32), (int)C); }); CK(cudaGetLastError());
      printf("2D 128 cols x 32 rows no swizzle (512 B rows)    : %7.1f us  %6.0f GB/s\n", ms * 1e3, GB / ms * 1e3);
    }
    cuuint32_t b2[2] = {256, 16};
    r = enc()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, d, s, b2, e, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) printf("enc wide2 failed %d\n", (int)r);
  }
  return 0;
}
