// Micro-experiment: cost of the intra-warp conflict-resolution primitives used by the backward scatter.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) k_match(int* out, int iters, int distinct) {
  const int lane = threadIdx.x & 31;
  unsigned acc = 0;
  unsigned key = distinct ? (unsigned)(lane * 7 + 3) : 5u;
  for (int i = 0; i < iters; ++i) {
    const unsigned peers = __match_any_sync(0xffffffffu, key + i);
    acc += peers;
    key += (peers >> 31);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void __launch_bounds__(256) k_redux(int* out, int iters) {
  unsigned acc = threadIdx.x;
  for (int i = 0; i < iters; ++i) acc += __reduce_max_sync(0xffffffffu, acc ^ i);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void __launch_bounds__(256) k_atoms(float* out, int iters) {
  __shared__ float s[8][1024];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int i = lane; i < 1024; i += 32) s[w][i] = 0.f;
  __syncwarp();
  int idx = lane * 17;
  for (int i = 0; i < iters; ++i) { atomicAdd(&s[w][idx & 1023], 1.0f); idx = idx * 5 + 1; }
  __syncwarp();
  out[blockIdx.x * blockDim.x + threadIdx.x] = s[w][lane];
}
__global__ void __launch_bounds__(256) k_tag(float* out, int iters) {
  __shared__ float s[8][1024];
  __shared__ int tag[8][256];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int i = lane; i < 1024; i += 32) s[w][i] = 0.f;
  __syncwarp();
  int idx = lane * 17;
  for (int i = 0; i < iters; ++i) {
    const int a = idx & 1023, h = a & 255;
    tag[w][h] = lane; __syncwarp();
    const bool lose = tag[w][h] != lane;
    if (!__any_sync(0xffffffffu, lose)) { s[w][a] += 1.0f; }
    else { for (int r = 0; r < 32; ++r) { if (lane == r) s[w][a] += 1.0f; __syncwarp(); } }
    __syncwarp();
    idx = idx * 5 + 1;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s[w][lane];
}
template <class F> void timeit(const char* name, F f, int iters) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  // 148*2 CTAs of 8 warps -> 16 warps/SM -> 4 warps per SMSP
  printf("%-28s %8.1f us  -> %.1f SM-cycles per op per warp-slot (at 1.9 GHz, 4 warps/SMSP): %s\n", name, ms * 1e3,
         ms * 1e-3 * 1.9e9 / iters, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  int* o; cudaMalloc(&o, 1 << 24);
  const int iters = 2000, grid = 148 * 2;
  timeit("match_any distinct keys", [&] { k_match<<<grid, 256>>>(o, iters, 1); }, iters);
  timeit("match_any same key", [&] { k_match<<<grid, 256>>>(o, iters, 0); }, iters);
  timeit("redux max", [&] { k_redux<<<grid, 256>>>(o, iters); }, iters);
  timeit("atoms f32 CAS spread", [&] { k_atoms<<<grid, 256>>>((float*)o, iters); }, iters);
  timeit("tag-check + RMW", [&] { k_tag<<<grid, 256>>>((float*)o, iters); }, iters);
  return 0;
}
