// Micro-experiment: dependent-load latency under the lookup kernel's access pattern.
// Each warp walks `chain` records of 392 B (coalesced float2 read by 32 lanes), the next record index
// depends on the loaded data.  Variants: mode 0 plain loads; mode 1 one lane first stores 8 B into the
// record (like the centre-tap zeroing); mode 2 store to a DIFFERENT array (no overlap).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) walk(float2* rec, float2* other, int nrec, int chain, int mode, float* out) {
  const int lane = threadIdx.x & 31;
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  unsigned idx = (unsigned)((w * 2654435761u) % (unsigned)nrec);
  float acc = 0.f;
  for (int c = 0; c < chain; ++c) {
    float2* R = rec + (size_t)idx * 49;
    if (mode == 1 && lane == 24) R[24] = make_float2(0.f, 0.f);
    if (mode == 2 && lane == 24) other[(size_t)idx * 49 + 24] = make_float2(0.f, 0.f);
    float2 v = (mode == 1 && lane == 24) ? make_float2(0.f, 0.f) : R[lane];
    acc += v.x + v.y;
    const float s = __shfl_sync(0xffffffffu, v.x, 3);
    idx = (unsigned)((idx * 1664525u + 1013904223u + (unsigned)(s != 12345.f)) % (unsigned)nrec);
  }
  if (acc == 1234567.f) out[w] = acc;
}

int main(int argc, char** argv) {
  const int nrec = argc > 1 ? atoi(argv[1]) : 147456;   // E=48 -> 57.8 MB
  const int chain = 64;
  float2 *rec, *other; float* out;
  cudaMalloc(&rec, (size_t)nrec * 392); cudaMalloc(&other, (size_t)nrec * 392); cudaMalloc(&out, 1 << 24);
  cudaMemset(rec, 0, (size_t)nrec * 392);
  for (int mode = 0; mode < 3; ++mode)
    for (int blocks_per_sm : {1, 2, 4, 8}) {
      const int grid = 148 * blocks_per_sm;
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
      walk<<<grid, 256>>>(rec, other, nrec, chain, mode, out);
      cudaEventRecord(a);
      walk<<<grid, 256>>>(rec, other, nrec, chain, mode, out);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      printf("mode %d warps/SM %2d: %.1f us total, %.0f ns per dependent load (%s)\n", mode, blocks_per_sm * 8,
             ms * 1e3, ms * 1e6 / chain, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
