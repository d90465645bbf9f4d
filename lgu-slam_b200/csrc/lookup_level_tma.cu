// lookup_level_tma.cu -- defCorr_index_forward / defCorr_index_backward for radius 3 with TMA-staged footprints.
//
// Same operators as lookup_fwd.cu / lookup_bwd.cu (reference: /root/reference/offersample_LGS/
// defCorrSample_kernel.cu:25-91 forward, :93-162 backward), same bit-exact index logic (quirks Q1, Q3, Q5), but
// the per-pixel footprint of the pyramid slice is staged in shared memory by ONE cp.async.bulk.tensor box load
// (20 x 16 floats, zero-filled out of bounds) instead of 4 scalar gathers per tap, and the backward accumulates into
// a box-shaped shared accumulator and streams the dense slice once -- the single-level versions of
// lookup_fused.cu / lookup_fused_bwd.cu.  The first-generation kernels were bound by L1 wavefronts (117 sectors
// requested per pixel-level) and run at 15-21 % of the HBM roofline; they remain the path for other radii, for
// W2 % 4 != 0 and for unaligned volumes.
#include "fused_common.cuh"

namespace lgu {

namespace lv {
using namespace fl;
constexpr int kBoxFloats = kBW01 * kBH01;                        // 320 floats = 1280 B
constexpr int kOutPitchL = kTile + 1;
constexpr int kLvBoxes = kWarps * kSlots * kBoxFloats * 4;     // 20,480 B
constexpr int kSmemTile = TAPS * kOutPitchL * 4;                 // 6,468 B (+ pad)
constexpr int kSmemTilePad = (kSmemTile + 15) & ~15;
constexpr int kFwdSmem = kLvBoxes + kSmemTilePad + kWarps * kSlots * 8;
constexpr int kBwdSmem = kLvBoxes + kWarps * kBoxFloats * 4 + kWarps * kSlots * 8 + 16;
}  // namespace lv

struct LevelParams {
  const float* volume;
  const float* coords;      // [E,2,P]  (PC: [E,P,2], x and y interleaved, as lowMem_defSample receives them)
  long long off_edge_stride;   // PC: float2 elements between the offset slabs of consecutive edges (0: all edges read slab 0, Q2)
  float* offset;            // [E,P,49,2]  centre tap zeroed in place (Q5)
  float* corr;              // fwd: out [E,49,P]
  const float* corr_grad;   // bwd: [E,49,P]
  float* volume_grad;       // bwd: [E,P,H2,W2]
  float* offset_grad;       // bwd: [E,P,49,2]
  int P, tiles_per_edge, H2, W2;
};

struct LevelMap {
  CUtensorMap m;
};

// PC (forward only): lowMem_defSample's semantics on a materialised volume -- every bilinear corner gated on its own
// (quirk Q4; the zero-filled box provides exactly that), fractions x - floor(x) (lowMem_defSample.cu:87-88), interleaved
// coords, offset slab indexing with a stride (0 = the reference's offset[b*n] with N = 1, quirk Q2).
template <bool BWD, bool PC = false>
__global__ void __launch_bounds__(fl::kThreads, BWD ? 3 : 4)
lookup_level_tma_kernel(const __grid_constant__ LevelMap map, const LevelParams prm) {
  using namespace lv;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* boxes = reinterpret_cast<float*>(smem) + warp * kSlots * kBoxFloats;
  float* acc = reinterpret_cast<float*>(smem + kLvBoxes) + warp * kBoxFloats;            // BWD only
  float* s_tile = reinterpret_cast<float*>(smem + kLvBoxes);                                // FWD only (output tile)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kLvBoxes + (BWD ? kWarps * kBoxFloats * 4 : kSmemTilePad)) +
                   warp * kSlots;

  const int P = prm.P, H2 = prm.H2, W2 = prm.W2;
  const int n = blockIdx.x / prm.tiles_per_edge;
  const int p0 = (blockIdx.x - n * prm.tiles_per_edge) * kTile;
  const int pw = p0 + warp * kPixPerWarp;
  const size_t Q = (size_t)H2 * W2;

  if (lane == 0) {
    fl_mbar_init(bars + 0, 1);
    fl_mbar_init(bars + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  float cxm = 0.0f, cym = 0.0f;                                 // lane k (< 4) holds pixel k's coords
  if (lane < kPixPerWarp) {
    const int p = min(pw + lane, P - 1);
    if (PC) {
      const float2 c = __ldg(reinterpret_cast<const float2*>(prm.coords) + (size_t)n * P + p);
      cxm = c.x; cym = c.y;
    } else {
      cxm = __ldg(prm.coords + (size_t)n * 2 * P + p);
      cym = __ldg(prm.coords + (size_t)n * 2 * P + P + p);
    }
  }
  auto issue = [&](int k, float cx, float cy) {                 // lane 0 only
    const int slot = k & 1;
    const int pix = n * P + min(pw + k, P - 1);
    fl_mbar_expect_tx(bars + slot, kBoxFloats * 4);
    const int xb = box_origin_x(floor_to_int(cx), 7, W2), yb = box_origin_y(floor_to_int(cy), 7, H2);
    fl_tma_box(boxes + slot * kBoxFloats, &map.m, bars + slot, xb, yb, pix);
  };
  {
    const float c0x = __shfl_sync(0xffffffffu, cxm, 0), c0y = __shfl_sync(0xffffffffu, cym, 0);
    const float c1x = __shfl_sync(0xffffffffu, cxm, 1), c1y = __shfl_sync(0xffffffffu, cym, 1);
    if (lane == 0) {
      issue(0, c0x, c0y);
      issue(1, c1x, c1y);
    }
  }
  const int t0 = lane, t1 = lane + 32;
  const int i0 = t0 / RD, j0 = t0 - i0 * RD;
  const int t1c = min(t1, TAPS - 1);
  const int i1 = t1c / RD, j1 = t1c - i1 * RD;
  const bool has1 = t1 < TAPS;
  constexpr int CENTER = R * RD + R;

  if (BWD) {
    for (int q = lane; q < kBoxFloats; q += 32) acc[q] = 0.0f;
    __syncwarp();
  }

  // BWD: the upstream gradients of this warp's 4 pixels, tap rows t0 / t1: one 16-byte load each (the 4 pixels are
  // consecutive in memory).  A CTA-wide shared-memory tile + __syncthreads made every CTA start with an exposed
  // DRAM round trip (ncu: 21 % of the samples on that store).
  float4 g4a = make_float4(0.f, 0.f, 0.f, 0.f), g4b = g4a;
  if (BWD) {
    const float* g = prm.corr_grad + (size_t)n * TAPS * P + pw;
    if (pw + 3 < P && (P & 3) == 0) {
      g4a = __ldg(reinterpret_cast<const float4*>(g + (size_t)t0 * P));
      g4b = __ldg(reinterpret_cast<const float4*>(g + (size_t)t1c * P));
    } else {
      float va[4], vb[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const bool in = pw + q < P;
        va[q] = in ? __ldg(g + (size_t)t0 * P + q) : 0.0f;
        vb[q] = in ? __ldg(g + (size_t)t1c * P + q) : 0.0f;
      }
      g4a = make_float4(va[0], va[1], va[2], va[3]);
      g4b = make_float4(vb[0], vb[1], vb[2], vb[3]);
    }
  }
  float2 a0, a1;
  const size_t oslab = PC ? (size_t)n * prm.off_edge_stride : (size_t)n * P * TAPS;
  auto load_offsets = [&](int k) {
    const float2* O = reinterpret_cast<const float2*>(prm.offset) + oslab + (size_t)min(pw + k, P - 1) * TAPS;
    a0 = O[t0];
    a1 = O[t1c];
  };
  load_offsets(0);

#pragma unroll 1
  for (int k = 0; k < kPixPerWarp; ++k) {
    const int slot = k & 1;
    const int p = pw + k;
    const bool live = p < P;                                    // warp-uniform
    const size_t pix = (size_t)n * P + min(p, P - 1);
    const float x0 = __shfl_sync(0xffffffffu, cxm, k), y0 = __shfl_sync(0xffffffffu, cym, k);
    float2 oa = a0, ob = a1;
    if (lane == CENTER) oa = make_float2(0.0f, 0.0f);           // Q5: reads as zero, zeroed in memory below
    if (k + 1 < kPixPerWarp) load_offsets(k + 1);

    fl_mbar_wait(bars + slot, (k >> 1) & 1);
    const float* bx = boxes + slot * kBoxFloats;
    const float* V = prm.volume + pix * Q;
    const int xb = box_origin_x(floor_to_int(x0), 7, W2), yb = box_origin_y(floor_to_int(y0), 7, H2);
    const int pl = warp * kPixPerWarp + k;

    Tap ta, tb;
    {
      const float px = __fadd_rn(oa.x, x0), py = __fadd_rn(oa.y, y0);          // defCorrSample_kernel.cu:56-61
      const int fx = floor_to_int(px), fy = floor_to_int(py);
      ta.dx = PC ? __fsub_rn(px, floorf(px)) : __fsub_rn(px, (float)fx);
      ta.dy = PC ? __fsub_rn(py, floorf(py)) : __fsub_rn(py, (float)fy);
      tap_fetch<kBW01, kBH01, PC>(ta, bx, xb, yb, fx, fy, i0, j0, R, H2, W2);
    }
    {
      const float px = __fadd_rn(ob.x, x0), py = __fadd_rn(ob.y, y0);
      const int fx = floor_to_int(px), fy = floor_to_int(py);
      tb.dx = PC ? __fsub_rn(px, floorf(px)) : __fsub_rn(px, (float)fx);
      tb.dy = PC ? __fsub_rn(py, floorf(py)) : __fsub_rn(py, (float)fy);
      tap_fetch<kBW01, kBH01, PC>(tb, bx, xb, yb, fx, fy, i1, j1, R, H2, W2);
    }
    if (__any_sync(0xffffffffu, ta.miss || tb.miss)) {          // |offset| >= 4 (or a far out-of-range pixel)
      tap_patch_from_global<PC>(ta, V, H2, W2);
      tap_patch_from_global<PC>(tb, V, H2, W2);
    }
    if (!BWD) {
      s_tile[t0 * kOutPitchL + pl] = tap_value(ta);
      if (has1) s_tile[t1 * kOutPitchL + pl] = tap_value(tb);
    } else {
      const float ga = k == 0 ? g4a.x : (k == 1 ? g4a.y : (k == 2 ? g4a.z : g4a.w));
      const float gb = k == 0 ? g4b.x : (k == 1 ? g4b.y : (k == 2 ? g4b.z : g4b.w));
      // offset gradient (defCorrSample_kernel.cu:156-157, the reference's SASS operation order); 0 for gated taps
      auto ograd = [](const Tap& t, float g) {
        const float omdx = __fsub_rn(1.0f, t.dx), omdy = __fsub_rn(1.0f, t.dy);
        float ty = __fmaf_rn(-t.q11, omdx, -__fmul_rn(t.dx, t.q21));
        ty = __fmaf_rn(omdx, t.q12, ty);
        ty = __fmaf_rn(t.dx, t.q22, ty);
        float tx = __fmaf_rn(omdy, t.q21, -__fmul_rn(t.q11, omdy));
        tx = __fmaf_rn(-t.dy, t.q12, tx);
        tx = __fmaf_rn(t.dy, t.q22, tx);
        return make_float2(__fmul_rn(tx, g), __fmul_rn(ty, g));
      };
      if (live) {
        float2* GO = reinterpret_cast<float2*>(prm.offset_grad) + pix * TAPS;
        GO[t0] = ta.gate ? ograd(ta, ga) : make_float2(0.0f, 0.0f);
        if (has1) GO[t1] = tb.gate ? ograd(tb, gb) : make_float2(0.0f, 0.0f);
      }
      // scatter the corner weights into the box accumulator (taps may collide: shared-memory reductions; a
      // tag-arbitrated plain read-modify-write scheme was measured SLOWER: 221 vs 163 us at level 3, E=48)
      auto scatter = [&](const Tap& t, bool active, float g) {
        const int x2 = wrap_inc(t.x1), y2 = wrap_inc(t.y1);
        const bool xo = (unsigned)x2 < (unsigned)W2, yo = (unsigned)y2 < (unsigned)H2;
        const float omdx = __fsub_rn(1.0f, t.dx), omdy = __fsub_rn(1.0f, t.dy);
        const float w11 = __fmul_rn(__fmul_rn(omdy, omdx), g), w21 = __fmul_rn(__fmul_rn(omdy, t.dx), g);
        const float w12 = __fmul_rn(__fmul_rn(t.dy, omdx), g), w22 = __fmul_rn(__fmul_rn(t.dy, t.dx), g);
        if (active && t.gate && !t.miss) {
          const unsigned idx = ((unsigned)t.y1 - (unsigned)yb) * kBW01 + ((unsigned)t.x1 - (unsigned)xb);
          atomicAdd(acc + idx, w11);
          if (xo) atomicAdd(acc + idx + 1, w21);
          if (yo) atomicAdd(acc + idx + kBW01, w12);
          if (xo && yo) atomicAdd(acc + idx + kBW01 + 1, w22);
        }
      };
      scatter(ta, true, ga);
      scatter(tb, has1, gb);
      __syncwarp();
      // dense slice: zeros outside the box, the accumulator (re-zeroed) inside; then the rare out-of-box taps
      float* G = prm.volume_grad + pix * Q;
      if (live) {
        const int W4 = W2 >> 2;
        float4* G4 = reinterpret_cast<float4*>(G);
        const float4 z = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (W4 <= 32 && (32 % W4) == 0) {
          const int rpi = 32 / W4, ly = lane / W4, lx = lane - ly * W4;
          const unsigned rx = (unsigned)((lx << 2) - xb);
          const bool col_in = rx < (unsigned)kBW01;
          float4* g4 = G4 + ly * W4 + lx;
          for (int y = ly; y < H2; y += rpi, g4 += 32) {
            const unsigned ry = (unsigned)(y - yb);
            float4 v = z;
            if (col_in && ry < (unsigned)kBH01) {
              float4* a = reinterpret_cast<float4*>(acc + ry * kBW01 + rx);
              v = *a;
              *a = z;
            }
            __stcs(g4, v);
          }
        } else {
          const int n4 = H2 * W4;
          for (int q4 = lane; q4 < n4; q4 += 32) {
            const int y = q4 / W4, x = (q4 - y * W4) << 2;
            const unsigned ry = (unsigned)(y - yb), rx = (unsigned)(x - xb);
            float4 v = z;
            if (ry < (unsigned)kBH01 && rx < (unsigned)kBW01) {
              float4* a = reinterpret_cast<float4*>(acc + ry * kBW01 + rx);
              v = *a;
              *a = z;
            }
            __stcs(G4 + q4, v);
          }
        }
      } else {
        for (int q = lane; q < kBoxFloats; q += 32) acc[q] = 0.0f;
      }
      __syncwarp();
      if (live && __any_sync(0xffffffffu, ta.miss || (has1 && tb.miss))) {
        __threadfence();
        auto scatter_global = [&](const Tap& t, bool active, float g) {
          if (active && t.miss) {                                   // miss implies gate
            const int x2 = wrap_inc(t.x1), y2 = wrap_inc(t.y1);
            const bool xo = (unsigned)x2 < (unsigned)W2, yo = (unsigned)y2 < (unsigned)H2;
            const float omdx = __fsub_rn(1.0f, t.dx), omdy = __fsub_rn(1.0f, t.dy);
            float* pq = G + (size_t)t.y1 * W2 + t.x1;
            atomicAdd(pq, __fmul_rn(__fmul_rn(omdy, omdx), g));
            if (xo) atomicAdd(pq + 1, __fmul_rn(__fmul_rn(omdy, t.dx), g));
            if (yo) atomicAdd(pq + W2, __fmul_rn(__fmul_rn(t.dy, omdx), g));
            if (xo && yo) atomicAdd(pq + W2 + 1, __fmul_rn(__fmul_rn(t.dy, t.dx), g));
          }
        };
        scatter_global(ta, true, ga);
        scatter_global(tb, has1, gb);
      }
    }
    // Q5: the centre offset tap is zeroed in the caller's tensor (after every load of this record)
    if (live && lane == CENTER)
      reinterpret_cast<float2*>(prm.offset)[oslab + (size_t)min(p, P - 1) * TAPS + CENTER] = make_float2(0.0f, 0.0f);
    __syncwarp();
    if (k + 2 < kPixPerWarp) {
      const float nx = __shfl_sync(0xffffffffu, cxm, k + 2), ny = __shfl_sync(0xffffffffu, cym, k + 2);
      if (lane == 0) issue(k + 2, nx, ny);
    }
  }
  if (!BWD) {
    __syncthreads();
    const bool live = (p0 + lane) < P;
    float* out = prm.corr + (size_t)n * TAPS * P + p0 + lane;
    for (int t = warp; t < TAPS; t += kWarps)
      if (live) out[(size_t)t * P] = s_tile[t * kOutPitchL + lane];
  }
}

// Returns LGU_OK / error, or -1 when this path does not apply (the caller falls back to the generic kernels).
int launch_level_tma(bool bwd, const float* volume, const float* coords, float* offset, float* corr,
                     const float* corr_grad, float* volume_grad, float* offset_grad, int E, int H1, int W1, int H2,
                     int W2, cudaStream_t st) {
  const long long P = (long long)H1 * W1;
  if ((W2 & 3) != 0 || (reinterpret_cast<uintptr_t>(volume) & 15) != 0 || E * P >= 2147483647LL) return -1;
  if (bwd && (reinterpret_cast<uintptr_t>(volume_grad) & 15) != 0) return -1;
  LevelMap map;
  int rc = make_slice_map(&map.m, volume, E * P, H2, W2, fl::kBW01, fl::kBH01);
  if (rc) return rc;
  LevelParams prm;
  prm.volume = volume; prm.coords = coords; prm.offset = offset; prm.corr = corr; prm.corr_grad = corr_grad;
  prm.volume_grad = volume_grad; prm.offset_grad = offset_grad; prm.off_edge_stride = 0;
  prm.P = (int)P; prm.H2 = H2; prm.W2 = W2;
  prm.tiles_per_edge = (int)((P + fl::kTile - 1) / fl::kTile);
  const long long nblk = (long long)E * prm.tiles_per_edge;
  if (nblk >= 2147483647LL) return -1;
  if (bwd) {
    if (int rc = optin_smem(reinterpret_cast<const void*>(lookup_level_tma_kernel<true>), lv::kBwdSmem, "lgu_defcorr_index_backward")) return rc;
    lookup_level_tma_kernel<true><<<(unsigned)nblk, fl::kThreads, lv::kBwdSmem, st>>>(map, prm);
    return check_launch("lgu_defcorr_index_backward(tma)");
  }
  if (int rc = optin_smem(reinterpret_cast<const void*>(lookup_level_tma_kernel<false>), lv::kFwdSmem, "lgu_defcorr_index_forward")) return rc;
  lookup_level_tma_kernel<false><<<(unsigned)nblk, fl::kThreads, lv::kFwdSmem, st>>>(map, prm);
  return check_launch("lgu_defcorr_index_forward(tma)");
}

// lowMem_defSample's sampling step on a materialised volume [B, P, H2*W2] (radius 3): see lowmem.cu
int launch_level_tma_pc(const float* volume, const float* coords, float* offset, float* corr, int B, int H1, int W1, int H2,
                        int W2, long long off_edge_stride, cudaStream_t st) {
  const long long P = (long long)H1 * W1;
  LGU_REQUIRE((W2 & 3) == 0 && (reinterpret_cast<uintptr_t>(volume) & 15) == 0 && B * P < 2147483647LL,
              "lowMem sampler: volume must be 16-byte aligned with W2 %% 4 == 0");
  LevelMap map;
  int rc = make_slice_map(&map.m, volume, B * P, H2, W2, fl::kBW01, fl::kBH01);
  if (rc) return rc;
  LevelParams prm;
  prm.volume = volume; prm.coords = coords; prm.offset = offset; prm.corr = corr; prm.corr_grad = nullptr;
  prm.volume_grad = nullptr; prm.offset_grad = nullptr; prm.off_edge_stride = off_edge_stride;
  prm.P = (int)P; prm.H2 = H2; prm.W2 = W2;
  prm.tiles_per_edge = (int)((P + fl::kTile - 1) / fl::kTile);
  const long long nblk = (long long)B * prm.tiles_per_edge;
  LGU_REQUIRE(nblk < 2147483647LL, "lowMem sampler: grid too large");
  if (int rc2 = optin_smem(reinterpret_cast<const void*>(lookup_level_tma_kernel<false, true>), lv::kFwdSmem, "lgu_lowmem_defsample_forward")) return rc2;
  lookup_level_tma_kernel<false, true><<<(unsigned)nblk, fl::kThreads, lv::kFwdSmem, st>>>(map, prm);
  return check_launch("lgu_lowmem_defsample_forward(volume)");
}

}  // namespace lgu
