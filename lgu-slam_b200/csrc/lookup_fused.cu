// lookup_fused.cu -- CorrBlock.__call__'s data path in ONE launch (sm_100a, TMA-staged pyramid patches).
//
// Replaces the op sequence of /root/reference/droid_slam/modules/corr.py:88-109:
//   permute+contiguous (coords) -> 4x coords/2^l -> corr_index_forward(lvl1, r=1) -> permute -> var -> sigmoid
//   -> offset[1] *= mask -> 4x defCorr_index_forward(lvl l, r=3) -> view -> cat
// (5 extension launches + ~12 small torch kernels, every level's result written, re-read by cat and written
// again) by one kernel that reads coords once, stages the <= 16x16 footprint of every pyramid level in shared
// memory with TMA box loads, and writes the 196-channel output exactly once.
//
// Work split: CTA = 32 consecutive source pixels of one edge (8 warps x 4 pixels, ONE WARP PER PIXEL,
// lane = tap).  Per pixel one elected lane issues four cp.async.bulk.tensor.3d loads (tensor maps over
// [E*P, H2, W2] per level; out-of-bounds elements are zero-filled by the TMA unit):
//     level 0,1 : box 20 x 16  (16 columns [fx-7, fx+8] for |offset| < 4, +3 because the box start must be
//                               16-byte aligned -- measured: an unaligned inner start coordinate raises
//                               "illegal instruction" on sm_100a)
//     level 2,3 : box 12 x 8   (zero offsets: columns [fx-3, fx+4])
// into a per-warp 2-slot ring (mbarrier complete_tx), so the gathers of the next pixel are in flight while the
// current one is blended.  Taps that fall outside the staged box (|offset| >= 4, never produced by the
// reference's 4*tanh heads) take a direct global-load path with identical gating.
// The 196 x 32 result tile is transposed through padded shared memory and stored as 128-byte rows.
//
// Index logic is the reference's, bit for bit (defCorrSample_kernel.cu:56-67, corrSample_kernel.cu:52-60;
// quirks Q1, Q3, Q5, Q7): levels 0, 2, 3 are bit-identical to defCorr_index_forward; level 1 differs only through
// the fp32 rounding of the 9-tap variance / sigmoid that scales its offsets.
#include "fused_common.cuh"

namespace lgu {

#ifndef LGU_FWD_SLOTS
#define LGU_FWD_SLOTS 2   // measured (E = 48): 2 slots 239 us, 3 slots 248 us, 4 slots (one CTA per SM) 409 us
#endif
namespace flf {
constexpr int kSlots = LGU_FWD_SLOTS;                               // TMA ring depth of the forward (boxes in flight per warp)
constexpr int kSmemBoxes = fl::kWarps * kSlots * fl::kSlotBytes;
constexpr int kSmemBytes = kSmemBoxes + fl::kSmemOut + fl::kWarps * kSlots * 8;
}  // namespace flf

struct FusedLookupParams {
  const float* lvl[4];
  const float* coords;   // [E,P,2] (x,y) level-0 units
  const float* off0;     // [E,P,49,2]  read only (its centre tap is read as 0, Q5)
  float* off1;           // [E,P,49,2]  <- off1 * mask (Q7), every tap; the centre tap is read as 0
  float* out;            // [E,196,P]
  float* mask_out;       // [E,P] or null: the sigmoid(var) mask of this call
  int P, tiles_per_edge;
  int H2[4], W2[4];
  long long off_edge_stride;   // float2 elements between the offset slabs of consecutive edges (0: every edge reads slab 0, Q2)
  int apply_mask;              // 1: off1 <- off1 * sigmoid(var) of this call (CorrBlock); 0: offsets are used as given
  const int32_t* slots;        // [E] or null: edge n lives in pyramid / offset slot slots[n] (edge-slot pool)
};

template <bool PC>   // PC: per-corner gating (lowMem / altcorr semantics, Q4) instead of top-left gating (Q3)
__global__ void __launch_bounds__(fl::kThreads, 2)
lookup_fused_kernel(const __grid_constant__ FusedMaps maps, const FusedLookupParams prm) {
  using namespace fl;
  constexpr int kSlots = flf::kSlots, kSmemBoxes = flf::kSmemBoxes;
  extern __shared__ __align__(1024) uint8_t smem[];              // no static shared memory: base is 1024-aligned
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* boxes = reinterpret_cast<float*>(smem) + warp * kSlots * kSlotFloats;
  float* s_out = reinterpret_cast<float*>(smem + kSmemBoxes);    // [CH][kOutPitch]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemBoxes + kSmemOut) + warp * kSlots;

  const int P = prm.P;
  const int n = blockIdx.x / prm.tiles_per_edge;
  const int p0 = (blockIdx.x - n * prm.tiles_per_edge) * kTile;
  const int pw = p0 + warp * kPixPerWarp;                       // first pixel of this warp
  const int ns = prm.slots != nullptr ? __ldg(prm.slots + n) : n;   // storage slot of this edge (pyramid, offsets)

  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < kSlots; ++q) fl_mbar_init(bars + q, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  // coords of the warp's pixels: lane k (< 4) loads pixel k, everyone gets them by shuffle
  float2 cmine = make_float2(0.0f, 0.0f);
  if (lane < kPixPerWarp)
    cmine = __ldg(reinterpret_cast<const float2*>(prm.coords) + (size_t)n * P + min(pw + lane, P - 1));

  auto issue = [&](int k, float cx, float cy) {                 // lane 0 only
    const int slot = k % kSlots;
    const int pix = ns * P + min(pw + k, P - 1);
    float* dst = boxes + slot * kSlotFloats;
    fl_mbar_expect_tx(bars + slot, kSlotBytes);
    float sx = cx, sy = cy;
#pragma unroll
    for (int l = 0; l < LEVELS; ++l) {
      const int fx = floor_to_int(sx), fy = floor_to_int(sy);
      const int reach = l < 2 ? 7 : 3;
      const int xb = box_origin_x(fx, reach, prm.W2[l]), yb = box_origin_y(fy, reach, prm.H2[l]);
      const int o = l == 0 ? kOff0 : (l == 1 ? kOff1 : (l == 2 ? kOff2 : kOff3));
      fl_tma_box(dst + o, &maps.m[l], bars + slot, xb, yb, pix);
      sx = __fmul_rn(sx, 0.5f);
      sy = __fmul_rn(sy, 0.5f);
    }
  };
#pragma unroll
  for (int q = 0; q < kSlots && q < kPixPerWarp; ++q) {
    const float cqx = __shfl_sync(0xffffffffu, cmine.x, q), cqy = __shfl_sync(0xffffffffu, cmine.y, q);
    if (lane == 0) issue(q, cqx, cqy);
  }

  const int t0 = lane, t1 = lane + 32;                          // this lane's taps (t1 valid for lane < 17)
  const int i0 = t0 / RD, j0 = t0 - i0 * RD;
  const int t1c = min(t1, TAPS - 1);
  const int i1 = t1c / RD, j1 = t1c - i1 * RD;
  const bool has1 = t1 < TAPS;
  constexpr int CENTER = R * RD + R;                            // tap 24 (lane 24, pass 0)
  const int mi = min(lane, 8) / 3, mj = min(lane, 8) - mi * 3;  // r=1 mask taps on lanes 0..8

  float2 a0, a1, b0, b1;                                        // level-0 / level-1 offsets of taps t0, t1
  auto load_offsets = [&](int k) {
    const size_t opix = (size_t)ns * prm.off_edge_stride + (size_t)min(pw + k, P - 1) * TAPS;
    const float2* O0 = reinterpret_cast<const float2*>(prm.off0) + opix;
    const float2* O1 = reinterpret_cast<const float2*>(prm.off1) + opix;
    a0 = O0[t0]; a1 = O0[t1c]; b0 = O1[t0]; b1 = O1[t1c];
  };
  load_offsets(0);

  // One deformable level: both tap passes fetched together (8 shared loads in flight), then blended.
  auto deform_level = [&](const float* bx, const float* V, int H2, int W2, float cx, float cy, float2 oa, float2 ob,
                          float* so) {
    const int xb = box_origin_x(floor_to_int(cx), 7, W2), yb = box_origin_y(floor_to_int(cy), 7, H2);
    Tap ta, tb;
    {
      const float px = __fadd_rn(oa.x, cx), py = __fadd_rn(oa.y, cy);        // defCorrSample_kernel.cu:56-61
      const int fx = floor_to_int(px), fy = floor_to_int(py);
      ta.dx = PC ? __fsub_rn(px, floorf(px)) : __fsub_rn(px, (float)fx);     // lowMem_defSample.cu:87-88 uses floor()
      ta.dy = PC ? __fsub_rn(py, floorf(py)) : __fsub_rn(py, (float)fy);
      tap_fetch<kBW01, kBH01, PC>(ta, bx, xb, yb, fx, fy, i0, j0, R, H2, W2);
    }
    {
      const float px = __fadd_rn(ob.x, cx), py = __fadd_rn(ob.y, cy);
      const int fx = floor_to_int(px), fy = floor_to_int(py);
      tb.dx = PC ? __fsub_rn(px, floorf(px)) : __fsub_rn(px, (float)fx);
      tb.dy = PC ? __fsub_rn(py, floorf(py)) : __fsub_rn(py, (float)fy);
      tap_fetch<kBW01, kBH01, PC>(tb, bx, xb, yb, fx, fy, i1, j1, R, H2, W2);
    }
    if (__any_sync(0xffffffffu, ta.miss || tb.miss)) {          // |offset| >= 4: rare
      tap_patch_from_global<PC>(ta, V, H2, W2);
      tap_patch_from_global<PC>(tb, V, H2, W2);
    }
    so[t0 * kOutPitch] = tap_value(ta);
    if (has1) so[t1 * kOutPitch] = tap_value(tb);
  };
  // One zero-offset level (levels 2, 3): px = 0 + c, so floor / fraction are warp-uniform.
  auto uniform_level = [&](const float* bx, const float* V, int H2, int W2, float cx, float cy, float* so) {
    const float px = __fadd_rn(0.0f, cx), py = __fadd_rn(0.0f, cy);
    const int fx = floor_to_int(px), fy = floor_to_int(py);
    const int xb = box_origin_x(fx, 3, W2), yb = box_origin_y(fy, 3, H2);
    Tap ta, tb;
    ta.dx = tb.dx = PC ? __fsub_rn(px, floorf(px)) : __fsub_rn(px, (float)fx);
    ta.dy = tb.dy = PC ? __fsub_rn(py, floorf(py)) : __fsub_rn(py, (float)fy);
    tap_fetch<kBW23, kBH23, PC>(ta, bx, xb, yb, fx, fy, i0, j0, R, H2, W2);
    tap_fetch<kBW23, kBH23, PC>(tb, bx, xb, yb, fx, fy, i1, j1, R, H2, W2);
    if (__any_sync(0xffffffffu, ta.miss || tb.miss)) {          // only for clamped (far out-of-range) coords
      tap_patch_from_global<PC>(ta, V, H2, W2);
      tap_patch_from_global<PC>(tb, V, H2, W2);
    }
    so[t0 * kOutPitch] = tap_value(ta);
    if (has1) so[t1 * kOutPitch] = tap_value(tb);
  };

#pragma unroll 1
  for (int k = 0; k < kPixPerWarp; ++k) {
    const int slot = k % kSlots;
    const int p = pw + k;
    const bool live = p < P;                                    // warp-uniform
    const size_t pix = (size_t)ns * P + min(p, P - 1);          // slice index in the pyramid storage
    const float x0 = __shfl_sync(0xffffffffu, cmine.x, k), y0 = __shfl_sync(0xffffffffu, cmine.y, k);
    float2 o00 = a0, o01 = a1, o10 = b0, o11 = b1;              // this pixel's offsets (loaded one iteration ahead)
    if (lane == CENTER) o00 = make_float2(0.0f, 0.0f);          // Q5: the centre tap reads as 0
    if (k + 1 < kPixPerWarp) load_offsets(k + 1);

    fl_mbar_wait(bars + slot, (k / kSlots) & 1);
    const float* box = boxes + slot * kSlotFloats;
    float* so = s_out + warp * kPixPerWarp + k;                 // column of this pixel in the output tile

    // ---------------- level 1: r=1 mask lookup (corrSample_kernel.cu:52-77) -> var -> sigmoid
    const float x1c = __fmul_rn(x0, 0.5f), y1c = __fmul_rn(y0, 0.5f);
    const float* V1 = prm.lvl[1] + pix * (size_t)(prm.H2[1] * prm.W2[1]);
    float m;
    {
      const int H2 = prm.H2[1], W2 = prm.W2[1];
      const int fx = floor_to_int(x1c), fy = floor_to_int(y1c);
      const int xb = box_origin_x(fx, 7, W2), yb = box_origin_y(fy, 7, H2);
      Tap tm;
      tm.dx = __fsub_rn(x1c, floorf(x1c));
      tm.dy = __fsub_rn(y1c, floorf(y1c));
      tap_fetch<kBW01, kBH01, PC>(tm, box + kOff1, xb, yb, fx, fy, mi, mj, 1, H2, W2);
      if (__any_sync(0xffffffffu, tm.miss)) tap_patch_from_global<PC>(tm, V1, H2, W2);
      const float v = lane < 9 ? tap_value(tm) : 0.0f;
      // unbiased variance over the 9 taps (torch.var default, corr.py:96), then sigmoid (corr.py:97)
      float s = v;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);    // lanes 9..15 contribute 0
      const float mean = __shfl_sync(0xffffffffu, s, 0) / 9.0f;
      const float d = lane < 9 ? (v - mean) : 0.0f;
      float ss = d * d;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      const float var = __shfl_sync(0xffffffffu, ss, 0) * 0.125f;
      m = 1.0f / (1.0f + expf(-var));
    }
    // offset[1] <- offset[1] * mask (Q7) for every tap; the lookup itself reads the centre tap as 0 (Q5)
    if (prm.apply_mask) {
      o10 = make_float2(__fmul_rn(o10.x, m), __fmul_rn(o10.y, m));
      o11 = make_float2(__fmul_rn(o11.x, m), __fmul_rn(o11.y, m));
    }
    const float2 o10_store = o10;
    if (lane == CENTER) o10 = make_float2(0.0f, 0.0f);

    deform_level(box + kOff1, V1, prm.H2[1], prm.W2[1], x1c, y1c, o10, o11, so + TAPS * kOutPitch);
    deform_level(box + kOff0, prm.lvl[0] + pix * (size_t)(prm.H2[0] * prm.W2[0]), prm.H2[0], prm.W2[0], x0, y0, o00,
                 o01, so);
    const float x2c = __fmul_rn(x1c, 0.5f), y2c = __fmul_rn(y1c, 0.5f);
    uniform_level(box + kOff2, prm.lvl[2] + pix * (size_t)(prm.H2[2] * prm.W2[2]), prm.H2[2], prm.W2[2], x2c, y2c,
                  so + 2 * TAPS * kOutPitch);
    uniform_level(box + kOff3, prm.lvl[3] + pix * (size_t)(prm.H2[3] * prm.W2[3]), prm.H2[3], prm.W2[3],
                  __fmul_rn(x2c, 0.5f), __fmul_rn(y2c, 0.5f), so + 3 * TAPS * kOutPitch);

    // ---------------- in-place side effects on the caller's offsets
    if (live) {
      if (prm.apply_mask) {
        float2* O1 = reinterpret_cast<float2*>(prm.off1) + (size_t)ns * prm.off_edge_stride + (size_t)min(p, P - 1) * TAPS;
        O1[t0] = o10_store;
        if (has1) O1[t1] = o11;
      }
      if (lane == 0 && prm.mask_out != nullptr) prm.mask_out[(size_t)n * P + p] = m;
    }
    __syncwarp();                                               // every lane is done with this slot
    if (k + kSlots < kPixPerWarp) {
      const float nx = __shfl_sync(0xffffffffu, cmine.x, k + kSlots), ny = __shfl_sync(0xffffffffu, cmine.y, k + kSlots);
      if (lane == 0) issue(k + kSlots, nx, ny);
    }
  }

  __syncthreads();
  const bool live = (p0 + lane) < P;
  float* out = prm.out + (size_t)n * CH * P + p0 + lane;
  const float* srow = s_out + lane;
#pragma unroll 4
  for (int ch = warp; ch < CH; ch += kWarps)
    if (live) __stcs(out + (size_t)ch * P, srow[ch * kOutPitch]);
}

}  // namespace lgu

namespace lgu {
static int launch_lookup_fused(const float* lvl0, const float* lvl1, const float* lvl2, const float* lvl3,
                               const float* coords, const float* off0, float* off1, float* corr, float* mask_out, int E,
                               int H, int W, int num_levels, int radius, int per_corner, int shared_offsets,
                               int apply_mask, const int32_t* slots, int num_slots, void* stream);
}
extern "C" int lgu_corr_lookup_fused(const float* lvl0, const float* lvl1, const float* lvl2, const float* lvl3,
                                     const float* coords, const float* off0, float* off1, float* corr, float* mask_out,
                                     int E, int H, int W, int num_levels, int radius, void* stream) {
  return lgu::launch_lookup_fused(lvl0, lvl1, lvl2, lvl3, coords, off0, off1, corr, mask_out, E, H, W, num_levels,
                                  radius, 0, 0, 1, nullptr, E, stream);
}
extern "C" int lgu_corr_lookup_fused_slots(const float* lvl0, const float* lvl1, const float* lvl2, const float* lvl3,
                                           const float* coords, const float* off0, float* off1, float* corr,
                                           float* mask_out, const int32_t* slots, int num_slots, int E, int H, int W,
                                           int num_levels, int radius, void* stream) {
  LGU_REQUIRE(E == 0 || slots != nullptr, "lgu_corr_lookup_fused_slots: null slot list");
  LGU_REQUIRE(num_slots > 0, "lgu_corr_lookup_fused_slots: bad pool size %d", num_slots);
  return lgu::launch_lookup_fused(lvl0, lvl1, lvl2, lvl3, coords, off0, off1, corr, mask_out, E, H, W, num_levels,
                                  radius, 0, 0, 1, slots, num_slots, stream);
}
extern "C" int lgu_altcorr_lookup_fused(const float* lvl0, const float* lvl1, const float* lvl2, const float* lvl3,
                                        const float* coords, const float* off0, float* off1, float* corr,
                                        float* mask_out, int E, int H, int W, int num_levels, int radius,
                                        int shared_offsets, int apply_mask, void* stream) {
  return lgu::launch_lookup_fused(lvl0, lvl1, lvl2, lvl3, coords, off0, off1, corr, mask_out, E, H, W, num_levels,
                                  radius, 1, shared_offsets, apply_mask, nullptr, E, stream);
}
static int lgu::launch_lookup_fused(const float* lvl0, const float* lvl1, const float* lvl2, const float* lvl3,
                                    const float* coords, const float* off0, float* off1, float* corr, float* mask_out,
                                    int E, int H, int W, int num_levels, int radius, int per_corner,
                                    int shared_offsets, int apply_mask, const int32_t* slots, int num_slots,
                                    void* stream) {
  using namespace lgu;
  if (E == 0) return LGU_OK;
  LGU_REQUIRE(lvl0 && lvl1 && lvl2 && lvl3 && coords && off0 && off1 && corr, "lgu_corr_lookup_fused: null pointer");
  LGU_REQUIRE(E > 0 && H > 0 && W > 0, "lgu_corr_lookup_fused: bad sizes E=%d H=%d W=%d", E, H, W);
  if (num_levels != 4 || radius != 3 || (W % 32) != 0 || (H % 8) != 0) {
    set_error("lgu_corr_lookup_fused: only num_levels=4, radius=3, W%%32==0, H%%8==0 are implemented "
              "(got levels=%d r=%d H=%d W=%d); use the per-level operators", num_levels, radius, H, W);
    return LGU_ERR_UNSUPPORTED;
  }
  const int P = H * W;
  const long long nslices = (long long)num_slots * P;        // slices held by the storage (== E without a pool)
  LGU_REQUIRE(nslices < 2147483647LL, "lgu_corr_lookup_fused: slots*H*W = %lld exceeds the TMA coordinate range", nslices);
  const float* lv[4] = {lvl0, lvl1, lvl2, lvl3};
  for (int l = 0; l < 4; ++l)
    LGU_REQUIRE((reinterpret_cast<uintptr_t>(lv[l]) & 15) == 0, "lgu_corr_lookup_fused: level %d is not 16-byte aligned", l);
  FusedMaps maps;
  FusedLookupParams prm;
  for (int l = 0; l < 4; ++l) {
    prm.lvl[l] = lv[l];
    prm.H2[l] = H >> l;
    prm.W2[l] = W >> l;
    const int rc = make_slice_map(&maps.m[l], lv[l], nslices, H >> l, W >> l, l < 2 ? fl::kBW01 : fl::kBW23,
                                  l < 2 ? fl::kBH01 : fl::kBH23);
    if (rc) return rc;
  }
  prm.coords = coords; prm.off0 = off0; prm.off1 = off1; prm.out = corr; prm.mask_out = mask_out;
  prm.P = P;
  prm.tiles_per_edge = (P + fl::kTile - 1) / fl::kTile;
  const long long nblk = (long long)E * prm.tiles_per_edge;
  LGU_REQUIRE(nblk < 2147483647LL, "lgu_corr_lookup_fused: grid too large (%lld CTAs)", nblk);
  prm.off_edge_stride = shared_offsets ? 0 : (long long)P * fl::TAPS;
  prm.apply_mask = apply_mask;
  prm.slots = slots;
  LGU_REQUIRE(!(shared_offsets && apply_mask), "lgu_*_lookup_fused: apply_mask needs per-edge offsets");
  auto kern = per_corner ? lookup_fused_kernel<true> : lookup_fused_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, flf::kSmemBytes);
  if (e != cudaSuccess) {
    set_error("lgu_corr_lookup_fused: cannot opt in to %d B of shared memory: %s", flf::kSmemBytes, cudaGetErrorString(e));
    return LGU_ERR_LAUNCH;
  }
  kern<<<(unsigned)nblk, fl::kThreads, flf::kSmemBytes, (cudaStream_t)stream>>>(maps, prm);
  return check_launch("lgu_corr_lookup_fused");
}
