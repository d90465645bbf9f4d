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
#include <cuda_fp16.h>
#include "fused_common.cuh"

namespace lgu {

// Round-2 structure (profiles/r02_fused_lookup.md).  The round-1 kernel ran 9 tap passes per pixel (2 per level with 17
// of 32 lanes idle in every second one, plus the mask pass) at ~1000 warp instructions per pixel and was ISSUE-limited
// (53 % issue-active with 4 warps per scheduler) while DRAM sat at 53 % of peak.  Here the 196 taps of a pixel are ONE
// flat index space g = level*49 + tap walked in 7 passes of 32 lanes (6.1 needed); the 9 taps of the r=1 mask lookup
// ride in the idle lanes of the last pass, which runs first.  Everything level-dependent (level coordinates, box
// origin, box address and pitch, level extent) comes from a per-warp table in shared memory that the TMA-issuing lanes
// fill, so the pass body is branch-free and identical for every lane: 2 LDS.128 + ~35 ALU + 4 LDS + 1 STS.
namespace flf {
constexpr int kRing = 2;                                            // TMA ring depth (boxes in flight per warp)
constexpr int kPasses = 7;                                           // ceil(196 / 32)
constexpr int kRecBytes = 32;                                        // one table record = 2 x 16 B
constexpr int kRingBytes = fl::kWarps * kRing * fl::kSlotBytes;     // 53,248 B
constexpr int kSmemTab = fl::kWarps * kRing * fl::LEVELS * kRecBytes;   // 2,048 B
constexpr int kSmemBars = fl::kWarps * kRing * 8;
constexpr int kFwdSmemBytes = kRingBytes + fl::kSmemOut + kSmemTab + kSmemBars;
}  // namespace flf

struct FusedLookupParams {
  const float* lvl[4];
  const float* coords;   // [E,P,2] (x,y) level-0 units
  const float* off0;     // [E,P,49,2]  read only (its centre tap is read as 0, Q5)
  float* off1;           // [E,P,49,2]  write_back: <- off1 * mask (Q7), every tap; the centre tap is read as 0
  void* out;             // [E_out,196,P] fp32 (or fp16, see HALF): edge n is written to row out_index[n] (or n)
  const int32_t* out_index;    // [E] or null: destination row of every edge -- lets a rank of the sharded backend store
                               // straight into the gathered buffer on another GPU (NVLink peer memory), sharded.py
  float* mask_out;       // [E,P] or null: the sigmoid(var) mask of this call
  float* cum_mask;       // [slots,P] or null: running product of the masks of all calls so far (in/out); off1 stays pristine
  int P, tiles_per_edge;
  int H2[4], W2[4];
  long long off_edge_stride;   // float2 elements between the offset slabs of consecutive edges (0: every edge reads slab 0, Q2)
  int apply_mask;              // 1: level-1 offsets are scaled by sigmoid(var) of this call (CorrBlock); 0: used as given
  const int32_t* slots;        // [E] or null: edge n lives in pyramid / offset slot slots[n] (edge-slot pool)
};

__device__ __forceinline__ float4 flf_lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float flf_lds(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

// PC: per-corner gating (lowMem / altcorr semantics, Q4) instead of top-left gating (Q3)
// HALF: the 196-channel rows are stored as fp16 (round-to-nearest) -- what `update_op` reads under autocast
//       (factor_graph.py:284-286); halves the output stream (and the NVLink traffic of the sharded backend)
template <bool PC, bool HALF>
__global__ void __launch_bounds__(fl::kThreads, 2)
lookup_fused_kernel(const __grid_constant__ FusedMaps maps, const FusedLookupParams prm) {
  using namespace fl;
  using namespace flf;
  extern __shared__ __align__(1024) uint8_t smem[];              // no static shared memory: base is 1024-aligned
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* s_out = reinterpret_cast<float*>(smem + kRingBytes);    // [CH][kOutPitch]
  const uint32_t box_base = fl_smem_u32(smem) + warp * kRing * kSlotBytes;
  const uint32_t tab_base = fl_smem_u32(smem) + kRingBytes + kSmemOut + warp * kRing * LEVELS * kRecBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kRingBytes + kSmemOut + kSmemTab) + warp * kRing;

  const int P = prm.P;
  const int n = blockIdx.x / prm.tiles_per_edge;
  const int p0 = (blockIdx.x - n * prm.tiles_per_edge) * kTile;
  const int pw = p0 + warp * kPixPerWarp;                       // first pixel of this warp (tiles are always full)
  const int ns = prm.slots != nullptr ? __ldg(prm.slots + n) : n;   // storage slot of this edge (pyramid, offsets)

  // ---- per-lane level record: lanes 0..3 own level `lane` of the warp's table (static half written once)
  const int myl = lane & 3;
  const int myW2 = prm.W2[myl], myH2 = prm.H2[myl];
  if (lane < LEVELS * kRing) {
    const int q = lane >> 2;                                    // slot
    const int boff = myl == 0 ? kOff0 : (myl == 1 ? kOff1 : (myl == 2 ? kOff2 : kOff3));
    const int bw = myl < 2 ? kBW01 : kBW23;
    const uint32_t rec = tab_base + (q * LEVELS + myl) * kRecBytes;
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rec + 16), "r"(box_base + q * kSlotBytes + boff * 4),
                 "r"(bw * 4), "r"(myW2), "r"(myH2)
                 : "memory");
  }
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < kRing; ++q) fl_mbar_init(bars + q, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  // coords of the warp's pixels: lane k (< 4) loads pixel k, everyone gets them by shuffle
  float2 cmine = make_float2(0.0f, 0.0f);
  if (lane < kPixPerWarp) cmine = __ldg(reinterpret_cast<const float2*>(prm.coords) + (size_t)n * P + pw + lane);

  // Issue the four boxes of pixel k: lane l (< 4) scales the coordinates to level l (coords / 2^l as successive exact
  // halvings, corr.py:103), derives the box origin and writes the dynamic half of its record; lane 0 launches the copies.
  auto issue = [&](int k, float cx, float cy) {
    const int slot = k % kRing;
#pragma unroll
    for (int q = 1; q < LEVELS; ++q)
      if (myl >= q) { cx = __fmul_rn(cx, 0.5f); cy = __fmul_rn(cy, 0.5f); }
    const int reach = myl < 2 ? 7 : 3;
    const int xb = box_origin_x(floor_to_int(cx), reach, myW2), yb = box_origin_y(floor_to_int(cy), reach, myH2);
    if (lane < LEVELS)
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tab_base + (slot * LEVELS + myl) * kRecBytes),
                   "r"(__float_as_int(cx)), "r"(__float_as_int(cy)), "r"(xb), "r"(yb)
                   : "memory");
    int xs[LEVELS], ys[LEVELS];
#pragma unroll
    for (int l = 0; l < LEVELS; ++l) { xs[l] = __shfl_sync(0xffffffffu, xb, l); ys[l] = __shfl_sync(0xffffffffu, yb, l); }
    if (lane == 0) {
      const int pix = ns * P + pw + k;
      const uint32_t dst = box_base + slot * kSlotBytes;
      fl_mbar_expect_tx(bars + slot, kSlotBytes);
#pragma unroll
      for (int l = 0; l < LEVELS; ++l) {
        const int o = l == 0 ? kOff0 : (l == 1 ? kOff1 : (l == 2 ? kOff2 : kOff3));
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                dst + o * 4),
            "l"(&maps.m[l]), "r"(fl_smem_u32(bars + slot)), "r"(xs[l]), "r"(ys[l]), "r"(pix)
            : "memory");
      }
    }
    __syncwarp();                                               // table records visible to the whole warp
  };
#pragma unroll
  for (int q = 0; q < kRing && q < kPixPerWarp; ++q)
    issue(q, __shfl_sync(0xffffffffu, cmine.x, q), __shfl_sync(0xffffffffu, cmine.y, q));

  // ---- per-lane tap constants of the 7 passes: g = pass*32 + lane -> (level, i - r, j - r); the mask taps (r = 1 on
  // level 1) sit on lanes 4..12 of the last pass, whose lanes 0..3 are taps 45..48 of level 3
  constexpr int CENTER = R * RD + R;                            // tap 24
  int recoff[kPasses], di[kPasses], dj[kPasses];
#pragma unroll
  for (int ps = 0; ps < kPasses; ++ps) {
    const int g = ps * 32 + lane;
    int l = min(g, CH - 1) / TAPS;
    int t = min(g, CH - 1) - l * TAPS;
    int i = t / RD, j = t - i * RD, r = R;
    if (ps == kPasses - 1 && lane >= 4) {                       // mask taps (lanes 13..31: harmless duplicates of tap 8)
      const int mt = min(lane - 4, 8);
      l = 1; r = 1; i = mt / 3; j = mt - i * 3;
    }
    recoff[ps] = l * kRecBytes;
    di[ps] = i - r;
    dj[ps] = j - r;
  }
  const bool is_mask_lane = lane >= 4 && lane < 13;

  // offsets of the deformable levels: flat taps g < 98 live in passes 0..3 (pass 3: lanes 0, 1 only)
  float2 onext[4];
  auto load_offsets = [&](int k) {
    const size_t opix = (size_t)ns * prm.off_edge_stride + (size_t)(pw + k) * TAPS;
    const float2* O0 = reinterpret_cast<const float2*>(prm.off0) + opix;
    const float2* O1 = reinterpret_cast<const float2*>(prm.off1) + opix;
    onext[0] = O0[lane];
    onext[1] = lane < 17 ? O0[32 + lane] : O1[lane - 17];
    onext[2] = O1[15 + lane];
    onext[3] = lane < 2 ? O1[47 + lane] : make_float2(0.0f, 0.0f);
  };
  load_offsets(0);

  // One bilinear tap of the flat index space against the staged boxes (defCorrSample_kernel.cu:56-86).
  auto tap = [&](int slot, int ps, float2 o, size_t pix) -> float {
    const uint32_t rec = tab_base + slot * LEVELS * kRecBytes + recoff[ps];
    const float4 a = flf_lds128(rec);
    const float4 b = flf_lds128(rec + 16);
    const int xb = __float_as_int(a.z), yb = __float_as_int(a.w);
    const uint32_t baddr = (uint32_t)__float_as_int(b.x);
    const int pitch = __float_as_int(b.y);                      // box row pitch in bytes (80 or 48)
    const int W2 = __float_as_int(b.z), H2 = __float_as_int(b.w);
    const float px = __fadd_rn(o.x, a.x), py = __fadd_rn(o.y, a.y);
    const int fx = floor_to_int(px), fy = floor_to_int(py);
    Tap t;
    t.dx = PC ? __fsub_rn(px, floorf(px)) : __fsub_rn(px, (float)fx);     // lowMem_defSample.cu:87-88 uses floor()
    t.dy = PC ? __fsub_rn(py, floorf(py)) : __fsub_rn(py, (float)fy);
    t.x1 = (int)((unsigned)fx + (unsigned)di[ps]);
    t.y1 = (int)((unsigned)fy + (unsigned)dj[ps]);
    t.gate = PC ? true : (((unsigned)t.x1 < (unsigned)W2) && ((unsigned)t.y1 < (unsigned)H2));   // Q3 / Q4
    const unsigned rx = (unsigned)t.x1 - (unsigned)xb, ry = (unsigned)t.y1 - (unsigned)yb;
    // box extents from the pitch: 20 x 16 (pitch 80) or 12 x 8 (pitch 48)
    const bool inbox = rx < (unsigned)((pitch >> 2) - 1) && ry < (unsigned)(pitch == kBW01 * 4 ? kBH01 - 1 : kBH23 - 1);
    t.miss = t.gate && !inbox;
    const uint32_t ad = baddr + (inbox ? ry * (unsigned)pitch + rx * 4u : 0u);
    t.q11 = flf_lds(ad);
    t.q21 = flf_lds(ad + 4);
    t.q12 = flf_lds(ad + pitch);
    t.q22 = flf_lds(ad + pitch + 4);
    if (__any_sync(0xffffffffu, t.miss)) {                      // |offset| >= 4 or clamped far-out coords: rare
      const int l = recoff[ps] / kRecBytes;
      tap_patch_from_global<PC>(t, prm.lvl[l] + pix * (size_t)(H2 * W2), H2, W2);
    }
    return tap_value(t);
  };

#pragma unroll 1
  for (int k = 0; k < kPixPerWarp; ++k) {
    const int slot = k % kRing;
    const int p = pw + k;
    const size_t pix = (size_t)ns * P + p;                      // slice index in the pyramid storage
    float2 o[4] = {onext[0], onext[1], onext[2], onext[3]};     // this pixel's offsets (loaded one iteration ahead)
    if (k + 1 < kPixPerWarp) load_offsets(k + 1);
    float cum = 1.0f;
    if (prm.cum_mask != nullptr) cum = __ldg(prm.cum_mask + (size_t)ns * P + p);

    fl_mbar_wait(bars + slot, (k / kRing) & 1);
    float* so = s_out + warp * kPixPerWarp + k;                 // column of this pixel in the output tile

    // ---------------- last pass first: level-3 taps 45..48 (lanes 0..3) + the r=1 mask taps on level 1 (lanes 4..12,
    // corrSample_kernel.cu:52-77) -> unbiased variance over the 9 taps (torch.var default, corr.py:96) -> sigmoid
    float m;
    {
      const float v = tap(slot, kPasses - 1, make_float2(0.0f, 0.0f), pix);
      if (lane < 4) so[((kPasses - 1) * 32 + lane) * kOutPitch] = v;
      const float vm = is_mask_lane ? v : 0.0f;
      float s = vm;
#pragma unroll
      for (int sh = 8; sh > 0; sh >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sh);    // lanes 0..3, 13..15 contribute 0
      const float mean = __shfl_sync(0xffffffffu, s, 0) / 9.0f;
      const float d = is_mask_lane ? (vm - mean) : 0.0f;
      float ss = d * d;
#pragma unroll
      for (int sh = 8; sh > 0; sh >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, sh);
      const float var = __shfl_sync(0xffffffffu, ss, 0) * 0.125f;
      m = 1.0f / (1.0f + expf(-var));
    }
    // offset[1] <- offset[1] * mask (Q7) for every tap of level 1 (flat taps 49..97: pass 1 lanes >= 17, pass 2, pass 3
    // lanes 0..1); with a cumulative-mask buffer the stored offsets stay pristine and the product of all masks so far
    // is applied instead
    if (prm.apply_mask) {
      const float mm = prm.cum_mask != nullptr ? __fmul_rn(cum, m) : m;
      if (lane >= 17) o[1] = make_float2(__fmul_rn(o[1].x, mm), __fmul_rn(o[1].y, mm));
      o[2] = make_float2(__fmul_rn(o[2].x, mm), __fmul_rn(o[2].y, mm));
      o[3] = make_float2(__fmul_rn(o[3].x, mm), __fmul_rn(o[3].y, mm));         // lanes >= 2 hold zeros
      if (prm.cum_mask != nullptr) {
        if (lane == 0) prm.cum_mask[(size_t)ns * P + p] = mm;
      } else {
        float2* O1 = reinterpret_cast<float2*>(prm.off1) + (size_t)ns * prm.off_edge_stride + (size_t)p * TAPS;
        if (lane >= 17) O1[lane - 17] = o[1];
        O1[15 + lane] = o[2];
        if (lane < 2) O1[47 + lane] = o[3];
      }
    }
    if (lane == 0 && prm.mask_out != nullptr) prm.mask_out[(size_t)n * P + p] = m;
    // Q5: the centre taps read as 0 (flat taps 24 and 49 + 24 = 73 = pass 2, lane 9)
    if (lane == CENTER) o[0] = make_float2(0.0f, 0.0f);
    if (lane == TAPS + CENTER - 64) o[2] = make_float2(0.0f, 0.0f);

#pragma unroll
    for (int ps = 0; ps < kPasses - 1; ++ps) {
      const float v = tap(slot, ps, ps < 4 ? o[ps] : make_float2(0.0f, 0.0f), pix);
      so[(ps * 32 + lane) * kOutPitch] = v;
    }

    __syncwarp();                                               // every lane is done with this slot
    if (k + kRing < kPixPerWarp)
      issue(k + kRing, __shfl_sync(0xffffffffu, cmine.x, k + kRing), __shfl_sync(0xffffffffu, cmine.y, k + kRing));
  }

  __syncthreads();
  const size_t row = (size_t)(prm.out_index != nullptr ? __ldg(prm.out_index + n) : n) * CH * P + p0 + lane;
  const float* srow = s_out + lane;
  if (HALF) {
    __half* out = reinterpret_cast<__half*>(prm.out) + row;
#pragma unroll 4
    for (int ch = warp; ch < CH; ch += kWarps) out[(size_t)ch * P] = __float2half_rn(srow[ch * kOutPitch]);
  } else {
    float* out = reinterpret_cast<float*>(prm.out) + row;
#pragma unroll 4
    for (int ch = warp; ch < CH; ch += kWarps) __stcs(out + (size_t)ch * P, srow[ch * kOutPitch]);
  }
}

}  // namespace lgu

namespace lgu {
static int launch_lookup_fused(const float* lvl0, const float* lvl1, const float* lvl2, const float* lvl3,
                               const float* coords, const float* off0, float* off1, float* corr, float* mask_out, int E,
                               int H, int W, int num_levels, int radius, int per_corner, int shared_offsets,
                               int apply_mask, const int32_t* slots, int num_slots, float* cum_mask,
                               const int32_t* out_index, int out_half, void* stream);
}
extern "C" int lgu_corr_lookup_fused(const float* lvl0, const float* lvl1, const float* lvl2, const float* lvl3,
                                     const float* coords, const float* off0, float* off1, float* corr, float* mask_out,
                                     int E, int H, int W, int num_levels, int radius, void* stream) {
  return lgu::launch_lookup_fused(lvl0, lvl1, lvl2, lvl3, coords, off0, off1, corr, mask_out, E, H, W, num_levels,
                                  radius, 0, 0, 1, nullptr, E, nullptr, nullptr, 0, stream);
}
extern "C" int lgu_corr_lookup_fused_cum(const float* lvl0, const float* lvl1, const float* lvl2, const float* lvl3,
                                         const float* coords, const float* off0, const float* off1, float* cum_mask,
                                         float* corr, float* mask_out, const int32_t* slots, int num_slots, int E, int H,
                                         int W, int num_levels, int radius, void* stream) {
  LGU_REQUIRE(E == 0 || cum_mask != nullptr, "lgu_corr_lookup_fused_cum: null cumulative-mask buffer");
  LGU_REQUIRE(slots == nullptr || num_slots > 0, "lgu_corr_lookup_fused_cum: bad pool size %d", num_slots);
  return lgu::launch_lookup_fused(lvl0, lvl1, lvl2, lvl3, coords, off0, const_cast<float*>(off1), corr, mask_out, E, H, W,
                                  num_levels, radius, 0, 0, 1, slots, slots != nullptr ? num_slots : E, cum_mask, nullptr, 0, stream);
}
extern "C" int lgu_corr_lookup_fused_slots(const float* lvl0, const float* lvl1, const float* lvl2, const float* lvl3,
                                           const float* coords, const float* off0, float* off1, float* corr,
                                           float* mask_out, const int32_t* slots, int num_slots, int E, int H, int W,
                                           int num_levels, int radius, void* stream) {
  LGU_REQUIRE(E == 0 || slots != nullptr, "lgu_corr_lookup_fused_slots: null slot list");
  LGU_REQUIRE(num_slots > 0, "lgu_corr_lookup_fused_slots: bad pool size %d", num_slots);
  return lgu::launch_lookup_fused(lvl0, lvl1, lvl2, lvl3, coords, off0, off1, corr, mask_out, E, H, W, num_levels,
                                  radius, 0, 0, 1, slots, num_slots, nullptr, nullptr, 0, stream);
}
extern "C" int lgu_altcorr_lookup_fused(const float* lvl0, const float* lvl1, const float* lvl2, const float* lvl3,
                                        const float* coords, const float* off0, float* off1, float* corr,
                                        float* mask_out, int E, int H, int W, int num_levels, int radius,
                                        int shared_offsets, int apply_mask, void* stream) {
  return lgu::launch_lookup_fused(lvl0, lvl1, lvl2, lvl3, coords, off0, off1, corr, mask_out, E, H, W, num_levels,
                                  radius, 1, shared_offsets, apply_mask, nullptr, E, nullptr, nullptr, 0, stream);
}
extern "C" int lgu_altcorr_lookup_fused_into(const float* lvl0, const float* lvl1, const float* lvl2, const float* lvl3,
                                             const float* coords, const float* off0, float* off1, void* corr,
                                             const int32_t* out_index, int out_half, float* mask_out, int E, int H, int W,
                                             int num_levels, int radius, int shared_offsets, int apply_mask,
                                             void* stream) {
  return lgu::launch_lookup_fused(lvl0, lvl1, lvl2, lvl3, coords, off0, off1, reinterpret_cast<float*>(corr), mask_out, E,
                                  H, W, num_levels, radius, 1, shared_offsets, apply_mask, nullptr, E, nullptr, out_index,
                                  out_half, stream);
}
static int lgu::launch_lookup_fused(const float* lvl0, const float* lvl1, const float* lvl2, const float* lvl3,
                                    const float* coords, const float* off0, float* off1, float* corr, float* mask_out,
                                    int E, int H, int W, int num_levels, int radius, int per_corner,
                                    int shared_offsets, int apply_mask, const int32_t* slots, int num_slots,
                                    float* cum_mask, const int32_t* out_index, int out_half, void* stream) {
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
  prm.cum_mask = cum_mask;
  prm.out_index = out_index;
  prm.P = P;
  prm.tiles_per_edge = (P + fl::kTile - 1) / fl::kTile;
  const long long nblk = (long long)E * prm.tiles_per_edge;
  LGU_REQUIRE(nblk < 2147483647LL, "lgu_corr_lookup_fused: grid too large (%lld CTAs)", nblk);
  prm.off_edge_stride = shared_offsets ? 0 : (long long)P * fl::TAPS;
  prm.apply_mask = apply_mask;
  prm.slots = slots;
  LGU_REQUIRE(!(shared_offsets && apply_mask), "lgu_*_lookup_fused: apply_mask needs per-edge offsets");
  auto kern = per_corner ? (out_half ? lookup_fused_kernel<true, true> : lookup_fused_kernel<true, false>)
                         : (out_half ? lookup_fused_kernel<false, true> : lookup_fused_kernel<false, false>);
  if (int rc = optin_smem(reinterpret_cast<const void*>(kern), flf::kFwdSmemBytes, "lgu_corr_lookup_fused")) return rc;
  kern<<<(unsigned)nblk, fl::kThreads, flf::kFwdSmemBytes, (cudaStream_t)stream>>>(maps, prm);
  return check_launch("lgu_corr_lookup_fused");
}
