// Fused resample kernel of the preprocessing path (the default for every layout and filter): horizontal pass,
// vertical pass, normalisation and the padded store in ONE persistent kernel, with the uint8 intermediate of Pillow's
// two-pass resample (src/libImaging/Resample.c ImagingResampleHorizontal_8bpc / Vertical_8bpc) kept in shared
// memory.  Replaces `transform(img)` at functions/data_curation.py:675.
//
// Work item = (image, band of TH output rows); TH in {32,16,8,4,2,1} is chosen per image so that the source rows a
// band needs fit the 64-row intermediate (resample_plan_kernel writes one 64-byte record per item).  Per item:
//   1. the source rows of the band are staged in shared memory with 16-byte cp.async copies (row starts aligned
//      down to 16 bytes, the per-row byte shift is recomputed arithmetically), double buffered in chunks;
//   2. horizontal pass: one thread per output column, its tap weights in registers; a row's 3*NT source bytes
//      are fetched as aligned 32-bit words, re-aligned with one funnel shift per word and unpacked with PRMT, so
//      the inner loop is PRMT + IMAD with no byte loads.  Bilinear weights are non-negative and sum to 2^22 +- a
//      few, so with the weights pre-shifted by 2 the rounded, clipped uint8 result IS the top byte of the 32-bit
//      accumulator (2^23 + sum px * 4w): no shift, no clamp.  The three channels go to the intermediate as one
//      RGBX word;
//   3. vertical pass: one thread per output pixel: one 32-bit load per tap, PRMT + IMAD per channel, 768-entry
//      normalisation table (exactly torch's fp32 (v/255 - mean)/std, rounded to bf16), one 8-byte store of the
//      (R,G,B,0) pixel; the 3-pixel zero border of the NHWC4P layout is written by the same item.
// Images with more than 6 horizontal taps (downscales beyond 2.5x) and the Lanczos filter (negative weights) take
// a generic horizontal loop (byte loads from the staged rows, signed accumulators with Pillow's clip8); their
// items are queued first so the few long ones do not become the kernel's tail.
#pragma once
#include <cuda_bf16.h>
#include <cstdint>

#include "../../include/irp_b200.h"

namespace irp {

constexpr int kFCrop = IRP_CROP;
constexpr int kFPad = IRP_PAD_HW;
constexpr int kFThreads = 448;                 // 2 row groups x 224 output columns (7 warps each)
constexpr int kFInterRows = 64;                // source rows of one band (intermediate capacity)
constexpr int kFStageBytes = 20 * 1024;        // one staging buffer (two are used)
constexpr int kFStageSlack = 64;               // zero-weight taps of the last columns may read a few bytes past a row
constexpr int kFVcoefInts = 256;               // TH * (vertical taps) bound
constexpr int kFMaxBand = 32;
constexpr int kFPrecision = 22;                // Pillow PRECISION_BITS
constexpr int kFVFastTaps = 8;                 // vertical taps of the unrolled path (weights padded to 8 per row)
constexpr int kFSmemBytes = kFInterRows * kFCrop * 4 + 2 * (kFStageBytes + kFStageSlack) + 2 * kFMaxBand * 4 +
                            kFVcoefInts * 4 + 768 * 2 + 16;

// one work item (a band of `th` output rows of one image), written by resample_plan_kernel: everything the fused
// kernel needs to start staging the band's source rows comes with ONE 64-byte load instead of a chain of lookups
struct alignas(16) FusedItem {
  int32_t img, band, th, nth;        // image, band index, band height, max horizontal taps of the image
  int32_t ntv, row_lo, n_rows, col_first;  // max vertical taps; first source row / row count of the band; first byte
  int32_t width, row_bytes;          // bytes per staged row (needed window), bytes per image row
  int64_t offset;                    // byte offset of the image in the packed buffer
  int32_t pad0, pad1, pad2, pad3;
};
static_assert(sizeof(FusedItem) == 64, "FusedItem is one 64-byte record");

struct FusedParams {
  const uint8_t* pixels;
  const int32_t* plan;
  const int32_t* img_info;      // [n][4]: TH (0 = not handled here), horizontal taps, vertical taps, -
  const int32_t* status;        // [0] != 0: max_taps was too small for some image (outputs are poisoned)
  int32_t* counters;            // [0] heavy items, [1] normal items, [2] next item
  const FusedItem* items_heavy;
  const FusedItem* items_normal;
  const __nv_bfloat16* lut;     // [3][256]
  void* out;
  int max_taps;
  int out_size;                 // rows = columns of the output (224; 64 for IRP_TRANSFORM_HASH_64)
};

__host__ __device__ inline size_t fused_plan_ints_per_axis(int max_taps) {
  return static_cast<size_t>(kFCrop) * (2 + max_taps);
}

// ------------------------------------------------------------------------------------------------------------
// horizontal pass
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(
                   static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))),
               "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ int fused_clip8(int acc) {
  const int v = acc >> kFPrecision;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

struct HRowGeom {
  const uint8_t* stage;  // staging buffer of this chunk
  int pitch;             // bytes per staged row
  int rows;              // rows in this chunk
  int row0;              // first row of the chunk, relative to the band's first source row
  int s0;                // (address of the band's first needed byte of image row 0) & 15
  int rb16;              // row_bytes & 15
  int row_lo;            // band's first source row (image coordinates)
};

// NT-tap register-weight path (bilinear, NT <= 6)
template <int NT>
__device__ __forceinline__ void hpass_fast(const HRowGeom& g, const int32_t* __restrict__ plan_h, int T, int x, int rg,
                                           int col_first, uint32_t* __restrict__ inter) {
  constexpr int NW = (3 + 3 * NT + 3) / 4;  // aligned words that can hold 3*NT bytes at any byte offset
  constexpr int NA = (3 * NT + 3) / 4;      // words after re-alignment
  const int hn = plan_h[kFCrop + x];
  const int b0 = plan_h[x] * 3 - col_first;
  uint32_t w[NT];
#pragma unroll
  for (int t = 0; t < NT; ++t)
    w[t] = t < hn ? static_cast<uint32_t>(plan_h[2 * kFCrop + static_cast<size_t>(x) * T + t]) << 2 : 0u;
  for (int rl = rg; rl < g.rows; rl += 2) {
    const int rr = g.row0 + rl;
    const int bo = ((g.s0 + (g.row_lo + rr) * g.rb16) & 15) + b0;
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(g.stage + rl * g.pitch + (bo & ~3));
    uint32_t W[NW];
#pragma unroll
    for (int k = 0; k < NW; ++k) W[k] = wp[k];
    const uint32_t shb = static_cast<uint32_t>(bo & 3) * 8u;
    uint32_t A[NA];
#pragma unroll
    for (int k = 0; k < NA; ++k) A[k] = __funnelshift_r(W[k], k + 1 < NW ? W[k + 1] : 0u, shb);
    uint32_t a0 = 1u << 23, a1 = 1u << 23, a2 = 1u << 23;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      constexpr uint32_t kSel = 0x4440u;
      const int j = 3 * t;
      a0 += __byte_perm(A[j >> 2], 0u, kSel + (j & 3)) * w[t];
      a1 += __byte_perm(A[(j + 1) >> 2], 0u, kSel + ((j + 1) & 3)) * w[t];
      a2 += __byte_perm(A[(j + 2) >> 2], 0u, kSel + ((j + 2) & 3)) * w[t];
    }
    // top bytes of the three accumulators -> one RGBX word
    const uint32_t rg2 = __byte_perm(a0, a1, 0x0073u);
    inter[rr * kFCrop + x] = __byte_perm(rg2, a2, 0x7710u);
  }
}

// any tap count, signed weights allowed (Pillow's arithmetic verbatim)
__device__ __forceinline__ void hpass_generic(const HRowGeom& g, const int32_t* __restrict__ plan_h, int T, int x,
                                              int rg, int col_first, uint32_t* __restrict__ inter) {
  const int hn = plan_h[kFCrop + x];
  const int b0 = plan_h[x] * 3 - col_first;
  const int32_t* cfp = plan_h + 2 * kFCrop + static_cast<size_t>(x) * T;
  for (int rl = rg; rl < g.rows; rl += 2) {
    const int rr = g.row0 + rl;
    const int bo = ((g.s0 + (g.row_lo + rr) * g.rb16) & 15) + b0;
    const uint8_t* rp = g.stage + rl * g.pitch + bo;
    int a0 = 1 << (kFPrecision - 1), a1 = a0, a2 = a0;
#pragma unroll 4
    for (int t = 0; t < hn; ++t) {
      const int c = __ldg(cfp + t);
      a0 += static_cast<int>(rp[3 * t + 0]) * c;
      a1 += static_cast<int>(rp[3 * t + 1]) * c;
      a2 += static_cast<int>(rp[3 * t + 2]) * c;
    }
    inter[rr * kFCrop + x] = static_cast<uint32_t>(fused_clip8(a0)) | (static_cast<uint32_t>(fused_clip8(a1)) << 8) |
                             (static_cast<uint32_t>(fused_clip8(a2)) << 16);
  }
}

// ------------------------------------------------------------------------------------------------------------
// vertical pass
// ------------------------------------------------------------------------------------------------------------
template <int LAYOUT>
__device__ __forceinline__ void store_pixel(void* out, int img, int y, int x, uint32_t v0, uint32_t v1, uint32_t v2,
                                            const uint16_t* __restrict__ lut_s, bool poisoned, int out_size) {
  if (LAYOUT == IRP_LAYOUT_U8_HWC) {  // uint8 [n, out_size, out_size, 3]
    if (x < out_size) {
      uint8_t* o8 = static_cast<uint8_t*>(out) + ((static_cast<size_t>(img) * out_size + y) * out_size + x) * 3;
      o8[0] = static_cast<uint8_t>(v0);
      o8[1] = static_cast<uint8_t>(v1);
      o8[2] = static_cast<uint8_t>(v2);
    }
    return;
  }
  uint32_t l0 = lut_s[v0], l1 = lut_s[256 + v1], l2 = lut_s[512 + v2];
  if (poisoned) l0 = l1 = l2 = 0x7FC0u;
  if (LAYOUT == IRP_LAYOUT_NHWC4P) {
    uint2* row = static_cast<uint2*>(out) + (static_cast<size_t>(img) * kFPad + 3 + y) * kFPad;
    row[3 + x] = make_uint2(l0 | (l1 << 16), l2);
  } else {
    uint16_t* o16 = static_cast<uint16_t*>(out) + (static_cast<size_t>(img) * 3 * kFCrop + y) * kFCrop + x;
    o16[0] = static_cast<uint16_t>(l0);
    o16[static_cast<size_t>(kFCrop) * kFCrop] = static_cast<uint16_t>(l1);
    o16[static_cast<size_t>(2) * kFCrop * kFCrop] = static_cast<uint16_t>(l2);
  }
}

// NTV-tap unrolled path (non-negative weights, pre-shifted by 2, zero-padded to 8 per row)
template <int LAYOUT, int NTV>
__device__ __forceinline__ void vpass_fast(const uint32_t* __restrict__ inter, const int32_t* __restrict__ voff,
                                           const int32_t* __restrict__ vcoef, int th, int y0, int x, int rg, int img,
                                           void* out, const uint16_t* __restrict__ lut_s, bool poisoned, int out_size) {
  for (int yl = rg; yl < th; yl += 2) {
    const uint32_t* ip = inter + voff[yl] + x;
    uint32_t w[8];
    *reinterpret_cast<uint4*>(&w[0]) = *reinterpret_cast<const uint4*>(vcoef + yl * kFVFastTaps);
    if (NTV > 4) *reinterpret_cast<uint4*>(&w[4]) = *reinterpret_cast<const uint4*>(vcoef + yl * kFVFastTaps + 4);
    uint32_t a0 = 1u << 23, a1 = 1u << 23, a2 = 1u << 23;
#pragma unroll
    for (int t = 0; t < NTV; ++t) {
      const uint32_t px = ip[t * kFCrop];  // rows past the output's own taps have weight 0
      a0 += __byte_perm(px, 0u, 0x4440u) * w[t];
      a1 += __byte_perm(px, 0u, 0x4441u) * w[t];
      a2 += __byte_perm(px, 0u, 0x4442u) * w[t];
    }
    store_pixel<LAYOUT>(out, img, y0 + yl, x, a0 >> 24, a1 >> 24, a2 >> 24, lut_s, poisoned, out_size);
  }
}

// any tap count, signed weights allowed
template <int LAYOUT, bool SIGNED>
__device__ __forceinline__ void vpass_generic(const uint32_t* __restrict__ inter, const int32_t* __restrict__ voff,
                                              const int32_t* __restrict__ vcount, const int32_t* __restrict__ vcoef,
                                              int ntv, int th, int y0, int x, int rg, int img, void* out,
                                              const uint16_t* __restrict__ lut_s, bool poisoned, int out_size) {
  for (int yl = rg; yl < th; yl += 2) {
    const int vn = vcount[yl];
    const uint32_t* ip = inter + voff[yl] + x;
    const int32_t* wc = vcoef + yl * ntv;
    uint32_t v0, v1, v2;
    if (!SIGNED) {
      uint32_t a0 = 1u << 23, a1 = 1u << 23, a2 = 1u << 23;
      for (int t = 0; t < vn; ++t) {
        const uint32_t px = ip[t * kFCrop];
        const uint32_t wgt = static_cast<uint32_t>(wc[t]);
        a0 += __byte_perm(px, 0u, 0x4440u) * wgt;
        a1 += __byte_perm(px, 0u, 0x4441u) * wgt;
        a2 += __byte_perm(px, 0u, 0x4442u) * wgt;
      }
      v0 = a0 >> 24;
      v1 = a1 >> 24;
      v2 = a2 >> 24;
    } else {
      int a0 = 1 << (kFPrecision - 1), a1 = a0, a2 = a0;
      for (int t = 0; t < vn; ++t) {
        const uint32_t px = ip[t * kFCrop];
        const int wgt = wc[t];
        a0 += static_cast<int>(px & 255u) * wgt;
        a1 += static_cast<int>((px >> 8) & 255u) * wgt;
        a2 += static_cast<int>((px >> 16) & 255u) * wgt;
      }
      v0 = fused_clip8(a0);
      v1 = fused_clip8(a1);
      v2 = fused_clip8(a2);
    }
    store_pixel<LAYOUT>(out, img, y0 + yl, x, v0, v1, v2, lut_s, poisoned, out_size);
  }
}

// ------------------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------------------
template <int LAYOUT, bool SIGNED>
__global__ void __launch_bounds__(kFThreads, 2) resample_fused_kernel(const FusedParams p) {
  extern __shared__ __align__(16) uint8_t fsmem[];
  uint32_t* inter = reinterpret_cast<uint32_t*>(fsmem);
  uint8_t* stage0 = fsmem + kFInterRows * kFCrop * 4;
  uint8_t* stage1 = stage0 + kFStageBytes + kFStageSlack;
  int32_t* vcoef = reinterpret_cast<int32_t*>(stage1 + kFStageBytes + kFStageSlack);  // 16-byte aligned
  int32_t* vcount = vcoef + kFVcoefInts;
  int32_t* voff = vcount + kFMaxBand;
  uint16_t* lut_s = reinterpret_cast<uint16_t*>(voff + kFMaxBand);
  int32_t* s_item = reinterpret_cast<int32_t*>(lut_s + 768);  // [2]

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int rg = tid >= kFCrop ? 1 : 0;
  const int x = tid - rg * kFCrop;
  const int T = p.max_taps;
  for (int i = tid; i < 768; i += kFThreads) lut_s[i] = reinterpret_cast<const uint16_t*>(p.lut)[i];
  const int n_heavy = p.counters[0];
  const int n_total = n_heavy + p.counters[1];
  // a caller-supplied tap bound that truncated a filter must not yield plausible pixels: bf16 outputs become NaN
  const bool poisoned = p.status[0] != 0;

  struct Geo {  // staging geometry of one item
    const uint8_t* src;
    int row_lo, n_rows, col_first, width, pitch, rows_cap, s0, rb16, row_bytes;
  };
  auto load_item = [&](int k) -> FusedItem {
    const FusedItem* rec = k < n_heavy ? p.items_heavy + k : p.items_normal + (k - n_heavy);
    FusedItem it;
    const int4* r4 = reinterpret_cast<const int4*>(rec);
    int4* d4 = reinterpret_cast<int4*>(&it);
    d4[0] = r4[0];
    d4[1] = r4[1];
    d4[2] = r4[2];
    return it;
  };
  auto geometry = [&](const FusedItem& it) -> Geo {
    Geo g;
    g.src = p.pixels + it.offset;
    g.row_lo = it.row_lo;
    g.n_rows = it.n_rows;
    g.col_first = it.col_first;
    g.width = it.width;
    g.pitch = ((it.width + 31) + 15) & ~15;
    g.rows_cap = kFStageBytes / g.pitch;
    g.row_bytes = it.row_bytes;
    g.s0 = static_cast<int>((reinterpret_cast<uintptr_t>(g.src) + static_cast<uintptr_t>(it.col_first)) & 15u);
    g.rb16 = it.row_bytes & 15;
    return g;
  };
  // stage rows [c0, c0 + rows) of the band: one warp per row, 16-byte cp.async copies
  auto issue = [&](const Geo& g, int c0, uint8_t* buf) {
    const int rows = min(g.rows_cap, g.n_rows - c0);
    for (int rl = warp; rl < rows; rl += kFThreads / 32) {
      const int r = g.row_lo + c0 + rl;
      const int sh = (g.s0 + r * g.rb16) & 15;
      // first needed byte of the row, aligned down to 16; the copy may run up to 15 bytes past the row's last
      // needed byte (images start on 16-byte boundaries of the packed buffer and are padded to them)
      const uint8_t* gp = g.src + static_cast<size_t>(r) * g.row_bytes + g.col_first - sh;
      uint8_t* d = buf + rl * g.pitch;
      const int nvec = (sh + g.width + 15) >> 4;
      for (int v = lane; v < nvec; v += 32) cp_async16(d + (v << 4), gp + (v << 4));
    }
    cp_async_commit();
  };

  if (tid == 0) s_item[0] = atomicAdd(&p.counters[2], 1);
  __syncthreads();
  int k = s_item[0];
  if (k >= n_total) return;
  FusedItem cur = load_item(k);
  issue(geometry(cur), 0, stage0);  // every item's first chunk goes to stage0

  for (int it = 0;; ++it) {
    const int img = cur.img, band = cur.band, th = cur.th, nth = cur.nth, ntv = cur.ntv;
    const int y0 = band * th;
    const Geo g = geometry(cur);
    const int32_t* plan_h = p.plan + (static_cast<size_t>(img) * 2 + 0) * fused_plan_ints_per_axis(T);
    const int32_t* plan_v = p.plan + (static_cast<size_t>(img) * 2 + 1) * fused_plan_ints_per_axis(T);
    const bool vfast = !SIGNED && ntv <= kFVFastTaps;
    const int vstride = vfast ? kFVFastTaps : ntv;
    if (tid < th) {
      vcount[tid] = plan_v[kFCrop + y0 + tid];
      voff[tid] = (plan_v[y0 + tid] - g.row_lo) * kFCrop;
    }
    for (int i = tid; i < th * vstride; i += kFThreads) {
      const int yl = i / vstride, t = i - yl * vstride;
      int c = 0;
      if (t < plan_v[kFCrop + y0 + yl]) c = plan_v[2 * kFCrop + static_cast<size_t>(y0 + yl) * T + t];
      vcoef[i] = SIGNED ? c : (c << 2);
    }
    // the NEXT item's index: fetched now, read after the first barrier below; its record is loaded during this
    // item's horizontal pass and its first chunk is staged while this item's vertical pass runs
    if (tid == 0) s_item[(it + 1) & 1] = atomicAdd(&p.counters[2], 1);

    int ci = 0;
    int k_next = n_total;
    FusedItem nxt;
    nxt.img = -1;
    for (int c0 = 0; c0 < g.n_rows; c0 += g.rows_cap, ++ci) {
      uint8_t* buf = (ci & 1) ? stage1 : stage0;
      if (c0 + g.rows_cap < g.n_rows) {
        issue(g, c0 + g.rows_cap, (ci & 1) ? stage0 : stage1);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();  // chunk landed (and, first time round, the vertical tables and the next index are written)
      if (ci == 0) {
        k_next = s_item[(it + 1) & 1];
        if (k_next < n_total) nxt = load_item(k_next);
      }
      HRowGeom hg;
      hg.stage = buf;
      hg.pitch = g.pitch;
      hg.rows = min(g.rows_cap, g.n_rows - c0);
      hg.row0 = c0;
      hg.s0 = g.s0;
      hg.rb16 = g.rb16;
      hg.row_lo = g.row_lo;
      if (SIGNED || nth > 6) hpass_generic(hg, plan_h, T, x, rg, g.col_first, inter);
      else if (nth <= 3) hpass_fast<3>(hg, plan_h, T, x, rg, g.col_first, inter);
      else if (nth == 4) hpass_fast<4>(hg, plan_h, T, x, rg, g.col_first, inter);
      else if (nth == 5) hpass_fast<5>(hg, plan_h, T, x, rg, g.col_first, inter);
      else hpass_fast<6>(hg, plan_h, T, x, rg, g.col_first, inter);
      __syncthreads();  // the staging buffer may be refilled; after the last chunk: the intermediate is complete
    }
    // both staging buffers are free: start the next item's first chunk before the vertical pass
    if (k_next < n_total) issue(geometry(nxt), 0, stage0);

    // ---- vertical pass + normalise + store ----
    if (!vfast) vpass_generic<LAYOUT, SIGNED>(inter, voff, vcount, vcoef, ntv, th, y0, x, rg, img, p.out, lut_s, poisoned, p.out_size);
    else if (ntv <= 3) vpass_fast<LAYOUT, 3>(inter, voff, vcoef, th, y0, x, rg, img, p.out, lut_s, poisoned, p.out_size);
    else if (ntv == 4) vpass_fast<LAYOUT, 4>(inter, voff, vcoef, th, y0, x, rg, img, p.out, lut_s, poisoned, p.out_size);
    else if (ntv <= 6) vpass_fast<LAYOUT, 6>(inter, voff, vcoef, th, y0, x, rg, img, p.out, lut_s, poisoned, p.out_size);
    else vpass_fast<LAYOUT, 8>(inter, voff, vcoef, th, y0, x, rg, img, p.out, lut_s, poisoned, p.out_size);
    if (LAYOUT == IRP_LAYOUT_NHWC4P) {
      // zero borders: 3 pixels left and right of every row of the band, 3 full rows above / below the image
      uint2* base = static_cast<uint2*>(p.out) + static_cast<size_t>(img) * kFPad * kFPad;
      for (int i = tid; i < th * 6; i += kFThreads) {
        const int yl = i / 6, j = i - yl * 6;
        base[(3 + y0 + yl) * kFPad + (j < 3 ? j : kFPad - 6 + j)] = make_uint2(0u, 0u);
      }
      if (band == 0)
        for (int i = tid; i < 3 * kFPad; i += kFThreads) base[i] = make_uint2(0u, 0u);
      if (y0 + th == kFCrop)
        for (int i = tid; i < 3 * kFPad; i += kFThreads) base[(3 + kFCrop) * kFPad + i] = make_uint2(0u, 0u);
    }
    if (k_next >= n_total) break;
    cur = nxt;
    __syncthreads();  // tables, intermediate and the item slot are reused by the next item
  }
}

}  // namespace irp
