// 7x7 / stride 2 / pad 3 stem convolution (3 -> 64) + folded BN + ReLU on tcgen05, without im2col traffic.
//
// Input: the padded NHWC4 bf16 batch [B,230,230,4] written by the preprocess kernel (3-pixel zero border, 4th
// channel zero).  For one CTA tile of 8 (wo) x 16 (ho) output pixels the 22 x 37 pixel input patch (6.5 KB) is
// fetched ONCE by TMA into shared memory with no swizzle.  The A operand of every MMA is then described straight
// on top of that patch:  for filter row r and pixel group h (4 pixels x 4 channels = K 16)
//     row i of an 8-row core matrix  = output pixel wo0+i   -> patch byte  (2*i)*8     = i*16   (core-matrix row pitch)
//     the second 16-byte K chunk     = the next two pixels  -> +16 bytes               (LBO = 16, overlapping)
//     the next 8-row group           = the next output row  -> +2 patch rows           (SBO = 2*176)
// which is exactly the no-swizzle K-major canonical layout, so the stride-2 window overlap is expressed by the
// descriptor strides instead of being materialised.  The folded weights (64 x 7 x 8 x 4, 28 KB, pre-arranged in
// core-matrix order) stay resident in shared memory for the whole persistent kernel.
//
// Replaces conv1/bn1/relu of torchvision/models/resnet.py:197-199 (called from functions/data_curation.py:677).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdint>

#include "ptx.cuh"

namespace irp {

constexpr int kStemWeightBytes = 7 * 2 * 8 * 256;           // (filter row, pixel group) x 8 n-groups x 256 B = 28672

// no-swizzle K-major descriptor: LBO = byte distance between the two 16-byte K chunks, SBO = between 8-row groups
__device__ __forceinline__ uint64_t umma_desc_noswz(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;  // descriptor version 1
  return d;         // layout_type 0 = SWIZZLE_NONE
}

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* desc, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}


// ------------------------------------------------------------------------------------------------------------
// Stem + 3x3/2 max pool fused (conv1/bn1/relu/maxpool of torchvision/models/resnet.py:197-200).
//
// Writing the 112x112x64 stem output (411 MB per 256 images) and reading it back for the pool costs more than
// the stem's arithmetic.  Here one CTA tile is a 7 x 7 block of POOLED outputs (56 = 8 x 7, so the tiles cover the
// image exactly): it needs the 16 x 16 block of stem pixels starting at stem coordinate 14*t - 1, computed as two
// 8 (w) x 16 (h) GEMM sub-tiles from one 38 x 37 pixel input patch (same overlapping no-swizzle descriptors as
// stem_conv_kernel; the patch pitch is padded to 38 pixels so that odd filter rows stay 16-byte aligned).  The
// epilogue adds bias, applies ReLU, zeroes the stem pixels at coordinate -1 (the pool's padding: all values are
// >= 0 after ReLU and every window holds a real pixel, so 0 is neutral), parks the 256 x 64 bf16 block in shared
// memory, and the same 256 threads then take the 3x3 maxima and write the 49 x 64 pooled block with 16-byte
// stores.  22 % of the stem pixels are computed twice; the 411 MB round trip disappears.
// ------------------------------------------------------------------------------------------------------------
constexpr int kSpThreads = 320;
constexpr int kSpEpiThreads = 256;
constexpr int kSpPatchW = 38, kSpPatchH = 37;
constexpr int kSpPitch = kSpPatchW * 8;                  // 304 bytes per patch row
constexpr int kSpPatchBytes = kSpPatchH * kSpPitch;      // 11248
constexpr int kSpPatchStride = 11264;                    // 128-byte aligned slot
constexpr int kSpSlots = 6;
constexpr int kSpScratchBytes = 256 * 128;               // 16 x 16 stem pixels x 64 channels bf16
constexpr int kSpSmemBytes = kStemWeightBytes + kSpSlots * kSpPatchStride + 2 * kSpScratchBytes + 512 + 1024;

struct alignas(64) StemPoolParams {
  CUtensorMap tmIn;              // (920 elements per row, 230 rows, B images), box (152, 37, 1), no swizzle
  const __nv_bfloat16* weights;  // kStemWeightBytes, core-matrix order (stem_fold_kernel)
  const float* bias;             // [64]
  __nv_bfloat16* out;            // [batch, 56, 56, 64] pooled output of this launch's first image onwards
  int n_base;                    // first image inside tmIn
  int num_tiles;                 // batch * 64
};

__global__ void __launch_bounds__(kSpThreads, 1) stem_pool_kernel(const __grid_constant__ StemPoolParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* smem_w = smem;
  uint8_t* smem_patch = smem_w + kStemWeightBytes;
  uint8_t* smem_scr = smem_patch + kSpSlots * kSpPatchStride;  // 2 scratch blocks
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_scr + 2 * kSpScratchBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kSpSlots;
  uint64_t* tfull_bar = bars + 2 * kSpSlots;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* sbias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);  // [64]
  constexpr uint32_t kTmemCols = 256;  // 2 buffers x 2 sub-tiles x 64 fp32 columns

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmIn);
    for (int i = 0; i < kSpSlots; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.weights);
    uint4* dst = reinterpret_cast<uint4*>(smem_w);
    for (int i = threadIdx.x; i < kStemWeightBytes / 16; i += kSpThreads) dst[i] = __ldg(src + i);
    if (threadIdx.x < 64) sbias[threadIdx.x] = __ldg(p.bias + threadIdx.x);
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();

  if (warp == 0) {
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int tx = tile & 7, ty = (tile >> 3) & 7, n = tile >> 6;
        mbar_wait(&empty_bar[slot], phase ^ 1);
        mbar_arrive_expect_tx(&full_bar[slot], kSpPatchBytes);
        // stem pixel sx reads padded-input columns 2*sx .. 2*sx+6; the block starts at stem coordinate 14*t - 1
        tma_load_3d(smem_patch + slot * kSpPatchStride, &p.tmIn, &full_bar[slot], (28 * tx - 2) * 4, 28 * ty - 2,
                    n + p.n_base);
        if (++slot == kSpSlots) {
          slot = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    {  // whole warp, converged; the tcgen05 instructions are issued by the elected lane (elect_one(), ptx.cuh)
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64);
      const uint32_t w_addr = smem_u32(smem_w);
      int slot = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        mbar_wait(&full_bar[slot], phase);
        tc_fence_after();
        const uint32_t patch = smem_u32(smem_patch + slot * kSpPatchStride);
        if (elect_one()) {
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            const uint32_t tmem_d = tmem_base + acc * 128 + sub * 64;
#pragma unroll
            for (int r = 0; r < 7; ++r) {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const uint64_t da = umma_desc_noswz(patch + r * kSpPitch + h * 32 + sub * 128, 16, 2 * kSpPitch);
                const uint64_t db = umma_desc_noswz(w_addr + (r * 2 + h) * 2048, 128, 256);
                umma_bf16(tmem_d, da, db, idesc, (r | h) != 0 ? 1u : 0u);
              }
            }
          }
          umma_commit(&empty_bar[slot]);
          umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++slot == kSpSlots) {
          slot = 0;
          phase ^= 1;
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    const int quarter = warp & 3;
    const int sub = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;            // GEMM row inside the sub-tile = ry * 8 + (rx - 8*sub)
    const int ry = row >> 3, rx = 8 * sub + (row & 7);
    const int spx = ry * 16 + rx;                   // stem pixel index inside the 16 x 16 block
    const uint32_t scr_row = static_cast<uint32_t>(spx) * 128u;
    const uint32_t swz = static_cast<uint32_t>(spx & 7);
    const int et = threadIdx.x - 64;                // 0..255
    int acc = 0;
    uint32_t acc_phase = 0;
    int j = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++j) {
      const int tx = tile & 7, ty = (tile >> 3) & 7, n = tile >> 6;
      uint8_t* scr = smem_scr + (j & 1) * kSpScratchBytes;
      const bool pad_px = (ry == 0 && ty == 0) || (rx == 0 && tx == 0);  // stem coordinate -1
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * 128 + sub * 64 + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll
      for (int c = 0; c < 64; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + c, v);
        const float4* bp = reinterpret_cast<const float4*>(sbias + c);
        float4 bv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) bv[i] = bp[i];
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 b0 = bv[2 * i], b1 = bv[2 * i + 1];
          float x0 = fmaxf(__uint_as_float(v[8 * i + 0]) + b0.x, 0.f);
          float x1 = fmaxf(__uint_as_float(v[8 * i + 1]) + b0.y, 0.f);
          float x2 = fmaxf(__uint_as_float(v[8 * i + 2]) + b0.z, 0.f);
          float x3 = fmaxf(__uint_as_float(v[8 * i + 3]) + b0.w, 0.f);
          float x4 = fmaxf(__uint_as_float(v[8 * i + 4]) + b1.x, 0.f);
          float x5 = fmaxf(__uint_as_float(v[8 * i + 5]) + b1.y, 0.f);
          float x6 = fmaxf(__uint_as_float(v[8 * i + 6]) + b1.z, 0.f);
          float x7 = fmaxf(__uint_as_float(v[8 * i + 7]) + b1.w, 0.f);
          uint4 o = make_uint4(pack_bf16x2(x0, x1), pack_bf16x2(x2, x3), pack_bf16x2(x4, x5), pack_bf16x2(x6, x7));
          if (pad_px) o = make_uint4(0u, 0u, 0u, 0u);
          *reinterpret_cast<uint4*>(scr + scr_row + ((((c >> 3) + i) ^ swz) << 4)) = o;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      named_bar_sync(1, kSpEpiThreads);  // the whole 16 x 16 block is in scratch
      // 3x3 / stride 2 maxima: work item = (pooled pixel 0..48, 8-channel group 0..7)
      __nv_bfloat16* out_img = p.out + static_cast<size_t>(n) * 56 * 56 * 64;
#pragma unroll 1
      for (int id = et; id < 49 * 8; id += kSpEpiThreads) {
        const int pp = id >> 3, cg = id & 7;
        const int py = pp / 7, px = pp - py * 7;
        __nv_bfloat162 m[4];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const int s = (2 * py + dy) * 16 + 2 * px + dx;
            const uint4 val = *reinterpret_cast<const uint4*>(scr + s * 128 + ((cg ^ (s & 7)) << 4));
            const __nv_bfloat162* pv = reinterpret_cast<const __nv_bfloat162*>(&val);
            if (dy == 0 && dx == 0) {
#pragma unroll
              for (int q = 0; q < 4; ++q) m[q] = pv[q];
            } else {
#pragma unroll
              for (int q = 0; q < 4; ++q) m[q] = __hmax2(m[q], pv[q]);
            }
          }
        }
        uint4 o;
        memcpy(&o, m, sizeof(o));
        *reinterpret_cast<uint4*>(out_img + (static_cast<size_t>(7 * ty + py) * 56 + 7 * tx + px) * 64 + cg * 8) = o;
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// Folded stem weights in the shared-memory order the kernel expects:
//   index = (((r*2 + h)*8 + n/8)*2 + kchunk)*64 + (n%8)*8 + j ,  pixel s = 4h + (kchunk*8+j)/4, channel c = j%4
__global__ void stem_fold_kernel(const float* __restrict__ w, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, const float* __restrict__ mean,
                                 const float* __restrict__ var, float eps, __nv_bfloat16* __restrict__ w_out,
                                 float* __restrict__ bias_out) {
  const int total = kStemWeightBytes / 2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int j = i & 7;
    const int nl = (i >> 3) & 7;
    const int kchunk = (i >> 6) & 1;
    const int ng = (i >> 7) & 7;
    const int rh = i >> 10;
    const int r = rh >> 1, h = rh & 1;
    const int n = ng * 8 + nl;
    const int kk = kchunk * 8 + j;
    const int s = 4 * h + (kk >> 2), c = kk & 3;
    const float scale = gamma[n] / sqrtf(var[n] + eps);
    float v = 0.f;
    if (c < 3 && s < 7) v = w[((n * 3 + c) * 7 + r) * 7 + s] * scale;
    w_out[i] = __float2bfloat16_rn(v);
    if (i < 64) bias_out[i] = beta[i] - mean[i] * (gamma[i] / sqrtf(var[i] + eps));
  }
}

}  // namespace irp
