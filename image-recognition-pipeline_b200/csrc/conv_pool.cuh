// The last bottleneck's conv3 (1x1, 512 -> 2048 @7x7) + folded BN + residual + ReLU + GLOBAL AVERAGE POOL in one
// kernel (torchvision/models/resnet.py:150-159 of layer4's last block, then :278-279 avgpool; called from
// functions/data_curation.py:677).  The 2048-channel activation is never written: the kernel's output is the fp32
// [B, 2048] embedding.
//
// The GEMM is computed TRANSPOSED (D^T = W . T2^T): M = 128 output channels (A operand = weight rows, K-major as
// stored), N = the 2 x 49 = 98 pixels of two whole images (B operand = activation rows, K-major as stored; N is
// rounded up to 112, the 14 surplus columns multiply stale shared-memory rows and are never read).  A TMEM lane is
// then a CHANNEL and its columns are that channel's pixels, so the average over an image's 49 pixels is a chain of
// additions inside one thread, in a fixed order -- no atomics, no cross-thread reduction, and an image's embedding
// does not depend on its batch or slot (the property the separate avgpool kernel was kept for in round 1).
//
// One CTA owns one 128-channel tile: its 128 x 512 weight slice (128 KB) is loaded ONCE and stays in shared memory;
// the CTA then walks image pairs, streaming their activation rows (eight 64-channel K blocks through a 3-slot
// ring) and the residual slice (98 x 128 bf16, no swizzle so that a warp's 32 channels are 64 contiguous bytes;
// double buffered: with a single buffer the next pair's residual load, and the activation loads queued behind it,
// waited for the current epilogue -- 9.9 k cycles per tile against 2.6 k of MMA).  Accumulators are double
// buffered in TMEM, so the epilogue of a pair overlaps the MMAs of the next.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..5 = epilogue.
#pragma once
#include "conv_params.cuh"

namespace irp {

constexpr int kCpThreads = 192;
constexpr int kCpImgs = 2;                    // images per tile
constexpr int kCpPix = 49;                    // pixels per image (7 x 7)
constexpr int kCpRows = kCpImgs * kCpPix;     // 98 activation rows per tile
constexpr int kCpN = 112;                     // UMMA N (multiple of 16 >= 98)
constexpr int kCpK = 512;                     // input channels
constexpr int kCpKBlocks = kCpK / 64;         // 8
constexpr int kCpSlots = 3;
constexpr int kCpABytes = 128 * 128;          // one K block of the weight slice (128 rows x 64 k)
constexpr int kCpBBytes = 15 * 1024;          // one K block of a pair's activations (112 row slots x 64 k, 1 KB aligned)
constexpr int kCpResBytes = 2 * kCpRows * 128;  // residual: two 64-channel planes of 98 rows x 128 B, no swizzle
constexpr int kCpResStride = ((kCpResBytes + 1023) / 1024) * 1024;
constexpr int kCpSmemBytes = kCpKBlocks * kCpABytes + kCpSlots * kCpBBytes + 2 * kCpResStride + 256 + 1024;

struct alignas(64) ConvPoolParams {
  CUtensorMap tmW;    // weights  [2048][512]  dims (512, Cout), box (64, 128), 128B swizzle
  CUtensorMap tmX;    // T2       [M][512]     dims (512, M),    box (64, 98), 128B swizzle
  CUtensorMap tmRes;  // residual [M][Cout]    dims (Cout, M),   box (64, 98), no swizzle
  const float* bias;  // [Cout]
  float* out;         // [batch][Cout] fp32 pooled embeddings
  int batch;
  int cout;
  int n_ctiles;       // Cout / 128
  int n_groups;       // ceil(batch / 2)
};

__global__ void __launch_bounds__(kCpThreads, 1) conv_pool_kernel(const __grid_constant__ ConvPoolParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* smem_w = smem;                                  // 8 x 16 KB, resident
  uint8_t* smem_x = smem_w + kCpKBlocks * kCpABytes;       // kCpSlots x 15 KB
  uint8_t* smem_res = smem_x + kCpSlots * kCpBBytes;       // 2 x 25 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_res + 2 * kCpResStride);
  uint64_t* full_bar = bars;                    // [kCpSlots]
  uint64_t* empty_bar = bars + kCpSlots;        // [kCpSlots]
  uint64_t* tfull_bar = bars + 2 * kCpSlots;    // [2]
  uint64_t* tempty_bar = tfull_bar + 2;         // [2]
  uint64_t* wfull_bar = tempty_bar + 2;         // [1] weight slice resident
  uint64_t* res_full = wfull_bar + 1;           // [2]
  uint64_t* res_empty = res_full + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_empty + 2);
  constexpr uint32_t kTmemCols = 256;           // two accumulators of 112 columns, 128 apart

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // work split: CTA -> (channel tile, first image group, group stride)
  const int ct = blockIdx.x % p.n_ctiles;
  const int g_first = blockIdx.x / p.n_ctiles;
  const int g_stride = gridDim.x / p.n_ctiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmW);
    tma_prefetch_desc(&p.tmX);
    tma_prefetch_desc(&p.tmRes);
    for (int i = 0; i < kCpSlots; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 128);
    }
    mbar_init(wfull_bar, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&res_full[i], 1);
      mbar_init(&res_empty[i], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();

  if (warp == 0) {
    if (lane == 0) {
      // weights are not written by the preceding kernels: fetch them before the dependency wait
      mbar_arrive_expect_tx(wfull_bar, kCpKBlocks * kCpABytes);
      for (int kb = 0; kb < kCpKBlocks; ++kb)
        tma_load_2d(smem_w + kb * kCpABytes, &p.tmW, wfull_bar, kb * 64, ct * 128);
      pdl_wait();
      int slot = 0;
      uint32_t phase = 0;
      int j = 0;
      for (int g = g_first; g < p.n_groups; g += g_stride, ++j) {
        const int row0 = g * kCpRows;
        // residual slice of this pair (two buffers: the epilogue of pair j - 2 must have read buffer j & 1)
        const int rb = j & 1;
        if (j >= 2) mbar_wait(&res_empty[rb], ((j >> 1) - 1) & 1);
        mbar_arrive_expect_tx(&res_full[rb], kCpResBytes);
        tma_load_2d(smem_res + rb * kCpResStride, &p.tmRes, &res_full[rb], ct * 128, row0);
        tma_load_2d(smem_res + rb * kCpResStride + kCpRows * 128, &p.tmRes, &res_full[rb], ct * 128 + 64, row0);
        for (int kb = 0; kb < kCpKBlocks; ++kb) {
          mbar_wait(&empty_bar[slot], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[slot], kCpRows * 128);
          tma_load_2d(smem_x + slot * kCpBBytes, &p.tmX, &full_bar[slot], kb * 64, row0);
          if (++slot == kCpSlots) {
            slot = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    {  // whole warp, converged; the tcgen05 instructions are issued by the elected lane (elect_one(), ptx.cuh)
      constexpr uint32_t idesc = umma_idesc_bf16(128, kCpN);
      mbar_wait(wfull_bar, 0);
      tc_fence_after();
      int slot = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int g = g_first; g < p.n_groups; g += g_stride) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * 128;  // accumulators at columns 0 and 128
        for (int kb = 0; kb < kCpKBlocks; ++kb) {
          mbar_wait(&full_bar[slot], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_w + kb * kCpABytes);
          const uint32_t b_addr = smem_u32(smem_x + slot * kCpBBytes);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_d, umma_smem_desc<128>(a_addr + k * 32), umma_smem_desc<128>(b_addr + k * 32), idesc,
                        (kb | k) != 0 ? 1u : 0u);
            umma_commit(&empty_bar[slot]);
            if (kb == kCpKBlocks - 1) umma_commit(&tfull_bar[acc]);
          }
          __syncwarp();
          if (++slot == kCpSlots) {
            slot = 0;
            phase ^= 1;
          }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ============================ epilogue: thread = output channel ============================
    const int quarter = warp & 3;
    const int ch = quarter * 32 + lane;  // channel inside the tile = TMEM lane
    const float bias = __ldg(p.bias + ct * 128 + ch);
    // residual element (pixel row r, channel ch): plane ch / 64, row r, 2 * (ch % 64) bytes into the 128-byte row
    const uint8_t* res_col0 = smem_res + (ch >> 6) * (kCpRows * 128) + (ch & 63) * 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    int j = 0;
    for (int g = g_first; g < p.n_groups; g += g_stride, ++j) {
      const int rb = j & 1;
      const uint8_t* res_col = res_col0 + rb * kCpResStride;
      mbar_wait(&tfull_bar[acc], acc_phase);
      mbar_wait(&res_full[rb], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * 128 + (static_cast<uint32_t>(quarter * 32) << 16);
      float sum[kCpImgs] = {0.f, 0.f};
#pragma unroll
      for (int c0 = 0; c0 < 128; c0 += 32) {  // 112 accumulator columns; the last load's surplus columns are ignored
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int pix = c0 + i;  // compile-time after unrolling: the image of a column is a constant
          if (pix < kCpRows) {
            const float r = __uint_as_float(static_cast<uint32_t>(*reinterpret_cast<const uint16_t*>(res_col + pix * 128))
                                            << 16);
            sum[pix / kCpPix] += fmaxf(__uint_as_float(v[i]) + bias + r, 0.f);
          }
        }
      }
      // accumulator and residual slice consumed
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
      mbar_arrive(&res_empty[rb]);
#pragma unroll
      for (int i = 0; i < kCpImgs; ++i) {
        const int img = g * kCpImgs + i;
        if (img < p.batch) p.out[static_cast<size_t>(img) * p.cout + ct * 128 + ch] = sum[i] * (1.0f / kCpPix);
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

}  // namespace irp
