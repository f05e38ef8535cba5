// 3x3 / stride 1 / pad 1 convolution with 64 input and 64 output channels (layer1's conv2 of every bottleneck,
// torchvision/models/resnet.py:146-148) + folded BN + ReLU, with the input patch and ALL weights resident in
// shared memory.
//
// The generic implicit-GEMM kernels fetch one 128-pixel A box per filter tap, i.e. every input pixel crosses
// L2 -> SM nine times; at N = 64 that operand stream (not the tensor pipe, not HBM) is the limiter.  Here one CTA
// tile is 8 (w) x 16 (h) output pixels of one image:
//   * the 10 x 18 pixel input patch (halo included; borders zero-filled by TMA = the conv's padding) is fetched
//     ONCE, 128 bytes (64 channels) per pixel, 128B-swizzled;
//   * the A operand of tap (r, s) is a shifted WINDOW of that patch: an 8-row core group is 8 consecutive
//     pixels of one patch row (1 KB contiguous), the next group is the next patch row, so the descriptor is
//     start = patch + ((r*10 + s)*128) bytes, SBO = 10*128 bytes.  The 128B swizzle is a function of the shared
//     memory address bits, so windows that start on any 128-byte row read back exactly what TMA wrote;
//   * the 9 x (64 x 64) weight tiles (72 KB) are loaded once per CTA and stay resident.
// Per tile the SM receives 23 KB instead of 9 x (16 + 8) KB.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..9 = epilogue.
#pragma once
#include "conv_params.cuh"

namespace irp {

constexpr int kC64Threads = 320;
constexpr int kC64EpiThreads = 256;
constexpr int kC64TileW = 8, kC64TileH = 16;
constexpr int kC64PatchW = kC64TileW + 2, kC64PatchH = kC64TileH + 2;  // 10 x 18
constexpr int kC64PatchBytes = kC64PatchW * kC64PatchH * 128;          // 23040
constexpr int kC64PatchStride = 23 * 1024;                             // 1024-aligned slot
constexpr int kC64Slots = 4;
constexpr int kC64WeightBytes = 9 * 64 * 128;  // 73728
constexpr int kC64Ring = 3;                    // staging buffers (16 KB each)
constexpr int kC64SmemBytes =
    kC64WeightBytes + kC64Slots * kC64PatchStride + kC64Ring * kStgChunkBytes + 1024 /*barriers + bias*/ + 1024;

// K-major SWIZZLE_128B descriptor with an explicit stride between 8-row groups
__device__ __forceinline__ uint64_t umma_smem_desc_sw128(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= 1ull << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}

// ConvParams use: tmA[0] = input (64, W, H, B) box (64, 10, 18, 1); tmB = weights (576, 64) box (64, 64);
// tmOut = output (64, W, H, B) box (64, 8, 16, 1); tiles_w/tiles_h/tiles_n; bias; relu.
__global__ void __launch_bounds__(kC64Threads, 1) conv3x3_c64_kernel(const __grid_constant__ ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* smem_w = smem;                                      // 9 x 8 KB
  uint8_t* smem_patch = smem_w + kC64WeightBytes;              // kC64Slots x 23 KB
  uint8_t* smem_stg = smem_patch + kC64Slots * kC64PatchStride;  // kC64Ring x 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stg + kC64Ring * kStgChunkBytes);
  uint64_t* full_bar = bars;                     // [kC64Slots]
  uint64_t* empty_bar = bars + kC64Slots;        // [kC64Slots]
  uint64_t* tfull_bar = bars + 2 * kC64Slots;    // [2]
  uint64_t* tempty_bar = tfull_bar + 2;          // [2]
  uint64_t* wfull_bar = tempty_bar + 2;          // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull_bar + 1);
  float* sbias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);  // [64]
  constexpr uint32_t kTmemCols = 128;  // 2 x 64 fp32 columns

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.tiles_w * p.tiles_h * p.tiles_n;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA[0]);
    tma_prefetch_desc(&p.tmB);
    tma_prefetch_desc(&p.tmOut);
    for (int i = 0; i < kC64Slots; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 8);  // one arrival per epilogue warp
    }
    mbar_init(wfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 128) sbias[threadIdx.x - 64] = __ldg(p.bias + threadIdx.x - 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      mbar_arrive_expect_tx(wfull_bar, kC64WeightBytes);
#pragma unroll 1
      for (int t = 0; t < 9; ++t) tma_load_2d(smem_w + t * 8192, &p.tmB, wfull_bar, t * 64, 0);
      pdl_wait();  // weights are constants; the activations below come from the previous kernel
      int slot = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int tw = tile % p.tiles_w;
        const int th = (tile / p.tiles_w) % p.tiles_h;
        const int n = tile / (p.tiles_w * p.tiles_h);
        mbar_wait(&empty_bar[slot], phase ^ 1);
        mbar_arrive_expect_tx(&full_bar[slot], kC64PatchBytes);
        tma_load_4d(smem_patch + slot * kC64PatchStride, &p.tmA[0], &full_bar[slot], 0, tw * kC64TileW - 1,
                    th * kC64TileH - 1, n);
        if (++slot == kC64Slots) {
          slot = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    {  // whole warp, converged; the tcgen05 instructions are issued by the elected lane (elect_one(), ptx.cuh)
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64);
      const uint32_t w_addr = smem_u32(smem_w);
      int slot = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      mbar_wait(wfull_bar, 0);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        mbar_wait(&full_bar[slot], phase);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * 64;
        const uint32_t patch = smem_u32(smem_patch + slot * kC64PatchStride);
        if (elect_one()) {
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const uint32_t a0 = patch + ((t / 3) * kC64PatchW + (t % 3)) * 128;
            const uint32_t b0 = w_addr + t * 8192;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t da = umma_smem_desc_sw128(a0 + k * 32, kC64PatchW * 128);
              const uint64_t db = umma_smem_desc_sw128(b0 + k * 32, 1024);
              umma_bf16(tmem_d, da, db, idesc, (t | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[slot]);
          umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++slot == kC64Slots) {
          slot = 0;
          phase ^= 1;
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ============================ epilogue (warps 2..9) ============================
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;  // GEMM row = ho_local * 8 + wo_local = the TMA-store box order
    const bool leader = (threadIdx.x == 64);
    const uint32_t row_off = static_cast<uint32_t>(row) * 128u;
    const uint32_t swz = static_cast<uint32_t>(row & 7);
    const int piece0 = half * 4;
    int acc = 0;
    uint32_t acc_phase = 0;
    int j = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++j) {
      const int tw = tile % p.tiles_w;
      const int th = (tile / p.tiles_w) % p.tiles_h;
      const int n = tile / (p.tiles_w * p.tiles_h);
      uint8_t* chunk = smem_stg + (j % kC64Ring) * kStgChunkBytes + row_off;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem_base + acc * 64 + (static_cast<uint32_t>(quarter * 32) << 16) + half * 32, v);
      const float4* bp = reinterpret_cast<const float4*>(sbias + half * 32);
      float4 bv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) bv[i] = bp[i];
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 b0 = bv[2 * i], b1 = bv[2 * i + 1];
        float x[8];
        x[0] = __uint_as_float(v[8 * i + 0]) + b0.x;
        x[1] = __uint_as_float(v[8 * i + 1]) + b0.y;
        x[2] = __uint_as_float(v[8 * i + 2]) + b0.z;
        x[3] = __uint_as_float(v[8 * i + 3]) + b0.w;
        x[4] = __uint_as_float(v[8 * i + 4]) + b1.x;
        x[5] = __uint_as_float(v[8 * i + 5]) + b1.y;
        x[6] = __uint_as_float(v[8 * i + 6]) + b1.z;
        x[7] = __uint_as_float(v[8 * i + 7]) + b1.w;
        if (p.relu) {
#pragma unroll
          for (int q = 0; q < 8; ++q) x[q] = fmaxf(x[q], 0.f);
        }
        *reinterpret_cast<uint4*>(chunk + (((piece0 + i) ^ swz) << 4)) =
            make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]),
                       pack_bf16x2(x[6], x[7]));
      }
      fence_proxy_async();
      // stores of tiles <= j-kC64Ring+1 have finished reading: tile j+1's buffer is free for everyone
      if (leader) tma_store_wait_read<kC64Ring - 2>();
      named_bar_sync(1, kC64EpiThreads);
      if (leader) {
        tma_store_4d(&p.tmOut, smem_stg + (j % kC64Ring) * kStgChunkBytes, 0, tw * kC64TileW, th * kC64TileH, n);
        tma_store_commit();
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (leader) tma_store_wait_all<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace irp
