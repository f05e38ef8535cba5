// Implicit-GEMM convolution for the ResNet-50 trunk on tcgen05 / TMEM, operands fed by TMA.
//
//   out[n,ho,wo,co] = act( sum_{r,s,ci} x[n, ho*stride+r-pad, wo*stride+s-pad, ci] * w[co,r,s,ci]
//                          + bias[co] (+ residual[n,ho,wo,co]) )
//
// GEMM view (torchvision/models/resnet.py:108-160 Bottleneck, BN folded into w/bias):
//   M = output pixels, N = Cout, K = taps * Cin.  Activations are NHWC bf16, weights [Cout][tap][Cin] bf16.
//   One CTA tile = 128 output pixels (a (bw,bh,bn) box in (wo,ho,n)) x BN output channels.  For each filter
//   tap the A operand is ONE shifted TMA box of the input (out-of-bounds rows/cols are zero-filled by TMA,
//   which is exactly the conv's zero padding); stride-2 convs read one of four "parity" views of the input,
//   each a plain strided tensor map.  Accumulators live in TMEM (double buffered) so the epilogue of tile i
//   overlaps the main loop of tile i+1.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..5 = epilogue.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdint>

#include "ptx.cuh"

namespace irp {

constexpr int kConvThreads = 192;
constexpr int kMaxTaps = 9;
constexpr int kTileM = 128;

struct alignas(64) ConvParams {
  CUtensorMap tmA[4];  // activation views (index = parity for stride 2; [0] only for stride 1 / stem)
  CUtensorMap tmB;     // weights [Cout][K] bf16, K-major
  // M tiling: box (bw,bh,bn) in output coordinates
  int bw, bh, bn;
  int tiles_w, tiles_h, tiles_n;
  int n_tiles_n;  // Cout / BN
  int num_tiles;
  int a_box_bytes;  // bytes one A TMA box delivers (rows_in_box * BK * 2)
  // problem
  int B, Ho, Wo, Cout;
  int ntaps, kc_blocks;  // K loop = ntaps * kc_blocks blocks of BK
  int cin;               // channels per tap in the weight matrix (K offset of tap t = t*cin)
  int8_t tap_map[kMaxTaps], tap_dw[kMaxTaps], tap_dh[kMaxTaps];
  // epilogue
  const float* bias;             // [Cout]
  const __nv_bfloat16* residual;  // NHWC [B,Ho,Wo,Cout] or nullptr
  __nv_bfloat16* out;             // NHWC [B,Ho,Wo,Cout]
  int relu;
};

template <int BN, int BK>
struct ConvSmem {
  static constexpr int kABytes = kTileM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBudget = 200 * 1024;
  static constexpr int kStagesRaw = kBudget / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kBarrierBytes = 256;
  static constexpr int kTotalBytes = kStages * kStageBytes + kBarrierBytes + 1024;  // +1024: manual alignment
};

template <int BN, int BK, bool STEM>
__global__ void __launch_bounds__(kConvThreads, 1) conv_gemm_kernel(const __grid_constant__ ConvParams p) {
  using S = ConvSmem<BN, BK>;
  constexpr int kStages = S::kStages;
  constexpr int kSwz = BK * 2;            // swizzle span in bytes == one K block row
  constexpr uint32_t kTmemCols = 2 * BN;  // double-buffered fp32 accumulator (power of two >= 32)
  static_assert(kTmemCols >= 32 && kTmemCols <= 512 && (kTmemCols & (kTmemCols - 1)) == 0, "TMEM columns");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * S::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * S::kStageBytes);
  uint64_t* full_bar = bars;                 // [kStages] TMA -> MMA
  uint64_t* empty_bar = bars + kStages;      // [kStages] MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * kStages;  // [2] MMA -> epilogue
  uint64_t* tempty_bar = tfull_bar + 2;      // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA[0]);
    tma_prefetch_desc(&p.tmB);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 128);
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

  const int k_blocks = p.ntaps * p.kc_blocks;

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.n_tiles_n;
        int m_tile = tile / p.n_tiles_n;
        const int tw = m_tile % p.tiles_w;
        m_tile /= p.tiles_w;
        const int th = m_tile % p.tiles_h;
        const int tn = m_tile / p.tiles_h;
        const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;
        for (int t = 0; t < p.ntaps; ++t) {
          for (int kc = 0; kc < p.kc_blocks; ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full_bar[stage], p.a_box_bytes + S::kBBytes);
            if constexpr (STEM) {
              // 5-D view (32 = 8 px * 4 ch, wo, row parity, row pair, n): filter row t of the 7x7/2 stem
              tma_load_5d(smem_a + stage * S::kABytes, &p.tmA[0], &full_bar[stage], 0, w0, t & 1, h0 + (t >> 1),
                          n0);
            } else {
              tma_load_4d(smem_a + stage * S::kABytes, &p.tmA[p.tap_map[t]], &full_bar[stage], kc * BK,
                          w0 + p.tap_dw[t], h0 + p.tap_dh[t], n0);
            }
            tma_load_2d(smem_b + stage * S::kBBytes, &p.tmB, &full_bar[stage], t * p.cin + kc * BK, n_tile * BN);
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kTileM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + stage * S::kABytes);
          const uint32_t b_addr = smem_u32(smem_b + stage * S::kBBytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = umma_smem_desc<kSwz>(a_addr + k * 32);
            const uint64_t db = umma_smem_desc<kSwz>(b_addr + k * 32);
            umma_bf16(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ============================ epilogue (warps 2..5) ============================
    const int quarter = warp & 3;  // TMEM lane quarter this warp may read
    const int row = quarter * 32 + lane;
    const int w_l = row % p.bw;
    const int h_l = (row / p.bw) % p.bh;
    const int n_l = row / (p.bw * p.bh);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int n_tile = tile % p.n_tiles_n;
      int m_tile = tile / p.n_tiles_n;
      const int tw = m_tile % p.tiles_w;
      m_tile /= p.tiles_w;
      const int th = m_tile % p.tiles_h;
      const int tn = m_tile / p.tiles_h;
      const int w = tw * p.bw + w_l, h = th * p.bh + h_l, n = tn * p.bn + n_l;
      const bool valid = (n_l < p.bn) && (n < p.B) && (h < p.Ho) && (w < p.Wo);
      const size_t pix = (static_cast<size_t>(n) * p.Ho + h) * p.Wo + w;
      const size_t off = pix * p.Cout + static_cast<size_t>(n_tile) * BN;

      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * BN + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t v[32];
        __syncwarp();
        tmem_ld_32x32b_x32(taddr + c, v);
        uint4 rv[4];
        const bool has_res = (p.residual != nullptr) && valid;
        if (has_res) {
          const uint4* rp = reinterpret_cast<const uint4*>(p.residual + off + c);
#pragma unroll
          for (int i = 0; i < 4; ++i) rv[i] = __ldg(rp + i);
        }
        const float4* bp = reinterpret_cast<const float4*>(p.bias + n_tile * BN + c);
        __syncwarp();
        tmem_ld_wait();
        uint32_t o[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 b = __ldg(bp + i);
          float x0 = __uint_as_float(v[4 * i + 0]) + b.x;
          float x1 = __uint_as_float(v[4 * i + 1]) + b.y;
          float x2 = __uint_as_float(v[4 * i + 2]) + b.z;
          float x3 = __uint_as_float(v[4 * i + 3]) + b.w;
          if (has_res) {
            const uint32_t* r32 = reinterpret_cast<const uint32_t*>(rv);
            x0 += bf16_lo(r32[2 * i]);
            x1 += bf16_hi(r32[2 * i]);
            x2 += bf16_lo(r32[2 * i + 1]);
            x3 += bf16_hi(r32[2 * i + 1]);
          }
          if (p.relu) {
            x0 = fmaxf(x0, 0.f);
            x1 = fmaxf(x1, 0.f);
            x2 = fmaxf(x2, 0.f);
            x3 = fmaxf(x3, 0.f);
          }
          o[2 * i] = pack_bf16x2(x0, x1);
          o[2 * i + 1] = pack_bf16x2(x2, x3);
        }
        if (valid) {
          uint4* op = reinterpret_cast<uint4*>(p.out + off + c);
#pragma unroll
          for (int i = 0; i < 4; ++i) op[i] = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
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
