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
// Epilogue: TMEM -> registers -> (+bias, +residual, ReLU, bf16) -> swizzled shared-memory staging tile -> TMA
// store with the same box geometry as the A loads (TMA clips rows outside the tensor, so ragged tiles need no
// predication).  The residual tile is prefetched by TMA into the staging buffer it will be overwritten in.
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
constexpr int kStgChunkBytes = kTileM * 128;  // one 64-channel column chunk of the output tile (128 B rows)

struct alignas(64) ConvParams {
  CUtensorMap tmA[4];  // activation views (index = parity for stride 2; [0] only for stride 1 / stem)
  CUtensorMap tmB;     // weights [Cout][K] bf16, K-major
  CUtensorMap tmOut;   // output  (Cout, Wo, Ho, B) / flat (Cout, M, 1, 1), box (64, bw, bh, bn)
  CUtensorMap tmRes;   // residual, same geometry as tmOut (valid only for the RES instances)
  // M tiling: box (bw,bh,bn) in output coordinates
  int bw, bh, bn;
  int tiles_w, tiles_h, tiles_n;
  int n_tiles_n;  // Cout / BN
  int num_tiles;
  int a_box_bytes;    // bytes one A TMA box delivers (rows_in_box * BK * 2)
  int out_box_bytes;  // bytes one 64-channel output/residual box moves (rows_in_box * 128)
  // problem
  int ntaps, kc_blocks;  // K loop = ntaps * kc_blocks blocks of BK
  int cin;               // channels per tap in the weight matrix (K offset of tap t = t*cin)
  int n_base;            // stem only: first image of this micro-batch inside the input tensor map
  int8_t tap_map[kMaxTaps], tap_dw[kMaxTaps], tap_dh[kMaxTaps];
  // epilogue
  const float* bias;  // [Cout]
  int relu;
};

template <int BN, int BK, int NB>
struct ConvSmem {
  static constexpr int kABytes = kTileM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStgBytes = (BN / 64) * kStgChunkBytes;  // one staging buffer
  static constexpr int kBudget = 220 * 1024 - NB * kStgBytes;
  static constexpr int kStagesRaw = kBudget / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kBarrierBytes = 1024;  // mbarriers + TMEM slot (256 B), then the tile's bias slice (fp32)
  static constexpr int kTotalBytes = kStages * kStageBytes + NB * kStgBytes + kBarrierBytes + 1024;
  static_assert(kStages >= 2, "not enough shared memory for a pipeline");
};

// NB = staging buffers: 3 with a residual (prefetch one tile ahead), else 1 (long K loops) or 2 (short ones).
// CM x CN = thread-block cluster shape: the CN CTAs of a cluster row share one A (activation) tile and the CM
// CTAs of a column share one B (weight) tile; each k-block's tile is fetched from L2 ONCE by one CTA of the group
// and TMA-multicast to the others, which divides the L2->SM operand traffic (the limiter at 128x128 tiles).
template <int BN, int BK, bool STEM, bool RES, int NB, int CM = 1, int CN = 1>
__global__ void __launch_bounds__(kConvThreads, 1) conv_gemm_kernel(const __grid_constant__ ConvParams p) {
  using S = ConvSmem<BN, BK, NB>;
  constexpr int kStages = S::kStages;
  constexpr int kSwz = BK * 2;            // swizzle span in bytes == one K block row
  constexpr uint32_t kTmemCols = 2 * BN;  // double-buffered fp32 accumulator (power of two >= 32)
  constexpr int kChunks = BN / 64;        // 64-channel column chunks per tile
  static_assert(kTmemCols >= 32 && kTmemCols <= 512 && (kTmemCols & (kTmemCols - 1)) == 0, "TMEM columns");
  static_assert(!RES || NB == 3, "residual prefetch needs three staging buffers");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * S::kABytes;
  uint8_t* smem_stg = smem + kStages * S::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stg + NB * S::kStgBytes);
  uint64_t* full_bar = bars;                 // [kStages] TMA -> MMA
  uint64_t* empty_bar = bars + kStages;      // [kStages] MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * kStages;  // [2] MMA -> epilogue
  uint64_t* tempty_bar = tfull_bar + 2;      // [2] epilogue -> MMA
  uint64_t* res_full = tempty_bar + 2;       // [NB] residual tile landed in staging buffer
  uint64_t* stg_empty = res_full + NB;       // [NB] staging buffer may be refilled by the producer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stg_empty + NB);
  float* sbias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);  // [BN]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  constexpr int kCluster = CM * CN;
  const int crank = kCluster > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  const int rm = crank % CM, rn = crank / CM;  // position inside the cluster
  const int cluster_id = blockIdx.x / kCluster;
  const int num_clusters = gridDim.x / kCluster;
  // tile groups: CM consecutive M tiles x CN consecutive N tiles
  const int groups_n = p.n_tiles_n / CN;
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int num_groups = ((m_tiles + CM - 1) / CM) * groups_n;
  // CTAs that may write into my smem slots / whose slots I may write: my cluster row and column
  uint16_t row_mask = 0, col_mask = 0;
#pragma unroll
  for (int i = 0; i < CN; ++i) row_mask |= static_cast<uint16_t>(1u << (rm + CM * i));  // same rm: share A
#pragma unroll
  for (int i = 0; i < CM; ++i) col_mask |= static_cast<uint16_t>(1u << (i + CM * rn));  // same rn: share B
  const uint16_t peer_mask = row_mask | col_mask;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA[0]);
    tma_prefetch_desc(&p.tmB);
    tma_prefetch_desc(&p.tmOut);
    if (RES) tma_prefetch_desc(&p.tmRes);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], CM + CN - 1);  // one MMA commit from every CTA sharing a tile with me
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 128);
    }
    for (int i = 0; i < NB; ++i) {
      mbar_init(&res_full[i], 1);
      mbar_init(&stg_empty[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();  // barrier inits visible before any peer multicasts / arrives
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int k_blocks = p.ntaps * p.kc_blocks;

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int j = 0;  // local tile counter
      int kb_global = 0;  // k-block counter (decides which CTA of a group fetches the shared tile)
      for (int g = cluster_id; g < num_groups; g += num_clusters, ++j) {
        const int n_tile = (g % groups_n) * CN + rn;
        int m_tile = (g / groups_n) * CM + rm;  // may be a phantom tile past the end: loads zero-fill, stores clip
        const int tw = m_tile % p.tiles_w;
        m_tile /= p.tiles_w;
        const int th = m_tile % p.tiles_h;
        const int tn = m_tile / p.tiles_h;
        const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;
        if (RES) {
          // residual tile -> the staging buffer the epilogue will overwrite in place
          const int b = j % NB;
          mbar_wait(&stg_empty[b], (j / NB) & 1);
          mbar_arrive_expect_tx(&res_full[b], kChunks * p.out_box_bytes);
#pragma unroll
          for (int cc = 0; cc < kChunks; ++cc)
            tma_load_4d(smem_stg + b * S::kStgBytes + cc * kStgChunkBytes, &p.tmRes, &res_full[b],
                        n_tile * BN + cc * 64, w0, h0, n0);
        }
        for (int t = 0; t < p.ntaps; ++t) {
          for (int kc = 0; kc < p.kc_blocks; ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full_bar[stage], p.a_box_bytes + S::kBBytes);
            if constexpr (STEM) {
              // 5-D view (32 = 8 px * 4 ch, wo, row parity, row pair, n): filter row t of the 7x7/2 stem
              tma_load_5d(smem_a + stage * S::kABytes, &p.tmA[0], &full_bar[stage], 0, w0, t & 1, h0 + (t >> 1),
                          n0 + p.n_base);
            } else if constexpr (CN > 1) {
              if (kb_global % CN == rn)
                tma_load_4d_mc(smem_a + stage * S::kABytes, &p.tmA[p.tap_map[t]], &full_bar[stage], kc * BK,
                               w0 + p.tap_dw[t], h0 + p.tap_dh[t], n0, row_mask);
            } else {
              tma_load_4d(smem_a + stage * S::kABytes, &p.tmA[p.tap_map[t]], &full_bar[stage], kc * BK,
                          w0 + p.tap_dw[t], h0 + p.tap_dh[t], n0);
            }
            if constexpr (CM > 1) {
              if (kb_global % CM == rm)
                tma_load_2d_mc(smem_b + stage * S::kBBytes, &p.tmB, &full_bar[stage], t * p.cin + kc * BK,
                               n_tile * BN, col_mask);
            } else {
              tma_load_2d(smem_b + stage * S::kBBytes, &p.tmB, &full_bar[stage], t * p.cin + kc * BK, n_tile * BN);
            }
            ++kb_global;
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
      for (int g = cluster_id; g < num_groups; g += num_clusters) {
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
          // frees the smem slot (here and in every CTA that multicasts into it) once these MMAs have read it
          if constexpr (kCluster > 1) umma_commit_mc(&empty_bar[stage], peer_mask);
          else umma_commit(&empty_bar[stage]);
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
    const bool leader = (threadIdx.x == 64);  // issues the TMA stores, owns their bulk groups
    const uint32_t row_off = static_cast<uint32_t>(row) * 128u;
    const uint32_t swz = static_cast<uint32_t>(row & 7);
    int acc = 0;
    uint32_t acc_phase = 0;
    int j = 0;
    if (RES && leader) mbar_arrive(&stg_empty[0]);  // first use of buffer 0 needs no predecessor
    // The tile's bias slice is staged in shared memory (no L1 is left beside ~200 KB of shared memory, so a
    // global load inside the column loop is an exposed L2 round trip); it is fetched one tile ahead.
    const int et = threadIdx.x - 64;  // 0..127
    float bias_next = 0.f;
    if (et < BN && cluster_id < num_groups) bias_next = __ldg(p.bias + ((cluster_id % groups_n) * CN + rn) * BN + et);
    for (int g = cluster_id; g < num_groups; g += num_clusters, ++j) {
      const int n_tile = (g % groups_n) * CN + rn;
      // every thread is past the previous tile's named barrier 2, i.e. done reading the old slice
      if (et < BN) sbias[et] = bias_next;
      named_bar_sync(3, 128);
      if (et < BN && g + num_clusters < num_groups)
        bias_next = __ldg(p.bias + (((g + num_clusters) % groups_n) * CN + rn) * BN + et);
      int m_tile = (g / groups_n) * CM + rm;
      const int tw = m_tile % p.tiles_w;
      m_tile /= p.tiles_w;
      const int th = m_tile % p.tiles_h;
      const int tn = m_tile / p.tiles_h;
      const int b = j % NB;
      uint8_t* stg = smem_stg + b * S::kStgBytes;

      if (RES) {
        // release the buffer of the NEXT tile (its previous store has finished reading), then wait for our residual
        if (leader) {
          tma_store_wait_read<(NB >= 2 ? NB - 2 : 0)>();
          mbar_arrive(&stg_empty[(j + 1) % NB]);
        }
        mbar_wait(&res_full[b], (j / NB) & 1);
      } else {
        if (leader) tma_store_wait_read<NB - 1>();
        named_bar_sync(1, 128);  // buffer b is free for everyone
      }

      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * BN + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll
      for (int c = 0; c < BN; c += 32) {
        uint32_t v[32];
        __syncwarp();
        tmem_ld_32x32b_x32(taddr + c, v);
        uint8_t* chunk = stg + (c >> 6) * kStgChunkBytes + row_off;
        const int piece0 = (c & 32) >> 3;  // first 16-byte piece of this half chunk: 0 or 4
        uint4 rv[4];
        if (RES) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            rv[i] = *reinterpret_cast<const uint4*>(chunk + (((piece0 + i) ^ swz) << 4));
        }
        const float4* bp = reinterpret_cast<const float4*>(sbias + c);
        __syncwarp();
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 b0 = bp[2 * i], b1 = bp[2 * i + 1];
          float x[8];
          x[0] = __uint_as_float(v[8 * i + 0]) + b0.x;
          x[1] = __uint_as_float(v[8 * i + 1]) + b0.y;
          x[2] = __uint_as_float(v[8 * i + 2]) + b0.z;
          x[3] = __uint_as_float(v[8 * i + 3]) + b0.w;
          x[4] = __uint_as_float(v[8 * i + 4]) + b1.x;
          x[5] = __uint_as_float(v[8 * i + 5]) + b1.y;
          x[6] = __uint_as_float(v[8 * i + 6]) + b1.z;
          x[7] = __uint_as_float(v[8 * i + 7]) + b1.w;
          if (RES) {
            const uint32_t* r32 = reinterpret_cast<const uint32_t*>(&rv[i]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              x[2 * q] += bf16_lo(r32[q]);
              x[2 * q + 1] += bf16_hi(r32[q]);
            }
          }
          if (p.relu) {
#pragma unroll
            for (int q = 0; q < 8; ++q) x[q] = fmaxf(x[q], 0.f);
          }
          *reinterpret_cast<uint4*>(chunk + (((piece0 + i) ^ swz) << 4)) =
              make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]),
                         pack_bf16x2(x[6], x[7]));
        }
      }
      // accumulator drained -> MMA may reuse it
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
      // staging tile complete -> async proxy -> TMA store
      fence_proxy_async();
      named_bar_sync(2, 128);
      if (leader) {
#pragma unroll
        for (int cc = 0; cc < kChunks; ++cc)
          tma_store_4d(&p.tmOut, stg + cc * kStgChunkBytes, n_tile * BN + cc * 64, tw * p.bw, th * p.bh, tn * p.bn);
        tma_store_commit();
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (leader) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();  // no CTA leaves while a peer may still multicast into it
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace irp
