// Second-generation implicit-GEMM convolution: CTA pairs (tcgen05.mma.cta_group::2), 256 x BN tiles.
//
// Same GEMM view, operand maps and fused epilogue as conv_gemm.cuh (out = act(conv(x, w) + bias (+ residual)),
// NHWC bf16, BN folded; torchvision/models/resnet.py:108-160), but one tile is computed by the two CTAs of a
// 2-CTA cluster (the two SMs of a TPC):
//
//   * CTA r of the pair owns M tile 2g+r (128 output pixels): it TMA-loads its own A boxes and HALF of the weight
//     rows (BN/2) of every k-block; the leader CTA issues ONE MMA of M = 256, N = BN per 16 channels that reads
//     both CTAs' shared memory and writes both CTAs' TMEM.  Operand bytes per FLOP are (256+BN)/(256*BN) instead
//     of (128+128)/(128*128): the L2 -> SM operand stream, which capped the single-CTA kernel at ~780 TFLOP/s,
//     is halved at BN = 256.
//   * all TMA loads of a stage (both CTAs) count their bytes on the LEADER's full barrier; the leader's
//     tcgen05.commit multicasts the slot release / accumulator-ready arrivals to both CTAs.
//   * eight epilogue warps per CTA (two per TMEM lane quarter) drain the 128 x BN accumulator in 64-channel
//     chunks through a ring of 16 KB swizzled staging buffers; each chunk leaves through its own TMA store.
//     With a residual, the producer warp prefetches the residual chunk by TMA into the staging buffer the
//     epilogue then overwrites in place.
//
//   * flat convolutions (1x1, stride 1: M is a plain row index) use the barrier-free "warp store" epilogue (WS):
//     the eight epilogue warps form two groups that take alternate chunks; a warp owns 32 rows x 64 channels of its
//     chunk (one whole 128-byte swizzled row per thread), and stores them with its OWN TMA store of a (64, 32)
//     box, so no CTA-wide barrier sits between the chunks and a warp's TMEM / shared-memory latencies overlap the
//     other warps' arithmetic.  ncu on the round-1 kernel: 20-27 % of the epilogue warps' samples were the per-chunk
//     named barrier.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM owner (+ MMA issuer in the leader), warps 2..9 =
// epilogue.
#pragma once
#include "conv_params.cuh"

namespace irp {

constexpr int kConv2Threads = 320;
constexpr int kConv2EpiThreads = 256;
constexpr int kConv2BK = 64;

constexpr int kConv2MaxCout = 2048;  // WS epilogue: the whole bias vector is staged in shared memory once

template <int BN, bool RES, bool WS>
struct Conv2Smem {
  static constexpr int kABytes = kTileM * kConv2BK * 2;    // this CTA's 128 rows
  static constexpr int kBBytes = (BN / 2) * kConv2BK * 2;  // this CTA's half of the weight rows
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kRing = RES ? 6 : 4;  // staging chunk buffers (residual prefetch needs look-ahead)
  static constexpr int kStgBytes = kRing * kStgChunkBytes;
  // mbarriers + TMEM slot, then the bias (fp32): the tile's slice, or the whole vector for the WS epilogue
  static constexpr int kBarrierBytes = 512 + (WS ? kConv2MaxCout : BN) * 4;
  static constexpr int kBudget = 227 * 1024 - 1024 - kBarrierBytes - kStgBytes;
  static constexpr int kStagesRaw = kBudget / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kTotalBytes = kStages * kStageBytes + kStgBytes + kBarrierBytes + 1024;
  static_assert(kStages >= 3, "not enough shared memory for a pipeline");
};

// one epilogue warp's share of a 64-channel chunk: 32 rows x 64 fp32 accumulator columns (+ bias, + residual read
// from the staging row it then overwrites, ReLU) -> bf16 -> the thread's 128-byte row of the swizzled staging chunk
template <bool RES>
__device__ __forceinline__ void epi_row64(uint32_t taddr, uint8_t* chunk_row, uint32_t swz, const float* sbias64,
                                          bool relu, uint64_t* res_bar, uint32_t res_parity) {
  uint32_t v[64];
  tmem_ld_32x32b_x32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
  tmem_ld_32x32b_x32(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
  uint4 rv[8];
  if (RES) {
    mbar_wait(res_bar, res_parity);
#pragma unroll
    for (int i = 0; i < 8; ++i) rv[i] = *reinterpret_cast<const uint4*>(chunk_row + ((static_cast<uint32_t>(i) ^ swz) << 4));
  }
  tmem_ld_wait();
  const float4* bp = reinterpret_cast<const float4*>(sbias64);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
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
      for (int j = 0; j < 4; ++j) {
        x[2 * j] += bf16_lo(r32[j]);
        x[2 * j + 1] += bf16_hi(r32[j]);
      }
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = fmaxf(x[j], 0.f);
    }
    *reinterpret_cast<uint4*>(chunk_row + ((static_cast<uint32_t>(i) ^ swz) << 4)) =
        make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]), pack_bf16x2(x[6], x[7]));
  }
}

template <int BN, bool RES, bool WS>
__global__ void __launch_bounds__(kConv2Threads, 1) conv_gemm2_kernel(const __grid_constant__ ConvParams p) {
  using S = Conv2Smem<BN, RES, WS>;
  constexpr int kStages = S::kStages;
  constexpr int kRing = S::kRing;
  constexpr int BK = kConv2BK;
  constexpr uint32_t kTmemCols = 2 * BN;  // double-buffered fp32 accumulator
  constexpr int kChunks = BN / 64;        // 64-channel column chunks per tile
  static_assert(kTmemCols >= 32 && kTmemCols <= 512 && (kTmemCols & (kTmemCols - 1)) == 0, "TMEM columns");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * S::kABytes;
  uint8_t* smem_stg = smem + kStages * S::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stg + S::kStgBytes);
  uint64_t* full_bar = bars;                 // [kStages] TMA (both CTAs) -> MMA; only the leader's is used
  uint64_t* empty_bar = bars + kStages;      // [kStages] MMA -> TMA, one per CTA (commit multicast)
  uint64_t* tfull_bar = bars + 2 * kStages;  // [2] MMA -> epilogue, one per CTA (commit multicast)
  uint64_t* tempty_bar = tfull_bar + 2;      // [2] epilogue (both CTAs) -> MMA; only the leader's is used
  uint64_t* res_full = tempty_bar + 2;       // [kRing] residual chunk landed in its staging buffer
  uint64_t* stg_empty = res_full + kRing;    // [kRing] staging buffer may be refilled by the producer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stg_empty + kRing);
  float* sbias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);  // [BN]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int num_groups = ((m_tiles + 1) >> 1) * p.n_tiles_n;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA[0]);
    tma_prefetch_desc(&p.tmB);
    tma_prefetch_desc(&p.tmOut);
    if (RES) tma_prefetch_desc(&p.tmRes);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      // one arrival per epilogue warp of both CTAs that reads the accumulator (WS with a single chunk per tile:
      // only one of the two warp groups touches a tile)
      mbar_init(&tempty_bar[i], (WS && kChunks == 1) ? 8 : 16);
    }
    for (int i = 0; i < kRing; ++i) {
      mbar_init(&res_full[i], 1);
      mbar_init(&stg_empty[i], WS ? 4 : 1);  // WS: the four warps that stored the chunk hand the buffer back
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_cg2(tmem_slot, kTmemCols);
    tmem_relinquish_cg2();
  }
  if (WS) {
    for (int i = threadIdx.x; i < p.cout; i += kConv2Threads) sbias[i] = __ldg(p.bias + i);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers are initialised before anyone signals across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();  // the next kernel in the stream may begin its own set-up

  const int k_blocks = p.ntaps * p.kc_blocks;

  if (warp == 0) {
    // ============================ TMA producer (both CTAs) ============================
    if (lane == 0) {
      pdl_wait();  // everything this kernel reads from the previous one goes through these TMA loads
      int stage = 0;
      uint32_t phase = 0;
      int q = 0;  // residual chunk counter
      for (int g = pair; g < num_groups; g += num_pairs) {
        const int n_tile = g % p.n_tiles_n;
        int m_tile = (g / p.n_tiles_n) * 2 + static_cast<int>(rank);  // may be a phantom tile: loads zero-fill
        const int tw = m_tile % p.tiles_w;
        m_tile /= p.tiles_w;
        const int th = m_tile % p.tiles_h;
        const int tn = m_tile / p.tiles_h;
        const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;
        const int nb0 = n_tile * BN + static_cast<int>(rank) * (BN / 2);
        for (int t = 0; t < p.ntaps; ++t) {
          for (int kc = 0; kc < p.kc_blocks; ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            const uint32_t full_leader = mapa_u32(&full_bar[stage], 0);
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * (p.a_box_bytes + S::kBBytes));
            tma_load_4d_cg2(smem_a + stage * S::kABytes, &p.tmA[p.tap_map[t]], full_leader, kc * BK,
                            w0 + p.tap_dw[t], h0 + p.tap_dh[t], n0);
            tma_load_2d_cg2(smem_b + stage * S::kBBytes, &p.tmB, full_leader, t * p.cin + kc * BK, nb0);
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
        if (RES) {
          // residual chunks of this tile -> the staging buffers the epilogue will overwrite in place
#pragma unroll 1
          for (int cc = 0; cc < kChunks; ++cc, ++q) {
            const int b = q % kRing;
            if (q >= kRing) mbar_wait(&stg_empty[b], ((q / kRing) - 1) & 1);
            mbar_arrive_expect_tx(&res_full[b], p.out_box_bytes);
            tma_load_4d(smem_stg + b * kStgChunkBytes, &p.tmRes, &res_full[b], n_tile * BN + cc * 64, w0, h0, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer (leader CTA only) ============================
    // The WHOLE warp runs the loop (converged: barrier waits, stage / phase counters and descriptor words are uniform);
    // the tcgen05 instructions are issued by the elected lane (see elect_one()).
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * kTileM, BN);
      const uint32_t a_lo0 = umma_desc_lo(smem_u32(smem_a));
      const uint32_t b_lo0 = umma_desc_lo(smem_u32(smem_b));
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int g = pair; g < num_groups; g += num_pairs) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + stage * (S::kABytes >> 4);
          const uint32_t b_lo = b_lo0 + stage * (S::kBBytes >> 4);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)  // 32 bytes of K per MMA = 2 descriptor address units
              umma_bf16_cg2_lohi(tmem_d, a_lo + 2 * k, b_lo + 2 * k, kUmmaDescHiSw128, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit_cg2(&empty_bar[stage], 3);  // frees the slot in both CTAs once these MMAs have read it
            if (kb == k_blocks - 1) umma_commit_cg2(&tfull_bar[acc], 3);  // accumulator complete -> both epilogues
          }
          __syncwarp();
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (WS) {
    // ============================ warp-store epilogue (warps 2..9, both CTAs; flat convs) ============================
    const int quarter = warp & 3;      // TMEM lane quarter this warp may read
    const int grp = (warp - 2) >> 2;   // chunks with (q & 1) == grp are this warp's
    const int row = quarter * 32 + lane;
    const uint32_t row_off = static_cast<uint32_t>(row) * 128u;
    const uint32_t swz = static_cast<uint32_t>(row & 7);
    int acc = 0;
    uint32_t acc_phase = 0;
    int q = 0;        // chunk counter (same sequence as the producer's)
    int prev_b = -1;  // staging buffer of this warp's previous store (RES: handed back once that store has been read)
    for (int g = pair; g < num_groups; g += num_pairs) {
      const int n_tile = g % p.n_tiles_n;
      const int m_tile = (g / p.n_tiles_n) * 2 + static_cast<int>(rank);
      const int last_cc = kChunks == 1 ? 0 : kChunks - 2 + grp;  // this warp's last chunk of the tile
      bool acc_ready = false;
#pragma unroll 1
      for (int cc = 0; cc < kChunks; ++cc, ++q) {
        if ((q & 1) != grp) continue;
        if (!acc_ready) {
          mbar_wait(&tfull_bar[acc], acc_phase);
          tc_fence_after();
          acc_ready = true;
        }
        const int b = q % kRing;
        uint8_t* chunk = smem_stg + b * kStgChunkBytes;
        if (!RES) {
          // this warp's previous store out of the same buffer (kRing / 2 of its stores ago) must have been read
          if (lane == 0) tma_store_wait_read<kRing / 2 - 1>();
          __syncwarp();
        }
        const uint32_t taddr = tmem_base + acc * BN + (static_cast<uint32_t>(quarter * 32) << 16) + cc * 64;
        epi_row64<RES>(taddr, chunk + row_off, swz, sbias + n_tile * BN + cc * 64, p.relu != 0, &res_full[b],
                       (q / kRing) & 1);
        if (cc == last_cc) {
          // accumulator drained by this warp -> one arrival on the leader's barrier (the TMEM reads completed at
          // the tcgen05.wait::ld inside epi_row64)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_u32(&tempty_bar[acc], 0));
        }
        fence_proxy_async();  // this thread's staging writes -> visible to the TMA store
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&p.tmOutW, chunk + quarter * 4096, n_tile * BN + cc * 64, m_tile * kTileM + quarter * 32, 0, 0);
          tma_store_commit();
          if (RES) {
            tma_store_wait_read<1>();  // every store of this warp but the one just committed has been read
            if (prev_b >= 0) mbar_arrive(&stg_empty[prev_b]);
            prev_b = b;
          }
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (lane == 0) tma_store_wait_all<0>();
  } else {
    // ============================ epilogue (warps 2..9, both CTAs) ============================
    const int quarter = warp & 3;         // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;     // which 32 columns of a 64-channel chunk
    const int row = quarter * 32 + lane;
    const bool leader = (threadIdx.x == 64);  // issues the TMA stores, owns their bulk groups
    const uint32_t row_off = static_cast<uint32_t>(row) * 128u;
    const uint32_t swz = static_cast<uint32_t>(row & 7);
    const int piece0 = half * 4;  // first 16-byte piece of this thread's 64 bytes inside the 128-byte row
    int acc = 0;
    uint32_t acc_phase = 0;
    int q = 0;  // chunk counter (same sequence as the producer's)
    // The bias slice of a tile is staged in shared memory (with ~227 KB of shared memory carved out there is no
    // L1 left, so a global load in the chunk loop is an exposed L2 round trip); it is fetched one tile ahead.
    const int et = threadIdx.x - 64;  // 0..255
    float bias_next = 0.f;
    if (et < BN && pair < num_groups) bias_next = __ldg(p.bias + (pair % p.n_tiles_n) * BN + et);
    for (int g = pair; g < num_groups; g += num_pairs) {
      const int n_tile = g % p.n_tiles_n;
      int m_tile = (g / p.n_tiles_n) * 2 + static_cast<int>(rank);
      const int tw = m_tile % p.tiles_w;
      m_tile /= p.tiles_w;
      const int th = m_tile % p.tiles_h;
      const int tn = m_tile / p.tiles_h;

      // every thread is past the previous tile's last chunk barrier, i.e. done reading the old slice
      if (et < BN) sbias[et] = bias_next;
      named_bar_sync(2, kConv2EpiThreads);
      if (et < BN && g + num_pairs < num_groups) bias_next = __ldg(p.bias + ((g + num_pairs) % p.n_tiles_n) * BN + et);

      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * BN + (static_cast<uint32_t>(quarter * 32) << 16) + half * 32;
#pragma unroll 1
      for (int cc = 0; cc < kChunks; ++cc, ++q) {
        const int b = q % kRing;
        uint8_t* chunk = smem_stg + b * kStgChunkBytes + row_off;
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + cc * 64, v);
        uint4 rv[4];
        if (RES) {
          mbar_wait(&res_full[b], (q / kRing) & 1);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            rv[i] = *reinterpret_cast<const uint4*>(chunk + (((piece0 + i) ^ swz) << 4));
        }
        const float4* bp = reinterpret_cast<const float4*>(sbias + cc * 64 + half * 32);
        float4 bv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) bv[i] = bp[i];
        tmem_ld_wait();
        if (cc == kChunks - 1) {
          // accumulator drained by this warp -> one arrival on the leader's barrier
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_u32(&tempty_bar[acc], 0));
        }
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
          if (RES) {
            const uint32_t* r32 = reinterpret_cast<const uint32_t*>(&rv[i]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              x[2 * j] += bf16_lo(r32[j]);
              x[2 * j + 1] += bf16_hi(r32[j]);
            }
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = fmaxf(x[j], 0.f);
          }
          *reinterpret_cast<uint4*>(chunk + (((piece0 + i) ^ swz) << 4)) =
              make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]),
                         pack_bf16x2(x[6], x[7]));
        }
        // staging chunk complete -> async proxy -> TMA store
        fence_proxy_async();
        if (leader) {
          if (RES) {
            // stores of chunks <= q-2 have finished reading: hand that buffer back to the producer
            tma_store_wait_read<1>();
            if (q >= 2) mbar_arrive(&stg_empty[(q - 2) % kRing]);
          } else {
            // stores of chunks <= q-kRing+1 have finished reading: chunk q+1's buffer is free for everyone
            tma_store_wait_read<kRing - 2>();
          }
        }
        named_bar_sync(1, kConv2EpiThreads);
        if (leader) {
          tma_store_4d(&p.tmOut, smem_stg + b * kStgChunkBytes, n_tile * BN + cc * 64, tw * p.bw, th * p.bh,
                       tn * p.bn);
          tma_store_commit();
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (leader) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still be reading this CTA's shared memory / signalling its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, kTmemCols);
  }
}

}  // namespace irp
