// Two chained 1x1 convolutions in one kernel: a bottleneck's conv3 (+ folded BN + residual + ReLU) and the NEXT
// bottleneck's conv1 (+ folded BN + ReLU)  (torchvision/models/resnet.py:150-159 then :142-144 of the next block).
//
//   Y   = relu(T2 . W3^T + b3 + R)      [M, N1]   written to HBM (it is the block output and the next residual)
//         or, for a block whose shortcut is a stride-1 1x1 convolution of X (layer1's first block):
//   Y   = relu([T2 | X] . [W3 | Wds]^T + (b3 + bds))   -- the shortcut is never written to or read from HBM
//   T1' = relu(Y  . W1^T + b1)          [M, N2]   written to HBM (input of the next block's 3x3)
//
// Unfused, conv1 re-reads the whole N1-channel block output from HBM (411 MB per 256 images in layer1) for a GEMM
// with almost no arithmetic; here the bf16 Y tile never leaves the SM between the two GEMMs: the 128B-swizzled
// 16 KB staging chunks the first epilogue writes (and TMA-stores) ARE canonical K-major A operands, so the second
// GEMM reads them in place.
//
// CTA pairs (tcgen05 cta_group::2, M = 256 rows per pair).  N1 is swept in passes of 128 columns:
//   GEMM1(pass)  : acc1[pass & 1] (128 TMEM columns, double buffered) = T2 tile . W3[pass]^T
//   epilogue1    : + bias + residual (TMA-prefetched into the ring buffer it overwrites), ReLU, bf16 -> two ring
//                  chunks -> TMA store to Y, and "chunk ready" to the MMA warp
//   GEMM2(pass)  : acc2 (N2 TMEM columns) += Y chunks . W1[:, pass]^T, software-pipelined one pass behind GEMM1
//   epilogue2    : after the last pass, acc2 + bias, ReLU -> ring chunks -> TMA store to T1'
// A ring buffer is recycled when BOTH its TMA store has finished reading it and the GEMM2 MMAs that used it have
// completed (two arrivals on its barrier).
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM owner (+ MMA issuer in the leader), warps 2..9 =
// epilogue.
#pragma once
#include "conv_gemm2.cuh"

namespace irp {

constexpr int kChainThreads = 320;
constexpr int kChainEpiThreads = 256;
constexpr int kChainBN1 = 128;  // columns of Y per pass
constexpr int kChainRing = 6;
constexpr int kChainMaxN1 = 1024;

struct alignas(64) ChainParams {
  CUtensorMap tmA;     // T2  [M, K1]  dims (K1, M), box (64, 128)
  CUtensorMap tmB1;    // W3  [N1, K1] dims (K1, N1), box (64, 64)  (half of a pass per CTA)
  CUtensorMap tmRes;   // R   [M, N1]  dims (N1, M), box (64, 128)
  CUtensorMap tmB2;    // W1' [N2, N1] dims (N1, N2), box (64, N2/2)
  CUtensorMap tmA2;    // X   [M, K2]  optional second GEMM1 operand (see k2_blocks)
  CUtensorMap tmYW;    // Y   [M, N1]  dims (N1, M), box (64, 32): one epilogue warp's rows of a chunk
  CUtensorMap tmOut2W; // T1' [M, N2]  dims (N2, M), box (64, 32)
  const float* bias1;  // [N1]
  const float* bias2;  // [N2]
  int k1_blocks;       // K1 / 64
  int k2_blocks;       // K2 / 64: GEMM1 continues over X . Wds^T (W3 and Wds concatenated along K in tmB1) -- the
                       // block's 1x1 downsample convolution computed in the same accumulator instead of being read
                       // back as a residual (first bottleneck of layer1, where it has stride 1)
  int has_res;         // 0: no residual tensor (k2_blocks > 0)
  int passes;          // N1 / 128
  int n1;              // N1
  int m_tiles;         // ceil(M / 128)
};

template <int N2>
struct ChainSmem {
  static constexpr int kStages = N2 == 256 ? 2 : (N2 == 128 ? 3 : 4);
  static constexpr int kB2Slots = 4;  // two passes of W1' k-blocks in flight
  static constexpr int kABytes = kTileM * 64 * 2;          // 16 KB
  static constexpr int kB1Bytes = (kChainBN1 / 2) * 128;   // 8 KB
  static constexpr int kStageBytes = kABytes + kB1Bytes;
  static constexpr int kB2Bytes = (N2 / 2) * 128;          // one 64-wide K block of this CTA's half of W1'
  static constexpr int kRingBytes = kChainRing * kStgChunkBytes;
  static constexpr int kBarrierBytes = 512;
  static constexpr int kBiasBytes = (kChainMaxN1 + N2) * 4;
  static constexpr int kTotalBytes = kStages * kStageBytes + kRingBytes + kB2Slots * kB2Bytes + kBarrierBytes + kBiasBytes + 1024;
  static_assert(kTotalBytes <= 227 * 1024, "shared memory budget");
};

__device__ __forceinline__ void tma_store_2d(const void* desc, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(desc)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}

template <int N2>
__global__ void __launch_bounds__(kChainThreads, 1) conv_chain_kernel(const __grid_constant__ ChainParams p) {
  using S = ChainSmem<N2>;
  constexpr int kStages = S::kStages;
  constexpr int R = kChainRing;
  constexpr int kOutChunks = N2 / 64;
  constexpr int kB2 = S::kB2Slots;
  constexpr uint32_t kTmemCols = 512;
  constexpr uint32_t kAcc2Col = 256;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b1 = smem + kStages * S::kABytes;
  uint8_t* smem_ring = smem + kStages * S::kStageBytes;
  uint8_t* smem_b2 = smem_ring + S::kRingBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b2 + S::kB2Slots * S::kB2Bytes);
  uint64_t* full1 = bars;                    // [kStages] leader's is used
  uint64_t* empty1 = full1 + kStages;        // [kStages]
  uint64_t* b2full = empty1 + kStages;       // [kB2] leader's is used
  uint64_t* b2empty = b2full + kB2;          // [kB2]
  uint64_t* acc1_full = b2empty + kB2;       // [2]
  uint64_t* acc1_empty = acc1_full + 2;      // [2] leader's is used
  uint64_t* acc2_full = acc1_empty + 2;      // [1]
  uint64_t* acc2_empty = acc2_full + 1;      // [1] leader's is used
  uint64_t* res_full = acc2_empty + 1;       // [R]
  uint64_t* stg_empty = res_full + R;        // [R] five arrivals per use: four warp stores read + GEMM2 completion
  uint64_t* ychunk_full = stg_empty + R;     // [R] leader's is used: one arrival per epilogue warp of both CTAs
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ychunk_full + R);
  float* sbias1 = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + S::kBarrierBytes);  // [n1]
  float* sbias2 = sbias1 + kChainMaxN1;                                                           // [N2]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int pair_tiles = (p.m_tiles + 1) >> 1;
  const int P = p.passes;
  const int L = 2 * P + kOutChunks;  // ring chunks consumed per tile
  // Ring-chunk order (identical in the producer, the MMA warp and the epilogue): the T1' chunks of tile i-1 are
  // written AFTER the first pass of tile i, so that the GEMM2 of a tile's last pass overlaps the next tile's first
  // epilogue instead of leaving the epilogue warps idle:
  //   tile 0: pass 0, pass 1, ...      tile i >= 1: pass 0, T1'(i-1), pass 1, ...      end: T1'(last)
  auto q_pass = [&](int i, int ps, int c) {
    return i == 0 ? 2 * ps + c : 2 * P + (i - 1) * L + (ps == 0 ? c : 2 + kOutChunks + 2 * (ps - 1) + c);
  };
  auto q_out = [&](int i, int j) { return 2 * P + i * L + 2 + j; };
  auto is_out_chunk = [&](int q) {
    if (q < 2 * P) return false;
    const int r = (q - 2 * P) % L;
    return r >= 2 && r < 2 + kOutChunks;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB1);
    tma_prefetch_desc(&p.tmRes);
    tma_prefetch_desc(&p.tmYW);
    tma_prefetch_desc(&p.tmB2);
    tma_prefetch_desc(&p.tmOut2W);
    if (p.k2_blocks > 0) tma_prefetch_desc(&p.tmA2);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full1[i], 1);
      mbar_init(&empty1[i], 1);
    }
    for (int i = 0; i < kB2; ++i) {
      mbar_init(&b2full[i], 1);
      mbar_init(&b2empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc1_full[i], 1);
      mbar_init(&acc1_empty[i], 16);
    }
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, kOutChunks >= 2 ? 16 : 8);  // the epilogue warps (both CTAs) that read acc2 of a tile
    for (int i = 0; i < R; ++i) {
      mbar_init(&res_full[i], 1);
      mbar_init(&stg_empty[i], 5);    // four warp stores read + (GEMM2 completion | one extra arrival for T1' chunks)
      mbar_init(&ychunk_full[i], 8);  // four warps of each CTA
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_cg2(tmem_slot, kTmemCols);
    tmem_relinquish_cg2();
  }
  for (int i = threadIdx.x; i < p.n1; i += kChainThreads) sbias1[i] = __ldg(p.bias1 + i);
  for (int i = threadIdx.x; i < N2; i += kChainThreads) sbias2[i] = __ldg(p.bias2 + i);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();

  if (warp == 0) {
    // ============================ TMA producer (both CTAs) ============================
    if (lane == 0) {
      pdl_wait();
      int stage = 0;
      uint32_t phase = 0;
      int b2n = 0;  // B2 k-blocks issued so far
      int i = 0;    // local tile counter
      const int kt = p.k1_blocks + p.k2_blocks;
      for (int mt = pair; mt < pair_tiles; mt += num_pairs, ++i) {
        const int row0 = (mt * 2 + static_cast<int>(rank)) * kTileM;  // may be past M: loads zero-fill, stores clip
        for (int ps = 0; ps < P; ++ps) {
          const int col0 = ps * kChainBN1;
          for (int kb = 0; kb < kt; ++kb) {
            mbar_wait(&empty1[stage], phase ^ 1);
            const uint32_t full_leader = mapa_u32(&full1[stage], 0);
            if (rank == 0) mbar_arrive_expect_tx(&full1[stage], 2 * S::kStageBytes);
            if (kb < p.k1_blocks)
              tma_load_2d_cg2(smem_a + stage * S::kABytes, &p.tmA, full_leader, kb * 64, row0);
            else
              tma_load_2d_cg2(smem_a + stage * S::kABytes, &p.tmA2, full_leader, (kb - p.k1_blocks) * 64, row0);
            tma_load_2d_cg2(smem_b1 + stage * S::kB1Bytes, &p.tmB1, full_leader, kb * 64,
                            col0 + static_cast<int>(rank) * (kChainBN1 / 2));
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
          // residual chunks of this pass -> the ring buffers epilogue1 overwrites in place
#pragma unroll 1
          for (int c = 0; c < 2 && p.has_res; ++c) {
            const int q = q_pass(i, ps, c);
            const int b = q % R;
            if (q >= R) mbar_wait(&stg_empty[b], ((q / R) - 1) & 1);
            mbar_arrive_expect_tx(&res_full[b], kStgChunkBytes);
            tma_load_2d(smem_ring + b * kStgChunkBytes, &p.tmRes, &res_full[b], col0 + c * 64, row0);
          }
          // the two 64-wide K blocks of W1' that GEMM2 of this pass multiplies with
#pragma unroll 1
          for (int c = 0; c < 2; ++c, ++b2n) {
            const int s2 = b2n % kB2;
            mbar_wait(&b2empty[s2], ((b2n / kB2) & 1) ^ 1);
            const uint32_t full_leader = mapa_u32(&b2full[s2], 0);
            if (rank == 0) mbar_arrive_expect_tx(&b2full[s2], 2 * S::kB2Bytes);
            tma_load_2d_cg2(smem_b2 + s2 * S::kB2Bytes, &p.tmB2, full_leader, col0 + c * 64,
                            static_cast<int>(rank) * (N2 / 2));
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer (leader CTA only) ============================
    // The whole warp runs the loop converged; the tcgen05 instructions are issued by the elected lane (elect_one()):
    // issued from an `if (lane == 0)` region, each MMA cost ~100 cycles of ELECT / R2UR waterfall and this thread's
    // 540 instructions per pass -- not the epilogue, the tensor pipe (21 %) or HBM -- set the pass time.
    if (rank == 0) {
      constexpr uint32_t idesc1 = umma_idesc_bf16(2 * kTileM, kChainBN1);
      constexpr uint32_t idesc2 = umma_idesc_bf16(2 * kTileM, N2);
      const uint32_t a_lo0 = umma_desc_lo(smem_u32(smem_a));
      const uint32_t b1_lo0 = umma_desc_lo(smem_u32(smem_b1));
      const uint32_t ring_lo0 = umma_desc_lo(smem_u32(smem_ring));
      const uint32_t b2_lo0 = umma_desc_lo(smem_u32(smem_b2));
      int stage = 0;
      uint32_t phase = 0;
      int my_tiles = 0;
      for (int mt = pair; mt < pair_tiles; mt += num_pairs) ++my_tiles;
      const int total_passes = my_tiles * P;
      const int kt = p.k1_blocks + p.k2_blocks;
      int i2 = 0, ps2 = 0, b2n = 0;  // GEMM2 runs one pass behind GEMM1: its (tile, pass) and W1' k-block counters
      // GEMM2 of the next pending pass (tile i2, pass ps2)
      auto gemm2 = [&]() {
        if (ps2 == 0) {
          mbar_wait(acc2_empty, (i2 & 1) ^ 1);
          tc_fence_after();
        }
#pragma unroll 1
        for (int c = 0; c < 2; ++c, ++b2n) {
          const int q = q_pass(i2, ps2, c);
          const int b = q % R;
          const int s2 = b2n % kB2;
          mbar_wait(&ychunk_full[b], (q / R) & 1);
          mbar_wait(&b2full[s2], (b2n / kB2) & 1);
          tc_fence_after();
          const uint32_t a_lo = ring_lo0 + b * (kStgChunkBytes >> 4);
          const uint32_t b_lo = b2_lo0 + s2 * (S::kB2Bytes >> 4);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_cg2_lohi(tmem_base + kAcc2Col, a_lo + 2 * k, b_lo + 2 * k, kUmmaDescHiSw128, idesc2,
                                 (ps2 | c | k) != 0 ? 1u : 0u);
            umma_commit_cg2(&b2empty[s2], 3);
            umma_commit_cg2(&stg_empty[b], 3);  // second arrival on the ring buffer (the first is its TMA store)
            if (c == 1 && ps2 == P - 1) umma_commit_cg2(acc2_full, 3);
          }
          __syncwarp();
        }
        if (++ps2 == P) {
          ps2 = 0;
          ++i2;
        }
      };
      for (int g = 0; g < total_passes; ++g) {
        const int a1 = g & 1;
        mbar_wait(&acc1_empty[a1], ((g >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + a1 * kChainBN1;
        for (int kb = 0; kb < kt; ++kb) {
          mbar_wait(&full1[stage], phase);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + stage * (S::kABytes >> 4);
          const uint32_t b_lo = b1_lo0 + stage * (S::kB1Bytes >> 4);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_cg2_lohi(tmem_d, a_lo + 2 * k, b_lo + 2 * k, kUmmaDescHiSw128, idesc1, (kb | k) != 0 ? 1u : 0u);
            umma_commit_cg2(&empty1[stage], 3);
            if (kb == kt - 1) umma_commit_cg2(&acc1_full[a1], 3);
          }
          __syncwarp();
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (g >= 1) gemm2();
      }
      if (total_passes > 0) gemm2();
    }
  } else {
    // ============================ epilogue (warps 2..9, both CTAs) ============================
    // Barrier-free "warp store" epilogue (see conv_gemm2.cuh): the two groups of four warps take the ring chunks of
    // even / odd q; a warp owns 32 rows x 64 channels of its chunk and stores them with its own (64, 32)-box TMA store.
    const int quarter = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t row_off = static_cast<uint32_t>(row) * 128u;
    const uint32_t swz = static_cast<uint32_t>(row & 7);
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    const bool has_res = p.has_res != 0;
    int g = 0;        // global pass counter
    int i = 0;        // local tile counter
    int prev_q = -1;  // this warp's previously stored chunk: its ring buffer is handed back once that store was read
    // lane 0, right after committing the store of chunk q
    auto after_store = [&](int q) {
      tma_store_wait_read<1>();  // every store of this warp but the one just committed has finished reading
      if (prev_q >= 0) {
        mbar_arrive(&stg_empty[prev_q % R]);
        // T1' chunks have no GEMM2 consumer: one warp supplies the fifth arrival
        if (quarter == 0 && is_out_chunk(prev_q)) mbar_arrive(&stg_empty[prev_q % R]);
      }
      prev_q = q;
    };
    // everything after the chunk's rows are in the staging buffer
    auto publish = [&](int q, const void* tmap, int col0, int trow0, bool out_chunk) {
      const int b = q % R;
      fence_proxy_async();  // generic-proxy writes -> visible to the TMA store and to the GEMM2 MMAs
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(tmap, smem_ring + b * kStgChunkBytes + quarter * 4096, col0, trow0 + quarter * 32);
        tma_store_commit();
        mbar_arrive_cluster(mapa_u32(&ychunk_full[b], 0));  // this warp's 32 rows of the chunk are in place
        // a T1' use of the buffer has no residual prefetch: complete that phase by hand so every barrier of
        // buffer b advances exactly once per use
        if (out_chunk && quarter == 0) mbar_arrive(&res_full[b]);
        after_store(q);
      }
    };
    // epilogue2 of tile ti (rows trow0): acc2 + bias, ReLU -> ring chunks -> TMA store to T1'
    auto epilogue2 = [&](int ti, int trow0) {
      int last_j = -1;
#pragma unroll
      for (int j = 0; j < kOutChunks; ++j)
        if ((q_out(ti, j) & 1) == grp) last_j = j;
      if (last_j < 0) return;
      mbar_wait(acc2_full, ti & 1);
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < kOutChunks; ++j) {
        const int q = q_out(ti, j);
        if ((q & 1) != grp) continue;
        const int b = q % R;
        if (q >= R) mbar_wait(&stg_empty[b], ((q / R) - 1) & 1);
        epi_row64<false>(tmem_base + kAcc2Col + j * 64 + lane_base, smem_ring + b * kStgChunkBytes + row_off, swz,
                         sbias2 + j * 64, true, nullptr, 0);
        if (j == last_j) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_u32(acc2_empty, 0));
        }
        publish(q, &p.tmOut2W, j * 64, trow0, true);
      }
    };
    int prev_row0 = 0;
    for (int mt = pair; mt < pair_tiles; mt += num_pairs, ++i) {
      const int row0 = (mt * 2 + static_cast<int>(rank)) * kTileM;
      for (int ps = 0; ps < P; ++ps, ++g) {
        const int a1 = g & 1;
        // exactly one of the pass's two chunks is this warp's
        const int c = ((q_pass(i, ps, 0) & 1) == grp) ? 0 : 1;
        const int q = q_pass(i, ps, c);
        const int b = q % R;
        mbar_wait(&acc1_full[a1], (g >> 1) & 1);
        tc_fence_after();
        uint8_t* chunk_row = smem_ring + b * kStgChunkBytes + row_off;
        const uint32_t taddr = tmem_base + a1 * kChainBN1 + c * 64 + lane_base;
        const float* bias = sbias1 + ps * kChainBN1 + c * 64;
        if (has_res) {
          epi_row64<true>(taddr, chunk_row, swz, bias, true, &res_full[b], (q / R) & 1);
        } else {
          // nothing was prefetched into the buffer: wait for its previous use (stores read + GEMM2) ourselves
          if (q >= R) mbar_wait(&stg_empty[b], ((q / R) - 1) & 1);
          epi_row64<false>(taddr, chunk_row, swz, bias, true, nullptr, 0);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_u32(&acc1_empty[a1], 0));
        publish(q, &p.tmYW, ps * kChainBN1 + c * 64, row0, false);
        // deferred T1' epilogue of the previous tile: its last GEMM2 ran while this tile's first pass drained
        if (ps == 0 && i > 0) epilogue2(i - 1, prev_row0);
      }
      prev_row0 = row0;
    }
    if (i > 0) epilogue2(i - 1, prev_row0);
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, kTmemCols);
  }
}

}  // namespace irp
