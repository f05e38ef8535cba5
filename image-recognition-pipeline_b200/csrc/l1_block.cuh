// One launch for the tail of a layer1 bottleneck and the head of the next one
// (torchvision/models/resnet.py:146-159 of block b, then :142-144 of block b+1):
//
//   T2  = relu(conv3x3(T1, W2) + b2)            64 -> 64 channels @56x56      (never leaves the SM)
//   Y   = relu(T2 . W3^T + b3 + R)              64 -> 256, R = block input    written to HBM
//   T1' = relu(Y . W1'^T + b1')                 256 -> N2 (64, or 128 into layer2)   written to HBM
//
// It combines conv3x3_c64.cuh (the 10 x 18 pixel input patch is loaded once, every tap's A operand is a shifted
// SW128 descriptor window of it) with conv_chain.cuh (the epilogue's swizzled 16 KB chunks ARE the next GEMM's A
// operand) in CTA pairs: CTA r of a pair owns the 8 x 16 pixel tile 2g+r, every weight matrix is split in halves
// between the two CTAs and stays RESIDENT in shared memory (W2 36 KB + W3 16 KB + W1' 16/32 KB per CTA), and all
// three GEMMs are tcgen05.mma.cta_group::2 with M = 256.  Per block the 64-channel T2 tensor (103 MB per 256
// images, written once and read once) and two kernel boundaries disappear.
//
// TMEM columns: [0,128) conv2 accumulator x2, [128,384) conv3 accumulator (128 columns per pass) x2, [384, 384+N2)
// conv1' accumulator.   Ring-chunk order, barriers and the deferred T1' epilogue are those of conv_chain.cuh.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM owner (+ MMA issuer in the leader), warps 2..9 =
// epilogue.
#pragma once
#include "conv3x3_c64.cuh"
#include "conv_chain.cuh"

namespace irp {

constexpr int kL1Threads = 320;
constexpr int kL1EpiThreads = 256;
constexpr int kL1N1 = 256;     // conv3 output channels
constexpr int kL1Passes = 2;   // of 128 columns

struct alignas(64) L1BlockParams {
  CUtensorMap tmIn;    // T1  (64, W, H, B)   box (64, 10, 18, 1)
  CUtensorMap tmW2;    // W2  (576, 64)       box (64, 32)      half of the output channels per CTA
  CUtensorMap tmW3;    // W3  (64, 256)       box (64, 64)      half of a 128-column pass per CTA
  CUtensorMap tmW1;    // W1' (256, N2)       box (64, N2/2)
  CUtensorMap tmRes;   // R   (256, W, H, B)  box (64, 8, 16, 1)
  CUtensorMap tmY;     // Y   same geometry
  CUtensorMap tmOut2;  // T1' (N2, W, H, B)   box (64, 8, 16, 1)
  const float* bias2;  // [64]
  const float* bias3;  // [256]
  const float* bias1;  // [N2]
  int tiles_w, tiles_h, tiles_n;
};

template <int N2>
struct L1Smem {
  static constexpr int kRing = N2 == 64 ? 5 : 4;
  static constexpr int kW2Bytes = 9 * 32 * 128;         // 36 KB
  static constexpr int kW3Bytes = kL1Passes * 64 * 128; // 16 KB
  static constexpr int kW1Bytes = 4 * (N2 / 2) * 128;   // 4 K blocks of this CTA's half
  static constexpr int kPatchSlots = 2;
  static constexpr int kT2Bytes = kStgChunkBytes;
  static constexpr int kRingBytes = kRing * kStgChunkBytes;
  static constexpr int kBarrierBytes = 512;
  static constexpr int kBiasBytes = (64 + kL1N1 + N2) * 4;
  static constexpr int kTotalBytes = kW2Bytes + kW3Bytes + kW1Bytes + kPatchSlots * kC64PatchStride + kT2Bytes +
                                     kRingBytes + kBarrierBytes + kBiasBytes + 1024;
  static_assert(kTotalBytes <= 227 * 1024, "shared memory budget");
};

template <int N2>
__global__ void __launch_bounds__(kL1Threads, 1) l1_block_kernel(const __grid_constant__ L1BlockParams p) {
  using S = L1Smem<N2>;
  constexpr int R = S::kRing;
  constexpr int P = kL1Passes;
  constexpr int kOutChunks = N2 / 64;
  constexpr int L = 2 * P + kOutChunks;
  constexpr uint32_t kTmemCols = 512;
  constexpr uint32_t kAcc1Col = 128, kAcc2Col = 384;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* smem_w2 = smem;                                   // [9 taps][32 rows][128 B]
  uint8_t* smem_w3 = smem_w2 + S::kW2Bytes;                  // [2 passes][64 rows][128 B]
  uint8_t* smem_w1 = smem_w3 + S::kW3Bytes;                  // [4 K blocks][N2/2 rows][128 B]
  uint8_t* smem_patch = smem_w1 + S::kW1Bytes;               // 2 x 23 KB
  uint8_t* smem_t2 = smem_patch + S::kPatchSlots * kC64PatchStride;  // 16 KB
  uint8_t* smem_ring = smem_t2 + S::kT2Bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_ring + S::kRingBytes);
  uint64_t* wfull = bars;                   // [1] leader's is used
  uint64_t* pfull = wfull + 1;              // [2] leader's is used
  uint64_t* pempty = pfull + 2;             // [2]
  uint64_t* c2_full = pempty + 2;           // [2]
  uint64_t* c2_empty = c2_full + 2;         // [2] leader's is used
  uint64_t* t2_full = c2_empty + 2;         // [1] leader's is used: one arrival per CTA
  uint64_t* t2_empty = t2_full + 1;         // [1]
  uint64_t* acc1_full = t2_empty + 1;       // [2]
  uint64_t* acc1_empty = acc1_full + 2;     // [2] leader's is used
  uint64_t* acc2_full = acc1_empty + 2;     // [1]
  uint64_t* acc2_empty = acc2_full + 1;     // [1] leader's is used
  uint64_t* res_full = acc2_empty + 1;      // [R]
  uint64_t* stg_empty = res_full + R;       // [R] two arrivals per use
  uint64_t* ychunk_full = stg_empty + R;    // [R] leader's is used: one arrival per CTA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ychunk_full + R);
  float* sbias2 = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + S::kBarrierBytes);  // [64]
  float* sbias3 = sbias2 + 64;                                                                    // [256]
  float* sbias1 = sbias3 + kL1N1;                                                                 // [N2]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int tiles_img = p.tiles_w * p.tiles_h;
  const int num_tiles = tiles_img * p.tiles_n;
  const int pair_tiles = (num_tiles + 1) >> 1;

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
    tma_prefetch_desc(&p.tmIn);
    tma_prefetch_desc(&p.tmW2);
    tma_prefetch_desc(&p.tmW3);
    tma_prefetch_desc(&p.tmW1);
    tma_prefetch_desc(&p.tmRes);
    tma_prefetch_desc(&p.tmY);
    tma_prefetch_desc(&p.tmOut2);
    mbar_init(wfull, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&pfull[i], 1);
      mbar_init(&pempty[i], 1);
      mbar_init(&c2_full[i], 1);
      mbar_init(&c2_empty[i], 16);
      mbar_init(&acc1_full[i], 1);
      mbar_init(&acc1_empty[i], 16);
    }
    mbar_init(t2_full, 2);
    mbar_init(t2_empty, 1);
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, 16);
    for (int i = 0; i < R; ++i) {
      mbar_init(&res_full[i], 1);
      mbar_init(&stg_empty[i], 2);
      mbar_init(&ychunk_full[i], 2);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_cg2(tmem_slot, kTmemCols);
    tmem_relinquish_cg2();
  }
  for (int i = threadIdx.x; i < 64; i += kL1Threads) sbias2[i] = __ldg(p.bias2 + i);
  for (int i = threadIdx.x; i < kL1N1; i += kL1Threads) sbias3[i] = __ldg(p.bias3 + i);
  for (int i = threadIdx.x; i < N2; i += kL1Threads) sbias1[i] = __ldg(p.bias1 + i);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();

  if (warp == 0) {
    // ============================ TMA producer (both CTAs) ============================
    if (lane == 0) {
      // resident weights: this CTA's halves; all bytes (both CTAs) are counted on the leader's barrier
      {
        const uint32_t wl = mapa_u32(wfull, 0);
        if (rank == 0) mbar_arrive_expect_tx(wfull, 2 * (S::kW2Bytes + S::kW3Bytes + S::kW1Bytes));
        const int r = static_cast<int>(rank);
#pragma unroll 1
        for (int t = 0; t < 9; ++t) tma_load_2d_cg2(smem_w2 + t * 4096, &p.tmW2, wl, t * 64, r * 32);
#pragma unroll 1
        for (int ps = 0; ps < P; ++ps) tma_load_2d_cg2(smem_w3 + ps * 8192, &p.tmW3, wl, 0, ps * 128 + r * 64);
#pragma unroll 1
        for (int kb = 0; kb < 4; ++kb)
          tma_load_2d_cg2(smem_w1 + kb * (N2 / 2) * 128, &p.tmW1, wl, kb * 64, r * (N2 / 2));
      }
      pdl_wait();  // T1 and the residual come from the previous kernels
      // patch of local tile i (pair tile mt) -> slot i & 1
      auto load_patch = [&](int i, int mt) {
        const int tile = mt * 2 + static_cast<int>(rank);  // may be a phantom tile: loads zero-fill, stores clip
        const int tw = tile % p.tiles_w, th = (tile / p.tiles_w) % p.tiles_h, n = tile / tiles_img;
        const int slot = i & 1;
        mbar_wait_dbg(&pempty[slot], ((i >> 1) & 1) ^ 1, __LINE__);
        const uint32_t pl = mapa_u32(&pfull[slot], 0);
        if (rank == 0) mbar_arrive_expect_tx(&pfull[slot], 2 * kC64PatchBytes);
        tma_load_4d_cg2(smem_patch + slot * kC64PatchStride, &p.tmIn, pl, 0, tw * kC64TileW - 1, th * kC64TileH - 1, n);
      };
      if (pair < pair_tiles) load_patch(0, pair);
      int i = 0;
      for (int mt = pair; mt < pair_tiles; mt += num_pairs, ++i) {
        const int tile = mt * 2 + static_cast<int>(rank);
        const int tw = tile % p.tiles_w, th = (tile / p.tiles_w) % p.tiles_h, n = tile / tiles_img;
        const int w0 = tw * kC64TileW, h0 = th * kC64TileH;
        // The NEXT tile's patch goes out before this tile's residuals: the MMA warp issues conv2(i+1) ahead of the
        // conv1' partial products whose completion frees the ring buffers those residual loads wait for (with the
        // 4-buffer ring a pass-1 residual reuses this tile's own pass-0 buffer) -- the other order deadlocks.
        if (mt + num_pairs < pair_tiles) load_patch(i + 1, mt + num_pairs);
        for (int ps = 0; ps < P; ++ps) {
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            const int q = q_pass(i, ps, c);
            const int b = q % R;
            if (q >= R) mbar_wait_dbg(&stg_empty[b], ((q / R) - 1) & 1, __LINE__, (N2 << 16) | q);
            mbar_arrive_expect_tx(&res_full[b], kStgChunkBytes);
            tma_load_4d(smem_ring + b * kStgChunkBytes, &p.tmRes, &res_full[b], ps * 128 + c * 64, w0, h0, n);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer (leader CTA only) ============================
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc_c2 = umma_idesc_bf16(2 * kTileM, 64);
      constexpr uint32_t idesc_1 = umma_idesc_bf16(2 * kTileM, 128);
      constexpr uint32_t idesc_2 = umma_idesc_bf16(2 * kTileM, N2);
      const uint32_t w2_addr = smem_u32(smem_w2), w3_addr = smem_u32(smem_w3), w1_addr = smem_u32(smem_w1);
      const uint32_t t2_addr = smem_u32(smem_t2);
      int my_tiles = 0;
      for (int mt = pair; mt < pair_tiles; mt += num_pairs) ++my_tiles;
      mbar_wait_dbg(wfull, 0, __LINE__);
      tc_fence_after();
      // conv2 of local tile i: 9 taps x 4 k-steps from the resident patch
      auto conv2 = [&](int i) {
        const int slot = i & 1, a = i & 1;
        mbar_wait_dbg(&c2_empty[a], ((i >> 1) & 1) ^ 1, __LINE__);
        mbar_wait_dbg(&pfull[slot], (i >> 1) & 1, __LINE__);
        tc_fence_after();
        const uint32_t patch = smem_u32(smem_patch + slot * kC64PatchStride);
        const uint32_t tmem_d = tmem_base + a * 64;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const uint32_t a0 = patch + ((t / 3) * kC64PatchW + (t % 3)) * 128;
          const uint32_t b0 = w2_addr + t * 4096;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_cg2(tmem_d, umma_smem_desc_sw128(a0 + k * 32, kC64PatchW * 128),
                          umma_smem_desc_sw128(b0 + k * 32, 1024), idesc_c2, (t | k) != 0 ? 1u : 0u);
        }
        umma_commit_cg2(&pempty[slot], 3);
        umma_commit_cg2(&c2_full[a], 3);
      };
      // conv1' partial product of global pass h (tile h / P, pass h % P): two 64-wide K blocks = two ring chunks
      auto gemm2 = [&](int h) {
        const int i = h / P, ps = h - i * P;
        if (ps == 0) {
          mbar_wait_dbg(acc2_empty, (i & 1) ^ 1, __LINE__);
          tc_fence_after();
        }
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          const int q = q_pass(i, ps, c);
          const int b = q % R;
          mbar_wait_dbg(&ychunk_full[b], (q / R) & 1, __LINE__);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_ring + b * kStgChunkBytes);
          const uint32_t b_addr = w1_addr + (2 * ps + c) * (N2 / 2) * 128;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_cg2(tmem_base + kAcc2Col, umma_smem_desc<128>(a_addr + k * 32),
                          umma_smem_desc<128>(b_addr + k * 32), idesc_2, (ps | c | k) != 0 ? 1u : 0u);
          umma_commit_cg2(&stg_empty[b], 3);  // second arrival on the ring buffer (the first is its TMA store)
        }
        if (ps == P - 1) umma_commit_cg2(acc2_full, 3);
      };
      if (my_tiles > 0) conv2(0);
      for (int i = 0; i < my_tiles; ++i) {
        for (int ps = 0; ps < P; ++ps) {
          const int g = i * P + ps;
          const int a1 = g & 1;
          mbar_wait_dbg(&acc1_empty[a1], ((g >> 1) & 1) ^ 1, __LINE__);
          if (ps == 0) mbar_wait_dbg(t2_full, i & 1, __LINE__);  // both CTAs' T2 chunks are in place
          tc_fence_after();
          const uint32_t b_addr = w3_addr + ps * 8192;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_cg2(tmem_base + kAcc1Col + a1 * 128, umma_smem_desc<128>(t2_addr + k * 32),
                          umma_smem_desc<128>(b_addr + k * 32), idesc_1, k != 0 ? 1u : 0u);
          if (ps == P - 1) umma_commit_cg2(t2_empty, 3);  // T2 may be overwritten once these MMAs have read it
          umma_commit_cg2(&acc1_full[a1], 3);
          if (g >= 1) gemm2(g - 1);
          if (ps == 0 && i + 1 < my_tiles) conv2(i + 1);  // next tile's 3x3 runs behind this tile's epilogues
        }
      }
      if (my_tiles > 0) gemm2(my_tiles * P - 1);
    }
  } else {
    // ============================ epilogue (warps 2..9, both CTAs) ============================
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    const bool leader = (threadIdx.x == 64);
    const uint32_t row_off = static_cast<uint32_t>(row) * 128u;
    const uint32_t swz = static_cast<uint32_t>(row & 7);
    const int piece0 = half * 4;
    const uint32_t lane_base = (static_cast<uint32_t>(quarter * 32) << 16) + half * 32;
    // Leader bookkeeping for the ring: hist1 = chunk whose store was committed last, hist0 = the one before.  At
    // the START of every chunk (before any wait on a ring barrier -- the buffer being waited for may be hist0's)
    // the older one is handed back: all committed stores but the newest have finished reading their source.
    int hist0 = -1, hist1 = -1;
    auto release_old = [&]() {
      if (hist0 >= 0) {
        tma_store_wait_read<1>();
        mbar_arrive(&stg_empty[hist0 % R]);
        if (is_out_chunk(hist0)) mbar_arrive(&stg_empty[hist0 % R]);  // T1' chunks have no GEMM2 consumer
        hist0 = -1;
      }
    };
    auto note_store = [&](int q) {
      hist0 = hist1;
      hist1 = q;
    };
    // bias + (residual) + ReLU on 32 accumulator columns -> 64 bytes of a swizzled 128-byte row
    auto finish_row = [&](const uint32_t (&v)[32], const float* bias, const uint4* rv, uint8_t* chunk) {
      const float4* bp = reinterpret_cast<const float4*>(bias);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 b0 = bp[2 * j], b1 = bp[2 * j + 1];
        float x[8];
        x[0] = __uint_as_float(v[8 * j + 0]) + b0.x;
        x[1] = __uint_as_float(v[8 * j + 1]) + b0.y;
        x[2] = __uint_as_float(v[8 * j + 2]) + b0.z;
        x[3] = __uint_as_float(v[8 * j + 3]) + b0.w;
        x[4] = __uint_as_float(v[8 * j + 4]) + b1.x;
        x[5] = __uint_as_float(v[8 * j + 5]) + b1.y;
        x[6] = __uint_as_float(v[8 * j + 6]) + b1.z;
        x[7] = __uint_as_float(v[8 * j + 7]) + b1.w;
        if (rv != nullptr) {
          const uint32_t* r32 = reinterpret_cast<const uint32_t*>(&rv[j]);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            x[2 * e] += bf16_lo(r32[e]);
            x[2 * e + 1] += bf16_hi(r32[e]);
          }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = fmaxf(x[e], 0.f);
        *reinterpret_cast<uint4*>(chunk + (((piece0 + j) ^ swz) << 4)) =
            make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]),
                       pack_bf16x2(x[6], x[7]));
      }
    };
    struct TileXY {
      int w0, h0, n;
    };
    auto tile_xy = [&](int mt) {
      const int tile = mt * 2 + static_cast<int>(rank);
      TileXY t;
      t.w0 = (tile % p.tiles_w) * kC64TileW;
      t.h0 = ((tile / p.tiles_w) % p.tiles_h) * kC64TileH;
      t.n = tile / tiles_img;
      return t;
    };
    // deferred epilogue of conv1': accumulator 2 of local tile ti -> T1'
    auto epilogue2 = [&](int ti, TileXY xy) {
      mbar_wait_dbg(acc2_full, ti & 1, __LINE__);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < kOutChunks; ++c) {
        const int q = q_out(ti, c);
        const int b = q % R;
        uint8_t* chunk = smem_ring + b * kStgChunkBytes + row_off;
        if (leader) release_old();
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + kAcc2Col + c * 64 + lane_base, v);
        if (q >= R) mbar_wait_dbg(&stg_empty[b], ((q / R) - 1) & 1, __LINE__, (N2 << 16) | q);
        tmem_ld_wait();
        if (c == kOutChunks - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_u32(acc2_empty, 0));
        }
        finish_row(v, sbias1 + c * 64 + half * 32, nullptr, chunk);
        fence_proxy_async();
        named_bar_sync(1, kL1EpiThreads);
        if (leader) {
          tma_store_4d(&p.tmOut2, smem_ring + b * kStgChunkBytes, c * 64, xy.w0, xy.h0, xy.n);
          tma_store_commit();
          note_store(q);
          mbar_arrive(&res_full[b]);  // keep every barrier of buffer b at one phase per use
          mbar_arrive_cluster(mapa_u32(&ychunk_full[b], 0));
        }
      }
    };
    int i = 0;
    TileXY prev = {0, 0, 0};
    for (int mt = pair; mt < pair_tiles; mt += num_pairs, ++i) {
      const TileXY xy = tile_xy(mt);
      // ---- conv2 epilogue: accumulator -> T2 chunk (A operand of conv3), never stored to HBM ----
      {
        const int a = i & 1;
        mbar_wait_dbg(&c2_full[a], (i >> 1) & 1, __LINE__);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + a * 64 + lane_base, v);
        mbar_wait_dbg(t2_empty, (i & 1) ^ 1, __LINE__);  // conv3 of the previous tile has read T2
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_u32(&c2_empty[a], 0));
        finish_row(v, sbias2 + half * 32, nullptr, smem_t2 + row_off);
        fence_proxy_async();
        named_bar_sync(1, kL1EpiThreads);
        if (leader) mbar_arrive_cluster(mapa_u32(t2_full, 0));
      }
      for (int ps = 0; ps < P; ++ps) {
        const int g = i * P + ps;
        const int a1 = g & 1;
        mbar_wait_dbg(&acc1_full[a1], (g >> 1) & 1, __LINE__);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          const int q = q_pass(i, ps, c);
          const int b = q % R;
          uint8_t* chunk = smem_ring + b * kStgChunkBytes + row_off;
          if (leader) release_old();
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_base + kAcc1Col + a1 * 128 + c * 64 + lane_base, v);
          mbar_wait_dbg(&res_full[b], (q / R) & 1, __LINE__);
          uint4 rv[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) rv[j] = *reinterpret_cast<const uint4*>(chunk + (((piece0 + j) ^ swz) << 4));
          tmem_ld_wait();
          if (c == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(&acc1_empty[a1], 0));
          }
          finish_row(v, sbias3 + ps * 128 + c * 64 + half * 32, rv, chunk);
          fence_proxy_async();
          named_bar_sync(1, kL1EpiThreads);
          if (leader) {
            tma_store_4d(&p.tmY, smem_ring + b * kStgChunkBytes, ps * 128 + c * 64, xy.w0, xy.h0, xy.n);
            tma_store_commit();
            note_store(q);
            mbar_arrive_cluster(mapa_u32(&ychunk_full[b], 0));
          }
        }
        if (ps == 0 && i > 0) epilogue2(i - 1, prev);
      }
      prev = xy;
    }
    if (i > 0) epilogue2(i - 1, prev);
    if (leader) tma_store_wait_all<0>();
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
