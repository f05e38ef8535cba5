// Host side of the ResNet-50 trunk: tile planning + tensor maps for conv_gemm_kernel, BN folding,
// max/avg pooling kernels, the irp_resnet50 handle and the single-conv parity hook.
// Architecture follows torchvision/models/resnet.py:108-160 (Bottleneck, stride on the 3x3), :197-205 (stem),
// :266-282 (_forward_impl) with the fc dropped as in functions/data_curation.py:658.
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "common.h"
#include "conv_params.cuh"
#include "conv_gemm2.cuh"
#include "conv3x3_c64.cuh"
#include "conv_chain.cuh"
#include "conv_pool.cuh"
#include "stem_conv.cuh"

namespace irp {

// ------------------------------------------------------------------------------------------------------------
// small kernels
// ------------------------------------------------------------------------------------------------------------

// Fold eval-mode BatchNorm into the conv: w'[co,r,s,ci] = w[co,ci,r,s] * g/sqrt(v+eps), b' = beta - mean*g/sqrt(v+eps).
// Output layout [Cout][kh][kw_pad][cin_pad] (zero filled where s >= kw or ci >= cin): the K-major B operand.
__global__ void fold_bn_kernel(const float* __restrict__ w, const float* __restrict__ gamma,
                               const float* __restrict__ beta, const float* __restrict__ mean,
                               const float* __restrict__ var, float eps, int cout, int cin, int kh, int kw,
                               int kw_pad, int cin_pad, __nv_bfloat16* __restrict__ w_out,
                               float* __restrict__ bias_out) {
  const int k_pad = kh * kw_pad * cin_pad;
  const long long total = static_cast<long long>(cout) * k_pad;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int co = static_cast<int>(i / k_pad);
    int rem = static_cast<int>(i % k_pad);
    const int ci = rem % cin_pad;
    rem /= cin_pad;
    const int s = rem % kw_pad;
    const int r = rem / kw_pad;
    const float scale = gamma[co] / sqrtf(var[co] + eps);
    float v = 0.f;
    if (ci < cin && s < kw) v = w[((static_cast<size_t>(co) * cin + ci) * kh + r) * kw + s] * scale;
    w_out[i] = __float2bfloat16_rn(v);
    if (i % k_pad == 0) bias_out[co] = beta[co] - mean[co] * scale;
  }
}

__global__ void add_bias_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                                int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + b[i];
}

// 3x3 / stride 2 / pad 1 max pooling, NHWC bf16, 8 channels (16 bytes) per thread.
__global__ void avgpool_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, int B, int HW, int C) {
  const int c2 = C / 2;
  const long long total = static_cast<long long>(B) * c2;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % c2);
    const int n = static_cast<int>(i / c2);
    const uint32_t* p = reinterpret_cast<const uint32_t*>(x + static_cast<size_t>(n) * HW * C) + c;
    float s0 = 0.f, s1 = 0.f;
    for (int j = 0; j < HW; ++j) {
      const uint32_t v = __ldg(p + static_cast<size_t>(j) * c2);
      s0 += bf16_lo(v);
      s1 += bf16_hi(v);
    }
    const float inv = 1.0f / static_cast<float>(HW);
    y[static_cast<size_t>(n) * C + 2 * c] = s0 * inv;
    y[static_cast<size_t>(n) * C + 2 * c + 1] = s1 * inv;
  }
}

// ------------------------------------------------------------------------------------------------------------
// conv planning
// ------------------------------------------------------------------------------------------------------------
enum ConvKind { kConvFlat = 0, kConvSpatial = 1, kConvPatch64 = 3 };

struct ConvPlan {
  ConvParams p;
  int bn_tile = 128;  // BN of the kernel instance
  ConvKind kind = kConvFlat;
  bool has_res = false;
  int H = 0, W = 0, Ho = 0, Wo = 0, cin = 0, cout = 0, ksize = 1, stride = 1;
  int max_batch = 0;
};

// Choose the (bw,bh,bn) output box with bw*bh*bn <= 128 that minimises the number of M tiles.
static void choose_box(int Wo, int Ho, int B, int* bw, int* bh, int* bn) {
  long long best_tiles = -1;
  int best[3] = {1, 1, 1};
  for (int w = 1; w <= Wo && w <= kTileM; ++w) {
    for (int h = 1; h <= Ho && w * h <= kTileM; ++h) {
      int n = kTileM / (w * h);
      if (n > B) n = B;
      if (n < 1) continue;
      const long long tiles = static_cast<long long>(ceil_div(Wo, w)) * ceil_div(Ho, h) * ceil_div(B, n);
      const bool better = best_tiles < 0 || tiles < best_tiles ||
                          (tiles == best_tiles && (w > best[0] || (w == best[0] && h > best[1])));
      if (better) {
        best_tiles = tiles;
        best[0] = w;
        best[1] = h;
        best[2] = n;
      }
    }
  }
  *bw = best[0];
  *bh = best[1];
  *bn = best[2];
}

static inline int floor_div2(int a) { return a >= 0 ? a / 2 : -((-a + 1) / 2); }

// Output / residual tensor maps: same (bw,bh,bn) box as the A operand, 64 channels per box.
static int encode_out_maps(ConvPlan* plan, void* out, const void* residual, int max_batch) {
  ConvParams& p = plan->p;
  const uint64_t C = static_cast<uint64_t>(plan->cout);
  uint64_t dims[4], strides[3];
  if (plan->kind == kConvFlat) {
    const uint64_t M = static_cast<uint64_t>(max_batch) * plan->Ho * plan->Wo;
    dims[0] = C; dims[1] = M; dims[2] = 1; dims[3] = 1;
    strides[0] = C * 2; strides[1] = M * C * 2; strides[2] = M * C * 2;
  } else {
    dims[0] = C; dims[1] = plan->Wo; dims[2] = plan->Ho; dims[3] = max_batch;
    strides[0] = C * 2; strides[1] = static_cast<uint64_t>(plan->Wo) * C * 2;
    strides[2] = static_cast<uint64_t>(plan->Ho) * plan->Wo * C * 2;
  }
  uint32_t box[4] = {64, static_cast<uint32_t>(p.bw), static_cast<uint32_t>(p.bh), static_cast<uint32_t>(p.bn)};
  IRP_TRY(encode_bf16_map(&p.tmOut, out, 4, dims, strides, box, 128));
  plan->has_res = residual != nullptr;
  if (residual) IRP_TRY(encode_bf16_map(&p.tmRes, const_cast<void*>(residual), 4, dims, strides, box, 128));
  if (plan->kind == kConvFlat) {
    uint32_t wbox[4] = {64, 32, 1, 1};  // one epilogue warp's rows (warp-store epilogue)
    IRP_TRY(encode_bf16_map(&p.tmOutW, out, 4, dims, strides, wbox, 128));
  }
  p.cout = plan->cout;
  p.out_box_bytes = p.bw * p.bh * p.bn * 128;
  return IRP_OK;
}

// Tile width of the CTA-pair kernel for a conv with `pair_m_tiles` 256-row tiles: the operand stream per tile is
// proportional to (256 + BN) bytes per unit of K and the tiles run in ceil(tiles / pairs) waves, so pick the BN
// that minimises waves * (256 + BN); ties go to the wider tile (fewer operand bytes per FLOP).
static int choose_bn2(int cout, long long pair_m_tiles) {
  const int pairs = num_sms() / 2;
  int best = 64;
  long long best_cost = -1;
  for (int bn = 64; bn <= 256; bn *= 2) {
    if (cout % bn != 0) continue;
    const long long tiles = pair_m_tiles * (cout / bn);
    const long long cost = ceil_div64(tiles, pairs) * (256 + bn);
    if (best_cost < 0 || cost <= best_cost) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

// Builds tensor maps + static fields. x/w/out pointers are baked into the maps / params.
static int plan_conv(ConvPlan* plan, const void* x, const void* w, const float* bias, const void* residual,
                     void* out, int max_batch, int H, int W, int Cin, int Cout, int ksize, int stride, int relu) {
  IRP_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0, "conv: Cin (%d) and Cout (%d) must be multiples of 64", Cin, Cout);
  IRP_REQUIRE((ksize == 1 || ksize == 3) && (stride == 1 || stride == 2), "conv: unsupported ksize %d stride %d",
              ksize, stride);
  const int pad = ksize == 3 ? 1 : 0;
  const int Ho = (H + 2 * pad - ksize) / stride + 1;
  const int Wo = (W + 2 * pad - ksize) / stride + 1;
  ConvParams& p = plan->p;
  memset(&p, 0, sizeof(p));
  plan->H = H;
  plan->W = W;
  plan->Ho = Ho;
  plan->Wo = Wo;
  plan->cin = Cin;
  plan->cout = Cout;
  plan->ksize = ksize;
  plan->stride = stride;
  plan->max_batch = max_batch;
  plan->bn_tile = (Cout % 128 == 0) ? 128 : 64;
  constexpr int BK = 64;
  p.cin = Cin;
  p.kc_blocks = Cin / BK;
  p.ntaps = ksize * ksize;
  p.bias = bias;
  p.relu = relu;
  p.n_tiles_n = Cout / plan->bn_tile;

  char* xb = static_cast<char*>(const_cast<void*>(x));
  if (ksize == 3 && stride == 1 && Cin == 64 && Cout == 64 && residual == nullptr) {
    // patch-resident kernel (conv3x3_c64.cuh): 8 x 16 output tiles, 10 x 18 input patches, resident weights
    plan->kind = kConvPatch64;
    plan->bn_tile = 64;
    p.n_tiles_n = 1;
    p.bw = kC64TileW;
    p.bh = kC64TileH;
    p.bn = 1;
    uint64_t dims[4] = {64, static_cast<uint64_t>(W), static_cast<uint64_t>(H), static_cast<uint64_t>(max_batch)};
    uint64_t strides[3] = {128, static_cast<uint64_t>(W) * 128, static_cast<uint64_t>(H) * W * 128};
    uint32_t ibox[4] = {64, kC64PatchW, kC64PatchH, 1};
    IRP_TRY(encode_bf16_map(&p.tmA[0], xb, 4, dims, strides, ibox, 128));
    uint64_t wdims[2] = {576, 64};
    uint64_t wstr[1] = {576 * 2};
    uint32_t wbox[2] = {64, 64};
    IRP_TRY(encode_bf16_map(&p.tmB, const_cast<void*>(w), 2, wdims, wstr, wbox, 128));
    IRP_TRY(encode_out_maps(plan, out, nullptr, max_batch));
    return IRP_OK;
  }
  if (ksize == 1 && stride == 1) {
    // flattened: A is the plain [M, Cin] matrix, M = B*H*W
    plan->kind = kConvFlat;
    const uint64_t M = static_cast<uint64_t>(max_batch) * H * W;
    uint64_t dims[4] = {static_cast<uint64_t>(Cin), M, 1, 1};
    uint64_t strides[3] = {static_cast<uint64_t>(Cin) * 2, M * Cin * 2, M * Cin * 2};
    uint32_t box[4] = {BK, kTileM, 1, 1};
    IRP_TRY(encode_bf16_map(&p.tmA[0], xb, 4, dims, strides, box, 128));
    p.bw = kTileM;
    p.bh = 1;
    p.bn = 1;
    p.tap_map[0] = 0;
    p.tap_dw[0] = 0;
    p.tap_dh[0] = 0;
  } else {
    plan->kind = kConvSpatial;
    choose_box(Wo, Ho, max_batch, &p.bw, &p.bh, &p.bn);
    uint32_t box[4] = {BK, static_cast<uint32_t>(p.bw), static_cast<uint32_t>(p.bh), static_cast<uint32_t>(p.bn)};
    if (stride == 1) {
      uint64_t dims[4] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                          static_cast<uint64_t>(max_batch)};
      uint64_t strides[3] = {static_cast<uint64_t>(Cin) * 2, static_cast<uint64_t>(W) * Cin * 2,
                             static_cast<uint64_t>(H) * W * Cin * 2};
      IRP_TRY(encode_bf16_map(&p.tmA[0], xb, 4, dims, strides, box, 128));
      for (int r = 0; r < ksize; ++r)
        for (int s = 0; s < ksize; ++s) {
          const int t = r * ksize + s;
          p.tap_map[t] = 0;
          p.tap_dw[t] = static_cast<int8_t>(s - pad);
          p.tap_dh[t] = static_cast<int8_t>(r - pad);
        }
    } else {
      // four parity views: element (c, ww, hh, n) of view (ph,pw) is x[n, 2*hh+ph, 2*ww+pw, c]
      for (int ph = 0; ph < 2; ++ph)
        for (int pw = 0; pw < 2; ++pw) {
          uint64_t dims[4] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>((W - pw + 1) / 2),
                              static_cast<uint64_t>((H - ph + 1) / 2), static_cast<uint64_t>(max_batch)};
          uint64_t strides[3] = {static_cast<uint64_t>(Cin) * 4, static_cast<uint64_t>(W) * Cin * 4,
                                 static_cast<uint64_t>(H) * W * Cin * 2};
          IRP_TRY(encode_bf16_map(&p.tmA[ph * 2 + pw], xb + (static_cast<size_t>(ph) * W + pw) * Cin * 2, 4, dims,
                                  strides, box, 128));
        }
      for (int r = 0; r < ksize; ++r)
        for (int s = 0; s < ksize; ++s) {
          const int t = r * ksize + s;
          const int dh = r - pad, dw = s - pad;
          p.tap_map[t] = static_cast<int8_t>((dh & 1) * 2 + (dw & 1));
          p.tap_dw[t] = static_cast<int8_t>(floor_div2(dw));
          p.tap_dh[t] = static_cast<int8_t>(floor_div2(dh));
        }
    }
  }
  p.a_box_bytes = p.bw * p.bh * p.bn * BK * 2;
  {
    const long long rows = static_cast<long long>(max_batch) * Ho * Wo;
    const long long m_tiles = plan->kind == kConvFlat
                                  ? ceil_div64(rows, kTileM)
                                  : static_cast<long long>(ceil_div(Wo, p.bw)) * ceil_div(Ho, p.bh) *
                                        ceil_div(max_batch, p.bn);
    plan->bn_tile = choose_bn2(Cout, (m_tiles + 1) / 2);
    p.n_tiles_n = Cout / plan->bn_tile;
  }
  IRP_TRY(encode_out_maps(plan, out, residual, max_batch));
  // weights: [Cout][taps*Cin]
  {
    const uint64_t K = static_cast<uint64_t>(p.ntaps) * Cin;
    uint64_t dims[2] = {K, static_cast<uint64_t>(Cout)};
    uint64_t strides[1] = {K * 2};
    // the CTA-pair kernel loads half of the tile's weight rows per CTA
    uint32_t box[2] = {BK, static_cast<uint32_t>(plan->bn_tile / 2)};
    IRP_TRY(encode_bf16_map(&p.tmB, const_cast<void*>(w), 2, dims, strides, box, 128));
  }
  return IRP_OK;
}

// Launch attributes shared by the trunk's CTA-pair kernels: 2-CTA clusters + programmatic dependent launch (a
// kernel's prologue overlaps its predecessor's tail; every kernel orders its reads with griddepcontrol.wait).
struct PairLaunch {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[2];
  PairLaunch(int grid, int threads, size_t smem, cudaStream_t stream, bool cluster2) {
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    int n = 0;
    if (cluster2) {
      attr[n].id = cudaLaunchAttributeClusterDimension;
      attr[n].val.clusterDim.x = 2;
      attr[n].val.clusterDim.y = 1;
      attr[n].val.clusterDim.z = 1;
      ++n;
    }
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
    cfg.attrs = attr;
    cfg.numAttrs = n;
  }
};

template <int BN, bool RES, bool WS>
static int launch_instance2(const ConvParams& p, cudaStream_t stream) {
  using S = Conv2Smem<BN, RES, WS>;
  auto kernel = conv_gemm2_kernel<BN, RES, WS>;
  IRP_TRY(ensure_smem(kernel, S::kTotalBytes));
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int num_groups = ((m_tiles + 1) / 2) * p.n_tiles_n;
  const int pairs = num_sms() / 2;
  const int grid = 2 * (num_groups < pairs ? num_groups : pairs);
  if (grid <= 0) return IRP_OK;
  PairLaunch l(grid, kConv2Threads, S::kTotalBytes, stream, true);
  IRP_CUDA_OK(cudaLaunchKernelEx(&l.cfg, kernel, p));
  return IRP_OK;
}

template <bool RES, bool WS>
static int launch_conv2(const ConvParams& p, int bn, cudaStream_t stream) {
  switch (bn) {
    case 256: return launch_instance2<256, RES, WS>(p, stream);
    case 128: return launch_instance2<128, RES, WS>(p, stream);
    default: return launch_instance2<64, RES, WS>(p, stream);
  }
}

// Launch a planned conv on `batch` images (batch <= plan->max_batch).
static int launch_conv(const ConvPlan& plan, int batch, cudaStream_t stream) {
  ConvParams p = plan.p;
  if (plan.kind == kConvFlat) {
    const long long M = static_cast<long long>(batch) * plan.Ho * plan.Wo;
    p.tiles_w = static_cast<int>(ceil_div64(M, kTileM));
    p.tiles_h = 1;
    p.tiles_n = 1;
  } else {
    p.tiles_w = ceil_div(plan.Wo, p.bw);
    p.tiles_h = ceil_div(plan.Ho, p.bh);
    p.tiles_n = ceil_div(batch, p.bn);
  }
  p.num_tiles = p.tiles_w * p.tiles_h * p.tiles_n * p.n_tiles_n;
  if (plan.kind == kConvPatch64) {
    IRP_TRY(ensure_smem(conv3x3_c64_kernel, kC64SmemBytes));
    const int tiles = p.tiles_w * p.tiles_h * p.tiles_n;
    const int grid = tiles < num_sms() ? tiles : num_sms();
    if (grid <= 0) return IRP_OK;
    PairLaunch l(grid, kC64Threads, kC64SmemBytes, stream, false);
    IRP_CUDA_OK(cudaLaunchKernelEx(&l.cfg, conv3x3_c64_kernel, p));
    return IRP_OK;
  }
  // 1x1 stride 1 with narrow tiles: barrier-free warp-store epilogue (measured: 40 -> 36 us on layer1's first conv1).
  // At BN = 256 it loses (46 -> 51 us on layer3's conv3 + residual): the whole-bias staging costs a pipeline stage
  // and those launches are bound by L2 <-> SM traffic, not by the epilogue's barrier.
  if (plan.kind == kConvFlat && plan.bn_tile < 256)
    return plan.has_res ? launch_conv2<true, true>(p, plan.bn_tile, stream)
                        : launch_conv2<false, true>(p, plan.bn_tile, stream);
  return plan.has_res ? launch_conv2<true, false>(p, plan.bn_tile, stream)
                      : launch_conv2<false, false>(p, plan.bn_tile, stream);
}

// conv3 of one bottleneck chained with conv1 of the next (conv_chain.cuh)
struct ChainPlan {
  ChainParams p;
  int n2 = 0;
  int rows_per_image = 0;
  bool valid = false;
};

static bool chain_supported(int K1, int N1, int N2) {
  return K1 % 64 == 0 && N1 % kChainBN1 == 0 && N1 <= kChainMaxN1 && (N2 == 64 || N2 == 128 || N2 == 256);
}

// x2 / K2: optional second GEMM1 operand (stride-1 1x1 shortcut convolution of x2 folded into the same accumulator);
// then w3 is [N1][K1 + K2] (conv3 and shortcut weights concatenated along K), b3 the sum of both biases, and
// `residual` must be null.
static int plan_chain(ChainPlan* plan, const void* t2, const void* w3, const float* b3, const void* residual, void* y,
                      const void* w1, const float* b1, void* t1, long long max_rows, int rows_per_image, int K1,
                      int N1, int N2, const void* x2 = nullptr, int K2 = 0) {
  IRP_REQUIRE(chain_supported(K1, N1, N2), "conv chain: unsupported shape K1 %d N1 %d N2 %d", K1, N1, N2);
  IRP_REQUIRE((x2 == nullptr) == (K2 == 0) && K2 % 64 == 0 && (x2 == nullptr) != (residual == nullptr),
              "conv chain: either a residual or a second operand (K2 %d)", K2);
  ChainParams& p = plan->p;
  memset(&p, 0, sizeof(p));
  const uint64_t M = static_cast<uint64_t>(max_rows);
  auto map2d = [&](CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint32_t box_outer) -> int {
    uint64_t dims[2] = {inner, outer};
    uint64_t strides[1] = {inner * 2};
    uint32_t box[2] = {64, box_outer};
    return encode_bf16_map(m, const_cast<void*>(base), 2, dims, strides, box, 128);
  };
  IRP_TRY(map2d(&p.tmA, t2, K1, M, kTileM));
  IRP_TRY(map2d(&p.tmB1, w3, K1 + K2, N1, kChainBN1 / 2));
  IRP_TRY(map2d(&p.tmRes, residual ? residual : y, N1, M, kTileM));
  IRP_TRY(map2d(&p.tmA2, x2 ? x2 : t2, x2 ? K2 : K1, M, kTileM));
  p.k2_blocks = K2 / 64;
  p.has_res = residual != nullptr ? 1 : 0;
  IRP_TRY(map2d(&p.tmYW, y, N1, M, 32));
  IRP_TRY(map2d(&p.tmB2, w1, N1, N2, N2 / 2));
  IRP_TRY(map2d(&p.tmOut2W, t1, N2, M, 32));
  p.bias1 = b3;
  p.bias2 = b1;
  p.k1_blocks = K1 / 64;
  p.passes = N1 / kChainBN1;
  p.n1 = N1;
  plan->n2 = N2;
  plan->rows_per_image = rows_per_image;
  plan->valid = true;
  return IRP_OK;
}

template <int N2>
static int launch_chain_instance(const ChainParams& p, cudaStream_t stream) {
  using S = ChainSmem<N2>;
  auto kernel = conv_chain_kernel<N2>;
  IRP_TRY(ensure_smem(kernel, S::kTotalBytes));
  const int pair_tiles = (p.m_tiles + 1) / 2;
  const int pairs = num_sms() / 2;
  const int grid = 2 * (pair_tiles < pairs ? pair_tiles : pairs);
  if (grid <= 0) return IRP_OK;
  PairLaunch l(grid, kChainThreads, S::kTotalBytes, stream, true);
  IRP_CUDA_OK(cudaLaunchKernelEx(&l.cfg, kernel, p));
  return IRP_OK;
}

static int launch_chain(const ChainPlan& plan, long long rows, cudaStream_t stream) {
  ChainParams p = plan.p;
  p.m_tiles = static_cast<int>(ceil_div64(rows, kTileM));
  switch (plan.n2) {
    case 64: return launch_chain_instance<64>(p, stream);
    case 128: return launch_chain_instance<128>(p, stream);
    default: return launch_chain_instance<256>(p, stream);
  }
}

// the last block's conv3 + residual + ReLU + global average pool (conv_pool.cuh)
struct PoolPlan {
  ConvPoolParams p;
  bool valid = false;
};

static bool pool_supported(int K, int Cout, int hw) { return K == kCpK && Cout % 128 == 0 && hw == kCpPix; }

static int plan_pool(PoolPlan* plan, const void* t2, const void* w, const float* bias, const void* residual,
                     int max_batch, int K, int Cout) {
  IRP_REQUIRE(pool_supported(K, Cout, kCpPix) && residual != nullptr, "conv + pool: unsupported shape K %d Cout %d", K, Cout);
  ConvPoolParams& p = plan->p;
  memset(&p, 0, sizeof(p));
  const uint64_t M = static_cast<uint64_t>(max_batch) * kCpPix;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(Cout)};
    uint64_t strides[1] = {static_cast<uint64_t>(K) * 2};
    uint32_t box[2] = {64, 128};
    IRP_TRY(encode_bf16_map(&p.tmW, const_cast<void*>(w), 2, dims, strides, box, 128));
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K), M};
    uint64_t strides[1] = {static_cast<uint64_t>(K) * 2};
    uint32_t box[2] = {64, kCpRows};
    IRP_TRY(encode_bf16_map(&p.tmX, const_cast<void*>(t2), 2, dims, strides, box, 128));
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(Cout), M};
    uint64_t strides[1] = {static_cast<uint64_t>(Cout) * 2};
    uint32_t box[2] = {64, kCpRows};
    IRP_TRY(encode_bf16_map(&p.tmRes, const_cast<void*>(residual), 2, dims, strides, box, 0));
  }
  p.bias = bias;
  p.cout = Cout;
  p.n_ctiles = Cout / 128;
  plan->valid = true;
  return IRP_OK;
}

static int launch_pool(const PoolPlan& plan, int batch, float* d_out, cudaStream_t stream) {
  ConvPoolParams p = plan.p;
  p.out = d_out;
  p.batch = batch;
  p.n_groups = ceil_div(batch, kCpImgs);
  IRP_TRY(ensure_smem(conv_pool_kernel, kCpSmemBytes));
  int per_tile = num_sms() / p.n_ctiles;  // CTAs that share one channel tile's image groups
  if (per_tile < 1) per_tile = 1;
  if (per_tile > p.n_groups) per_tile = p.n_groups;
  PairLaunch l(p.n_ctiles * per_tile, kCpThreads, kCpSmemBytes, stream, false);
  IRP_CUDA_OK(cudaLaunchKernelEx(&l.cfg, conv_pool_kernel, p));
  return IRP_OK;
}

static int grid_for(long long total, int threads) {
  long long g = (total + threads - 1) / threads;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  return static_cast<int>(g < cap ? (g < 1 ? 1 : g) : cap);
}

}  // namespace irp

// ------------------------------------------------------------------------------------------------------------
// irp_resnet50 handle
// ------------------------------------------------------------------------------------------------------------
using namespace irp;

namespace {

struct ConvSpec {
  int cin, cout, ksize, stride;
  int H, W;        // input spatial size
  int block;       // bottleneck index (-1 for the stem)
  int role;        // 0 stem, 1 conv1, 2 conv2, 3 conv3, 4 downsample
};

// execution-independent index order: conv1, then per block conv1, conv2, conv3[, downsample]
static std::vector<ConvSpec> build_specs() {
  std::vector<ConvSpec> v;
  v.push_back({3, 64, 7, 2, 224, 224, -1, 0});
  const int planes[4] = {64, 128, 256, 512};
  const int blocks[4] = {3, 4, 6, 3};
  int inplanes = 64, hw = 56, blk = 0;
  for (int l = 0; l < 4; ++l) {
    for (int b = 0; b < blocks[l]; ++b, ++blk) {
      const int stride = (b == 0 && l > 0) ? 2 : 1;
      const int width = planes[l];
      v.push_back({inplanes, width, 1, 1, hw, hw, blk, 1});
      v.push_back({width, width, 3, stride, hw, hw, blk, 2});
      const int hw_out = hw / stride;
      v.push_back({width, width * 4, 1, 1, hw_out, hw_out, blk, 3});
      if (b == 0) v.push_back({inplanes, width * 4, 1, stride, hw, hw, blk, 4});
      inplanes = width * 4;
      hw = hw_out;
    }
  }
  return v;
}

static const std::vector<ConvSpec>& specs() {
  static const std::vector<ConvSpec> s = build_specs();
  return s;
}

}  // namespace

struct irp_resnet50 {
  StemPoolParams stem;                 // stem + max pool fused (stem_pool_kernel)
  __nv_bfloat16* stem_w = nullptr;     // its weights (core-matrix order, stem_fold_kernel)
  int max_batch = 0;
  std::vector<ConvPlan> plans;
  std::vector<ChainPlan> chains;     // indexed by the conv3 of the first block of a fused junction
  std::vector<ChainPlan> chains_ds;  // same junction with the block's stride-1 shortcut conv folded into GEMM1
  PoolPlan pool;                     // the last block's conv3 + residual + ReLU + global average pool
  int cat_c3 = -1, cat_ds = -1;      // conv3 / shortcut conv whose folded weights are also kept concatenated along K
  __nv_bfloat16* wcat = nullptr;     // [cout][cin_c3 + cin_ds]
  float* bcat = nullptr;             // bias_c3 + bias_ds
  std::vector<__nv_bfloat16*> weights;
  std::vector<float*> biases;
  std::vector<int> out_buf;  // arena buffer id holding each conv's output
  // arena
  __nv_bfloat16* buf[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // A, B, T1, T2, DS
  const void* planned_input = nullptr;
  bool planned = false;
};

// Junctions conv3 (+ residual) -> next conv1 are chained in one kernel up to this many conv3 output channels:
// layer1 and layer2.  Layer3 (N1 = 1024) was measured slower chained (eight 128-column passes per tile: the second
// GEMM's weight stream and ring traffic cost more than the 104 MB re-read they save).
constexpr int kChainMaxCout = 512;

// Everything in the plan that depends on the INPUT pointer (the stem's tensor map): re-encoded alone when a call
// brings a different input buffer.
static int plan_stem_input(irp_resnet50* net, const void* d_x) {
  enum { A = 0 };
  StemPoolParams& sp = net->stem;
  memset(&sp, 0, sizeof(sp));
  constexpr uint64_t P = IRP_PAD_HW;
  uint64_t idims[3] = {P * 4, P, static_cast<uint64_t>(net->max_batch)};
  uint64_t istr[2] = {P * 8, P * P * 8};
  uint32_t ibox[3] = {kSpPatchW * 4, kSpPatchH, 1};
  IRP_TRY(encode_bf16_map(&sp.tmIn, const_cast<void*>(d_x), 3, idims, istr, ibox, 0));
  sp.weights = net->stem_w;
  sp.bias = net->biases[0];
  sp.out = net->buf[A];
  net->planned_input = d_x;
  return IRP_OK;
}

static int resnet50_plan(irp_resnet50* net, const void* d_x) {
  const auto& sp = specs();
  const int B = net->max_batch;
  enum { A = 0, Bb = 1, T1 = 2, T2 = 3, DS = 4 };
  IRP_TRY(plan_stem_input(net, d_x));
  net->out_buf[0] = A;  // conv 0 + max pool: the pooled stem output
  int cur = A, other = Bb;
  size_t i = 1;
  while (i < sp.size()) {
    const bool has_ds = (i + 3 < sp.size()) && sp[i + 3].role == 4;
    const ConvSpec &c1 = sp[i], &c2 = sp[i + 1], &c3 = sp[i + 2];
    IRP_TRY(plan_conv(&net->plans[i], net->buf[cur], net->weights[i], net->biases[i], nullptr, net->buf[T1], B, c1.H,
                      c1.W, c1.cin, c1.cout, 1, 1, 1));
    net->out_buf[i] = T1;
    IRP_TRY(plan_conv(&net->plans[i + 1], net->buf[T1], net->weights[i + 1], net->biases[i + 1], nullptr,
                      net->buf[T2], B, c2.H, c2.W, c2.cin, c2.cout, 3, c2.stride, 1));
    net->out_buf[i + 1] = T2;
    const __nv_bfloat16* res = net->buf[cur];
    if (has_ds) {
      const ConvSpec& d = sp[i + 3];
      IRP_TRY(plan_conv(&net->plans[i + 3], net->buf[cur], net->weights[i + 3], net->biases[i + 3], nullptr,
                        net->buf[DS], B, d.H, d.W, d.cin, d.cout, 1, d.stride, 0));
      net->out_buf[i + 3] = DS;
      res = net->buf[DS];
    }
    IRP_TRY(plan_conv(&net->plans[i + 2], net->buf[T2], net->weights[i + 2], net->biases[i + 2], res,
                      net->buf[other], B, c3.H, c3.W, c3.cin, c3.cout, 1, 1, 1));
    net->out_buf[i + 2] = other;
    {
      // junction with the next bottleneck: its conv1 consumes this block's output at the same resolution
      const size_t nxt = i + (has_ds ? 4 : 3);
      net->chains[i + 2].valid = false;
      if (nxt < sp.size() && c3.cout <= kChainMaxCout && chain_supported(c3.cin, c3.cout, sp[nxt].cout)) {
        const ConvSpec& n1 = sp[nxt];
        IRP_TRY(plan_chain(&net->chains[i + 2], net->buf[T2], net->weights[i + 2], net->biases[i + 2], res,
                           net->buf[other], net->weights[nxt], net->biases[nxt], net->buf[T1],
                           static_cast<long long>(B) * c3.H * c3.W, c3.H * c3.W, c3.cin, c3.cout, n1.cout));
      }
      // the same junction with the stride-1 shortcut conv computed inside GEMM1 (no DS tensor, no residual read)
      net->chains_ds[i + 2].valid = false;
      if (net->chains[i + 2].valid && has_ds && static_cast<int>(i + 3) == net->cat_ds && sp[i + 3].cin % 64 == 0) {
        IRP_TRY(plan_chain(&net->chains_ds[i + 2], net->buf[T2], net->wcat, net->bcat, nullptr, net->buf[other],
                           net->weights[nxt], net->biases[nxt], net->buf[T1],
                           static_cast<long long>(B) * c3.H * c3.W, c3.H * c3.W, c3.cin, c3.cout, sp[nxt].cout,
                           net->buf[cur], sp[i + 3].cin));
      }
    }
    if (i + (has_ds ? 4 : 3) >= sp.size() && pool_supported(c3.cin, c3.cout, c3.H * c3.W))
      IRP_TRY(plan_pool(&net->pool, net->buf[T2], net->weights[i + 2], net->biases[i + 2], res, B, c3.cin, c3.cout));
    const int t = cur;
    cur = other;
    other = t;
    i += has_ds ? 4 : 3;
  }
  net->planned_input = d_x;
  net->planned = true;
  return IRP_OK;
}

extern "C" {

int irp_resnet50_conv_shape(int index, int* cout, int* cin, int* kh, int* kw, int* stride) {
  const auto& sp = specs();
  IRP_REQUIRE(index >= 0 && index < static_cast<int>(sp.size()), "conv index %d out of range", index);
  if (cout) *cout = sp[index].cout;
  if (cin) *cin = sp[index].cin;
  if (kh) *kh = sp[index].ksize;
  if (kw) *kw = sp[index].ksize;
  if (stride) *stride = sp[index].stride;
  return IRP_OK;
}

int irp_resnet50_create(irp_resnet50** out, int max_batch) {
  IRP_REQUIRE(out != nullptr && max_batch > 0, "irp_resnet50_create: bad arguments");
  IRP_REQUIRE(max_batch <= 4096, "irp_resnet50_create: max_batch %d > 4096", max_batch);
  irp_resnet50* net = new (std::nothrow) irp_resnet50();
  if (!net) return IRP_ERR_NOMEM;
  const auto& sp = specs();
  net->max_batch = max_batch;
  net->plans.resize(sp.size());
  net->chains.resize(sp.size());
  net->chains_ds.resize(sp.size());
  for (size_t i = 0; i < sp.size(); ++i)
    if (sp[i].role == 4 && sp[i].stride == 1 && net->cat_ds < 0) {  // conv order in a block: conv1, conv2, conv3, ds
      net->cat_ds = static_cast<int>(i);
      net->cat_c3 = static_cast<int>(i) - 1;
    }
  net->weights.assign(sp.size(), nullptr);
  net->biases.assign(sp.size(), nullptr);
  net->out_buf.assign(sp.size(), -1);
  // activation arena (NHWC bf16 elements per image): A, B (block outputs, ping-pong), T1, T2, DS
  const size_t per_img[5] = {56 * 56 * 256, 56 * 56 * 256, 56 * 56 * 128, 56 * 56 * 64, 56 * 56 * 256};
  cudaError_t e = cudaSuccess;
  for (int i = 0; i < 5 && e == cudaSuccess; ++i)
    e = cudaMalloc(reinterpret_cast<void**>(&net->buf[i]), per_img[i] * max_batch * sizeof(__nv_bfloat16));
  for (size_t i = 1; i < sp.size() && e == cudaSuccess; ++i) {
    e = cudaMalloc(reinterpret_cast<void**>(&net->weights[i]),
                   static_cast<size_t>(sp[i].cout) * sp[i].ksize * sp[i].ksize * sp[i].cin * sizeof(__nv_bfloat16));
  }
  for (size_t i = 0; i < sp.size() && e == cudaSuccess; ++i)
    e = cudaMalloc(reinterpret_cast<void**>(&net->biases[i]), sp[i].cout * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&net->stem_w), kStemWeightBytes);
  if (e == cudaSuccess && net->cat_ds >= 0) {
    const ConvSpec &c3 = sp[net->cat_c3], &d = sp[net->cat_ds];
    e = cudaMalloc(reinterpret_cast<void**>(&net->wcat),
                   static_cast<size_t>(c3.cout) * (c3.cin + d.cin) * sizeof(__nv_bfloat16));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&net->bcat), c3.cout * sizeof(float));
    if (e == cudaSuccess) e = cudaMemset(net->biases[net->cat_c3], 0, c3.cout * sizeof(float));
    if (e == cudaSuccess) e = cudaMemset(net->biases[net->cat_ds], 0, d.cout * sizeof(float));
  }
  if (e != cudaSuccess) {
    set_last_error("irp_resnet50_create: cudaMalloc failed: %s", cudaGetErrorString(e));
    irp_resnet50_destroy(net);
    return IRP_ERR_NOMEM;
  }
  *out = net;
  return IRP_OK;
}

void irp_resnet50_destroy(irp_resnet50* net) {
  if (!net) return;
  for (auto& b : net->buf) cudaFree(b);
  cudaFree(net->stem_w);
  cudaFree(net->wcat);
  cudaFree(net->bcat);
  for (auto* w : net->weights) cudaFree(w);
  for (auto* b : net->biases) cudaFree(b);
  delete net;
}

int irp_resnet50_load_conv(irp_resnet50* net, int index, const float* d_weight_oihw, const float* d_gamma,
                           const float* d_beta, const float* d_mean, const float* d_var, float eps, void* stream) {
  const auto& sp = specs();
  IRP_REQUIRE(net != nullptr && index >= 0 && index < static_cast<int>(sp.size()), "load_conv: bad index %d", index);
  const ConvSpec& s = sp[index];
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (s.role == 0) {
    stem_fold_kernel<<<grid_for(kStemWeightBytes / 2, 256), 256, 0, st>>>(d_weight_oihw, d_gamma, d_beta, d_mean,
                                                                          d_var, eps, net->stem_w, net->biases[0]);
    IRP_CUDA_OK(cudaGetLastError());
    return IRP_OK;
  }
  const long long total = static_cast<long long>(s.cout) * s.ksize * s.ksize * s.cin;
  fold_bn_kernel<<<grid_for(total, 256), 256, 0, st>>>(d_weight_oihw, d_gamma, d_beta, d_mean, d_var, eps, s.cout,
                                                       s.cin, s.ksize, s.ksize, s.ksize, s.cin, net->weights[index],
                                                       net->biases[index]);
  IRP_CUDA_OK(cudaGetLastError());
  if (index == net->cat_c3 || index == net->cat_ds) {
    // keep [W3 | Wds] and b3 + bds current for the junction kernel that folds the shortcut conv into its GEMM1
    const ConvSpec &c3 = sp[net->cat_c3], &d = sp[net->cat_ds];
    const size_t pitch = static_cast<size_t>(c3.cin + d.cin) * sizeof(__nv_bfloat16);
    const size_t col0 = index == net->cat_c3 ? 0 : static_cast<size_t>(c3.cin);
    IRP_CUDA_OK(cudaMemcpy2DAsync(net->wcat + col0, pitch, net->weights[index], s.cin * sizeof(__nv_bfloat16),
                                  s.cin * sizeof(__nv_bfloat16), s.cout, cudaMemcpyDeviceToDevice, st));
    add_bias_kernel<<<grid_for(c3.cout, 256), 256, 0, st>>>(net->biases[net->cat_c3], net->biases[net->cat_ds],
                                                           net->bcat, c3.cout);
    IRP_CUDA_OK(cudaGetLastError());
  }
  return IRP_OK;
}

// One pass of `batch` images through the trunk.  capture_index >= 0 (with d_capture): also copy that convolution's
// output tensor; every conv then runs as its own launch (and the last block's activation is materialised and pooled
// by avgpool_kernel instead of the fused conv + pool kernel) (the junction kernel that folds layer1's shortcut conv into
// its accumulator never materialises the shortcut tensor), and index 0 yields the stem output AFTER the fused
// 3x3/2 max pool, [batch,56,56,64].
static int resnet50_forward(irp_resnet50* net, const void* d_x, int batch, float* d_embed, int capture_index,
                            void* d_capture, size_t capture_capacity, cudaStream_t st) {
  IRP_REQUIRE(net != nullptr && d_x != nullptr && d_embed != nullptr, "embed: null argument");
  IRP_REQUIRE(batch > 0 && batch <= net->max_batch, "embed: batch %d not in [1,%d]", batch, net->max_batch);
  if (!net->planned)
    IRP_TRY(resnet50_plan(net, d_x));
  else if (net->planned_input != d_x)
    IRP_TRY(plan_stem_input(net, d_x));
  const auto& sp = specs();
  auto capture = [&](int idx) -> int {
    if (idx != capture_index || d_capture == nullptr) return IRP_OK;
    const ConvSpec& s = sp[idx];
    const int ho = s.role == 0 ? 56 : s.H / s.stride, wo = s.role == 0 ? 56 : s.W / s.stride;
    const size_t per_img = static_cast<size_t>(ho) * wo * s.cout;
    IRP_REQUIRE(per_img * batch <= capture_capacity, "capture buffer too small: need %zu elements", per_img * batch);
    IRP_CUDA_OK(cudaMemcpyAsync(d_capture, net->buf[net->out_buf[idx]], per_img * batch * sizeof(__nv_bfloat16),
                                cudaMemcpyDeviceToDevice, st));
    return IRP_OK;
  };
  {
    StemPoolParams sp3 = net->stem;
    sp3.n_base = 0;
    sp3.num_tiles = batch * 64;
    IRP_TRY(ensure_smem(stem_pool_kernel, kSpSmemBytes));
    const int grid = sp3.num_tiles < num_sms() ? sp3.num_tiles : num_sms();
    stem_pool_kernel<<<grid, kSpThreads, kSpSmemBytes, st>>>(sp3);
    IRP_CUDA_OK(cudaGetLastError());
  }
  IRP_TRY(capture(0));
  size_t i = 1;
  int last = 0;
  bool conv1_done = false;  // this block's conv1 was already produced by the previous block's chained conv3
  while (i < sp.size()) {
    const bool has_ds = (i + 3 < sp.size()) && sp[i + 3].role == 4;
    if (!conv1_done) IRP_TRY(launch_conv(net->plans[i], batch, st));
    IRP_TRY(capture(static_cast<int>(i)));
    IRP_TRY(launch_conv(net->plans[i + 1], batch, st));
    IRP_TRY(capture(static_cast<int>(i + 1)));
    const bool ds_folded = has_ds && net->chains_ds[i + 2].valid && d_capture == nullptr;
    if (has_ds && !ds_folded) {
      IRP_TRY(launch_conv(net->plans[i + 3], batch, st));
      IRP_TRY(capture(static_cast<int>(i + 3)));
    }
    const ChainPlan& ch = ds_folded ? net->chains_ds[i + 2] : net->chains[i + 2];
    const bool is_last = i + (has_ds ? 4 : 3) >= sp.size();
    if (is_last && net->pool.valid && d_capture == nullptr) {
      // conv3 + residual + ReLU + global average pool in one kernel: the embedding is written directly
      IRP_TRY(launch_pool(net->pool, batch, d_embed, st));
      return IRP_OK;
    }
    if (ch.valid) {
      IRP_TRY(launch_chain(ch, static_cast<long long>(batch) * ch.rows_per_image, st));
      conv1_done = true;
    } else {
      IRP_TRY(launch_conv(net->plans[i + 2], batch, st));
      conv1_done = false;
    }
    IRP_TRY(capture(static_cast<int>(i + 2)));
    last = static_cast<int>(i + 2);
    i += has_ds ? 4 : 3;
  }
  const long long total = static_cast<long long>(batch) * (2048 / 2);
  avgpool_kernel<<<grid_for(total, 128), 128, 0, st>>>(net->buf[net->out_buf[last]], d_embed, batch, 49, 2048);
  IRP_CUDA_OK(cudaGetLastError());
  return IRP_OK;
}

int irp_resnet50_embed(irp_resnet50* net, const void* d_x_nhwc4p, int batch, float* d_embed, void* stream) {
  return resnet50_forward(net, d_x_nhwc4p, batch, d_embed, -1, nullptr, 0, static_cast<cudaStream_t>(stream));
}

int irp_resnet50_embed_capture(irp_resnet50* net, const void* d_x_nhwc4p, int batch, float* d_embed,
                               int capture_index, void* d_capture_bf16, size_t capacity_elems, void* stream) {
  return resnet50_forward(net, d_x_nhwc4p, batch, d_embed, capture_index, d_capture_bf16, capacity_elems,
                          static_cast<cudaStream_t>(stream));
}

int irp_conv1x1_chain(const void* d_t2, const void* d_w3, const float* d_b3, const void* d_residual, void* d_y,
                      const void* d_w1, const float* d_b1, void* d_t1, int64_t rows, int K1, int N1, int N2,
                      void* stream) {
  IRP_REQUIRE(d_t2 && d_w3 && d_b3 && d_residual && d_y && d_w1 && d_b1 && d_t1 && rows > 0, "conv chain: bad argument");
  ChainPlan plan;
  IRP_TRY(plan_chain(&plan, d_t2, d_w3, d_b3, d_residual, d_y, d_w1, d_b1, d_t1, rows, 1, K1, N1, N2));
  return launch_chain(plan, rows, static_cast<cudaStream_t>(stream));
}

int irp_conv1x1_chain_ds(const void* d_t2, const void* d_x, const void* d_wcat, const float* d_bias, void* d_y,
                         const void* d_w1, const float* d_b1, void* d_t1, int64_t rows, int K1, int K2, int N1, int N2,
                         void* stream) {
  IRP_REQUIRE(d_t2 && d_x && d_wcat && d_bias && d_y && d_w1 && d_b1 && d_t1 && rows > 0 && K2 > 0,
              "conv chain ds: bad argument");
  ChainPlan plan;
  IRP_TRY(plan_chain(&plan, d_t2, d_wcat, d_bias, nullptr, d_y, d_w1, d_b1, d_t1, rows, 1, K1, N1, N2, d_x, K2));
  return launch_chain(plan, rows, static_cast<cudaStream_t>(stream));
}

int irp_conv1x1_pool(const void* d_t2, const void* d_w, const float* d_bias, const void* d_residual, float* d_out,
                     int batch, int K, int Cout, void* stream) {
  IRP_REQUIRE(d_t2 && d_w && d_bias && d_residual && d_out && batch > 0, "conv + pool: bad argument");
  PoolPlan plan;
  IRP_TRY(plan_pool(&plan, d_t2, d_w, d_bias, d_residual, batch, K, Cout));
  return launch_pool(plan, batch, d_out, static_cast<cudaStream_t>(stream));
}

int irp_conv2d_nhwc(const void* d_x, const void* d_w, const float* d_bias, const void* d_residual, void* d_out,
                    int B, int H, int W, int Cin, int Cout, int ksize, int stride, int relu, void* stream) {
  IRP_REQUIRE(d_x && d_w && d_bias && d_out && B > 0 && H > 0 && W > 0, "conv2d: bad arguments");
  ConvPlan plan;
  IRP_TRY(plan_conv(&plan, d_x, d_w, d_bias, d_residual, d_out, B, H, W, Cin, Cout, ksize, stride, relu));
  return launch_conv(plan, B, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
