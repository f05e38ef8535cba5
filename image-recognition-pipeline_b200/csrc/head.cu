// Classifier head of the reference's AnimalClassifier:
//   /root/reference/functions/model.py:29-35   Sequential(Dropout, Linear(2048,512), ReLU, Dropout, Linear(512,C))
//   /root/reference/functions/train.py:208-216 loss = CrossEntropyLoss(outputs, labels); predicted = argmax
// on the pooled fp32 features the trunk produces; eval mode, so the dropouts are identities.
//
// The head is 2 MFLOP per image against the trunk's 8 GFLOP, so it stays in fp32 on the CUDA cores (the reference
// computes it in fp32): a split-K register-tiled SGEMM for Linear(2048,512) whose K-slices are summed in a fixed
// order by the second kernel (deterministic), which also applies bias + ReLU, the 512 -> C layer (one warp per
// image), argmax and the per-image cross-entropy.
#include <cfloat>
#include <cmath>

#include "common.h"

namespace irp {

constexpr int kHeadTile = 64;     // output tile (rows x hidden units) per CTA
constexpr int kHeadKChunk = 16;   // K elements staged per iteration
constexpr int kHeadSplitK = 4;    // K slices (grid.z): 4 x 8 x 4 = 128 CTAs at batch 256
constexpr int kHeadMaxClasses = 64;

// partial[z][m][n] = sum_{k in slice z} x[m][k] * w[n][k]
__global__ void __launch_bounds__(256) head_linear1_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                           int M, int N, int K, float* __restrict__ partial) {
  __shared__ float xs[kHeadKChunk][kHeadTile + 4];
  __shared__ float ws[kHeadKChunk][kHeadTile + 4];
  const int m0 = blockIdx.x * kHeadTile, n0 = blockIdx.y * kHeadTile;
  const int k_per = K / kHeadSplitK;
  const int k_begin = blockIdx.z * k_per, k_end = k_begin + k_per;
  const int tid = threadIdx.x;
  const int lr = tid >> 2, lc = (tid & 3) * 4;  // loader: row 0..63 of the tile, 4 consecutive k
  const int tm = (tid >> 4) * 4, tn = (tid & 15) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const bool xrow_ok = m0 + lr < M, wrow_ok = n0 + lr < N;
  const float* xp = x + static_cast<size_t>(xrow_ok ? m0 + lr : 0) * K + lc;
  const float* wp = w + static_cast<size_t>(wrow_ok ? n0 + lr : 0) * K + lc;
  for (int k0 = k_begin; k0 < k_end; k0 += kHeadKChunk) {
    const float4 xv = xrow_ok ? *reinterpret_cast<const float4*>(xp + k0) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 wv = wrow_ok ? *reinterpret_cast<const float4*>(wp + k0) : make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    xs[lc + 0][lr] = xv.x;
    xs[lc + 1][lr] = xv.y;
    xs[lc + 2][lr] = xv.z;
    xs[lc + 3][lr] = xv.w;
    ws[lc + 0][lr] = wv.x;
    ws[lc + 1][lr] = wv.y;
    ws[lc + 2][lr] = wv.z;
    ws[lc + 3][lr] = wv.w;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kHeadKChunk; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&xs[k][tm]);
      const float4 b = *reinterpret_cast<const float4*>(&ws[k][tn]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }
  float* out = partial + static_cast<size_t>(blockIdx.z) * M * N;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + tm + i;
    if (m >= M) continue;
    if (n0 + tn + 3 < N) {
      *reinterpret_cast<float4*>(out + static_cast<size_t>(m) * N + n0 + tn) =
          make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    } else {
      for (int j = 0; j < 4; ++j)
        if (n0 + tn + j < N) out[static_cast<size_t>(m) * N + n0 + tn + j] = acc[i][j];
    }
  }
}

// one warp per image: hidden = relu(sum_z partial[z] + b1) (kept in shared memory); logits = hidden . w2^T + b2,
// eight classes at a time in registers; argmax (first maximum, like torch.max)
__global__ void __launch_bounds__(256) head_linear2_kernel(const float* __restrict__ partial,
                                                           const float* __restrict__ b1, const float* __restrict__ w2,
                                                           const float* __restrict__ b2, int M, int N, int C,
                                                           float* __restrict__ logits, int32_t* __restrict__ pred) {
  extern __shared__ float sh_hidden[];  // [8][N]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * 8 + warp;
  if (m >= M) return;
  float* hid = sh_hidden + static_cast<size_t>(warp) * N;
  for (int n = lane; n < N; n += 32) {
    float h = b1[n];
#pragma unroll
    for (int z = 0; z < kHeadSplitK; ++z) h += partial[(static_cast<size_t>(z) * M + m) * N + n];
    hid[n] = fmaxf(h, 0.f);
  }
  __syncwarp();
  float best = -FLT_MAX;
  int best_c = 0;
  for (int c0 = 0; c0 < C; c0 += 8) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int n = lane; n < N; n += 32) {
      const float h = hid[n];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (c0 + j < C) acc[j] = fmaf(h, w2[static_cast<size_t>(c0 + j) * N + n], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (c0 + j >= C) break;
      float v = acc[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      v += b2[c0 + j];
      if (lane == 0) logits[static_cast<size_t>(m) * C + c0 + j] = v;
      if (v > best) {
        best = v;
        best_c = c0 + j;
      }
    }
  }
  if (lane == 0 && pred != nullptr) pred[m] = best_c;
}

// stats[0] = sum_i w[y_i] * ce_i, stats[1] = sum_i w[y_i], stats[2] = #(argmax == y)   (single CTA, fixed order)
__global__ void __launch_bounds__(256) cross_entropy_kernel(const float* __restrict__ logits,
                                                            const int64_t* __restrict__ labels, int M, int C,
                                                            const float* __restrict__ class_weights,
                                                            double* __restrict__ stats) {
  __shared__ double s_loss[256], s_w[256], s_ok[256];
  double loss = 0.0, wsum = 0.0, ok = 0.0;
  for (int m = threadIdx.x; m < M; m += 256) {
    const float* row = logits + static_cast<size_t>(m) * C;
    float mx = row[0];
    int arg = 0;
    for (int c = 1; c < C; ++c)
      if (row[c] > mx) {
        mx = row[c];
        arg = c;
      }
    double se = 0.0;
    for (int c = 0; c < C; ++c) se += exp(static_cast<double>(row[c]) - static_cast<double>(mx));
    const int y = static_cast<int>(labels[m]);
    if (y < 0 || y >= C) continue;  // ignore_index-style rows contribute nothing
    const double ce = log(se) + static_cast<double>(mx) - static_cast<double>(row[y]);
    const double w = class_weights != nullptr ? static_cast<double>(class_weights[y]) : 1.0;
    loss += w * ce;
    wsum += w;
    ok += (arg == y) ? 1.0 : 0.0;
  }
  s_loss[threadIdx.x] = loss;
  s_w[threadIdx.x] = wsum;
  s_ok[threadIdx.x] = ok;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s_loss[threadIdx.x] += s_loss[threadIdx.x + o];
      s_w[threadIdx.x] += s_w[threadIdx.x + o];
      s_ok[threadIdx.x] += s_ok[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    stats[0] = s_loss[0];
    stats[1] = s_w[0];
    stats[2] = s_ok[0];
  }
}

}  // namespace irp

using namespace irp;

extern "C" {

size_t irp_classifier_head_workspace_bytes(int batch, int hidden) {
  if (batch <= 0 || hidden <= 0) return 0;
  return static_cast<size_t>(kHeadSplitK) * batch * hidden * sizeof(float) + 256;
}

int irp_classifier_head(const float* d_features, int batch, int in_dim, const float* d_w1, const float* d_b1,
                        int hidden, const float* d_w2, const float* d_b2, int num_classes, float* d_logits,
                        int32_t* d_pred, void* d_workspace, size_t workspace_bytes, void* stream) {
  IRP_REQUIRE(d_features && d_w1 && d_b1 && d_w2 && d_b2 && d_logits && d_workspace, "classifier_head: null argument");
  IRP_REQUIRE(batch > 0 && in_dim > 0 && in_dim % (kHeadSplitK * kHeadKChunk) == 0 && hidden > 0 && hidden % 4 == 0,
              "classifier_head: batch %d in_dim %d hidden %d unsupported (in_dim %% 64 == 0, hidden %% 4 == 0)", batch,
              in_dim, hidden);
  IRP_REQUIRE(num_classes >= 1 && num_classes <= kHeadMaxClasses * 16 && hidden <= 1536,
              "classifier_head: num_classes %d / hidden %d unsupported (<= %d classes, hidden <= 1536)", num_classes,
              hidden, kHeadMaxClasses * 16);
  IRP_REQUIRE(workspace_bytes >= irp_classifier_head_workspace_bytes(batch, hidden),
              "classifier_head: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* partial = reinterpret_cast<float*>((reinterpret_cast<size_t>(d_workspace) + 255) & ~static_cast<size_t>(255));
  const dim3 grid(ceil_div(batch, kHeadTile), ceil_div(hidden, kHeadTile), kHeadSplitK);
  head_linear1_kernel<<<grid, 256, 0, st>>>(d_features, d_w1, batch, hidden, in_dim, partial);
  IRP_CUDA_OK(cudaGetLastError());
  head_linear2_kernel<<<ceil_div(batch, 8), 256, static_cast<size_t>(hidden) * 8 * sizeof(float), st>>>(partial, d_b1, d_w2, d_b2, batch, hidden, num_classes,
                                                          d_logits, d_pred);
  IRP_CUDA_OK(cudaGetLastError());
  return IRP_OK;
}

int irp_cross_entropy_stats(const float* d_logits, const int64_t* d_labels, int batch, int num_classes,
                            const float* d_class_weights, double* d_stats, void* stream) {
  IRP_REQUIRE(d_logits && d_labels && d_stats && batch > 0 && num_classes >= 1, "cross_entropy_stats: bad argument");
  cross_entropy_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_logits, d_labels, batch, num_classes,
                                                                          d_class_weights, d_stats);
  IRP_CUDA_OK(cudaGetLastError());
  return IRP_OK;
}

}  // extern "C"
