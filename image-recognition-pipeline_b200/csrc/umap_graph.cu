// N4: the graph-construction half of the supervised UMAP at functions/data_curation.py:704-705
// (`umap.UMAP(**umap_params).fit_transform(features_pca, y=y_numeric)`), i.e. what umap-learn 0.5.7 (requirements.txt:175,
// not vendored under /root/reference) computes from the k-NN arrays before its layout optimisation:
//   umap_.py smooth_knn_dist               -> per-sample rho (distance to the nearest neighbour, local_connectivity
//                                             interpolated) and sigma (binary search so that
//                                             sum_j exp(-(d_ij - rho_i) / sigma_i) = log2(k) * bandwidth)
//   umap_.py compute_membership_strengths  -> directed edge weights exp(-(d_ij - rho_i) / sigma_i)
// The k-NN arrays themselves come from irp_knn_graph (lof.cu).  The symmetrisation (fuzzy set union) and everything after
// it (categorical intersection with the labels, spectral initialisation, SGD layout) stay host-side with UMAP.
// Arithmetic is float32 like umap-learn's numba kernels (psum / lo / mid / hi are declared float32 there); one warp per
// sample, the k distances of a row in registers.  oracle: oracle/umap_graph_ref.py (parity unpinned: umap-learn is not
// installable in the build container).
#include <cstdint>

#include "common.h"

namespace irp {

constexpr int kUgMaxK = 160;           // columns of the k-NN arrays (n_neighbors, the sample itself included)
constexpr int kUgSlots = kUgMaxK / 32;  // entries per lane
constexpr float kUgSmoothTol = 1e-5f;   // SMOOTH_K_TOLERANCE
constexpr float kUgMinScale = 1e-3f;    // MIN_K_DIST_SCALE

// mean of all n*k distances, accumulated in fp64 in ONE fixed order (a single block: strided partial sums, then a
// tree), so the MIN_K_DIST_SCALE floor is the same on every run and rank
__global__ void __launch_bounds__(1024) ug_mean_kernel(const float* __restrict__ dist, long long total,
                                                       double* __restrict__ mean_out) {
  __shared__ double part[1024];
  double s = 0.0;
  for (long long i = threadIdx.x; i < total; i += 1024) s += static_cast<double>(dist[i]);
  part[threadIdx.x] = s;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *mean_out = total > 0 ? part[0] / static_cast<double>(total) : 0.0;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__global__ void __launch_bounds__(256) ug_fuzzy_kernel(const int32_t* __restrict__ idx, const float* __restrict__ dist,
                                                       long long n, int k, float local_connectivity, float bandwidth,
                                                       int n_iter, const double* __restrict__ mean_all,
                                                       float* __restrict__ sigma_out, float* __restrict__ rho_out,
                                                       float* __restrict__ vals) {
  const int lane = threadIdx.x & 31;
  const long long row = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  if (row >= n) return;
  const float* dr = dist + row * k;
  float d[kUgSlots];
#pragma unroll
  for (int s = 0; s < kUgSlots; ++s) {
    const int j = lane + 32 * s;
    d[s] = j < k ? dr[j] : 0.f;
  }
  // ---- rho: the local_connectivity-th positive distance of the row (in column order), interpolated ----
  const int index = static_cast<int>(floorf(local_connectivity));
  const float interpolation = local_connectivity - static_cast<float>(index);
  int n_pos = 0;                      // positive entries of the row
  float v_lo = 0.f, v_hi = 0.f;       // non_zero_dists[index - 1], non_zero_dists[index]
  float v_first = 0.f, v_max = 0.f;   // non_zero_dists[0], max(non_zero_dists)
#pragma unroll
  for (int s = 0; s < kUgSlots; ++s) {
    const int j = lane + 32 * s;
    const bool pos = j < k && d[s] > 0.f;
    const unsigned m = __ballot_sync(0xffffffffu, pos);
    const int rank = n_pos + __popc(m & ((1u << lane) - 1u));  // position among the positive entries
    float c_lo = (pos && rank == index - 1) ? d[s] : 0.f;
    float c_hi = (pos && rank == index) ? d[s] : 0.f;
    float c_first = (pos && rank == 0) ? d[s] : 0.f;
    v_lo += warp_sum(c_lo);      // exactly one lane contributes: the sum IS the value
    v_hi += warp_sum(c_hi);
    v_first += warp_sum(c_first);
    v_max = fmaxf(v_max, warp_max(pos ? d[s] : 0.f));
    n_pos += __popc(m);
  }
  float rho = 0.f;
  if (static_cast<float>(n_pos) >= local_connectivity) {
    if (index > 0) {
      rho = v_lo;
      if (interpolation > kUgSmoothTol) rho += interpolation * (v_hi - v_lo);
    } else {
      rho = interpolation * v_first;
    }
  } else if (n_pos > 0) {
    rho = v_max;
  }
  // ---- sigma: binary search on sum_{j >= 1} exp(-(max(d_j - rho, 0)) / sigma) = log2(k) * bandwidth ----
  const float target = log2f(static_cast<float>(k)) * bandwidth;
  float lo = 0.f, hi = INFINITY, mid = 1.f;
  for (int it = 0; it < n_iter; ++it) {
    float psum = 0.f;
#pragma unroll
    for (int s = 0; s < kUgSlots; ++s) {
      const int j = lane + 32 * s;
      if (j >= 1 && j < k) {
        const float dd = d[s] - rho;
        psum += dd > 0.f ? expf(-(dd / mid)) : 1.f;
      }
    }
    psum = warp_sum(psum);
    if (fabsf(psum - target) < kUgSmoothTol) break;
    if (psum > target) {
      hi = mid;
      mid = (lo + hi) / 2.f;
    } else {
      lo = mid;
      if (hi == INFINITY) mid *= 2.f;
      else mid = (lo + hi) / 2.f;
    }
  }
  float sigma = mid;
  if (rho > 0.f) {
    float s_row = 0.f;
#pragma unroll
    for (int s = 0; s < kUgSlots; ++s) s_row += (lane + 32 * s < k) ? d[s] : 0.f;
    const float mean_row = warp_sum(s_row) / static_cast<float>(k);
    if (sigma < kUgMinScale * mean_row) sigma = kUgMinScale * mean_row;
  } else {
    const float mean_g = static_cast<float>(*mean_all);
    if (sigma < kUgMinScale * mean_g) sigma = kUgMinScale * mean_g;
  }
  if (lane == 0) {
    sigma_out[row] = sigma;
    rho_out[row] = rho;
  }
  // ---- membership strengths of the row's directed edges ----
  const int32_t* ir = idx + row * k;
#pragma unroll
  for (int s = 0; s < kUgSlots; ++s) {
    const int j = lane + 32 * s;
    if (j >= k) continue;
    const int32_t nb = ir[j];
    float v;
    if (nb < 0 || nb == row) v = 0.f;                              // missing neighbour / the sample itself
    else if (d[s] - rho <= 0.f || sigma == 0.f) v = 1.f;
    else v = expf(-((d[s] - rho) / sigma));
    vals[row * k + j] = v;
  }
}

}  // namespace irp

using namespace irp;

extern "C" {

int irp_umap_fuzzy_weights(const int32_t* d_idx, const float* d_dist, int64_t n_rows, int k, float local_connectivity,
                           float bandwidth, int n_iter, float* d_sigma, float* d_rho, float* d_vals, void* d_workspace,
                           size_t workspace_bytes, void* stream) {
  IRP_REQUIRE(d_idx && d_dist && d_sigma && d_rho && d_vals && d_workspace, "umap_fuzzy_weights: null argument");
  IRP_REQUIRE(n_rows >= 1 && n_rows < (1ll << 31), "umap_fuzzy_weights: n_rows %lld", static_cast<long long>(n_rows));
  IRP_REQUIRE(k >= 2 && k <= kUgMaxK, "umap_fuzzy_weights: n_neighbors %d not in [2,%d]", k, kUgMaxK);
  IRP_REQUIRE(local_connectivity >= 0.f && bandwidth > 0.f && n_iter >= 1, "umap_fuzzy_weights: bad parameters");
  IRP_REQUIRE(workspace_bytes >= sizeof(double), "umap_fuzzy_weights: workspace of %zu bytes needed", sizeof(double));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* mean_all = static_cast<double*>(d_workspace);
  ug_mean_kernel<<<1, 1024, 0, st>>>(d_dist, static_cast<long long>(n_rows) * k, mean_all);
  const long long threads = static_cast<long long>(n_rows) * 32;
  ug_fuzzy_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, st>>>(
      d_idx, d_dist, n_rows, k, local_connectivity, bandwidth, n_iter, mean_all, d_sigma, d_rho, d_vals);
  IRP_CUDA_OK(cudaGetLastError());
  return IRP_OK;
}

}  // extern "C"
