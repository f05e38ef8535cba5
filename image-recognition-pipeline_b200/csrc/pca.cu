// A3 PCA: shifted scatter matrix on tensor cores (split-bf16 x3, fp32 TMEM accumulate per K chunk, fp64 combine),
// fp64 symmetric eigensolver for the top-k pairs (Householder tridiagonalisation with the matrix resident in L2,
// multisection Sturm bisection, inverse iteration, reflector back-transform), sklearn's sign convention, and the
// fp64-accumulated projection.
//
// Replaces PCA(n_components).fit_transform at functions/data_curation.py:700-701; mirrors the exact solver
// sklearn offers for the same class: sklearn/decomposition/_pca.py:587-640 (covariance_eigh), _base.py:151-159
// (transform), utils/extmath.py:973-981 (svd_flip, u_based_decision=False).  oracle/pca_ref.py is the fp64 numpy
// restatement the parity tests compare against.
#include <cuda_bf16.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.h"
#include "ptx.cuh"

namespace irp {

// ============================================================================================================
// 1. split + transpose pre-pass:  y = x - shift;  y = a + b + c with a,b,c bf16 (8 mantissa bits each);
//    writes A^T, B^T, C^T as [dim][n_pad] (K-major: samples contiguous) and accumulates column sums.
// ============================================================================================================
constexpr int kSplitTile = 64;

__global__ void __launch_bounds__(256) split_transpose_kernel(const float* __restrict__ x, long long n_rows, int dim,
                                                              const float* __restrict__ shift, long long n_pad,
                                                              __nv_bfloat16* __restrict__ at,
                                                              __nv_bfloat16* __restrict__ bt,
                                                              __nv_bfloat16* __restrict__ ct,
                                                              double* __restrict__ col_part) {
  __shared__ float tile[kSplitTile][kSplitTile + 1];
  const long long r0 = static_cast<long long>(blockIdx.x) * kSplitTile;
  const int c0 = blockIdx.y * kSplitTile;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;  // 64 x 4
  for (int i = ty; i < kSplitTile; i += 4) {
    const long long r = r0 + i;
    float v = 0.f;
    if (r < n_rows) v = x[r * dim + c0 + tx] - shift[c0 + tx];
    tile[i][tx] = v;
  }
  __syncthreads();
  // column sums of this row tile in fp64, rows in order; col_sum_reduce_kernel adds the tiles in order
  if (ty == 0) {
    double s = 0.0;
#pragma unroll 8
    for (int i = 0; i < kSplitTile; ++i) s += static_cast<double>(tile[i][tx]);
    col_part[static_cast<size_t>(blockIdx.x) * dim + c0 + tx] = s;
  }
  // transposed write: thread tx walks samples (contiguous in the output), ty strides features
  for (int j = ty; j < kSplitTile; j += 4) {
    const float v = tile[tx][j];
    const __nv_bfloat16 a = __float2bfloat16_rn(v);
    const float ra = v - __bfloat162float(a);
    const __nv_bfloat16 b = __float2bfloat16_rn(ra);
    const float rb = ra - __bfloat162float(b);
    const __nv_bfloat16 c = __float2bfloat16_rn(rb);
    const size_t o = static_cast<size_t>(c0 + j) * n_pad + r0 + tx;
    at[o] = a;
    bt[o] = b;
    ct[o] = c;
  }
}

// col_sum[c] += sum over row tiles of col_part[tile][c].  Block = 32 columns x 8 tile lanes: lane g adds tiles
// g, g+8, ... in ascending order, the 8 lane sums are then added in lane order by one thread per column -- the
// order of every addition is fixed by construction, so the result does not depend on the block schedule.
__global__ void __launch_bounds__(256) col_sum_reduce_kernel(const double* __restrict__ col_part, int n_tiles, int dim,
                                                             double* __restrict__ col_sum) {
  __shared__ double part[8][33];
  const int cx = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  double s = 0.0;
  if (c < dim) {
#pragma unroll 4
    for (int t = g; t < n_tiles; t += 8) s += col_part[static_cast<size_t>(t) * dim + c];
  }
  part[g][cx] = s;
  __syncthreads();
  if (g == 0 && c < dim) {
    double tot = col_sum[c];
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += part[i][cx];
    col_sum[c] = tot;
  }
}

// ============================================================================================================
// 2. scatter GEMM on tcgen05:  S[i,j] += sum_k (a_i a_j + a_i b_j + b_i a_j + b_i b_j + a_i c_j + c_i a_j)[k]
//    upper-triangular 128x128 tiles only.  A CTA owns whole tiles and walks their K chunks in ascending order
//    (fp32 TMEM accumulation per chunk, fp64 read-modify-write of S per chunk by the same thread), so every
//    element of S sees its additions in one fixed order: no atomics, results independent of the block schedule.
// ============================================================================================================
constexpr int kCovBN = 128, kCovBK = 64, kCovStages = 6, kCovThreads = 192;
constexpr int kCovABytes = 128 * kCovBK * 2, kCovBBytes = kCovBN * kCovBK * 2;
constexpr int kCovSmem = kCovStages * (kCovABytes + kCovBBytes) + 256 + 1024;
constexpr int kCovPasses = 6;

struct alignas(64) CovParams {
  CUtensorMap tm[3];  // A^T, B^T, C^T: [dim][n_pad] bf16
  int tiles_1d;       // dim / 128
  int n_tri;          // tiles_1d*(tiles_1d+1)/2
  int k_blocks_total; // n_pad / 64
  int k_blocks_per_chunk;
  int n_chunks;
  int dim;
  double* scatter;    // [dim][dim] fp64, upper-triangular tiles accumulate here
};

__device__ __forceinline__ void tri_decode(int t, int nt, int* ti, int* tj) {
  // row-major enumeration of the upper triangle (ti <= tj)
  int i = 0;
  int rem = t;
  while (rem >= nt - i) {
    rem -= nt - i;
    ++i;
  }
  *ti = i;
  *tj = i + rem;
}

__global__ void __launch_bounds__(kCovThreads, 1) cov_gemm_kernel(const __grid_constant__ CovParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kCovStages * kCovABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kCovStages * (kCovABytes + kCovBBytes));
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kCovStages;
  uint64_t* tfull_bar = bars + 2 * kCovStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  // two accumulators per buffer: [hi*hi] and [all correction products] -- the small terms must not be
  // rounded against the large hi*hi partial sums (the tensor core truncates when it aligns addends)
  constexpr uint32_t kTmemCols = 4 * kCovBN;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 3; ++i) tma_prefetch_desc(&p.tm[i]);
    for (int i = 0; i < kCovStages; ++i) {
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

  // pass -> (A-side array, B-side array)
  const int pa[kCovPasses] = {0, 0, 1, 1, 0, 2};
  const int pb[kCovPasses] = {0, 1, 0, 1, 2, 0};

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < p.n_tri; t += gridDim.x)
      for (int chunk = 0; chunk < p.n_chunks; ++chunk) {
        int ti, tj;
        tri_decode(t, p.tiles_1d, &ti, &tj);
        const int kb0 = chunk * p.k_blocks_per_chunk;
        const int kb1 = min(p.k_blocks_total, kb0 + p.k_blocks_per_chunk);
        for (int pass = 0; pass < kCovPasses; ++pass) {
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full_bar[stage], kCovABytes + kCovBBytes);
            tma_load_2d(smem_a + stage * kCovABytes, &p.tm[pa[pass]], &full_bar[stage], kb * kCovBK, ti * 128);
            tma_load_2d(smem_b + stage * kCovBBytes, &p.tm[pb[pass]], &full_bar[stage], kb * kCovBK, tj * 128);
            if (++stage == kCovStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, kCovBN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < p.n_tri; t += gridDim.x)
      for (int chunk = 0; chunk < p.n_chunks; ++chunk) {
        const int kb0 = chunk * p.k_blocks_per_chunk;
        const int kb1 = min(p.k_blocks_total, kb0 + p.k_blocks_per_chunk);
        const int per_pass = kb1 - kb0;
        const int total = kCovPasses * per_pass;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_big = tmem_base + acc * 2 * kCovBN;
        const uint32_t tmem_small = tmem_big + kCovBN;
        for (int it = 0; it < total; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + stage * kCovABytes);
          const uint32_t b_addr = smem_u32(smem_b + stage * kCovBBytes);
          const bool big = it < per_pass;  // pass 0 = hi*hi
          const uint32_t tmem_d = big ? tmem_big : tmem_small;
          const int first = big ? 0 : per_pass;
#pragma unroll
          for (int k = 0; k < kCovBK / 16; ++k)
            umma_bf16(tmem_d, umma_smem_desc<128>(a_addr + k * 32), umma_smem_desc<128>(b_addr + k * 32), idesc,
                      (it != first || k != 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (++stage == kCovStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < p.n_tri; t += gridDim.x)
    for (int chunk = 0; chunk < p.n_chunks; ++chunk) {
      int ti, tj;
      tri_decode(t, p.tiles_1d, &ti, &tj);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * 2 * kCovBN + (static_cast<uint32_t>(quarter * 32) << 16);
      double* dst = p.scatter + static_cast<size_t>(ti * 128 + row) * p.dim + tj * 128;
#pragma unroll 1
      for (int c = 0; c < kCovBN; c += 32) {
        uint32_t v[32], s[32];
        __syncwarp();
        tmem_ld_32x32b_x32(taddr + c, v);
        tmem_ld_32x32b_x32(taddr + kCovBN + c, s);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          dst[c + i] += static_cast<double>(__uint_as_float(v[i])) + static_cast<double>(__uint_as_float(s[i]));
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

// ============================================================================================================
// 3. covariance assembly:  mean = shift + sum/n;  C = (S - n * delta delta^T) / (n-1), delta = sum/n; symmetric.
// ============================================================================================================
__global__ void assemble_cov_kernel(const double* __restrict__ count, const double* __restrict__ sum,
                                    const double* __restrict__ scatter, const float* __restrict__ shift, int dim,
                                    double* __restrict__ mean, double* __restrict__ cov) {
  const double n = count[0];
  const long long total = static_cast<long long>(dim) * dim;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int i = static_cast<int>(idx / dim), j = static_cast<int>(idx % dim);
    const int a = i < j ? i : j, b = i < j ? j : i;
    // tiles are stored for tile_row <= tile_col; inside a diagonal tile both triangles are valid
    const double s = scatter[static_cast<size_t>(a) * dim + b];
    const double di = sum[i] / n, dj = sum[j] / n;
    const double c = (s - n * di * dj) / (n - 1.0);
    cov[idx] = c;
    if (i == j) mean[i] = static_cast<double>(shift[i]) + di;
  }
}

// trace of the covariance: one block, each thread sums its strided share of the diagonal in ascending order and
// the block combines the 1024 partials in a fixed tree, so the total variance is bitwise repeatable.
__global__ void __launch_bounds__(1024) trace_kernel(const double* __restrict__ cov, int dim,
                                                     double* __restrict__ trace) {
  __shared__ double part[1024];
  double s = 0.0;
  for (int i = threadIdx.x; i < dim; i += 1024) s += cov[static_cast<size_t>(i) * dim + i];
  part[threadIdx.x] = s;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) trace[0] = part[0];
}

// ============================================================================================================
// 4. Householder tridiagonalisation (fp64, matrix [n][n] symmetric, both triangles kept).
//    tridiag_step(j) applies reflector j (rank-2 update of the trailing block) and, fused in the same pass over
//    the matrix, forms reflector j+1 and p_{j+1} = tau A v.  One launch per column; the 33 MB matrix stays in L2.
//      V    [n][n]: row t holds reflector t (entries t+1..n-1; v[t+1] = 1)
//      tau  [n], diag [n], off [n] (off[t] = sub-diagonal between t and t+1)
//      pbuf [2][n]: p vectors, ping-pong
// ============================================================================================================
constexpr int kTriThreads = 256;

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double s = 0.0;
  const int nw = blockDim.x >> 5;
  for (int i = 0; i < nw; ++i) s += red[i];
  return s;
}

__global__ void __launch_bounds__(kTriThreads) tridiag_step_kernel(double* __restrict__ A, int n, int j,
                                                                    double* __restrict__ V, double* __restrict__ tau,
                                                                    double* __restrict__ diag,
                                                                    double* __restrict__ off,
                                                                    double* __restrict__ pbuf) {
  extern __shared__ double sh[];
  double* w = sh;           // [n] w_j           (indices j+1..n-1 valid)
  double* v = sh + n;       // [n] v_j
  double* vn = sh + 2 * n;  // [n] v_{j+1}       (indices t+1..n-1 valid)
  __shared__ double red[kTriThreads / 32];
  __shared__ double s_tau_next, s_scale, s_beta;

  const int tid = threadIdx.x;
  const int t = j + 1;  // row whose reflector is formed in this launch
  const double* p_cur = pbuf + static_cast<size_t>((j & 1)) * n;
  double* p_next = pbuf + static_cast<size_t>((t & 1)) * n;

  // ---- w_j = p_j - (tau_j/2 * p_j.v_j) v_j ----
  if (j >= 0) {
    const double tj = tau[j];
    double part = 0.0;
    for (int c = t + tid; c < n; c += kTriThreads) {
      const double vv = V[static_cast<size_t>(j) * n + c];
      const double pp = p_cur[c];
      v[c] = vv;
      w[c] = pp;
      part += vv * pp;
    }
    const double alpha = 0.5 * tj * block_sum(part, red);
    for (int c = t + tid; c < n; c += kTriThreads) w[c] -= alpha * v[c];
  } else {
    for (int c = t + tid; c < n; c += kTriThreads) {
      v[c] = 0.0;
      w[c] = 0.0;
    }
  }
  __syncthreads();

  // ---- updated row t -> d[t], x = A_new[t][t+1:] -> reflector t ----
  const double vt = v[t], wt = w[t];
  double norm2 = 0.0;
  for (int c = t + 1 + tid; c < n; c += kTriThreads) {
    const double a = A[static_cast<size_t>(t) * n + c] - vt * w[c] - wt * v[c];
    vn[c] = a;
    if (c > t + 1) norm2 += a * a;
  }
  norm2 = block_sum(norm2, red);
  if (tid == 0) {
    const double alpha = vn[t + 1];
    double beta, tn, scale;
    if (norm2 == 0.0) {
      beta = alpha;
      tn = 0.0;
      scale = 0.0;
    } else {
      beta = -copysign(sqrt(alpha * alpha + norm2), alpha);
      tn = (beta - alpha) / beta;
      scale = 1.0 / (alpha - beta);
    }
    s_tau_next = tn;
    s_scale = scale;
    s_beta = beta;
    if (blockIdx.x == 0) {
      diag[t] = A[static_cast<size_t>(t) * n + t] - 2.0 * vt * wt;
      off[t] = beta;
      tau[t] = tn;
    }
  }
  __syncthreads();
  const double tn = s_tau_next, scale = s_scale;
  for (int c = t + 1 + tid; c < n; c += kTriThreads) {
    const double val = (c == t + 1) ? 1.0 : vn[c] * scale;
    vn[c] = val;
    if (blockIdx.x == 0) V[static_cast<size_t>(t) * n + c] = val;
  }
  __syncthreads();

  // ---- trailing block rows i >= t+1: rank-2 update fused with p_{t} = tau_t * A_new[i, t+1:] . v_t ----
  const int warp = tid >> 5, lane = tid & 31;
  const int warps_total = gridDim.x * (kTriThreads / 32);
  for (int i = t + 1 + blockIdx.x * (kTriThreads / 32) + warp; i < n; i += warps_total) {
    double* row = A + static_cast<size_t>(i) * n;
    const double vi = v[i], wi = w[i];
    double acc = 0.0;
    for (int c = t + 1 + lane; c < n; c += 32) {
      const double a = row[c] - vi * w[c] - wi * v[c];
      row[c] = a;
      acc += a * vn[c];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) p_next[i] = tn * acc;
  }
}

// ============================================================================================================
// 4b. Lanczos with full re-orthogonalisation: only the top-k eigenpairs of the covariance are needed, and the
//     Krylov space of a generic start vector captures them after a few hundred steps, while the Householder
//     reduction above always pays for all n columns (n^3/3 * 16 bytes of matrix traffic and n launches).
//       step j:  w = C q_j;  h1 = Q^T w; w -= Q h1;  h2 = Q^T w; w -= Q h2   (classical Gram-Schmidt, twice)
//                alpha_j = h1[j] + h2[j];  beta_j = ||w||;  q_{j+1} = w / beta_j
//     T_m = tridiag(alpha, beta) goes through the same bisection / inverse-iteration kernels; a Ritz pair
//     (theta, Q s) has residual |beta_{m-1} s_{m-1}|, which the host checks after m steps (more steps if needed;
//     at m = n the recurrence IS the exact tridiagonalisation, and the caller falls back to section 4 long before).
// ============================================================================================================
__global__ void lz_start_kernel(double* __restrict__ q0, int n) {
  __shared__ double red[32];
  pdl_trigger();
  pdl_wait();
  double part = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    uint32_t h = static_cast<uint32_t>(i) * 2654435761u + 0x9E3779B9u;  // fixed pseudo-random start vector
    h ^= h >> 16;
    h *= 0x85EBCA6Bu;
    h ^= h >> 13;
    const double v = static_cast<double>(h) * (1.0 / 4294967296.0) - 0.5;
    q0[i] = v;
    part += v * v;
  }
  const double nrm = sqrt(block_sum(part, red));
  for (int i = threadIdx.x; i < n; i += blockDim.x) q0[i] /= nrm;
}

// All Lanczos kernels are launched with programmatic stream serialisation: the step is a chain of five short
// dependent kernels, and letting each one become resident while its predecessor drains removes most of the
// launch gap.  pdl_wait() precedes the first access to anything a predecessor wrote.
//
// Step j:  matvec   q_j = v / ||v||  (v = the previous step's orthogonalised vector; every CTA normalises its own
//                   shared-memory copy, CTA 0 stores q_j, alpha_{j-1}, beta_{j-1});  w0 = C q_j
//          dots     h1 = Q^T w0            (one CTA per basis vector)
//          axpy     w1 = w0 - Q h1         (32 columns x 8 row shares per CTA, shares combined in a fixed order)
//          dots     h2 = Q^T w1
//          axpy     w2 = w1 - Q h2  -> v of step j+1

// normalise v into shared memory; returns ||v||
__device__ __forceinline__ double lz_normalise(const double* __restrict__ v, int n, double* sh_q, double* red) {
  double part = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double x = v[i];
    sh_q[i] = x;
    part += x * x;
  }
  const double beta = sqrt(block_sum(part, red));
  const double inv = beta > 0.0 ? 1.0 / beta : 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) sh_q[i] *= inv;  // own elements
  return beta;
}

__device__ __forceinline__ void lz_store_basis(const double* sh_q, int n, int j, double beta,
                                               const double* __restrict__ h1, const double* __restrict__ h2,
                                               double* __restrict__ Q, double* __restrict__ diag,
                                               double* __restrict__ off) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) Q[static_cast<size_t>(j) * n + i] = sh_q[i];
  if (threadIdx.x == 0 && j > 0) {
    diag[j - 1] = h1[j - 1] + h2[j - 1];
    off[j - 1] = beta;
  }
}

__global__ void __launch_bounds__(256) lz_matvec_kernel(const double* __restrict__ C, const double* __restrict__ v,
                                                        int n, int j, const double* __restrict__ h1,
                                                        const double* __restrict__ h2, double* __restrict__ Q,
                                                        double* __restrict__ diag, double* __restrict__ off,
                                                        double* __restrict__ w) {
  extern __shared__ double sh_q[];  // [n]
  __shared__ double red[32];
  pdl_trigger();
  pdl_wait();
  const double beta = lz_normalise(v, n, sh_q, red);
  if (blockIdx.x == 0) lz_store_basis(sh_q, n, j, beta, h1, h2, Q, diag, off);
  __syncthreads();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const double2* c2 = reinterpret_cast<const double2*>(C + static_cast<size_t>(row) * n);
  const double2* q2 = reinterpret_cast<const double2*>(sh_q);
  double acc0 = 0.0, acc1 = 0.0;
#pragma unroll 8
  for (int i = lane; i < (n >> 1); i += 32) {
    const double2 a = c2[i], b = q2[i];
    acc0 = fma(a.x, b.x, acc0);
    acc1 = fma(a.y, b.y, acc1);
  }
  double acc = acc0 + acc1;
  if ((n & 1) && lane == 0) acc = fma(C[static_cast<size_t>(row) * n + n - 1], sh_q[n - 1], acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) w[row] = acc;
}

// closes a batch of steps: q_m, alpha_{m-1}, beta_{m-1} (the next batch's first matvec recomputes the same values)
__global__ void __launch_bounds__(256) lz_close_kernel(const double* __restrict__ v, int n, int j,
                                                       const double* __restrict__ h1, const double* __restrict__ h2,
                                                       double* __restrict__ Q, double* __restrict__ diag,
                                                       double* __restrict__ off) {
  extern __shared__ double sh_q[];
  __shared__ double red[32];
  pdl_trigger();
  pdl_wait();
  const double beta = lz_normalise(v, n, sh_q, red);
  lz_store_basis(sh_q, n, j, beta, h1, h2, Q, diag, off);
}

// h[r] = Q[r] . w, one CTA per basis vector
__global__ void __launch_bounds__(256) lz_dots_kernel(const double* __restrict__ Q, const double* __restrict__ w,
                                                      int n, double* __restrict__ h) {
  __shared__ double red[32];
  pdl_trigger();
  pdl_wait();
  const double* qr = Q + static_cast<size_t>(blockIdx.x) * n;
  double acc = 0.0;
#pragma unroll 8
  for (int i = threadIdx.x; i < n; i += 256) acc = fma(qr[i], w[i], acc);
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) h[blockIdx.x] = acc;
}

// w_out = w_in - sum_{r < cnt} h[r] Q[r]: 32 columns per CTA, the basis rows dealt round-robin to 8 shares
// (thread = share * 32 + column) and the shares combined in a fixed order: deterministic, no atomics.
__global__ void __launch_bounds__(256) lz_axpy_kernel(const double* __restrict__ Q, const double* __restrict__ h,
                                                      const double* __restrict__ w_in, int n, int cnt,
                                                      double* __restrict__ w_out) {
  __shared__ double sh_part[8][33];
  pdl_trigger();
  pdl_wait();
  const int col = threadIdx.x & 31, share = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + col;
  double acc = 0.0;
  if (i < n) {
#pragma unroll 8
    for (int r = share; r < cnt; r += 8) acc = fma(h[r], Q[static_cast<size_t>(r) * n + i], acc);
  }
  sh_part[share][col] = acc;
  __syncthreads();
  if (share == 0 && i < n) {
    double sum = 0.0;
#pragma unroll
    for (int y = 0; y < 8; ++y) sum += sh_part[y][col];
    w_out[i] = w_in[i] - sum;
  }
}

// convergence data for the host: out[0] = max_i |beta_{m-1} S[i][m-1]|, out[1] = min_j beta_j (j < m-1), out[2] = ||T||
__global__ void lz_residual_kernel(const double* __restrict__ Z, const double* __restrict__ off, int m, int k,
                                   const double* __restrict__ tnorm, double* __restrict__ out) {
  double r = 0.0, bmin = INFINITY;
  for (int i = threadIdx.x; i < k; i += blockDim.x) r = fmax(r, fabs(off[m - 1] * Z[static_cast<size_t>(i) * m + m - 1]));
  for (int j = threadIdx.x; j < m - 1; j += blockDim.x) bmin = fmin(bmin, off[j]);
  __shared__ double sr[256], sb[256];
  sr[threadIdx.x] = r;
  sb[threadIdx.x] = bmin;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int t = 1; t < blockDim.x; ++t) {
      r = fmax(r, sr[t]);
      bmin = fmin(bmin, sb[t]);
    }
    out[0] = r;
    out[1] = bmin;
    out[2] = tnorm[0];
  }
}

// Ritz vectors comps[which] = sum_j S[which][j] Q[j] (grid: component x 256-column block), then sklearn's sign
// convention (entry of largest magnitude positive, first occurrence on ties) in a second kernel
__global__ void __launch_bounds__(256) lz_ritz_kernel(const double* __restrict__ Q, const double* __restrict__ S,
                                                      int n, int m, double* __restrict__ comps) {
  extern __shared__ double sh_s[];  // S row [m]
  const int which = blockIdx.x;
  for (int j = threadIdx.x; j < m; j += 256) sh_s[j] = S[static_cast<size_t>(which) * m + j];
  __syncthreads();
  const int c = blockIdx.y * 256 + threadIdx.x;
  if (c >= n) return;
  double acc0 = 0.0, acc1 = 0.0;
  int j = 0;
#pragma unroll 4
  for (; j + 1 < m; j += 2) {
    acc0 = fma(sh_s[j], Q[static_cast<size_t>(j) * n + c], acc0);
    acc1 = fma(sh_s[j + 1], Q[static_cast<size_t>(j + 1) * n + c], acc1);
  }
  if (j < m) acc0 = fma(sh_s[j], Q[static_cast<size_t>(j) * n + c], acc0);
  comps[static_cast<size_t>(which) * n + c] = acc0 + acc1;
}

__global__ void __launch_bounds__(256) lz_sign_kernel(double* __restrict__ comps, int n) {
  __shared__ double wb[8];
  __shared__ int wi[8];
  __shared__ double s_sign;
  double* row = comps + static_cast<size_t>(blockIdx.x) * n;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  double best = -1.0;
  int bidx = n;
  for (int c = tid; c < n; c += 256) {
    const double a = fabs(row[c]);
    if (a > best) {  // c ascends per thread: first occurrence kept on ties
      best = a;
      bidx = c;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
    if (ob > best || (ob == best && oi < bidx)) {
      best = ob;
      bidx = oi;
    }
  }
  if (lane == 0) {
    wb[warp] = best;
    wi[warp] = bidx;
  }
  __syncthreads();
  if (tid == 0) {
    double bb = wb[0];
    int bi = wi[0];
    for (int i = 1; i < 8; ++i)
      if (wb[i] > bb || (wb[i] == bb && wi[i] < bi)) {
        bb = wb[i];
        bi = wi[i];
      }
    s_sign = row[bi] < 0.0 ? -1.0 : 1.0;
  }
  __syncthreads();
  if (s_sign < 0.0)
    for (int c = tid; c < n; c += 256) row[c] = -row[c];
}

// ============================================================================================================
// 5. top-k eigenvalues of the tridiagonal by multisection: one CTA per eigenvalue, one shift per thread and round.
//    Sturm count = sign changes of the leading-minor polynomials p_i(x) (three-term recurrence, division free,
//    rescaled every 8 steps; a zero takes the sign opposite to its predecessor -- Wilkinson).  The matrix is scaled
//    by 1/||T|| so the recurrence cannot overflow between rescalings.
// ============================================================================================================
constexpr int kBisThreads = 128;

__global__ void __launch_bounds__(kBisThreads) bisect_topk_kernel(const double* __restrict__ diag,
                                                                  const double* __restrict__ off, int n, int k,
                                                                  double* __restrict__ evals,
                                                                  double* __restrict__ tnorm) {
  extern __shared__ double sh[];
  double* d = sh;
  double* e2 = sh + n;  // e2[i] = (off[i-1]/||T||)^2 for i>=1
  __shared__ double s_gl[kBisThreads / 32], s_gu[kBisThreads / 32];
  double gl = INFINITY, gu = -INFINITY;
  for (int i = threadIdx.x; i < n; i += kBisThreads) {
    const double di = diag[i];
    const double el = i > 0 ? fabs(off[i - 1]) : 0.0, er = i < n - 1 ? fabs(off[i]) : 0.0;
    gl = fmin(gl, di - el - er);
    gu = fmax(gu, di + el + er);
  }
  for (int o = 16; o > 0; o >>= 1) {
    gl = fmin(gl, __shfl_xor_sync(0xffffffffu, gl, o));
    gu = fmax(gu, __shfl_xor_sync(0xffffffffu, gu, o));
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    s_gl[warp] = gl;
    s_gu[warp] = gu;
  }
  __syncthreads();
  gl = s_gl[0];
  gu = s_gu[0];
  for (int w = 1; w < kBisThreads / 32; ++w) {
    gl = fmin(gl, s_gl[w]);
    gu = fmax(gu, s_gu[w]);
  }
  double tn = fmax(fabs(gl), fabs(gu));
  if (!(tn > 0.0)) tn = 1.0;
  const double inv = 1.0 / tn;
  for (int i = threadIdx.x; i < n; i += kBisThreads) {
    d[i] = diag[i] * inv;
    const double e = i > 0 ? off[i - 1] * inv : 0.0;
    e2[i] = e * e;
  }
  __syncthreads();
  const int which = blockIdx.x;  // which-th largest
  if (which == 0 && threadIdx.x == 0) tnorm[0] = tn;
  if (which >= k) return;
  const int idx = n - 1 - which;  // ascending 0-based index
  const double slack = 4.0 * 2.220446049250313e-16 * n;
  double lo = gl * inv - slack, hi = gu * inv + slack;
  for (int round = 0; round < 9; ++round) {
    const double x = lo + (hi - lo) * (static_cast<double>(threadIdx.x + 1) / (kBisThreads + 1));
    double pm1 = 1.0, p = d[0] - x;
    int sgn = (p > 0.0) ? 1 : -1;  // p == 0 counts as a sign change against p_0 = 1
    int cnt = sgn < 0;
    for (int i = 1; i < n; ++i) {
      const double pn = fma(d[i] - x, p, -e2[i] * pm1);
      const int s = pn > 0.0 ? 1 : (pn < 0.0 ? -1 : -sgn);
      cnt += (s != sgn);
      sgn = s;
      pm1 = p;
      p = pn;
      if ((i & 7) == 0) {
        const double a = fabs(p);
        if (a > 1e100) {
          p *= 1e-100;
          pm1 *= 1e-100;
        } else if (a < 1e-100) {
          p *= 1e100;
          pm1 *= 1e100;
        }
      }
    }
    // threads whose shift is at or below the eigenvalue see cnt <= idx; monotone in the thread index
    const int nb = __syncthreads_count(cnt <= idx);
    const double w = (hi - lo) / (kBisThreads + 1);
    const double new_lo = nb > 0 ? lo + w * nb : lo;
    const double new_hi = nb < kBisThreads ? lo + w * (nb + 1) : hi;
    lo = new_lo;
    hi = new_hi;
  }
  if (threadIdx.x == 0) evals[which] = 0.5 * (lo + hi) * tn;
}

// ============================================================================================================
// 6. inverse iteration on the tridiagonal: one warp per eigenvector, everything in shared memory.
//    (T - lambda I) is factored ONCE with partial pivoting (LAPACK dgttrf), then three solves (dgttrs) from a
//    pseudo-random start, normalising in between.  Lane 0 walks the recurrences, the warp does the vector parts.
//    dynamic smem: dd, rd, du, du2, fl (multipliers) [n doubles each], b [n], piv [n bytes]
// ============================================================================================================
__device__ __forceinline__ double hash_unit(uint32_t a, uint32_t b) {
  uint32_t h = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA77u;
  h ^= h >> 15;
  h *= 0x2C1B3C6Du;
  h ^= h >> 12;
  h *= 0x297A2D39u;
  h ^= h >> 15;
  return (static_cast<double>(h) / 4294967296.0) - 0.5;
}

__global__ void __launch_bounds__(32) inverse_iteration_kernel(const double* __restrict__ diag,
                                                               const double* __restrict__ off, int n, int k,
                                                               const double* __restrict__ evals,
                                                               const double* __restrict__ tnorm,
                                                               double* __restrict__ Z) {
  extern __shared__ double shi[];
  double* dd = shi;
  double* rd = dd + n;
  double* du = rd + n;
  double* du2 = du + n;
  double* fl = du2 + n;
  double* b = fl + n;
  uint8_t* piv = reinterpret_cast<uint8_t*>(b + n);
  const int which = blockIdx.x;
  if (which >= k) return;
  const int lane = threadIdx.x;
  const double lam = evals[which];
  const double pivmin = fmax(tnorm[0], 1e-300) * 2.220446049250313e-16;
  for (int i = lane; i < n; i += 32) {
    dd[i] = diag[i] - lam;
    du[i] = i < n - 1 ? off[i] : 0.0;
    fl[i] = i < n - 1 ? off[i] : 0.0;  // sub-diagonal, overwritten by the multipliers
    du2[i] = 0.0;
    piv[i] = 0;
    b[i] = hash_unit(static_cast<uint32_t>(which), static_cast<uint32_t>(i));
  }
  __syncwarp();
  if (lane == 0) {
    // dgttrf
    double di = dd[0];
    for (int i = 0; i < n - 1; ++i) {
      const double li = fl[i];
      double dn = dd[i + 1];
      if (fabs(di) >= fabs(li)) {
        if (fabs(di) < pivmin) di = copysign(pivmin, di == 0.0 ? 1.0 : di);
        const double f = li / di;
        fl[i] = f;
        dd[i] = di;
        rd[i] = 1.0 / di;
        dn -= f * du[i];
      } else {
        const double f = di / li;
        dd[i] = li;
        rd[i] = 1.0 / li;
        fl[i] = f;
        const double tmp = du[i];
        du[i] = dn;
        dn = tmp - f * dn;
        if (i < n - 2) {
          du2[i] = du[i + 1];
          du[i + 1] = -f * du[i + 1];
        }
        piv[i] = 1;
      }
      di = dn;
    }
    if (fabs(di) < pivmin) di = copysign(pivmin, di == 0.0 ? 1.0 : di);
    dd[n - 1] = di;
    rd[n - 1] = 1.0 / di;
  }
  __syncwarp();
  for (int iter = 0; iter < 3; ++iter) {
    if (lane == 0) {
      // dgttrs: L solve
      double bi = b[0];
      for (int i = 0; i < n - 1; ++i) {
        const double bn = b[i + 1];
        if (!piv[i]) {
          b[i] = bi;
          bi = bn - fl[i] * bi;
        } else {
          b[i] = bn;
          bi = bi - fl[i] * bn;
        }
      }
      b[n - 1] = bi;
      // U solve
      double x1 = b[n - 1] * rd[n - 1];
      b[n - 1] = x1;
      double x2 = 0.0;
      if (n > 1) {
        const double x0 = (b[n - 2] - du[n - 2] * x1) * rd[n - 2];
        b[n - 2] = x0;
        x2 = x1;
        x1 = x0;
      }
      for (int i = n - 3; i >= 0; --i) {
        const double x0 = (b[i] - du[i] * x1 - du2[i] * x2) * rd[i];
        b[i] = x0;
        x2 = x1;
        x1 = x0;
      }
    }
    __syncwarp();
    // normalise (max-abs first to stay in range)
    double mx = 0.0;
    for (int i = lane; i < n; i += 32) mx = fmax(mx, fabs(b[i]));
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const double invm = mx > 0.0 ? 1.0 / mx : 1.0;
    double ss = 0.0;
    for (int i = lane; i < n; i += 32) {
      const double v = b[i] * invm;
      ss += v * v;
    }
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const double rn = invm / sqrt(ss);
    for (int i = lane; i < n; i += 32) b[i] *= rn;
    __syncwarp();
  }
  for (int i = lane; i < n; i += 32) Z[static_cast<size_t>(which) * n + i] = b[i];
}

// Re-orthogonalise eigenvectors whose eigenvalues are numerically degenerate (|gap| < 1e-8 ||T||); single CTA.
__global__ void __launch_bounds__(256) cluster_mgs_kernel(double* __restrict__ Z, int n, int k,
                                                          const double* __restrict__ evals,
                                                          const double* __restrict__ tnorm) {
  __shared__ double red[8];
  const double tol = 1e-8 * tnorm[0];
  int cluster_start = 0;
  for (int i = 1; i < k; ++i) {
    if (fabs(evals[i - 1] - evals[i]) >= tol) {
      cluster_start = i;
      continue;
    }
    double* zi = Z + static_cast<size_t>(i) * n;
    for (int j = cluster_start; j < i; ++j) {
      const double* zj = Z + static_cast<size_t>(j) * n;
      double part = 0.0;
      for (int c = threadIdx.x; c < n; c += blockDim.x) part += zi[c] * zj[c];
      const double dot = block_sum(part, red);
      for (int c = threadIdx.x; c < n; c += blockDim.x) zi[c] -= dot * zj[c];
      __syncthreads();
    }
    double part = 0.0;
    for (int c = threadIdx.x; c < n; c += blockDim.x) part += zi[c] * zi[c];
    const double nn = block_sum(part, red);
    const double rn = nn > 0.0 ? 1.0 / sqrt(nn) : 0.0;
    for (int c = threadIdx.x; c < n; c += blockDim.x) zi[c] *= rn;
    __syncthreads();
  }
}

// ============================================================================================================
// 7. back-transform y = H_0 H_1 ... H_{n-2} z (one CTA per eigenvector; z lives in registers, the next reflector
//    row is prefetched while the current one is reduced), then sklearn's sign convention: the entry of largest
//    magnitude (first on ties) is made positive.
// ============================================================================================================
constexpr int kBtThreads = 256;
constexpr int kBtPerThread = 16;  // supports n <= 4096

__global__ void __launch_bounds__(kBtThreads) back_transform_kernel(const double* __restrict__ V,
                                                                    const double* __restrict__ tau, int n,
                                                                    const double* __restrict__ Z,
                                                                    double* __restrict__ comps) {
  __shared__ double red[2][kBtThreads / 32];
  __shared__ double wb[kBtThreads / 32];
  __shared__ int wi[kBtThreads / 32];
  __shared__ int s_idx;
  const int which = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  double z[kBtPerThread], vc[kBtPerThread], vnx[kBtPerThread];
#pragma unroll
  for (int j = 0; j < kBtPerThread; ++j) {
    const int c = tid + kBtThreads * j;
    z[j] = c < n ? Z[static_cast<size_t>(which) * n + c] : 0.0;
  }
  auto load_row = [&](int t, double (&dst)[kBtPerThread]) {
    const double* vt = V + static_cast<size_t>(t) * n;
#pragma unroll
    for (int j = 0; j < kBtPerThread; ++j) {
      const int c = tid + kBtThreads * j;
      dst[j] = (c > t && c < n) ? vt[c] : 0.0;
    }
  };
  if (n >= 2) load_row(n - 2, vc);
  int parity = 0;
  for (int t = n - 2; t >= 0; --t) {
    if (t > 0) load_row(t - 1, vnx);
    const double tt = tau[t];
    double part = 0.0;
#pragma unroll
    for (int j = 0; j < kBtPerThread; ++j) part += vc[j] * z[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) red[parity][warp] = part;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kBtThreads / 32; ++w) s += red[parity][w];
    s *= tt;
#pragma unroll
    for (int j = 0; j < kBtPerThread; ++j) {
      z[j] -= s * vc[j];
      vc[j] = vnx[j];
    }
    parity ^= 1;
  }
  // argmax |z| (first occurrence)
  double best = -1.0;
  int bidx = n;
#pragma unroll
  for (int j = 0; j < kBtPerThread; ++j) {
    const int c = tid + kBtThreads * j;
    const double a = fabs(z[j]);
    if (c < n && (a > best || (a == best && c < bidx))) {
      best = a;
      bidx = c;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
    if (ob > best || (ob == best && oi < bidx)) {
      best = ob;
      bidx = oi;
    }
  }
  if (lane == 0) {
    wb[warp] = best;
    wi[warp] = bidx;
  }
  __syncthreads();
  if (tid == 0) {
    double bb = wb[0];
    int bi = wi[0];
    for (int i = 1; i < kBtThreads / 32; ++i)
      if (wb[i] > bb || (wb[i] == bb && wi[i] < bi)) {
        bb = wb[i];
        bi = wi[i];
      }
    s_idx = bi;
  }
  __syncthreads();
  // the owner of the arg-max entry publishes the sign
  __shared__ double s_sign;
  {
    const int c = s_idx;
    if (c % kBtThreads == tid) {
      double val = 0.0;
#pragma unroll
      for (int j = 0; j < kBtPerThread; ++j)
        if (tid + kBtThreads * j == c) val = z[j];
      s_sign = val < 0.0 ? -1.0 : 1.0;
    }
  }
  __syncthreads();
  const double sign = s_sign;
#pragma unroll
  for (int j = 0; j < kBtPerThread; ++j) {
    const int c = tid + kBtThreads * j;
    if (c < n) comps[static_cast<size_t>(which) * n + c] = sign * z[j];
  }
}

// clip negative eigenvalues to zero (sklearn _pca.py:626)
__global__ void clip_evals_kernel(double* evals, int k) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < k && evals[i] < 0.0) evals[i] = 0.0;
}

// ============================================================================================================
// 8. projection  Z = (X - mean) V^T  as a split-bf16 tensor-core GEMM (fp32 out)
//    pre-pass : y = float(double(x) - mean) split EXACTLY into three bf16 terms (a + b + c = y), rows K-major;
//               the components (fp64) are split the same way, padded to N rows
//    GEMM     : one CTA per 128-row tile, N = 64 or 128 columns; per 64-wide K block the three A tiles and the
//               three B tiles are loaded ONCE and feed the six products a.va | a.vb + b.va + b.vb + a.vc + c.va
//               (fp32 accumulation in TMEM; the hi*hi term has its own accumulator so the small terms are not
//               rounded against its partial sums); the epilogue adds the two accumulators and stores fp32.
// ============================================================================================================
__global__ void __launch_bounds__(256) project_split_kernel(const float* __restrict__ x, long long n_rows, int dim,
                                                            const double* __restrict__ mean,
                                                            __nv_bfloat16* __restrict__ ya,
                                                            __nv_bfloat16* __restrict__ yb,
                                                            __nv_bfloat16* __restrict__ yc, long long n_pad) {
  // one thread per 4 consecutive features: 16-byte load, three 8-byte stores; rows >= n_rows are written as zeros
  const long long total = n_pad * (dim / 4);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / (dim / 4);
    const int c4 = static_cast<int>(i - r * (dim / 4)) * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (r < n_rows) {
      const float4 xv = *reinterpret_cast<const float4*>(x + r * dim + c4);
      v[0] = static_cast<float>(static_cast<double>(xv.x) - mean[c4 + 0]);
      v[1] = static_cast<float>(static_cast<double>(xv.y) - mean[c4 + 1]);
      v[2] = static_cast<float>(static_cast<double>(xv.z) - mean[c4 + 2]);
      v[3] = static_cast<float>(static_cast<double>(xv.w) - mean[c4 + 3]);
    }
    __nv_bfloat16 a[4], b[4], c[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      a[t] = __float2bfloat16_rn(v[t]);
      const float ra = v[t] - __bfloat162float(a[t]);
      b[t] = __float2bfloat16_rn(ra);
      c[t] = __float2bfloat16_rn(ra - __bfloat162float(b[t]));
    }
    const size_t o = static_cast<size_t>(r) * dim + c4;
    *reinterpret_cast<uint2*>(ya + o) = *reinterpret_cast<const uint2*>(a);
    *reinterpret_cast<uint2*>(yb + o) = *reinterpret_cast<const uint2*>(b);
    *reinterpret_cast<uint2*>(yc + o) = *reinterpret_cast<const uint2*>(c);
  }
}

__global__ void __launch_bounds__(256) project_split_comps_kernel(const double* __restrict__ comps, int k, int dim,
                                                                  int n_cols, __nv_bfloat16* __restrict__ va,
                                                                  __nv_bfloat16* __restrict__ vb,
                                                                  __nv_bfloat16* __restrict__ vc) {
  const int total = n_cols * dim;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / dim;
    const double v = r < k ? comps[i] : 0.0;
    const __nv_bfloat16 a = __double2bfloat16(v);
    const double ra = v - static_cast<double>(__bfloat162float(a));
    const __nv_bfloat16 b = __double2bfloat16(ra);
    const double rb = ra - static_cast<double>(__bfloat162float(b));
    va[i] = a;
    vb[i] = b;
    vc[i] = __double2bfloat16(rb);
  }
}

constexpr int kPjBK = 64, kPjThreads = 192;
constexpr int kPjABytes = 128 * kPjBK * 2;

template <int N>
struct PjSmem {
  static constexpr int kBBytes = N * kPjBK * 2;
  static constexpr int kStageBytes = 3 * (kPjABytes + kBBytes);
  static constexpr int kStages = (220 * 1024) / kStageBytes;
  static constexpr int kTotal = kStages * kStageBytes + 256 + 1024;
  static_assert(kStages >= 2, "projection GEMM needs two pipeline stages");
};

struct alignas(64) ProjParams {
  CUtensorMap tmY[3];  // y terms [n_pad][dim] bf16, box (64, 128)
  CUtensorMap tmV[3];  // component terms [N][dim] bf16, box (64, N)
  float* z;            // [n_rows][k]
  long long n_rows;
  int k, k_blocks, n_tiles;
};

template <int N>
__global__ void __launch_bounds__(kPjThreads, 1) proj_gemm_kernel(const __grid_constant__ ProjParams p) {
  using S = PjSmem<N>;
  constexpr int kStages = S::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * S::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tfull_bar = bars + 2 * kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  constexpr uint32_t kTmemCols = 4 * N;  // two buffers x (hi*hi | corrections)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 3; ++i) {
      tma_prefetch_desc(&p.tmY[i]);
      tma_prefetch_desc(&p.tmV[i]);
    }
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

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], S::kStageBytes);
          uint8_t* st = smem + stage * S::kStageBytes;
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            tma_load_2d(st + i * kPjABytes, &p.tmY[i], &full_bar[stage], kb * kPjBK, t * 128);
            tma_load_2d(st + 3 * kPjABytes + i * S::kBBytes, &p.tmV[i], &full_bar[stage], kb * kPjBK, 0);
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, N);
      // (A term, B term) of the six products; the first one accumulates alone
      const int pa[6] = {0, 0, 1, 1, 0, 2};
      const int pb[6] = {0, 1, 0, 1, 2, 0};
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_big = tmem_base + acc * 2 * N;
        const uint32_t tmem_small = tmem_big + N;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t st = smem_u32(smem + stage * S::kStageBytes);
#pragma unroll
          for (int c = 0; c < 6; ++c) {
            const uint32_t a_addr = st + pa[c] * kPjABytes;
            const uint32_t b_addr = st + 3 * kPjABytes + pb[c] * S::kBBytes;
#pragma unroll
            for (int k = 0; k < kPjBK / 16; ++k) {
              const uint32_t accum = c == 0 ? ((kb | k) != 0 ? 1u : 0u) : ((kb != 0 || c > 1 || k != 0) ? 1u : 0u);
              umma_bf16(c == 0 ? tmem_big : tmem_small, umma_smem_desc<128>(a_addr + k * 32),
                        umma_smem_desc<128>(b_addr + k * 32), idesc, accum);
            }
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    const int quarter = warp & 3;
    const long long row_in_tile = quarter * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * 2 * N + (static_cast<uint32_t>(quarter * 32) << 16);
      const long long row = static_cast<long long>(t) * 128 + row_in_tile;
      float* dst = p.z + row * p.k;
#pragma unroll 1
      for (int c = 0; c < N; c += 32) {
        uint32_t v[32], s[32];
        __syncwarp();
        tmem_ld_32x32b_x32(taddr + c, v);
        tmem_ld_32x32b_x32(taddr + N + c, s);
        tmem_ld_wait();
        if (row < p.n_rows) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c + i < p.k) dst[c + i] = __uint_as_float(v[i]) + __uint_as_float(s[i]);
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

__global__ void add_count_kernel(double* count, double v) { count[0] += v; }

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace irp

using namespace irp;

extern "C" {

size_t irp_cov_workspace_bytes(int64_t n_rows, int dim) {
  if (n_rows <= 0 || dim <= 0) return 0;
  const size_t n_pad = align_up(static_cast<size_t>(n_rows), 64);
  return 3 * align_up(static_cast<size_t>(dim) * n_pad * sizeof(__nv_bfloat16), 1024) +
         align_up(n_pad / kSplitTile * static_cast<size_t>(dim) * sizeof(double), 1024) + 1024;
}

int irp_cov_accumulate(const float* d_x, int64_t n_rows, int dim, const float* d_shift, double* d_count,
                       double* d_sum, double* d_scatter, void* d_workspace, size_t workspace_bytes, void* stream) {
  IRP_REQUIRE(d_x && d_shift && d_count && d_sum && d_scatter && d_workspace, "cov_accumulate: null argument");
  IRP_REQUIRE(n_rows > 0 && dim > 0 && dim % 128 == 0, "cov_accumulate: n_rows %lld, dim %d (dim must be a multiple of 128)",
              static_cast<long long>(n_rows), dim);
  IRP_REQUIRE(workspace_bytes >= irp_cov_workspace_bytes(n_rows, dim), "cov_accumulate: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t n_pad = align_up(static_cast<size_t>(n_rows), 64);
  const size_t arr = align_up(static_cast<size_t>(dim) * n_pad * sizeof(__nv_bfloat16), 1024);
  uint8_t* ws = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<size_t>(d_workspace), 1024));
  __nv_bfloat16* at = reinterpret_cast<__nv_bfloat16*>(ws);
  __nv_bfloat16* bt = reinterpret_cast<__nv_bfloat16*>(ws + arr);
  __nv_bfloat16* ct = reinterpret_cast<__nv_bfloat16*>(ws + 2 * arr);
  double* col_part = reinterpret_cast<double*>(ws + 3 * arr);
  dim3 grid(static_cast<unsigned>(n_pad / kSplitTile), dim / kSplitTile);
  split_transpose_kernel<<<grid, 256, 0, st>>>(d_x, n_rows, dim, d_shift, static_cast<long long>(n_pad), at, bt, ct,
                                               col_part);
  IRP_CUDA_OK(cudaGetLastError());
  col_sum_reduce_kernel<<<ceil_div(dim, 32), 256, 0, st>>>(col_part, static_cast<int>(n_pad / kSplitTile), dim, d_sum);
  IRP_CUDA_OK(cudaGetLastError());
  add_count_kernel<<<1, 1, 0, st>>>(d_count, static_cast<double>(n_rows));
  IRP_CUDA_OK(cudaGetLastError());

  CovParams p;
  memset(&p, 0, sizeof(p));
  uint64_t dims[2] = {n_pad, static_cast<uint64_t>(dim)};
  uint64_t strides[1] = {n_pad * 2};
  uint32_t box[2] = {kCovBK, 128};
  IRP_TRY(encode_bf16_map(&p.tm[0], at, 2, dims, strides, box, 128));
  IRP_TRY(encode_bf16_map(&p.tm[1], bt, 2, dims, strides, box, 128));
  IRP_TRY(encode_bf16_map(&p.tm[2], ct, 2, dims, strides, box, 128));
  p.tiles_1d = dim / 128;
  p.n_tri = p.tiles_1d * (p.tiles_1d + 1) / 2;
  p.k_blocks_total = static_cast<int>(n_pad / kCovBK);
  // 512 samples per fp32 accumulation chain: measured 9e-5 rad subspace error at k=50 (32 blocks: 5e-4)
  p.k_blocks_per_chunk = 8;
  p.n_chunks = ceil_div(p.k_blocks_total, p.k_blocks_per_chunk);
  p.dim = dim;
  p.scatter = d_scatter;
  IRP_TRY(ensure_smem(cov_gemm_kernel, kCovSmem));
  const int grid_g = p.n_tri < num_sms() ? p.n_tri : num_sms();
  cov_gemm_kernel<<<grid_g, kCovThreads, kCovSmem, st>>>(p);
  IRP_CUDA_OK(cudaGetLastError());
  return IRP_OK;
}

}  // extern "C"

template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Top-k eigenpairs by Lanczos (section 4b).  *done = false asks the caller for the exact Householder path
// (breakdown, or no convergence within dim/2 steps).  Synchronises the stream once per convergence check.
static int lanczos_topk(const double* A, int n, int k, double* Q, double* diag, double* off, double* tnorm,
                        double* work, double* Z, double* d_eigenvalues, double* d_components, cudaStream_t st,
                        int first_check, int* steps_out, int* checks_out, bool* done) {
  *done = false;
  double* w0 = work;
  double* w1 = w0 + n;
  double* w2 = w1 + n;
  double* h1 = w2 + n;
  double* h2 = h1 + n;
  double* out = h2 + n;
  const double *cQ = Q, *cw0 = w0, *cw1 = w1, *cw2 = w2, *ch1 = h1, *ch2 = h2;
  int m_target = 3 * k + 10 > 96 ? 3 * k + 10 : 96;
  if (first_check > k) m_target = first_check;  // caller-chosen first convergence check (tests force an early one)
  const int m_cap = n / 2;
  if (m_target > m_cap) m_target = m_cap;
  int m = 0;
  IRP_CUDA_OK(launch_pdl(lz_start_kernel, dim3(1), dim3(1024), 0, st, w2, n));
  const size_t vec_smem = static_cast<size_t>(n) * sizeof(double);
  IRP_TRY(ensure_smem(lz_matvec_kernel, vec_smem));
  IRP_TRY(ensure_smem(lz_close_kernel, vec_smem));
  const dim3 axpy_grid(ceil_div(n, 32));
  int checks = 0;
  for (;;) {
    for (int j = m; j < m_target; ++j) {
      IRP_CUDA_OK(launch_pdl(lz_matvec_kernel, dim3(ceil_div(n, 8)), dim3(256), vec_smem, st, A, cw2, n, j, ch1, ch2, Q,
                             diag, off, w0));
      IRP_CUDA_OK(launch_pdl(lz_dots_kernel, dim3(j + 1), dim3(256), 0, st, cQ, cw0, n, h1));
      IRP_CUDA_OK(launch_pdl(lz_axpy_kernel, axpy_grid, dim3(256), 0, st, cQ, ch1, cw0, n, j + 1, w1));
      IRP_CUDA_OK(launch_pdl(lz_dots_kernel, dim3(j + 1), dim3(256), 0, st, cQ, cw1, n, h2));
      IRP_CUDA_OK(launch_pdl(lz_axpy_kernel, axpy_grid, dim3(256), 0, st, cQ, ch2, cw1, n, j + 1, w2));
    }
    IRP_CUDA_OK(launch_pdl(lz_close_kernel, dim3(1), dim3(256), vec_smem, st, cw2, n, m_target, ch1, ch2, Q, diag, off));
    IRP_CUDA_OK(cudaGetLastError());
    m = m_target;
    const size_t bis_smem = 2 * static_cast<size_t>(m) * sizeof(double);
    IRP_TRY(ensure_smem(bisect_topk_kernel, bis_smem));
    bisect_topk_kernel<<<k, kBisThreads, bis_smem, st>>>(diag, off, m, k, d_eigenvalues, tnorm);
    const size_t ii_smem = 6 * static_cast<size_t>(m) * sizeof(double) + static_cast<size_t>(m) + 16;
    IRP_TRY(ensure_smem(inverse_iteration_kernel, ii_smem));
    inverse_iteration_kernel<<<k, 32, ii_smem, st>>>(diag, off, m, k, d_eigenvalues, tnorm, Z);
    cluster_mgs_kernel<<<1, 256, 0, st>>>(Z, m, k, d_eigenvalues, tnorm);
    lz_residual_kernel<<<1, 256, 0, st>>>(Z, off, m, k, tnorm, out);
    IRP_CUDA_OK(cudaGetLastError());
    double host[3] = {0.0, 0.0, 0.0};
    IRP_CUDA_OK(cudaMemcpyAsync(host, out, sizeof(host), cudaMemcpyDeviceToHost, st));
    IRP_CUDA_OK(cudaStreamSynchronize(st));
    const double tn = host[2] > 0.0 ? host[2] : 1.0;
    ++checks;
    *steps_out = m;
    *checks_out = checks;
    if (!(host[1] > 1e-10 * tn) || !(host[0] == host[0])) return IRP_OK;  // breakdown (invariant subspace) or NaN
    if (host[0] <= 1e-12 * tn) {
      const size_t ritz_smem = static_cast<size_t>(m) * sizeof(double);
      IRP_TRY(ensure_smem(lz_ritz_kernel, ritz_smem));
      lz_ritz_kernel<<<dim3(k, ceil_div(n, 256)), 256, ritz_smem, st>>>(Q, Z, n, m, d_components);
      lz_sign_kernel<<<k, 256, 0, st>>>(d_components, n);
      IRP_CUDA_OK(cudaGetLastError());
      *done = true;
      return IRP_OK;
    }
    if (m >= m_cap) return IRP_OK;
    m_target = m + 64 < m_cap ? m + 64 : m_cap;
  }
}

extern "C" {

size_t irp_pca_fit_workspace_bytes(int dim, int k) {
  if (dim <= 0 || k <= 0) return 0;
  const size_t n = static_cast<size_t>(dim);
  size_t b = 0;
  b += n * n * 8;                             // covariance / working matrix
  b += n * n * 8;                             // reflectors V
  b += 8 * n * 8;                             // tau, diag, off, pbuf[2], spare
  b += static_cast<size_t>(k) * 5 * n * 8;    // inverse-iteration scratch
  b += static_cast<size_t>(k) * n * 8;        // Z (tridiagonal eigenvectors)
  b += 1024;
  return b;
}

int irp_pca_fit(const double* d_count, const double* d_sum, const double* d_scatter, const float* d_shift, int dim,
                int k, double* d_mean, double* d_components, double* d_eigenvalues, void* d_workspace,
                size_t workspace_bytes, void* stream) {
  return irp_pca_fit_ex(d_count, d_sum, d_scatter, d_shift, dim, k, d_mean, d_components, d_eigenvalues, d_workspace,
                        workspace_bytes, IRP_PCA_SOLVER_AUTO, 0, nullptr, stream);
}

int irp_pca_fit_ex(const double* d_count, const double* d_sum, const double* d_scatter, const float* d_shift, int dim,
                   int k, double* d_mean, double* d_components, double* d_eigenvalues, void* d_workspace,
                   size_t workspace_bytes, int solver, int lanczos_first_check, int32_t* h_info, void* stream) {
  IRP_REQUIRE(solver == IRP_PCA_SOLVER_AUTO || solver == IRP_PCA_SOLVER_LANCZOS || solver == IRP_PCA_SOLVER_HOUSEHOLDER,
              "pca_fit: unknown solver %d", solver);
  IRP_REQUIRE(d_count && d_sum && d_scatter && d_shift && d_mean && d_components && d_eigenvalues && d_workspace,
              "pca_fit: null argument");
  IRP_REQUIRE(dim >= 4 && dim <= 4096 && k >= 1 && k <= dim && k <= 512, "pca_fit: dim %d / k %d unsupported", dim, k);
  IRP_REQUIRE(workspace_bytes >= irp_pca_fit_workspace_bytes(dim, k), "pca_fit: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int n = dim;
  double* ws = reinterpret_cast<double*>(align_up(reinterpret_cast<size_t>(d_workspace), 256));
  double* A = ws;
  double* V = A + static_cast<size_t>(n) * n;
  double* tau = V + static_cast<size_t>(n) * n;
  double* diag = tau + n;
  double* off = diag + n;
  double* pbuf = off + n;       // [2][n]
  double* scal = pbuf + 2 * n;  // [n]: scal[0] = trace, scal[1] = ||T||
  double* work = scal + 3 * n;
  double* Z = work + static_cast<size_t>(k) * 5 * n;

  IRP_CUDA_OK(cudaMemsetAsync(tau, 0, 8 * static_cast<size_t>(n) * sizeof(double), st));
  assemble_cov_kernel<<<num_sms() * 4, 256, 0, st>>>(d_count, d_sum, d_scatter, d_shift, n, d_mean, A);
  IRP_CUDA_OK(cudaGetLastError());
  trace_kernel<<<1, 1024, 0, st>>>(A, n, scal);
  IRP_CUDA_OK(cudaGetLastError());

  bool done = false;
  int lz_steps = 0, lz_checks = 0;
  const bool lanczos_fits = n >= 512 && k * 8 <= n && k >= 5;
  IRP_REQUIRE(solver != IRP_PCA_SOLVER_LANCZOS || lanczos_fits,
              "pca_fit: the Lanczos solver needs dim >= 512 and 5 <= k <= dim / 8 (dim %d, k %d)", dim, k);
  if (solver != IRP_PCA_SOLVER_HOUSEHOLDER && lanczos_fits)
    IRP_TRY(lanczos_topk(A, n, k, V, diag, off, scal + 1, work, Z, d_eigenvalues, d_components, st,
                         lanczos_first_check, &lz_steps, &lz_checks, &done));
  if (h_info) {
    h_info[0] = done ? IRP_PCA_SOLVER_LANCZOS : IRP_PCA_SOLVER_HOUSEHOLDER;
    h_info[1] = lz_steps;
    h_info[2] = lz_checks;
    h_info[3] = 0;
  }
  if (!done) {
  // ---- tridiagonalisation: launches j = -1 .. n-3 ----
  const size_t tri_smem = 3 * static_cast<size_t>(n) * sizeof(double);
  IRP_TRY(ensure_smem(tridiag_step_kernel, tri_smem));
  const int rows_per_cta = kTriThreads / 32;
  for (int j = -1; j <= n - 3; ++j) {
    const int rows = n - (j + 2);
    int grid = ceil_div(rows > 0 ? rows : 1, rows_per_cta);
    const int cap = num_sms() * 2;
    if (grid > cap) grid = cap;
    tridiag_step_kernel<<<grid, kTriThreads, tri_smem, st>>>(A, n, j, V, tau, diag, off, pbuf);
  }
  IRP_CUDA_OK(cudaGetLastError());
  // last diagonal entry: after launch j = n-3 the (n-1,n-1) element is final
  IRP_CUDA_OK(cudaMemcpyAsync(diag + (n - 1), A + static_cast<size_t>(n - 1) * n + (n - 1), sizeof(double),
                              cudaMemcpyDeviceToDevice, st));
  // d[0] is produced by launch j=-1 (t=0) as A[0][0] ✓

  // ---- eigenvalues (top-k) ----
  const size_t bis_smem = 2 * static_cast<size_t>(n) * sizeof(double);
  IRP_TRY(ensure_smem(bisect_topk_kernel, bis_smem));
  bisect_topk_kernel<<<k, kBisThreads, bis_smem, st>>>(diag, off, n, k, d_eigenvalues, scal + 1);
  IRP_CUDA_OK(cudaGetLastError());
  // ---- eigenvectors of T ----
  const size_t ii_smem = 6 * static_cast<size_t>(n) * sizeof(double) + static_cast<size_t>(n) + 16;
  IRP_TRY(ensure_smem(inverse_iteration_kernel, ii_smem));
  (void)work;
  inverse_iteration_kernel<<<k, 32, ii_smem, st>>>(diag, off, n, k, d_eigenvalues, scal + 1, Z);
  IRP_CUDA_OK(cudaGetLastError());
  cluster_mgs_kernel<<<1, 256, 0, st>>>(Z, n, k, d_eigenvalues, scal + 1);
  IRP_CUDA_OK(cudaGetLastError());
  // ---- back-transform + sign ----
  back_transform_kernel<<<k, kBtThreads, 0, st>>>(V, tau, n, Z, d_components);
  IRP_CUDA_OK(cudaGetLastError());
  }
  clip_evals_kernel<<<ceil_div(k, 128), 128, 0, st>>>(d_eigenvalues, k);
  IRP_CUDA_OK(cudaGetLastError());
  // total variance (trace of C) is returned right after the k eigenvalues
  IRP_CUDA_OK(cudaMemcpyAsync(d_eigenvalues + k, scal, sizeof(double), cudaMemcpyDeviceToDevice, st));
  return IRP_OK;
}

constexpr long long kPjChunkRows = 32768;  // rows per pass through the workspace

size_t irp_pca_transform_workspace_bytes(int64_t n_rows, int dim, int k) {
  if (n_rows <= 0 || dim <= 0 || k <= 0) return 0;
  const size_t rows = static_cast<size_t>(n_rows < kPjChunkRows ? (n_rows + 127) / 128 * 128 : kPjChunkRows);
  const size_t n_cols = k <= 64 ? 64 : 128;
  return 3 * align_up(rows * dim * sizeof(__nv_bfloat16), 1024) + 3 * align_up(n_cols * dim * sizeof(__nv_bfloat16), 1024) +
         1024;
}

}  // extern "C"

template <int N>
static int launch_proj(const ProjParams& p, cudaStream_t st) {
  IRP_TRY(ensure_smem(proj_gemm_kernel<N>, PjSmem<N>::kTotal));
  const int grid = p.n_tiles < num_sms() ? p.n_tiles : num_sms();
  proj_gemm_kernel<N><<<grid, kPjThreads, PjSmem<N>::kTotal, st>>>(p);
  IRP_CUDA_OK(cudaGetLastError());
  return IRP_OK;
}

extern "C" {

int irp_pca_transform(const float* d_x, int64_t n_rows, int dim, const double* d_mean, const double* d_components,
                      int k, float* d_z, void* d_workspace, size_t workspace_bytes, void* stream) {
  IRP_REQUIRE(d_x && d_mean && d_components && d_z && d_workspace, "pca_transform: null argument");
  IRP_REQUIRE(n_rows > 0 && dim > 0 && dim % kPjBK == 0 && k >= 1 && k <= 128,
              "pca_transform: n_rows %lld dim %d k %d unsupported (k <= 128, dim %% 64 == 0)",
              static_cast<long long>(n_rows), dim, k);
  IRP_REQUIRE(workspace_bytes >= irp_pca_transform_workspace_bytes(n_rows, dim, k), "pca_transform: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t rows_cap = static_cast<size_t>(n_rows < kPjChunkRows ? (n_rows + 127) / 128 * 128 : kPjChunkRows);
  const int n_cols = k <= 64 ? 64 : 128;
  const size_t y_bytes = align_up(rows_cap * dim * sizeof(__nv_bfloat16), 1024);
  const size_t v_bytes = align_up(static_cast<size_t>(n_cols) * dim * sizeof(__nv_bfloat16), 1024);
  uint8_t* ws = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<size_t>(d_workspace), 1024));
  __nv_bfloat16* y[3];
  __nv_bfloat16* v[3];
  for (int i = 0; i < 3; ++i) {
    y[i] = reinterpret_cast<__nv_bfloat16*>(ws + i * y_bytes);
    v[i] = reinterpret_cast<__nv_bfloat16*>(ws + 3 * y_bytes + i * v_bytes);
  }
  project_split_comps_kernel<<<num_sms(), 256, 0, st>>>(d_components, k, dim, n_cols, v[0], v[1], v[2]);
  IRP_CUDA_OK(cudaGetLastError());
  ProjParams p;
  memset(&p, 0, sizeof(p));
  for (int i = 0; i < 3; ++i) {
    uint64_t vd[2] = {static_cast<uint64_t>(dim), static_cast<uint64_t>(n_cols)};
    uint64_t vs[1] = {static_cast<uint64_t>(dim) * 2};
    uint32_t vbox[2] = {kPjBK, static_cast<uint32_t>(n_cols)};
    IRP_TRY(encode_bf16_map(&p.tmV[i], v[i], 2, vd, vs, vbox, 128));
  }
  p.k = k;
  p.k_blocks = dim / kPjBK;
  for (long long r0 = 0; r0 < n_rows; r0 += kPjChunkRows) {
    const long long rows = n_rows - r0 < kPjChunkRows ? n_rows - r0 : kPjChunkRows;
    const long long n_pad = (rows + 127) / 128 * 128;
    const long long total = n_pad * (dim / 4);
    const unsigned blocks = static_cast<unsigned>(total / 256 < static_cast<long long>(num_sms()) * 32
                                                      ? (total + 255) / 256
                                                      : static_cast<long long>(num_sms()) * 32);
    project_split_kernel<<<blocks, 256, 0, st>>>(d_x + r0 * dim, rows, dim, d_mean, y[0], y[1], y[2], n_pad);
    IRP_CUDA_OK(cudaGetLastError());
    for (int i = 0; i < 3; ++i) {
      uint64_t yd[2] = {static_cast<uint64_t>(dim), static_cast<uint64_t>(n_pad)};
      uint64_t ys[1] = {static_cast<uint64_t>(dim) * 2};
      uint32_t ybox[2] = {kPjBK, 128};
      IRP_TRY(encode_bf16_map(&p.tmY[i], y[i], 2, yd, ys, ybox, 128));
    }
    p.z = d_z + r0 * k;
    p.n_rows = rows;
    p.n_tiles = static_cast<int>(n_pad / 128);
    IRP_TRY(n_cols == 64 ? launch_proj<64>(p, st) : launch_proj<128>(p, st));
  }
  return IRP_OK;
}

}  // extern "C"
