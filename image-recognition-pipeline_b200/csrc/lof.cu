// A4 outlier scoring: Local Outlier Factor (per group and global) and the distance-to-centroid z-score scorer.
//
// Replaces detect_outliers (functions/data_curation.py:709-728).  LOF arithmetic follows scikit-learn
// (sklearn/neighbors/_lof.py:286-323 fit, :498-523 local reachability density): k = max(1, min(n_neighbors, n-1)),
// neighbours exclude the sample itself, reach = max(d, k-distance(neighbour)), lrd = 1/(mean reach + 1e-10),
// score = -mean(lrd[nbr]/lrd[i]), offset = np.percentile(score, 100*contamination) (linear interpolation),
// outlier iff score < offset.  Everything is evaluated in fp64; oracle/lof_ref.py is the numpy restatement.
//
// Pipeline: stable counting sort of the rows by group -> squared norms -> tiled brute-force k-NN (64x64 distance
// tiles in registers, per-query sorted candidate lists in shared memory) -> lrd -> lof -> per-group radix select
// of the two order statistics around the percentile -> flags.
#include <cfloat>
#include <cstdlib>
#include <cstring>

#include <cub/device/device_radix_sort.cuh>

#include "common.h"

namespace irp {

constexpr int kSortThreads = 1024;
constexpr int kKnnThreads = 256;
constexpr int kKnnTile = 64;   // queries per CTA, candidates per step
constexpr int kKnnDChunk = 32; // feature chunk staged in shared memory
constexpr int kMaxK = 128;

// ------------------------------------------------------------------------------------------------------------
// stable counting sort by group: thread t owns the contiguous row range [t*L, (t+1)*L)
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSortThreads) group_count_kernel(const int32_t* __restrict__ group, long long n,
                                                                   int n_groups, int32_t* __restrict__ counts) {
  // counts: [n_groups][kSortThreads]
  const long long L = (n + kSortThreads - 1) / kSortThreads;
  const int t = threadIdx.x;
  for (int g = 0; g < n_groups; ++g) counts[static_cast<size_t>(g) * kSortThreads + t] = 0;
  const long long lo = t * L, hi = min(n, lo + L);
  for (long long i = lo; i < hi; ++i) {
    const int g = group ? group[i] : 0;
    if (g >= 0 && g < n_groups) counts[static_cast<size_t>(g) * kSortThreads + t] += 1;
  }
}

// exclusive scan over counts in (group, thread) order; gstart[g] = first sorted position of group g
__global__ void __launch_bounds__(kSortThreads) group_scan_kernel(int32_t* __restrict__ counts, int n_groups,
                                                                  int32_t* __restrict__ gstart) {
  __shared__ int32_t warp_tot[kSortThreads / 32];
  __shared__ int32_t carry;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (t == 0) carry = 0;
  __syncthreads();
  for (int g = 0; g < n_groups; ++g) {
    const int32_t v = counts[static_cast<size_t>(g) * kSortThreads + t];
    int32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    int32_t base = carry;
    for (int w = 0; w < warp; ++w) base += warp_tot[w];
    counts[static_cast<size_t>(g) * kSortThreads + t] = base + x - v;  // exclusive
    __syncthreads();
    if (t == kSortThreads - 1) {
      gstart[g] = carry;
      carry = base + x;
    }
    __syncthreads();
  }
  if (t == 0) gstart[n_groups] = carry;
}

__global__ void __launch_bounds__(kSortThreads) group_scatter_kernel(const int32_t* __restrict__ group, long long n,
                                                                     int n_groups, int32_t* __restrict__ counts,
                                                                     int32_t* __restrict__ order) {
  const long long L = (n + kSortThreads - 1) / kSortThreads;
  const int t = threadIdx.x;
  const long long lo = t * L, hi = min(n, lo + L);
  for (long long i = lo; i < hi; ++i) {
    const int g = group ? group[i] : 0;
    if (g >= 0 && g < n_groups) {
      const int32_t pos = counts[static_cast<size_t>(g) * kSortThreads + t]++;
      order[pos] = static_cast<int32_t>(i);
    }
  }
}

__global__ void iota_kernel(int32_t* order, long long n, int32_t* gstart) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) order[i] = static_cast<int32_t>(i);
  if (i == 0) {
    gstart[0] = 0;
    gstart[1] = static_cast<int32_t>(n);
  }
}

// squared norms in sorted order (fp64)
__global__ void sqnorm_kernel(const float* __restrict__ z, const int32_t* __restrict__ order, long long n, int dim,
                              double* __restrict__ sq, unsigned long long* __restrict__ max_sq_bits) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const float* r = z + static_cast<size_t>(order[i]) * dim;
  double s = 0.0;
  for (int j = 0; j < dim; ++j) {
    const double v = r[j];
    s += v * v;
  }
  sq[i] = s;
  // non-negative doubles order like their bit patterns
  if (max_sq_bits != nullptr) atomicMax(max_sq_bits, static_cast<unsigned long long>(__double_as_longlong(s)));
}

// sort key of a row: (group, first coordinate) -- rows of a group end up ordered by z[.,0], which lets the search
// stop scanning once the gap in that coordinate alone exceeds the current k-th distance
__global__ void sort_key_kernel(const float* __restrict__ z, const int32_t* __restrict__ group, long long n, int dim,
                                unsigned long long* __restrict__ keys, int32_t* __restrict__ vals) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  uint32_t u = __float_as_uint(z[static_cast<size_t>(i) * dim]);
  u ^= (u >> 31) ? 0xFFFFFFFFu : 0x80000000u;  // order-preserving map of float to uint32
  const unsigned long long g = group ? static_cast<unsigned long long>(static_cast<uint32_t>(group[i])) : 0ull;
  keys[i] = (g << 32) | u;
  vals[i] = static_cast<int32_t>(i);
}

// ------------------------------------------------------------------------------------------------------------
// brute-force k-NN inside each group
// ------------------------------------------------------------------------------------------------------------
// warp-cooperative insertion of (dc, jc) into the ascending list (ld, li) of length K held in shared memory
__device__ __forceinline__ void list_insert(double* ld, int32_t* li, int K, double dc, int32_t jc, int lane) {
  int pos = 0;
  for (int s0 = 0; s0 < K; s0 += 32) {
    const int s = s0 + lane;
    // position by (distance, candidate index): the result does not depend on the order candidates arrive in
    const bool le = (s < K) && (ld[s] < dc || (ld[s] == dc && li[s] < jc));
    pos += __popc(__ballot_sync(0xffffffffu, le));
  }
  // shift [pos, K-2] -> [pos+1, K-1]
  double vd[kMaxK / 32];
  int32_t vi[kMaxK / 32];
#pragma unroll
  for (int m = 0; m < kMaxK / 32; ++m) {
    const int s = m * 32 + lane;
    if (s >= pos && s < K - 1) {
      vd[m] = ld[s];
      vi[m] = li[s];
    }
  }
  __syncwarp();
#pragma unroll
  for (int m = 0; m < kMaxK / 32; ++m) {
    const int s = m * 32 + lane;
    if (s >= pos && s < K - 1) {
      ld[s + 1] = vd[m];
      li[s + 1] = vi[m];
    }
  }
  if (lane == 0) {
    ld[pos] = dc;
    li[pos] = jc;
  }
  __syncwarp();
}

// dynamic smem: Qs[kKnnDChunk][64], Cs[kKnnDChunk][64], Dt[64][65], ld[64][K], li[64][K]
__global__ void __launch_bounds__(kKnnThreads) knn_kernel(const float* __restrict__ z,
                                                          const int32_t* __restrict__ order,
                                                          const int32_t* __restrict__ gstart, int n_groups, int dim,
                                                          int k, const double* __restrict__ sq, int part,
                                                          int n_parts, double* __restrict__ knn_d,
                                                          int32_t* __restrict__ knn_i, double* __restrict__ kdist) {
  extern __shared__ double shk[];
  double* Qs = shk;
  double* Cs = Qs + kKnnDChunk * kKnnTile;
  double* Dt = Cs + kKnnDChunk * kKnnTile;
  double* ld = Dt + kKnnTile * (kKnnTile + 1);
  int32_t* li = reinterpret_cast<int32_t*>(ld + kKnnTile * k);

  // query tiles are dealt round-robin to the parts (ranks); this launch only runs the tiles of `part`
  if (static_cast<int>(blockIdx.x % n_parts) != part) return;
  // locate this CTA's (group, query tile)
  int tile = blockIdx.x;
  int g = 0, g0 = 0, g1 = 0;
  for (; g < n_groups; ++g) {
    g0 = gstart[g];
    g1 = gstart[g + 1];
    const int tiles = (g1 - g0 + kKnnTile - 1) / kKnnTile;
    if (tile < tiles) break;
    tile -= tiles;
  }
  if (g >= n_groups) return;
  const int ng = g1 - g0;
  const int q0 = g0 + tile * kKnnTile;
  const int K = max(1, min(k, ng - 1));  // _lof.py:293

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < kKnnTile * k; i += kKnnThreads) {
    ld[i] = INFINITY;
    li[i] = -1;
  }
  const int ty = tid >> 4, tx = tid & 15;
  for (int c0 = g0; c0 < g1; c0 += kKnnTile) {
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
    for (int j0 = 0; j0 < dim; j0 += kKnnDChunk) {
      __syncthreads();
      for (int i = tid; i < kKnnTile * kKnnDChunk; i += kKnnThreads) {
        const int r = i / kKnnDChunk, jj = i % kKnnDChunk;
        const int j = j0 + jj;
        double qv = 0.0, cv = 0.0;
        if (j < dim) {
          if (q0 + r < g1) qv = z[static_cast<size_t>(order[q0 + r]) * dim + j];
          if (c0 + r < g1) cv = z[static_cast<size_t>(order[c0 + r]) * dim + j];
        }
        Qs[jj * kKnnTile + r] = qv;
        Cs[jj * kKnnTile + r] = cv;
      }
      __syncthreads();
#pragma unroll 8
      for (int jj = 0; jj < kKnnDChunk; ++jj) {
        const double2 qa = *reinterpret_cast<const double2*>(&Qs[jj * kKnnTile + ty * 4]);
        const double2 qb = *reinterpret_cast<const double2*>(&Qs[jj * kKnnTile + ty * 4 + 2]);
        const double2 ca = *reinterpret_cast<const double2*>(&Cs[jj * kKnnTile + tx * 4]);
        const double2 cb = *reinterpret_cast<const double2*>(&Cs[jj * kKnnTile + tx * 4 + 2]);
        const double q[4] = {qa.x, qa.y, qb.x, qb.y};
        const double c[4] = {ca.x, ca.y, cb.x, cb.y};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = fma(q[a], c[b], acc[a][b]);
      }
    }
    // squared distances -> Dt
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int qi = q0 + ty * 4 + a;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int ci = c0 + tx * 4 + b;
        double d2 = INFINITY;
        if (qi < g1 && ci < g1 && ci != qi) d2 = fmax(sq[qi] + sq[ci] - 2.0 * acc[a][b], 0.0);
        Dt[(ty * 4 + a) * (kKnnTile + 1) + tx * 4 + b] = d2;
      }
    }
    __syncthreads();
    // selection: warp w owns queries w*8 .. w*8+7
    for (int qq = 0; qq < 8; ++qq) {
      const int ql = warp * 8 + qq;
      if (q0 + ql >= g1) break;
      double* qld = ld + ql * k;
      int32_t* qli = li + ql * k;
      double tau = qld[K - 1];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const double dc = Dt[ql * (kKnnTile + 1) + half * 32 + lane];
        unsigned pending = __ballot_sync(0xffffffffu, dc < tau);
        while (pending) {
          const int src = __ffs(pending) - 1;
          pending &= pending - 1;
          const double dv = __shfl_sync(0xffffffffu, dc, src);
          if (dv < tau) {
            list_insert(qld, qli, K, dv, c0 + half * 32 + src, lane);
            tau = qld[K - 1];
          }
        }
      }
    }
  }
  __syncthreads();
  // write out: distances (sqrt) + neighbour positions (sorted-order indices)
  for (int i = tid; i < kKnnTile * k; i += kKnnThreads) {
    const int ql = i / k, s = i % k;
    if (q0 + ql < g1) {
      knn_d[static_cast<size_t>(q0 + ql) * k + s] = s < K ? sqrt(ld[ql * k + s]) : INFINITY;
      knn_i[static_cast<size_t>(q0 + ql) * k + s] = s < K ? li[ql * k + s] : -1;
      if (s == K - 1) kdist[q0 + ql] = sqrt(ld[ql * k + s]);  // k-distance of the row (_lof.py:305)
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Fast path (dim <= kKnnMaxResidentDim): rows are first gathered into group-sorted order as fp64 [n][dpad] (zero
// padded to a multiple of 8 features), so the search kernel reads contiguous rows without the order[] indirection;
// the CTA's 64 query rows stay resident in shared memory for its whole life, candidate tiles are streamed with a
// register prefetch of the next tile behind the 64 x 64 x dpad fp64 FMA block of the current one.
// ------------------------------------------------------------------------------------------------------------
constexpr int kKnnMaxResidentDim = 128;

__global__ void gather_sorted_kernel(const float* __restrict__ z, const int32_t* __restrict__ order, long long n,
                                     int dim, int dpad, double* __restrict__ zs, float* __restrict__ zs32) {
  const long long total = n * dpad;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / dpad;
    const int j = static_cast<int>(i - r * dpad);
    const float v = j < dim ? z[static_cast<size_t>(order[r]) * dim + j] : 0.f;
    zs[i] = static_cast<double>(v);
    zs32[i] = v;
  }
}

// ------------------------------------------------------------------------------------------------------------
// Filtered search (default): the 64 x 128 x dpad distance block is evaluated in fp32 (FFMA, 4 x 8 register tiles)
// and used only as a CONSERVATIVE FILTER: a candidate is re-evaluated in fp64 -- with exactly the arithmetic of
// an exhaustive fp64 search, so the neighbour sets are identical -- iff  d2_fp32 <= tau + E (|q|^2 + |c|^2), where tau is the
// query's current k-th best squared distance and E bounds the fp32 evaluation error of |q|^2 + |c|^2 - 2 q.c
// (inputs are exact: the rows ARE float32; the dot product contributes at most dpad 2^-24 sum|q_j c_j| <=
// dpad 2^-25 (|q|^2+|c|^2), the norms and the final three operations a few ulps of |q|^2+|c|^2; E = 4 (dpad+8)
// 2^-24 leaves a factor > 4).  Every true neighbour passes (its exact d2 <= final tau <= current tau), so the
// k-NN lists equal those of the exhaustive fp64 search; after the first tile only ~k ln(n/k) candidates per query
// take the fp64 path.
// dynamic smem: Qs[dpad][64] f32, Cs[dpad][128] f32, csq[128] f32, tauf[64] f32, wl[8][512] u16, wlc[8] i32,
//               ld[64][k] f64, li[64][k] i32
// ------------------------------------------------------------------------------------------------------------
// k-best set kept UNSORTED with its maximum tracked: an accepted candidate overwrites the maximum and the new
// maximum is found by the whole warp (k/32 elements per lane + 5 shuffle steps).  Order is (distance, index).
struct KBestTop {
  double d;
  int32_t idx;
  int32_t pos;
};
__device__ __forceinline__ bool kb_greater(double d0, int32_t i0, double d1, int32_t i1) {
  return d0 > d1 || (d0 == d1 && i0 > i1);
}
__device__ __forceinline__ void kbest_replace_max(double* ld, int32_t* li, int K, double dv, int32_t jc, int lane,
                                                  KBestTop* top) {
  if (lane == 0) {
    ld[top->pos] = dv;
    li[top->pos] = jc;
  }
  __syncwarp();
  double md = -1.0;  // squared distances are >= 0
  int32_t mi = -2, mp = 0;
  for (int s = lane; s < K; s += 32) {
    const double d = ld[s];
    const int32_t i = li[s];
    if (kb_greater(d, i, md, mi)) {
      md = d;
      mi = i;
      mp = s;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double od = __shfl_xor_sync(0xffffffffu, md, o);
    const int32_t oi = __shfl_xor_sync(0xffffffffu, mi, o);
    const int32_t op = __shfl_xor_sync(0xffffffffu, mp, o);
    if (kb_greater(od, oi, md, mi) || (od == md && oi == mi && op < mp)) {
      md = od;
      mi = oi;
      mp = op;
    }
  }
  if (lane == 0) {
    top->d = md;
    top->idx = mi;
    top->pos = mp;
  }
  __syncwarp();
}

constexpr int kFltCand = 128;   // candidates per step
constexpr int kFltCap = 256;    // worklist entries per warp

__global__ void __launch_bounds__(kKnnThreads, 2) knn_filter_kernel(const double* __restrict__ zs,
                                                                 const float* __restrict__ zs32,
                                                                 const int32_t* __restrict__ gstart, int n_groups,
                                                                 int dpad, int k, const double* __restrict__ sq,
                                                                 const double* __restrict__ max_sq, int part,
                                                                 int n_parts, double* __restrict__ knn_d,
                                                                 int32_t* __restrict__ knn_i,
                                                                 double* __restrict__ kdist) {
  extern __shared__ double shk[];
  double* ld = shk;                                          // [64][k]
  int32_t* li = reinterpret_cast<int32_t*>(ld + kKnnTile * k);  // [64][k]
  float* Qs = reinterpret_cast<float*>(li + kKnnTile * k + ((kKnnTile * k) & 1));  // 8-byte aligned
  float* Cs = Qs + dpad * kKnnTile;
  float* csq = Cs + dpad * kFltCand;
  float* tauf = csq + kFltCand;
  uint16_t* wl = reinterpret_cast<uint16_t*>(tauf + kKnnTile);
  int* wlc = reinterpret_cast<int*>(wl + 8 * kFltCap);
  KBestTop* tops = reinterpret_cast<KBestTop*>(wlc + 8);  // [64], 16-byte entries
  double* wmax = reinterpret_cast<double*>(tops + kKnnTile);  // [8] per-warp max of the current k-th distances

  if (static_cast<int>(blockIdx.x % n_parts) != part) return;
  int tile = blockIdx.x;
  int g = 0, g0 = 0, g1 = 0;
  for (; g < n_groups; ++g) {
    g0 = gstart[g];
    g1 = gstart[g + 1];
    const int tiles = (g1 - g0 + kKnnTile - 1) / kKnnTile;
    if (tile < tiles) break;
    tile -= tiles;
  }
  if (g >= n_groups) return;
  const int ng = g1 - g0;
  const int q0 = g0 + tile * kKnnTile;
  const int K = max(1, min(k, ng - 1));  // _lof.py:293
  const float E = 4.0f * static_cast<float>(dpad + 8) * 5.9604645e-8f;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < kKnnTile * k; i += kKnnThreads) {
    ld[i] = INFINITY;
    li[i] = -1;
  }
  if (tid < kKnnTile) {
    tauf[tid] = INFINITY;
    tops[tid].d = INFINITY;
    tops[tid].idx = -1;
    tops[tid].pos = 0;
  }
  if (tid < 8) wlc[tid] = 0;
  // query tile, transposed: thread (r = tid % 64, h = tid / 64) moves float4 index h + 4u of row r
  const int n4 = dpad >> 2;
  {
    const int r = tid & 63, h = tid >> 6;
    const float* row = zs32 + static_cast<size_t>(min(q0 + r, g1 - 1)) * dpad;
    for (int jq = h; jq < n4; jq += 4) {
      const float4 v = *reinterpret_cast<const float4*>(row + 4 * jq);
      Qs[(4 * jq + 0) * kKnnTile + r] = v.x;
      Qs[(4 * jq + 1) * kKnnTile + r] = v.y;
      Qs[(4 * jq + 2) * kKnnTile + r] = v.z;
      Qs[(4 * jq + 3) * kKnnTile + r] = v.w;
    }
  }
  // candidate tiles: thread (r = tid % 128, h = tid / 128) moves float4 index h + 2u of row r
  constexpr int kMaxU = kKnnMaxResidentDim / 8;
  const int cr = tid & 127, ch = tid >> 7;
  const int n_u = dpad >> 3;
  float4 pre[kMaxU];
  float pre_sq = 0.f;
  auto prefetch = [&](int c0) {
    const int rowi = min(c0 + cr, g1 - 1);
    const float* row = zs32 + static_cast<size_t>(rowi) * dpad;
#pragma unroll
    for (int u = 0; u < kMaxU; ++u)
      if (u < n_u) pre[u] = __ldg(reinterpret_cast<const float4*>(row + 4 * (ch + 2 * u)));
    if (ch == 0) pre_sq = static_cast<float>(sq[rowi]);
  };
  const int ty = tid >> 4, tx = tid & 15;  // 4 queries x 8 candidates per thread
  float sqq[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) sqq[a] = static_cast<float>(sq[min(q0 + ty * 4 + a, g1 - 1)]);

  // Rows of a group are sorted by their first coordinate z0, and d^2 >= (z0_q - z0_c)^2.  Candidate tiles are
  // visited outward from the tile that holds the queries, alternating right / left; a side is finished for good as
  // soon as the z0 gap between its next tile and the query range exceeds the largest current k-th distance of the
  // tile's queries (with a margin far above the rounding error of the evaluated distances).  On PCA scores z0 is
  // the max-variance direction, so most of the O(n^2) work disappears -- and what remains does not grow with n.
  const int n_ctiles = (ng + kFltCand - 1) / kFltCand;
  const int tq = (q0 - g0) / kFltCand;
  const int q_last = min(q0 + kKnnTile, g1) - 1;
  const double z_lo = zs[static_cast<size_t>(q0) * dpad], z_hi = zs[static_cast<size_t>(q_last) * dpad];
  const double abs_margin = 1e-9 * (4.0 * max_sq[0] + 1.0);
  if (tid < 8) wmax[tid] = INFINITY;
  int t_left = tq, t_right = tq;       // tiles [t_left, t_right] have been visited
  bool go_right = true;                // side to try first at the next step
  bool left_done = (tq == 0), right_done = (tq == n_ctiles - 1);
  // next tile to visit given the pruning threshold, or -1
  auto next_tile = [&](double tau_max) {
    const double thr = tau_max * (1.0 + 1e-6) + abs_margin;
    if (!right_done) {
      const double gap = zs[static_cast<size_t>(g0 + (t_right + 1) * kFltCand) * dpad] - z_hi;
      if (gap > 0.0 && gap * gap > thr) right_done = true;
    }
    if (!left_done) {
      const double gap = z_lo - zs[static_cast<size_t>(g0 + t_left * kFltCand - 1) * dpad];
      if (gap > 0.0 && gap * gap > thr) left_done = true;
    }
    int t = -1;
    if (go_right ? !right_done : left_done && !right_done) t = ++t_right;
    else if (!left_done) t = --t_left;
    if (t >= 0) {
      go_right = !go_right;
      if (t_right == n_ctiles - 1) right_done = true;
      if (t_left == 0) left_done = true;
    }
    return t;
  };
  int cur = tq;
  prefetch(g0 + cur * kFltCand);
  bool first = true;
  while (cur >= 0) {
    const int c0 = g0 + cur * kFltCand;
    __syncthreads();  // previous tile: FMA block done with Cs, selection done with tauf / worklists / wmax
#pragma unroll
    for (int u = 0; u < kMaxU; ++u)
      if (u < n_u) {
        const int jq = ch + 2 * u;
        Cs[(4 * jq + 0) * kFltCand + cr] = pre[u].x;
        Cs[(4 * jq + 1) * kFltCand + cr] = pre[u].y;
        Cs[(4 * jq + 2) * kFltCand + cr] = pre[u].z;
        Cs[(4 * jq + 3) * kFltCand + cr] = pre[u].w;
      }
    if (ch == 0) csq[cr] = pre_sq;
    __syncthreads();
    // the next tile is chosen with the k-th distances as of the END of the previous step (one step stale, i.e.
    // conservative), so that its rows can be prefetched behind this tile's arithmetic
    double tau_max = 0.0;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) tau_max = fmax(tau_max, wmax[w8]);
    const int nxt = next_tile(tau_max);
    if (nxt >= 0) prefetch(g0 + nxt * kFltCand);
    const bool dense = first;  // lists are empty: everything passes, skip the filter
    first = false;
    if (!dense) {
      float acc[4][8];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;
#pragma unroll 4
      for (int jj = 0; jj < dpad; ++jj) {
        const float4 qv = *reinterpret_cast<const float4*>(&Qs[jj * kKnnTile + ty * 4]);
        const float4 ca = *reinterpret_cast<const float4*>(&Cs[jj * kFltCand + tx * 8]);
        const float4 cb = *reinterpret_cast<const float4*>(&Cs[jj * kFltCand + tx * 8 + 4]);
        const float q[4] = {qv.x, qv.y, qv.z, qv.w};
        const float c[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(q[a], c[b], acc[a][b]);
      }
      // filter -> per-warp worklists (the warp that owns query ql is ql / 8)
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int ql = ty * 4 + a;
        const int qi = q0 + ql;
        if (qi >= g1) continue;
        const float t = tauf[ql];
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          const int cl = tx * 8 + b;
          const int ci = c0 + cl;
          const float s = sqq[a] + csq[cl];
          const float d2 = s - 2.0f * acc[a][b];
          if (ci < g1 && ci != qi && d2 <= t + E * s) {
            const int w = ql >> 3;
            const int idx = atomicAdd(&wlc[w], 1);
            if (idx < kFltCap) wl[w * kFltCap + idx] = static_cast<uint16_t>((ql << 7) | cl);
          }
        }
      }
    }
    __syncthreads();
    // ---- exact fp64 evaluation + insertion: warp w owns queries 8w .. 8w+7 ----
    {
      const int cnt = dense ? 0 : wlc[warp];
      const bool all_pairs = dense || cnt > kFltCap;  // overflow: fall back to every pair of this warp's queries
      const int total = all_pairs ? 8 * kFltCand : cnt;
      for (int base = 0; base < total; base += 32) {
        const int e = base + lane;
        int ql = -1, cl = 0;
        if (e < total) {
          if (all_pairs) {
            ql = warp * 8 + (e >> 7);
            cl = e & 127;
          } else {
            const int ent = wl[warp * kFltCap + e];
            ql = ent >> 7;
            cl = ent & 127;
          }
        }
        double dc = INFINITY;
        const int qi = q0 + ql, ci = c0 + cl;
        if (ql >= 0 && qi < g1 && ci < g1 && ci != qi) {
          // exact fp64 distance: sequential fma over the (zero padded) features.  The rows are
          // float32 values, so the fp32 tiles already in shared memory convert to the fp64 operands exactly.
          double acc = 0.0;
#pragma unroll 4
          for (int jj = 0; jj < dpad; ++jj)
            acc = fma(static_cast<double>(Qs[jj * kKnnTile + ql]), static_cast<double>(Cs[jj * kFltCand + cl]), acc);
          dc = fmax(sq[qi] + sq[ci] - 2.0 * acc, 0.0);
        }
        // worklist entries arrive in no particular order: a candidate is accepted iff it precedes the heap root in
        // (distance, candidate index) order, so the final set is the k smallest such pairs whatever the arrival
        // order -- the same neighbours the in-order exhaustive kernels find.
        bool cand = false;
        if (dc < INFINITY) {  // lane-local pre-check against the query's current k-th best (it only ever tightens)
          const KBestTop t = tops[ql];
          cand = dc < t.d || (dc == t.d && ci < t.idx);
        }
        unsigned pending = __ballot_sync(0xffffffffu, cand);
        while (pending) {
          const int src = __ffs(pending) - 1;
          pending &= pending - 1;
          const double dv = __shfl_sync(0xffffffffu, dc, src);
          const int sq_l = __shfl_sync(0xffffffffu, ql, src);
          const int jc = c0 + __shfl_sync(0xffffffffu, cl, src);
          KBestTop* top = &tops[sq_l];
          if (dv < top->d || (dv == top->d && jc < top->idx))
            kbest_replace_max(ld + sq_l * k, li + sq_l * k, K, dv, jc, lane, top);
        }
      }
      __syncwarp();
      double t = 0.0;
      if (lane < 8) {
        t = tops[warp * 8 + lane].d;
        float tf = static_cast<float>(t);
        if (static_cast<double>(tf) < t) tf = nextafterf(tf, INFINITY);  // round up: the filter must not tighten tau
        tauf[warp * 8 + lane] = tf;
        if (q0 + warp * 8 + lane >= g1) t = 0.0;  // rows past the group do not hold the scan open
      }
      t = fmax(t, __shfl_xor_sync(0xffffffffu, t, 4));
      t = fmax(t, __shfl_xor_sync(0xffffffffu, t, 2));
      t = fmax(t, __shfl_xor_sync(0xffffffffu, t, 1));
      if (lane == 0) {
        wmax[warp] = t;
        wlc[warp] = 0;
      }
    }
    cur = nxt;
  }
  __syncthreads();
  // write out in ascending (distance, index) order: rank of an element = number of smaller elements
  for (int qq = 0; qq < 8; ++qq) {
    const int ql = warp * 8 + qq;
    if (q0 + ql >= g1) break;
    const double* qld = ld + ql * k;
    const int32_t* qli = li + ql * k;
    const size_t out = static_cast<size_t>(q0 + ql) * k;
    for (int s = lane; s < k; s += 32) {
      if (s >= K) {
        knn_d[out + s] = INFINITY;
        knn_i[out + s] = -1;
        continue;
      }
      const double d = qld[s];
      const int32_t ix = qli[s];
      int rank = 0;
      for (int t = 0; t < K; ++t) rank += kb_greater(d, ix, qld[t], qli[t]) ? 1 : 0;
      knn_d[out + rank] = sqrt(d);
      knn_i[out + rank] = ix;
      if (rank == K - 1) kdist[q0 + ql] = sqrt(d);
    }
  }
}

// index of the k-NN CTA (query tile) that owns sorted row i: tiles are numbered group by group
__device__ __forceinline__ int owner_tile(const int32_t* gstart, int g, int i) {
  int t = 0;
  for (int gg = 0; gg < g; ++gg) t += (gstart[gg + 1] - gstart[gg] + kKnnTile - 1) / kKnnTile;
  return t + (i - gstart[g]) / kKnnTile;
}

// per-row group lookup in sorted order (binary search over gstart) -> K for that row
__device__ __forceinline__ int row_group(const int32_t* gstart, int n_groups, int i) {
  int lo = 0, hi = n_groups;  // gstart[lo] <= i < gstart[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (gstart[mid] <= i) lo = mid; else hi = mid;
  }
  return lo;
}

// rows owned by other parts are written as 0 so that a sum all-reduce assembles the full vector
__global__ void lrd_kernel(const double* __restrict__ knn_d, const int32_t* __restrict__ knn_i,
                           const double* __restrict__ kdist_all, const int32_t* __restrict__ gstart, int n_groups,
                           long long n, int k, int part, int n_parts, double* __restrict__ lrd) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const int g = row_group(gstart, n_groups, static_cast<int>(i));
  if (n_parts > 1 && owner_tile(gstart, g, static_cast<int>(i)) % n_parts != part) {
    lrd[i] = 0.0;
    return;
  }
  const int ng = gstart[g + 1] - gstart[g];
  if (ng < 2) {
    lrd[i] = 1.0;
    return;
  }
  const int K = max(1, min(k, ng - 1));
  double s = 0.0;
  for (int t = 0; t < K; ++t) {
    const int nb = knn_i[i * k + t];
    s += fmax(knn_d[i * k + t], kdist_all[nb]);  // reachability distance (_lof.py:519-521)
  }
  lrd[i] = 1.0 / (s / K + 1e-10);
}

__global__ void lof_score_kernel(const double* __restrict__ lrd, const int32_t* __restrict__ knn_i,
                                 const int32_t* __restrict__ gstart, int n_groups, long long n, int k, int part,
                                 int n_parts, double* __restrict__ score_sorted) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const int g = row_group(gstart, n_groups, static_cast<int>(i));
  if (n_parts > 1 && owner_tile(gstart, g, static_cast<int>(i)) % n_parts != part) {
    score_sorted[i] = 0.0;
    return;
  }
  const int ng = gstart[g + 1] - gstart[g];
  double sc = -1.0;
  if (ng >= 2) {
    const int K = max(1, min(k, ng - 1));
    const double li = lrd[i];
    double s = 0.0;
    for (int t = 0; t < K; ++t) s += lrd[knn_i[i * k + t]] / li;
    sc = -(s / K);
  }
  score_sorted[i] = sc;
}

__global__ void unsort_kernel(const double* __restrict__ v_sorted, const int32_t* __restrict__ order, long long n,
                              double* __restrict__ out) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) out[order[i]] = v_sorted[i];
}

// ------------------------------------------------------------------------------------------------------------
// np.percentile (method="linear") per group via MSB-first radix select on order-preserving 64-bit keys
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long f64_key(double v) {
  const unsigned long long u = static_cast<unsigned long long>(__double_as_longlong(v));
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_f64(unsigned long long k) {
  const unsigned long long u = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
  return __longlong_as_double(static_cast<long long>(u));
}

// rank-th smallest (0-based) of v[0..n); whole block participates
__device__ double block_select(const double* __restrict__ v, int n, int rank, int* hist) {
  unsigned long long prefix = 0, mask = 0;
  int r = rank;
  for (int shift = 56; shift >= 0; shift -= 8) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const unsigned long long key = f64_key(v[i]);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 0xFF], 1);
    }
    __syncthreads();
    // every thread walks the 256 bins (cheap, keeps control flow uniform)
    int cum = 0, bin = 0;
    for (; bin < 256; ++bin) {
      const int h = hist[bin];
      if (cum + h > r) break;
      cum += h;
    }
    r -= cum;
    prefix |= static_cast<unsigned long long>(bin) << shift;
    mask |= 0xFFull << shift;
    __syncthreads();
  }
  return key_f64(prefix);
}

// q in [0,1]; returns numpy's linear-interpolated quantile of v[0..n)
__device__ double block_percentile(const double* __restrict__ v, int n, double q, int* hist, double* red) {
  // numpy/lib/_function_base_impl.py _compute_virtual_index(n, q, alpha=1, beta=1), _get_gamma, _lerp
  const double vi = (static_cast<double>(n) * q + (1.0 + q * (1.0 - 1.0 - 1.0))) - 1.0;
  double prev = floor(vi);
  double gamma = vi - prev;
  int lo = static_cast<int>(prev);
  if (lo < 0) lo = 0;
  if (lo > n - 1) lo = n - 1;
  int hi = lo + 1;
  if (hi > n - 1) hi = n - 1;
  const double a = block_select(v, n, lo, hi == lo ? hist : hist);
  double b = a;
  if (hi != lo) {
    // b = next order statistic: a again if it has duplicates beyond rank lo, else the smallest value > a
    int cnt_le = 0;
    double mn = INFINITY;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const double x = v[i];
      if (x <= a) ++cnt_le;
      else mn = fmin(mn, x);
    }
    for (int o = 16; o > 0; o >>= 1) {
      cnt_le += __shfl_xor_sync(0xffffffffu, cnt_le, o);
      mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    __shared__ int s_cnt[32];
    __shared__ double s_mn[32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) {
      s_cnt[warp] = cnt_le;
      s_mn[warp] = mn;
    }
    __syncthreads();
    int tc = 0;
    double tm = INFINITY;
    for (int w = 0; w < (blockDim.x >> 5); ++w) {
      tc += s_cnt[w];
      tm = fmin(tm, s_mn[w]);
    }
    b = (tc >= lo + 2) ? a : tm;
  }
  (void)red;
  const double diff = b - a;
  double res = a + diff * gamma;
  if (gamma >= 0.5) res = b - diff * (1.0 - gamma);
  return res;
}

__global__ void __launch_bounds__(1024) group_percentile_kernel(const double* __restrict__ values_sorted,
                                                                const int32_t* __restrict__ gstart, double q,
                                                                double* __restrict__ out) {
  __shared__ int hist[256];
  __shared__ double red[32];
  const int g = blockIdx.x;
  const int g0 = gstart[g], n = gstart[g + 1] - g0;
  if (n <= 0) {
    if (threadIdx.x == 0) out[g] = nan("");
    return;
  }
  const double r = block_percentile(values_sorted + g0, n, q, hist, red);
  if (threadIdx.x == 0) out[g] = r;
}

__global__ void flag_kernel(const double* __restrict__ value_sorted, const int32_t* __restrict__ gstart, int n_groups,
                            const int32_t* __restrict__ order, long long n, const double* __restrict__ thr,
                            int greater, uint8_t* __restrict__ flags) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const int g = row_group(gstart, n_groups, static_cast<int>(i));
  const double v = value_sorted[i], t = thr[g];
  flags[order[i]] = greater ? (v > t) : (v < t);
}

// ------------------------------------------------------------------------------------------------------------
// centroid / z-score scorer
// ------------------------------------------------------------------------------------------------------------
// one CTA per (group, feature chunk): centroid[g][j] = mean over the group's rows (fixed summation order)
__global__ void __launch_bounds__(256) centroid_kernel(const float* __restrict__ z, const int32_t* __restrict__ order,
                                                       const int32_t* __restrict__ gstart, int dim,
                                                       double* __restrict__ centroid) {
  const int g = blockIdx.x;
  const int g0 = gstart[g], g1 = gstart[g + 1];
  __shared__ double part[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int j0 = blockIdx.y * 32; j0 < dim; j0 += gridDim.y * 32) {
    const int j = j0 + lane;
    double s = 0.0;
    if (j < dim)
      for (int i = g0 + warp; i < g1; i += 8) s += static_cast<double>(z[static_cast<size_t>(order[i]) * dim + j]);
    part[warp][lane] = s;
    __syncthreads();
    if (warp == 0 && j < dim) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += part[w][lane];
      centroid[static_cast<size_t>(g) * dim + j] = g1 > g0 ? t / (g1 - g0) : 0.0;
    }
    __syncthreads();
  }
}

__global__ void centroid_dist_kernel(const float* __restrict__ z, const int32_t* __restrict__ order,
                                     const int32_t* __restrict__ gstart, int n_groups, long long n, int dim,
                                     const double* __restrict__ centroid, double* __restrict__ dist_sorted,
                                     double* __restrict__ dist_out) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const int g = row_group(gstart, n_groups, static_cast<int>(i));
  const float* r = z + static_cast<size_t>(order[i]) * dim;
  const double* c = centroid + static_cast<size_t>(g) * dim;
  double s = 0.0;
  for (int j = 0; j < dim; ++j) {
    const double d = static_cast<double>(r[j]) - c[j];
    s += d * d;
  }
  s = sqrt(s);
  dist_sorted[i] = s;
  dist_out[order[i]] = s;
}

// per group: mean / population std of the distances (two-pass, fixed order), then z-scores
__global__ void __launch_bounds__(1024) group_zscore_kernel(const double* __restrict__ dist_sorted,
                                                            const int32_t* __restrict__ gstart,
                                                            const int32_t* __restrict__ order,
                                                            double* __restrict__ zscore_out) {
  __shared__ double red[32];
  __shared__ double s_mean, s_std;
  const int g = blockIdx.x;
  const int g0 = gstart[g], n = gstart[g + 1] - g0;
  if (n <= 0) return;
  const double* v = dist_sorted + g0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += v[i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 32; ++w) t += red[w];
    s_mean = t / n;
  }
  __syncthreads();
  const double mean = s_mean;
  double q = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double d = v[i] - mean;
    q += d * d;
  }
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  __syncthreads();
  if (lane == 0) red[warp] = q;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 32; ++w) t += red[w];
    s_std = sqrt(t / n);
  }
  __syncthreads();
  const double sd = s_std;
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    zscore_out[order[g0 + i]] = sd > 0.0 ? (v[i] - mean) / sd : 0.0;
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct SortedRows {
  int32_t* order;
  int32_t* gstart;
  int32_t* counts;
};

static size_t sort_bytes(long long n, int n_groups) {
  return align_up(static_cast<size_t>(n) * 4, 256) + align_up(static_cast<size_t>(n_groups + 1) * 4, 256) +
         align_up(static_cast<size_t>(n_groups) * kSortThreads * 4, 256);
}

static int sort_rows(const int32_t* d_group, long long n, int n_groups, uint8_t*& ws, SortedRows* out,
                     cudaStream_t st) {
  out->order = reinterpret_cast<int32_t*>(ws);
  ws += align_up(static_cast<size_t>(n) * 4, 256);
  out->gstart = reinterpret_cast<int32_t*>(ws);
  ws += align_up(static_cast<size_t>(n_groups + 1) * 4, 256);
  out->counts = reinterpret_cast<int32_t*>(ws);
  ws += align_up(static_cast<size_t>(n_groups) * kSortThreads * 4, 256);
  if (d_group == nullptr || n_groups == 1) {
    iota_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(out->order, n, out->gstart);
  } else {
    group_count_kernel<<<1, kSortThreads, 0, st>>>(d_group, n, n_groups, out->counts);
    group_scan_kernel<<<1, kSortThreads, 0, st>>>(out->counts, n_groups, out->gstart);
    group_scatter_kernel<<<1, kSortThreads, 0, st>>>(d_group, n, n_groups, out->counts, out->order);
    IRP_CUDA_OK(cudaGetLastError());
    // A group id outside [0, n_groups) would leave rows out of every group while the later kernels still walk all
    // n rows: refuse the call instead (gstart[n_groups] = number of rows that landed in a group).
    int32_t placed = 0;
    IRP_CUDA_OK(cudaMemcpyAsync(&placed, out->gstart + n_groups, sizeof(placed), cudaMemcpyDeviceToHost, st));
    IRP_CUDA_OK(cudaStreamSynchronize(st));
    IRP_REQUIRE(placed == static_cast<int32_t>(n), "group ids must lie in [0, %d): %lld of %lld rows do not", n_groups,
                n - static_cast<long long>(placed), n);
  }
  IRP_CUDA_OK(cudaGetLastError());
  return IRP_OK;
}

// ------------------------------------------------------------------------------------------------------------
// N4: UMAP's k-NN arrays from the neighbour lists of the search above.  umap-learn's
// nearest_neighbors returns, per sample, ITSELF (distance 0) followed by its n_neighbors - 1 nearest other samples,
// ascending; rows come back in the caller's order and indices refer to the caller's rows.
// ------------------------------------------------------------------------------------------------------------
__global__ void knn_graph_emit_kernel(const double* __restrict__ knn_d, const int32_t* __restrict__ knn_i,
                                      const int32_t* __restrict__ order, long long n, int k_other,
                                      int32_t* __restrict__ idx, float* __restrict__ dist) {
  const long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const int k = k_other + 1;
  if (t >= n * k) return;
  const long long p = t / k;  // sorted position
  const int j = static_cast<int>(t - p * k);
  const long long r = order[p];
  if (j == 0) {
    idx[r * k] = static_cast<int32_t>(r);
    dist[r * k] = 0.f;
    return;
  }
  const int32_t nb = knn_i[p * k_other + j - 1];
  idx[r * k + j] = nb < 0 ? -1 : order[nb];
  dist[r * k + j] = static_cast<float>(knn_d[p * k_other + j - 1]);
}

}  // namespace irp

using namespace irp;

extern "C" {

// scratch reserved for cub::DeviceRadixSort (histograms / look-back state; far below this bound for 64-bit keys)
static size_t lof_sort_tmp_bytes(size_t n) { return (4u << 20) + n * 16; }  // cub radix sort of (u64, i32) pairs: measured 12.2 B per row at 1 M rows

size_t irp_lof_workspace_bytes(int64_t n_rows, int dim, int k) {
  if (n_rows <= 0 || dim <= 0 || k <= 0) return 0;
  const size_t n = static_cast<size_t>(n_rows);
  size_t b = sort_bytes(n_rows, 1024);
  b += align_up(n * 8, 256);          // sq
  b += align_up(n * k * 8, 256);      // knn_d
  b += align_up(n * k * 4, 256);      // knn_i
  b += 3 * align_up(n * 8, 256);      // kdist, lrd, score_sorted (single-part path)
  if (dim <= kKnnMaxResidentDim) {
    b += align_up(n * static_cast<size_t>((dim + 7) / 8 * 8) * 12, 256);  // sorted fp64 + fp32 rows
    b += 256 + n * 16 + align_up(n * 4, 256) + lof_sort_tmp_bytes(n);      // max norm, sort keys x2, values, cub scratch
  }
  return b + 1024;
}

namespace {

// carve-up of the caller's workspace; identical for every phase of one (n_rows, k) problem
struct LofWs {
  SortedRows sr;
  double* sq;
  double* knn_d;
  int32_t* knn_i;
  double* kdist;
  double* lrd;
  double* score_sorted;
  double* zs;  // [n][dpad] group-sorted fp64 rows (knn phase only; last, so the other offsets do not depend on dim)
};

static LofWs lof_layout(void* d_workspace, size_t n, int k) {
  LofWs w;
  uint8_t* ws = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<size_t>(d_workspace), 256));
  w.sr.order = reinterpret_cast<int32_t*>(ws);
  ws += align_up(n * 4, 256);
  w.sr.gstart = reinterpret_cast<int32_t*>(ws);
  ws += align_up(static_cast<size_t>(1024 + 1) * 4, 256);
  w.sr.counts = reinterpret_cast<int32_t*>(ws);
  ws += align_up(static_cast<size_t>(1024) * kSortThreads * 4, 256);
  w.sq = reinterpret_cast<double*>(ws);
  ws += align_up(n * 8, 256);
  w.knn_d = reinterpret_cast<double*>(ws);
  ws += align_up(n * k * 8, 256);
  w.knn_i = reinterpret_cast<int32_t*>(ws);
  ws += align_up(n * k * 4, 256);
  w.kdist = reinterpret_cast<double*>(ws);
  ws += align_up(n * 8, 256);
  w.lrd = reinterpret_cast<double*>(ws);
  ws += align_up(n * 8, 256);
  w.score_sorted = reinterpret_cast<double*>(ws);
  ws += align_up(n * 8, 256);
  w.zs = reinterpret_cast<double*>(ws);
  return w;
}

static int lof_check(int64_t n_rows, int k, int n_groups, int part, int n_parts, const void* ws, size_t ws_bytes) {
  IRP_REQUIRE(ws != nullptr, "lof: null workspace");
  IRP_REQUIRE(n_rows >= 2 && n_rows < (1ll << 31), "lof: n_rows %lld", static_cast<long long>(n_rows));
  IRP_REQUIRE(k >= 1 && k <= kMaxK, "lof: n_neighbors %d not in [1,%d]", k, kMaxK);
  IRP_REQUIRE(n_groups >= 1 && n_groups <= 1024, "lof: n_groups %d not in [1,1024]", n_groups);
  IRP_REQUIRE(n_parts >= 1 && part >= 0 && part < n_parts, "lof: part %d of %d", part, n_parts);
  IRP_REQUIRE(ws_bytes >= irp_lof_workspace_bytes(n_rows, 1, k), "lof: workspace too small");
  return IRP_OK;
}

}  // namespace

int irp_lof_knn_part(const float* d_z, int64_t n_rows, int dim, const int32_t* d_group, int n_groups, int k, int part,
                     int n_parts, double* d_kdist, void* d_workspace, size_t workspace_bytes, void* stream) {
  IRP_REQUIRE(d_z && d_kdist && dim >= 1, "lof_knn_part: bad argument");
  IRP_TRY(lof_check(n_rows, k, n_groups, part, n_parts, d_workspace, workspace_bytes));
  IRP_REQUIRE(workspace_bytes >= irp_lof_workspace_bytes(n_rows, dim, k), "lof: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t n = static_cast<size_t>(n_rows);
  LofWs w = lof_layout(d_workspace, n, k);
  uint8_t* cursor = reinterpret_cast<uint8_t*>(w.sr.order);
  SortedRows sr;
  IRP_TRY(sort_rows(d_group, n_rows, n_groups, cursor, &sr, st));
  w.sr = sr;
  const unsigned blocks = static_cast<unsigned>((n + 255) / 256);
  const bool resident = dim <= kKnnMaxResidentDim;
  double* max_sq = nullptr;
  if (resident) {
    // re-order the rows of every group by their first coordinate (stable radix sort of (group, z0) keys); the
    // group boundaries computed above stay valid.  Scratch lives behind the fp64 / fp32 row copies.
    const size_t dpad = static_cast<size_t>((dim + 7) / 8 * 8);
    uint8_t* scratch = reinterpret_cast<uint8_t*>(w.zs) + align_up(n * dpad * 12, 256);
    max_sq = reinterpret_cast<double*>(scratch);
    unsigned long long* keys_in = reinterpret_cast<unsigned long long*>(scratch + 256);
    unsigned long long* keys_out = keys_in + n;
    int32_t* vals_in = reinterpret_cast<int32_t*>(keys_out + n);
    void* cub_tmp = reinterpret_cast<uint8_t*>(vals_in) + align_up(n * 4, 256);
    size_t cub_bytes = 0;
    int end_bit = 32;
    for (int gmax = n_groups - 1; gmax > 0; gmax >>= 1) ++end_bit;
    IRP_CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, keys_in, keys_out, vals_in, w.sr.order,
                                                static_cast<long long>(n), 0, end_bit, st));
    IRP_REQUIRE(cub_bytes <= lof_sort_tmp_bytes(n), "lof: radix sort needs %zu bytes of scratch", cub_bytes);
    sort_key_kernel<<<blocks, 256, 0, st>>>(d_z, n_groups > 1 ? d_group : nullptr, n_rows, dim, keys_in, vals_in);
    IRP_CUDA_OK(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, keys_in, keys_out, vals_in, w.sr.order,
                                                static_cast<long long>(n), 0, end_bit, st));
    IRP_CUDA_OK(cudaMemsetAsync(max_sq, 0, sizeof(double), st));
  }
  sqnorm_kernel<<<blocks, 256, 0, st>>>(d_z, w.sr.order, n_rows, dim, w.sq,
                                        reinterpret_cast<unsigned long long*>(max_sq));
  IRP_CUDA_OK(cudaMemsetAsync(d_kdist, 0, n * sizeof(double), st));  // rows of other parts stay 0
  const unsigned knn_grid = static_cast<unsigned>((n + kKnnTile - 1) / kKnnTile + n_groups);
  if (dim <= kKnnMaxResidentDim) {
    const int dpad = (dim + 7) / 8 * 8;
    const long long total = static_cast<long long>(n) * dpad;
    const unsigned gblocks = static_cast<unsigned>(total / 256 + 1 < 148 * 16 ? total / 256 + 1 : 148 * 16);
    float* zs32 = reinterpret_cast<float*>(w.zs + n * dpad);
    gather_sorted_kernel<<<gblocks, 256, 0, st>>>(d_z, w.sr.order, n_rows, dim, dpad, w.zs, zs32);
    {
      const size_t lists = static_cast<size_t>(kKnnTile) * k * 12 + 8;
      const size_t smem = lists + (static_cast<size_t>(dpad) * (kKnnTile + kFltCand) + kFltCand + kKnnTile) * 4 +
                          8 * kFltCap * 2 + 64 + kKnnTile * 16 + 64;
      IRP_TRY(ensure_smem(knn_filter_kernel, smem));
      knn_filter_kernel<<<knn_grid, kKnnThreads, smem, st>>>(w.zs, zs32, w.sr.gstart, n_groups, dpad, k, w.sq, max_sq,
                                                             part, n_parts, w.knn_d, w.knn_i, d_kdist);
    }
  } else {
    const size_t smem = (2 * kKnnDChunk * kKnnTile + kKnnTile * (kKnnTile + 1) + static_cast<size_t>(kKnnTile) * k) * 8 +
                        static_cast<size_t>(kKnnTile) * k * 4;
    IRP_TRY(ensure_smem(knn_kernel, smem));
    knn_kernel<<<knn_grid, kKnnThreads, smem, st>>>(d_z, w.sr.order, w.sr.gstart, n_groups, dim, k, w.sq, part, n_parts,
                                                    w.knn_d, w.knn_i, d_kdist);
  }
  IRP_CUDA_OK(cudaGetLastError());
  return IRP_OK;
}

int irp_lof_lrd_part(int64_t n_rows, int n_groups, int k, int part, int n_parts, const double* d_kdist_all,
                     double* d_lrd, void* d_workspace, size_t workspace_bytes, void* stream) {
  IRP_REQUIRE(d_kdist_all && d_lrd, "lof_lrd_part: null argument");
  IRP_TRY(lof_check(n_rows, k, n_groups, part, n_parts, d_workspace, workspace_bytes));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t n = static_cast<size_t>(n_rows);
  LofWs w = lof_layout(d_workspace, n, k);
  const unsigned blocks = static_cast<unsigned>((n + 255) / 256);
  lrd_kernel<<<blocks, 256, 0, st>>>(w.knn_d, w.knn_i, d_kdist_all, w.sr.gstart, n_groups, n_rows, k, part, n_parts,
                                     d_lrd);
  IRP_CUDA_OK(cudaGetLastError());
  return IRP_OK;
}

int irp_lof_score_part(int64_t n_rows, int n_groups, int k, int part, int n_parts, const double* d_lrd_all,
                       double* d_score_sorted, void* d_workspace, size_t workspace_bytes, void* stream) {
  IRP_REQUIRE(d_lrd_all && d_score_sorted, "lof_score_part: null argument");
  IRP_TRY(lof_check(n_rows, k, n_groups, part, n_parts, d_workspace, workspace_bytes));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t n = static_cast<size_t>(n_rows);
  LofWs w = lof_layout(d_workspace, n, k);
  const unsigned blocks = static_cast<unsigned>((n + 255) / 256);
  lof_score_kernel<<<blocks, 256, 0, st>>>(d_lrd_all, w.knn_i, w.sr.gstart, n_groups, n_rows, k, part, n_parts,
                                           d_score_sorted);
  IRP_CUDA_OK(cudaGetLastError());
  return IRP_OK;
}

int irp_lof_finish(int64_t n_rows, int n_groups, int k, double contamination, const double* d_score_sorted_all,
                   double* d_scores, double* d_offsets, uint8_t* d_flags, void* d_workspace, size_t workspace_bytes,
                   void* stream) {
  IRP_REQUIRE(d_score_sorted_all && d_scores && d_offsets && d_flags, "lof_finish: null argument");
  IRP_REQUIRE(contamination > 0.0 && contamination <= 0.5, "lof: contamination %g not in (0,0.5]", contamination);
  IRP_TRY(lof_check(n_rows, k, n_groups, 0, 1, d_workspace, workspace_bytes));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t n = static_cast<size_t>(n_rows);
  LofWs w = lof_layout(d_workspace, n, k);
  const unsigned blocks = static_cast<unsigned>((n + 255) / 256);
  unsort_kernel<<<blocks, 256, 0, st>>>(d_score_sorted_all, w.sr.order, n_rows, d_scores);
  // sklearn passes 100*contamination to np.percentile, which divides by 100 again
  const double q = (100.0 * contamination) / 100.0;
  group_percentile_kernel<<<n_groups, 1024, 0, st>>>(d_score_sorted_all, w.sr.gstart, q, d_offsets);
  flag_kernel<<<blocks, 256, 0, st>>>(d_score_sorted_all, w.sr.gstart, n_groups, w.sr.order, n_rows, d_offsets, 0,
                                      d_flags);
  IRP_CUDA_OK(cudaGetLastError());
  return IRP_OK;
}

int irp_lof(const float* d_z, int64_t n_rows, int dim, const int32_t* d_group, int n_groups, int k,
            double contamination, double* d_scores, double* d_offsets, uint8_t* d_flags, void* d_workspace,
            size_t workspace_bytes, void* stream) {
  IRP_REQUIRE(d_z && d_scores && d_offsets && d_flags && d_workspace, "lof: null argument");
  IRP_REQUIRE(dim >= 1, "lof: dim %d", dim);
  IRP_REQUIRE(contamination > 0.0 && contamination <= 0.5, "lof: contamination %g not in (0,0.5]", contamination);
  IRP_TRY(lof_check(n_rows, k, n_groups, 0, 1, d_workspace, workspace_bytes));
  LofWs w = lof_layout(d_workspace, static_cast<size_t>(n_rows), k);
  IRP_TRY(irp_lof_knn_part(d_z, n_rows, dim, d_group, n_groups, k, 0, 1, w.kdist, d_workspace, workspace_bytes, stream));
  IRP_TRY(irp_lof_lrd_part(n_rows, n_groups, k, 0, 1, w.kdist, w.lrd, d_workspace, workspace_bytes, stream));
  IRP_TRY(irp_lof_score_part(n_rows, n_groups, k, 0, 1, w.lrd, w.score_sorted, d_workspace, workspace_bytes, stream));
  return irp_lof_finish(n_rows, n_groups, k, contamination, w.score_sorted, d_scores, d_offsets, d_flags, d_workspace,
                        workspace_bytes, stream);
}

size_t irp_knn_graph_workspace_bytes(int64_t n_rows, int dim, int k) {
  if (n_rows <= 0 || dim <= 0 || k < 2) return 0;
  return irp_lof_workspace_bytes(n_rows, dim, k - 1);
}

int irp_knn_graph(const float* d_z, int64_t n_rows, int dim, int k, int32_t* d_idx, float* d_dist, void* d_workspace,
                  size_t workspace_bytes, void* stream) {
  IRP_REQUIRE(d_z && d_idx && d_dist && d_workspace, "knn_graph: null argument");
  IRP_REQUIRE(k >= 2 && k - 1 <= kMaxK, "knn_graph: n_neighbors %d not in [2,%d]", k, kMaxK + 1);
  IRP_REQUIRE(n_rows >= k, "knn_graph: %lld rows < n_neighbors %d", static_cast<long long>(n_rows), k);
  IRP_REQUIRE(workspace_bytes >= irp_knn_graph_workspace_bytes(n_rows, dim, k), "knn_graph: workspace too small");
  const int ko = k - 1;
  LofWs w = lof_layout(d_workspace, static_cast<size_t>(n_rows), ko);
  // the exact (fp64, ties by index) neighbour search of the LOF scorer, all rows in one group
  IRP_TRY(irp_lof_knn_part(d_z, n_rows, dim, nullptr, 1, ko, 0, 1, w.kdist, d_workspace, workspace_bytes, stream));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long total = static_cast<long long>(n_rows) * k;
  knn_graph_emit_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(w.knn_d, w.knn_i, w.sr.order,
                                                                                   n_rows, ko, d_idx, d_dist);
  IRP_CUDA_OK(cudaGetLastError());
  return IRP_OK;
}

size_t irp_centroid_workspace_bytes(int64_t n_rows, int dim, int n_groups) {
  if (n_rows <= 0 || dim <= 0 || n_groups <= 0) return 0;
  const size_t n = static_cast<size_t>(n_rows);
  return sort_bytes(n_rows, n_groups) + align_up(static_cast<size_t>(n_groups) * dim * 8, 256) +
         align_up(n * 8, 256) + 1024;
}

int irp_centroid_zscore(const float* d_z, int64_t n_rows, int dim, const int32_t* d_group, int n_groups,
                        double contamination, double* d_dist, double* d_zscore, double* d_thresholds,
                        uint8_t* d_flags, void* d_workspace, size_t workspace_bytes, void* stream) {
  IRP_REQUIRE(d_z && d_dist && d_zscore && d_thresholds && d_flags && d_workspace, "centroid: null argument");
  IRP_REQUIRE(n_rows >= 1 && n_rows < (1ll << 31) && dim >= 1, "centroid: n_rows %lld dim %d",
              static_cast<long long>(n_rows), dim);
  IRP_REQUIRE(n_groups >= 1 && n_groups <= 1024, "centroid: n_groups %d not in [1,1024]", n_groups);
  IRP_REQUIRE(contamination > 0.0 && contamination < 1.0, "centroid: contamination %g", contamination);
  IRP_REQUIRE(workspace_bytes >= irp_centroid_workspace_bytes(n_rows, dim, n_groups), "centroid: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t n = static_cast<size_t>(n_rows);
  uint8_t* ws = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<size_t>(d_workspace), 256));
  SortedRows sr;
  IRP_TRY(sort_rows(d_group, n_rows, n_groups, ws, &sr, st));
  double* centroid = reinterpret_cast<double*>(ws);
  ws += align_up(static_cast<size_t>(n_groups) * dim * 8, 256);
  double* dist_sorted = reinterpret_cast<double*>(ws);
  const unsigned blocks = static_cast<unsigned>((n + 255) / 256);
  dim3 cgrid(n_groups, (dim + 31) / 32 < 64 ? (dim + 31) / 32 : 64);
  centroid_kernel<<<cgrid, 256, 0, st>>>(d_z, sr.order, sr.gstart, dim, centroid);
  centroid_dist_kernel<<<blocks, 256, 0, st>>>(d_z, sr.order, sr.gstart, n_groups, n_rows, dim, centroid, dist_sorted,
                                               d_dist);
  group_zscore_kernel<<<n_groups, 1024, 0, st>>>(dist_sorted, sr.gstart, sr.order, d_zscore);
  const double q = (100.0 * (1.0 - contamination)) / 100.0;
  group_percentile_kernel<<<n_groups, 1024, 0, st>>>(dist_sorted, sr.gstart, q, d_thresholds);
  flag_kernel<<<blocks, 256, 0, st>>>(dist_sorted, sr.gstart, n_groups, sr.order, n_rows, d_thresholds, 1, d_flags);
  IRP_CUDA_OK(cudaGetLastError());
  return IRP_OK;
}

}  // extern "C"
