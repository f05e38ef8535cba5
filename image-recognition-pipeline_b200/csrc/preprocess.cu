// A1 preprocessing: ragged uint8 HWC batch -> resize(232, Pillow antialiased bilinear) -> center crop 224 ->
// /255 -> (x-mean)/std -> bf16, in one pass over the pixels.
//
// Replaces `transform(img)` at functions/data_curation.py:675 (ResNet50_Weights.DEFAULT.transforms(),
// torchvision/transforms/_presets.py ImageClassification.forward; output size per
// torchvision/transforms/functional.py:359-384, crop offsets per :592-593).  The resample arithmetic restates
// Pillow's ImagingResample for 8-bit images (src/libImaging/Resample.c in Pillow 11/12: precompute_coeffs,
// normalize_coeffs_8bpc with PRECISION_BITS = 22, horizontal pass then vertical pass, each rounded and clipped
// to uint8).  oracle/pil_resample.py is the numpy restatement checked bit-for-bit against Pillow itself.
//
// Two kernels:
//   resample_plan_kernel : per image, per axis, for the 224 output indices inside the crop window: first source
//                          index, tap count and the 22-bit fixed-point tap weights (fp64 maths, like Pillow).
//   resample_kernel      : one CTA per (image, band of 8 output rows): horizontal pass of the source rows the
//                          band needs into shared memory (uint8, rounded like Pillow), vertical pass into
//                          registers, LUT normalise, staged in shared memory and written with 16-byte stores.
#include <cuda_bf16.h>

#include <cstdlib>

#include "common.h"
#include "preprocess_fused.cuh"

namespace irp {

constexpr int kCrop = IRP_CROP;      // 224
constexpr int kResize = IRP_RESIZE;  // 232
constexpr int kPad = IRP_PAD_HW;     // 230
constexpr int kPrecisionBits = 22;   // Pillow: 32 - 8 - 2
constexpr int kBandRows = 8;         // output rows per CTA
constexpr int kChunkRows = 24;       // source rows horizontally filtered per pass
constexpr int kRowElems = kCrop * 3; // 672 (x, c) elements per filtered row
constexpr int kThreads = 256;
constexpr int kElemsPerThread = 3;   // ceil(672 / 256)

// plan layout per image (int32): [axis 0 | axis 1], each axis: first[224], count[224], coef[224][max_taps]
__host__ __device__ inline size_t plan_ints_per_axis(int max_taps) { return static_cast<size_t>(kCrop) * (2 + max_taps); }

struct Geometry {
  int out_h, out_w;  // resized size
  int top, left;     // crop offsets
};

// torchvision _compute_resized_output_size (short side -> 232, long = int(232*long/short)) + center_crop offsets
// (Python round = half to even).
__host__ __device__ inline int round_half_even_div2(int d) {
  // round(d / 2.0) with ties to even, d >= 0
  const int q = d >> 1;
  if ((d & 1) == 0) return q;
  return (q & 1) ? q + 1 : q;
}
__host__ __device__ inline Geometry compute_geometry(int h, int w, int transform = IRP_TRANSFORM_WEIGHTS_DEFAULT) {
  Geometry g;
  if (transform == IRP_TRANSFORM_VAL_256) {
    // functions/dataload.py:51-56 val_transform: Resize((256, 256)) ignores the aspect ratio, CenterCrop(224)
    g.out_h = g.out_w = IRP_VAL_RESIZE;
    g.top = g.left = round_half_even_div2(IRP_VAL_RESIZE - kCrop);
    return g;
  }
  if (transform == IRP_TRANSFORM_HASH_64) {
    // functions/data_curation.py:283-292 compute_image_hash: img.resize((64, 64)), aspect ratio not kept, no crop
    g.out_h = g.out_w = IRP_HASH_SIZE;
    g.top = g.left = 0;
    return g;
  }
  if (transform == IRP_TRANSFORM_WDS_LANCZOS) {
    // functions/data_curation.py:896-913 resize_and_crop_image: smaller side -> 224, the other one
    // int(side * (224 / smaller)) (Python float arithmetic), crop offsets by floor division
    const double t = static_cast<double>(kCrop);
    if (w < h) {
      g.out_w = kCrop;
      g.out_h = static_cast<int>(static_cast<double>(h) * (t / static_cast<double>(w)));
    } else {
      g.out_h = kCrop;
      g.out_w = static_cast<int>(static_cast<double>(w) * (t / static_cast<double>(h)));
    }
    g.top = (g.out_h - kCrop) / 2;
    g.left = (g.out_w - kCrop) / 2;
    return g;
  }
  if (w <= h) {
    g.out_w = kResize;
    g.out_h = static_cast<int>(static_cast<double>(static_cast<long long>(kResize) * h) / static_cast<double>(w));
  } else {
    g.out_h = kResize;
    g.out_w = static_cast<int>(static_cast<double>(static_cast<long long>(kResize) * w) / static_cast<double>(h));
  }
  g.top = round_half_even_div2(g.out_h - kCrop);
  g.left = round_half_even_div2(g.out_w - kCrop);
  return g;
}

// Pillow's filter kernels (src/libImaging/Resample.c: bilinear_filter, sinc_filter, lanczos_filter)
__device__ __forceinline__ double triangle_weight(double x) {
  if (x < 0.0) x = -x;
  return x < 1.0 ? __dsub_rn(1.0, x) : 0.0;
}
__device__ __forceinline__ double sinc_weight(double x) {
  if (x == 0.0) return 1.0;
  x = __dmul_rn(x, 3.14159265358979323846);
  return __ddiv_rn(sin(x), x);
}
__device__ __forceinline__ double lanczos3_weight(double x) {
  if (-3.0 <= x && x < 3.0) return __dmul_rn(sinc_weight(x), sinc_weight(__ddiv_rn(x, 3.0)));
  return 0.0;
}
// bicubic_filter with a = -0.5, Pillow's expression order
__device__ __forceinline__ double bicubic_weight(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0)
    return __dadd_rn(__dmul_rn(__dmul_rn(__dsub_rn(__dmul_rn(__dadd_rn(a, 2.0), x), __dadd_rn(a, 3.0)), x), x), 1.0);
  if (x < 2.0)
    return __dmul_rn(__dsub_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dsub_rn(x, 5.0), x), 8.0), x), 4.0), a);
  return 0.0;
}
// filter of a transform: 0 triangle (support 1), 1 Lanczos (support 3), 2 bicubic (support 2)
__host__ __device__ inline int transform_filter(int transform) {
  return transform == IRP_TRANSFORM_WDS_LANCZOS ? 1 : (transform == IRP_TRANSFORM_HASH_64 ? 2 : 0);
}
__host__ __device__ inline int transform_out_size(int transform) {
  return transform == IRP_TRANSFORM_HASH_64 ? IRP_HASH_SIZE : kCrop;
}
__device__ __forceinline__ double filter_weight(int filt, double x) {
  return filt == 1 ? lanczos3_weight(x) : (filt == 2 ? bicubic_weight(x) : triangle_weight(x));
}

// One CTA (448 threads) per image: thread (axis, o) computes the taps of output index o; the CTA then chooses the
// fused kernel's band height for the image and appends one 64-byte work record per band to the heavy / normal list.
__global__ void __launch_bounds__(2 * kCrop) resample_plan_kernel(const int32_t* __restrict__ hw,
                                                                   const int64_t* __restrict__ offsets, int n_images,
                                                                   int max_taps, int transform,
                                                                   int32_t* __restrict__ plan,
                                                                   int32_t* __restrict__ status,
                                                                   __nv_bfloat16* __restrict__ lut,
                                                                   int32_t* __restrict__ img_info,
                                                                   int32_t* __restrict__ counters,
                                                                   FusedItem* __restrict__ items_heavy,
                                                                   FusedItem* __restrict__ items_normal) {
  __shared__ int s_first[2][kCrop], s_count[2][kCrop];
  __shared__ int s_red[4];  // max h taps, max v taps, band height, list base
  const int img = blockIdx.x;
  const int j = threadIdx.x;
  if (img == 0 && lut != nullptr) {
    // normalisation table: lut[c*256 + v] = bf16(((v / 255) - mean_c) / std_c), torch's fp32 operations in order
    for (int i = j; i < 768; i += blockDim.x) {
      const int c = i >> 8, v = i & 255;
      const float mean = c == 0 ? 0.485f : (c == 1 ? 0.456f : 0.406f);
      const float stdv = c == 0 ? 0.229f : (c == 1 ? 0.224f : 0.225f);
      const float f = __fdiv_rn(static_cast<float>(v), 255.0f);
      lut[i] = __float2bfloat16_rn(__fdiv_rn(__fsub_rn(f, mean), stdv));
    }
  }
  if (img >= n_images) return;
  if (j < 4) s_red[j] = 0;
  const int axis = j / kCrop;  // 0: horizontal (x), 1: vertical (y)
  const int o = j % kCrop;
  const int n_out = transform_out_size(transform);  // outputs per axis; the tables keep 224 entries, the rest is empty
  const int h = hw[2 * img], w = hw[2 * img + 1];
  const Geometry g = compute_geometry(h, w, transform);
  const int in_size = axis == 0 ? w : h;
  const int out_size = axis == 0 ? g.out_w : g.out_h;
  const int xx = (o < n_out ? o : n_out - 1) + (axis == 0 ? g.left : g.top);

  int32_t* base = plan + (static_cast<size_t>(img) * 2 + axis) * plan_ints_per_axis(max_taps);
  int32_t* first = base;
  int32_t* count = base + kCrop;
  int32_t* coef = base + 2 * kCrop + static_cast<size_t>(o) * max_taps;

  int xmin = xx, n = 1;
  if (in_size == out_size) {  // Pillow skips the pass: identity tap
    coef[0] = 1 << kPrecisionBits;
  } else {
    // Pillow precompute_coeffs (bilinear: support 1.0, Lanczos: support 3.0), evaluated in fp64 without FMA
    // contraction.
    const int filt = transform_filter(transform);
    const double scale = __ddiv_rn(static_cast<double>(in_size), static_cast<double>(out_size));
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = filt == 1 ? __dmul_rn(3.0, filterscale) : (filt == 2 ? __dmul_rn(2.0, filterscale) : filterscale);
    const double center = __dmul_rn(__dadd_rn(static_cast<double>(xx), 0.5), scale);
    const double ss = __ddiv_rn(1.0, filterscale);
    xmin = static_cast<int>(__dadd_rn(__dsub_rn(center, support), 0.5));
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(__dadd_rn(__dadd_rn(center, support), 0.5));
    if (xmax > in_size) xmax = in_size;
    n = xmax - xmin;
    if (n > max_taps) {
      atomicExch(status, 1);
      n = max_taps;
    }
    double ww = 0.0;
    for (int x = 0; x < n; ++x) {
      const double a = __dmul_rn(__dadd_rn(__dsub_rn(static_cast<double>(x + xmin), center), 0.5), ss);
      ww = __dadd_rn(ww, filter_weight(filt, a));
    }
    for (int x = 0; x < n; ++x) {
      const double a = __dmul_rn(__dadd_rn(__dsub_rn(static_cast<double>(x + xmin), center), 0.5), ss);
      double wgt = filter_weight(filt, a);
      if (ww != 0.0) wgt = __ddiv_rn(wgt, ww);
      // normalize_coeffs_8bpc: round half away from zero into 22-bit fixed point
      const double scaled = __dmul_rn(wgt, static_cast<double>(1 << kPrecisionBits));
      coef[x] = wgt < 0.0 ? static_cast<int>(__dadd_rn(-0.5, scaled)) : static_cast<int>(__dadd_rn(0.5, scaled));
    }
  }
  if (o >= n_out) n = 0;  // table entries past the transform's output size: no taps, nothing is read for them
  first[o] = xmin;
  count[o] = n;
  s_first[axis][o] = xmin;
  s_count[axis][o] = n;
  __syncthreads();
  atomicMax(&s_red[axis], n);
  __syncthreads();
  const int nth = s_red[0], ntv = s_red[1];
  // ---- schedule of the fused kernel: the largest band height whose source rows fit its intermediate ----
  const int col_first = s_first[0][0] * 3;
  const int width = (s_first[0][n_out - 1] + s_count[0][n_out - 1]) * 3 - col_first;
  const int pitch = ((width + 31) + 15) & ~15;
  const bool heavy = transform_filter(transform) != 0 || nth > 6;  // generic horizontal loop: queue first
  if (j < 32) {
    int th = 0;
    if (pitch <= kFStageBytes) {
      for (int cand = kFMaxBand; cand >= 1 && th == 0; cand >>= 1) {
        if (cand * ntv > kFVcoefInts) continue;
        int mx = 0;
        for (int b = j; b < n_out / cand; b += 32) {
          const int y0 = b * cand, y1 = y0 + cand - 1;
          mx = max(mx, s_first[1][y1] + s_count[1][y1] - s_first[1][y0]);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
        if (mx <= kFInterRows) th = cand;
      }
    }
    if (j == 0) {
      s_red[2] = th;
      s_red[3] = th > 0 ? atomicAdd(&counters[heavy ? 0 : 1], n_out / th) : 0;
      img_info[4 * img + 0] = th;
      img_info[4 * img + 1] = nth;
      img_info[4 * img + 2] = ntv;
      img_info[4 * img + 3] = 0;
    }
  }
  __syncthreads();
  const int th = s_red[2];
  if (th > 0 && j < n_out / th) {
    FusedItem it;
    const int y0 = j * th;
    it.img = img;
    it.band = j;
    it.th = th;
    it.nth = nth;
    it.ntv = ntv;
    it.row_lo = s_first[1][y0];
    it.n_rows = s_first[1][y0 + th - 1] + s_count[1][y0 + th - 1] - it.row_lo;
    it.col_first = col_first;
    it.width = width;
    it.row_bytes = w * 3;
    it.offset = offsets[img];
    it.pad0 = it.pad1 = it.pad2 = it.pad3 = 0;
    (heavy ? items_heavy : items_normal)[s_red[3] + j] = it;
  }
}

__device__ __forceinline__ int clip8_fixed(int acc) {
  const int v = acc >> kPrecisionBits;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// dynamic smem: int32 hfirst[224], hcount[224], hcoef[224*T], vfirst[8], vcount[8], vcoef[8*T];
//               bf16 lut[768]; uint8 hbuf[kChunkRows*672]; bf16 obuf[...]
template <int LAYOUT>
__global__ void __launch_bounds__(kThreads) resample_kernel(const uint8_t* __restrict__ pixels,
                                                            const int64_t* __restrict__ offsets,
                                                            const int32_t* __restrict__ hw, int max_taps,
                                                            const int32_t* __restrict__ plan,
                                                            const int32_t* __restrict__ img_info,
                                                            __nv_bfloat16* __restrict__ out, int n_out) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int T = max_taps;
  if (img_info[4 * blockIdx.y] != 0) return;  // the fused kernel takes this image
  if (static_cast<int>(blockIdx.x) * kBandRows >= n_out) return;  // band past the transform's output rows
  int32_t* hfirst = reinterpret_cast<int32_t*>(smem);
  int32_t* hcount = hfirst + kCrop;
  int32_t* hcoef = hcount + kCrop;
  int32_t* vfirst = hcoef + kCrop * T;
  int32_t* vcount = vfirst + kBandRows;
  int32_t* vcoef = vcount + kBandRows;
  __nv_bfloat16* lut = reinterpret_cast<__nv_bfloat16*>(vcoef + kBandRows * T);
  // keep 16-byte alignment for the staged output
  size_t used = reinterpret_cast<uint8_t*>(lut + 768) - smem;
  used = (used + 15) & ~static_cast<size_t>(15);
  uint8_t* hbuf = smem + used;
  used += static_cast<size_t>(kChunkRows) * kRowElems;
  used = (used + 15) & ~static_cast<size_t>(15);
  __nv_bfloat16* obuf = reinterpret_cast<__nv_bfloat16*>(smem + used);

  const int img = blockIdx.y;
  const int band = blockIdx.x;
  const int y0 = band * kBandRows;
  const int tid = threadIdx.x;
  const int w = hw[2 * img + 1];
  const uint8_t* src = pixels + offsets[img];
  const size_t row_bytes = static_cast<size_t>(w) * 3;

  // ---- load plan tables + build the normalisation LUT ----
  const int32_t* plan_h = plan + (static_cast<size_t>(img) * 2 + 0) * plan_ints_per_axis(T);
  const int32_t* plan_v = plan + (static_cast<size_t>(img) * 2 + 1) * plan_ints_per_axis(T);
  for (int i = tid; i < kCrop * (2 + T); i += kThreads) hfirst[i] = plan_h[i];
  if (tid < kBandRows) {
    vfirst[tid] = plan_v[y0 + tid];
    vcount[tid] = plan_v[kCrop + y0 + tid];
  }
  for (int i = tid; i < kBandRows * T; i += kThreads) vcoef[i] = plan_v[2 * kCrop + static_cast<size_t>(y0) * T + i];
  for (int i = tid; i < 768; i += kThreads) {
    const int c = i >> 8, v = i & 255;
    const float mean = c == 0 ? 0.485f : (c == 1 ? 0.456f : 0.406f);
    const float stdv = c == 0 ? 0.229f : (c == 1 ? 0.224f : 0.225f);
    const float f = __fdiv_rn(static_cast<float>(v), 255.0f);  // to_tensor: u8 -> float / 255
    lut[i] = __float2bfloat16_rn(__fdiv_rn(__fsub_rn(f, mean), stdv));
  }
  __syncthreads();

  // source rows this band needs: [row_lo, row_hi)
  int row_lo = vfirst[0], row_hi = vfirst[0] + vcount[0];
#pragma unroll
  for (int y = 1; y < kBandRows; ++y) {
    row_lo = min(row_lo, vfirst[y]);
    row_hi = max(row_hi, vfirst[y] + vcount[y]);
  }

  int acc[kBandRows][kElemsPerThread];
#pragma unroll
  for (int y = 0; y < kBandRows; ++y)
#pragma unroll
    for (int k = 0; k < kElemsPerThread; ++k) acc[y][k] = 1 << (kPrecisionBits - 1);

  for (int chunk = row_lo; chunk < row_hi; chunk += kChunkRows) {
    const int rows = min(kChunkRows, row_hi - chunk);
    // ---- horizontal pass: rows x 672 elements, rounded to uint8 like Pillow ----
    for (int idx = tid; idx < rows * kRowElems; idx += kThreads) {
      const int r = idx / kRowElems;
      const int e = idx - r * kRowElems;
      const int x = e / 3;
      const int c = e - x * 3;
      const int first = hfirst[x], n = hcount[x];
      const uint8_t* sp = src + static_cast<size_t>(chunk + r) * row_bytes + static_cast<size_t>(first) * 3 + c;
      const int32_t* cf = hcoef + x * T;
      int a = 1 << (kPrecisionBits - 1);
      for (int t = 0; t < n; ++t) a += static_cast<int>(__ldg(sp + 3 * t)) * cf[t];
      hbuf[idx] = static_cast<uint8_t>(clip8_fixed(a));
    }
    __syncthreads();
    // ---- vertical pass: accumulate the taps that fall in this chunk ----
#pragma unroll
    for (int y = 0; y < kBandRows; ++y) {
      const int vf = vfirst[y], vn = vcount[y];
      const int t_lo = max(0, chunk - vf), t_hi = min(vn, chunk + rows - vf);
      for (int t = t_lo; t < t_hi; ++t) {
        const int cf = vcoef[y * T + t];
        const uint8_t* hp = hbuf + (vf + t - chunk) * kRowElems;
#pragma unroll
        for (int k = 0; k < kElemsPerThread; ++k) {
          const int e = tid + k * kThreads;
          if (e < kRowElems) acc[y][k] += static_cast<int>(hp[e]) * cf;
        }
      }
    }
    __syncthreads();
  }

  // ---- normalise + stage ----
  if (LAYOUT == IRP_LAYOUT_U8_HWC) {
    // uint8 [n, n_out, n_out, 3]
    uint8_t* o8 = reinterpret_cast<uint8_t*>(out) + (static_cast<size_t>(img) * n_out + y0) * n_out * 3;
#pragma unroll
    for (int y = 0; y < kBandRows; ++y)
#pragma unroll
      for (int k = 0; k < kElemsPerThread; ++k) {
        const int e = tid + k * kThreads;
        if (e < n_out * 3) o8[static_cast<size_t>(y) * n_out * 3 + e] = static_cast<uint8_t>(clip8_fixed(acc[y][k]));
      }
  } else if (LAYOUT == IRP_LAYOUT_NHWC4P) {
    // obuf[8][230][4]; zero everything first (borders + pad channel)
    uint32_t* z = reinterpret_cast<uint32_t*>(obuf);
    for (int i = tid; i < kBandRows * kPad * 2; i += kThreads) z[i] = 0u;
    __syncthreads();
#pragma unroll
    for (int y = 0; y < kBandRows; ++y)
#pragma unroll
      for (int k = 0; k < kElemsPerThread; ++k) {
        const int e = tid + k * kThreads;
        if (e < kRowElems) {
          const int x = e / 3, c = e - x * 3;
          obuf[(y * kPad + 3 + x) * 4 + c] = lut[c * 256 + clip8_fixed(acc[y][k])];
        }
      }
    __syncthreads();
    // rows y0..y0+7 of image -> padded rows 3+y0.. ; each padded row is 230*8 = 1840 bytes = 115 uint4
    constexpr int kVecPerRow = kPad * 8 / 16;
    uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<size_t>(img) * kPad + 3 + y0) * kPad * 4);
    const uint4* sv = reinterpret_cast<const uint4*>(obuf);
    for (int i = tid; i < kBandRows * kVecPerRow; i += kThreads) dst[i] = sv[i];
    // top / bottom zero borders
    if (band == 0) {
      uint4* top = reinterpret_cast<uint4*>(out + static_cast<size_t>(img) * kPad * kPad * 4);
      for (int i = tid; i < 3 * kVecPerRow; i += kThreads) top[i] = make_uint4(0, 0, 0, 0);
    }
    if (band == gridDim.x - 1) {
      uint4* bot = reinterpret_cast<uint4*>(out + (static_cast<size_t>(img) * kPad + 3 + kCrop) * kPad * 4);
      for (int i = tid; i < 3 * kVecPerRow; i += kThreads) bot[i] = make_uint4(0, 0, 0, 0);
    }
  } else {
    // obuf[3][8][224]
#pragma unroll
    for (int y = 0; y < kBandRows; ++y)
#pragma unroll
      for (int k = 0; k < kElemsPerThread; ++k) {
        const int e = tid + k * kThreads;
        if (e < kRowElems) {
          const int x = e / 3, c = e - x * 3;
          obuf[(c * kBandRows + y) * kCrop + x] = lut[c * 256 + clip8_fixed(acc[y][k])];
        }
      }
    __syncthreads();
    constexpr int kVecPerRow = kCrop * 2 / 16;  // 28
    const uint4* sv = reinterpret_cast<const uint4*>(obuf);
    for (int i = tid; i < 3 * kBandRows * kVecPerRow; i += kThreads) {
      const int v = i % kVecPerRow;
      const int y = (i / kVecPerRow) % kBandRows;
      const int c = i / (kVecPerRow * kBandRows);
      uint4* dst = reinterpret_cast<uint4*>(out + ((static_cast<size_t>(img) * 3 + c) * kCrop + y0 + y) * kCrop);
      dst[v] = sv[i];
    }
  }
}


static size_t resample_smem_bytes(int max_taps, int layout) {
  size_t b = static_cast<size_t>(kCrop) * (2 + max_taps) * 4 + static_cast<size_t>(kBandRows) * (2 + max_taps) * 4 +
             768 * 2;
  b = (b + 15) & ~static_cast<size_t>(15);
  b += static_cast<size_t>(kChunkRows) * kRowElems;
  b = (b + 15) & ~static_cast<size_t>(15);
  b += layout == IRP_LAYOUT_NHWC4P ? static_cast<size_t>(kBandRows) * kPad * 4 * 2
                                   : static_cast<size_t>(3) * kBandRows * kCrop * 2;
  return b;
}

}  // namespace irp

using namespace irp;

// Images the fused kernel cannot take (a band of ONE output row would need more than its 64 intermediate rows:
// downscales beyond ~30x) run through the band kernel on a side stream, concurrently with the fused kernel (fork
// after the schedule kernel, join after the fused kernel); the two kernels write disjoint images.
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
static int side_stream(SideStream** out) {
  static SideStream per_device[64];
  int dev = 0;
  IRP_CUDA_OK(cudaGetDevice(&dev));
  IRP_REQUIRE(dev >= 0 && dev < 64, "preprocess: device index %d", dev);
  SideStream& s = per_device[dev];
  if (s.stream == nullptr) {
    int least = 0, greatest = 0;
    IRP_CUDA_OK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    IRP_CUDA_OK(cudaStreamCreateWithPriority(&s.stream, cudaStreamNonBlocking, greatest));
    IRP_CUDA_OK(cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming));
    IRP_CUDA_OK(cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming));
  }
  *out = &s;
  return IRP_OK;
}

// smallest tap bound at which an image can exceed the fused kernel's envelope (2*ceil(scale)+1 with 2*scale+1
// source rows per output row > 64, or a staged row wider than a staging buffer)
constexpr int kFusedSafeTaps = 57;

template <int LAYOUT>
static int launch_generic(const FusedParams& fp, const int64_t* d_offsets, const int32_t* d_hw, int n_images,
                          cudaStream_t st) {
  const size_t smem = resample_smem_bytes(fp.max_taps, LAYOUT);
  IRP_REQUIRE(smem <= 227 * 1024, "preprocess: max_taps %d needs %zu bytes of shared memory", fp.max_taps, smem);
  auto k = resample_kernel<LAYOUT>;
  IRP_TRY(ensure_smem(k, smem));
  dim3 grid(kCrop / kBandRows, n_images);
  k<<<grid, kThreads, smem, st>>>(fp.pixels, d_offsets, d_hw, fp.max_taps, fp.plan, fp.img_info,
                                  static_cast<__nv_bfloat16*>(fp.out), fp.out_size);
  IRP_CUDA_OK(cudaGetLastError());
  return IRP_OK;
}

template <int LAYOUT, bool SIGNED>
static int launch_fused(const FusedParams& fp, const int64_t* d_offsets, const int32_t* d_hw, int n_images,
                        cudaStream_t st) {
  auto k = resample_fused_kernel<LAYOUT, SIGNED>;
  IRP_TRY(ensure_smem(k, kFSmemBytes));
  // persistent: two CTAs per SM, work items fetched from the schedule's lists
  const long long max_items = static_cast<long long>(n_images) * kCrop;
  const long long want = 2LL * num_sms();
  const int grid = static_cast<int>(max_items < want ? max_items : want);
  SideStream* side = nullptr;
  const bool may_overflow = fp.max_taps >= kFusedSafeTaps;
  if (may_overflow) {
    IRP_TRY(side_stream(&side));
    IRP_CUDA_OK(cudaEventRecord(side->fork, st));
    IRP_CUDA_OK(cudaStreamWaitEvent(side->stream, side->fork, 0));
    IRP_TRY(launch_generic<LAYOUT>(fp, d_offsets, d_hw, n_images, side->stream));
    IRP_CUDA_OK(cudaEventRecord(side->join, side->stream));
  }
  k<<<grid, kFThreads, kFSmemBytes, st>>>(fp);
  IRP_CUDA_OK(cudaGetLastError());
  if (may_overflow) IRP_CUDA_OK(cudaStreamWaitEvent(st, side->join, 0));
  return IRP_OK;
}

namespace {
struct PreWs {
  int32_t* plan;
  int32_t* status;    // [4]
  int32_t* counters;  // [4]
  int32_t* img_info;  // [n][4]
  FusedItem* items_heavy;
  FusedItem* items_normal;
  __nv_bfloat16* lut;
  size_t zero_bytes;  // status + counters (contiguous)
  size_t total;
};
PreWs pre_layout(void* ws, int n_images, int max_taps) {
  PreWs w;
  uint8_t* p = static_cast<uint8_t*>(ws);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    uint8_t* r = p ? p + off : nullptr;
    off += (bytes + 255) & ~static_cast<size_t>(255);
    return r;
  };
  w.plan = reinterpret_cast<int32_t*>(take(static_cast<size_t>(n_images) * 2 * plan_ints_per_axis(max_taps) * 4));
  w.status = reinterpret_cast<int32_t*>(take(32));
  w.counters = w.status ? w.status + 4 : nullptr;
  w.zero_bytes = 32;
  w.img_info = reinterpret_cast<int32_t*>(take(static_cast<size_t>(n_images) * 16));
  w.items_heavy = reinterpret_cast<FusedItem*>(take(static_cast<size_t>(n_images) * kCrop * sizeof(FusedItem)));
  w.items_normal = reinterpret_cast<FusedItem*>(take(static_cast<size_t>(n_images) * kCrop * sizeof(FusedItem)));
  w.lut = reinterpret_cast<__nv_bfloat16*>(take(768 * 2));
  w.total = off;
  return w;
}
}  // namespace

extern "C" {

size_t irp_preprocess_workspace_bytes(int n_images, int max_taps) {
  if (n_images <= 0 || max_taps <= 0) return 0;
  return pre_layout(nullptr, n_images, max_taps).total;
}

int irp_preprocess(const uint8_t* d_pixels, const int64_t* d_offsets, const int32_t* d_hw, int n_images,
                   int max_taps, void* d_workspace, size_t workspace_bytes, void* d_out, int out_layout,
                   void* stream) {
  return irp_preprocess_ex(d_pixels, d_offsets, d_hw, n_images, max_taps, d_workspace, workspace_bytes, d_out,
                           out_layout, IRP_TRANSFORM_WEIGHTS_DEFAULT, stream);
}

int irp_preprocess_ex(const uint8_t* d_pixels, const int64_t* d_offsets, const int32_t* d_hw, int n_images,
                      int max_taps, void* d_workspace, size_t workspace_bytes, void* d_out, int out_layout,
                      int transform, void* stream) {
  IRP_REQUIRE(d_pixels && d_offsets && d_hw && d_workspace && d_out, "preprocess: null argument");
  IRP_REQUIRE(transform == IRP_TRANSFORM_WEIGHTS_DEFAULT || transform == IRP_TRANSFORM_VAL_256 ||
                  transform == IRP_TRANSFORM_WDS_LANCZOS || transform == IRP_TRANSFORM_HASH_64,
              "preprocess: unknown transform %d", transform);
  IRP_REQUIRE(transform != IRP_TRANSFORM_HASH_64 || out_layout == IRP_LAYOUT_U8_HWC,
              "preprocess: IRP_TRANSFORM_HASH_64 writes uint8 [n,64,64,3] (IRP_LAYOUT_U8_HWC) only");
  IRP_REQUIRE(n_images > 0 && n_images < (1 << 23), "preprocess: n_images %d", n_images);
  IRP_REQUIRE(max_taps >= 3 && max_taps <= 513, "preprocess: max_taps %d out of range", max_taps);
  IRP_REQUIRE(out_layout == IRP_LAYOUT_NCHW || out_layout == IRP_LAYOUT_NHWC4P || out_layout == IRP_LAYOUT_U8_HWC,
              "preprocess: bad layout %d", out_layout);
  IRP_REQUIRE(workspace_bytes >= irp_preprocess_workspace_bytes(n_images, max_taps),
              "preprocess: workspace %zu < %zu bytes", workspace_bytes,
              irp_preprocess_workspace_bytes(n_images, max_taps));
  IRP_REQUIRE((reinterpret_cast<uintptr_t>(d_pixels) & 15u) == 0, "preprocess: d_pixels must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const PreWs w = pre_layout(d_workspace, n_images, max_taps);
  IRP_CUDA_OK(cudaMemsetAsync(w.status, 0, w.zero_bytes, st));
  resample_plan_kernel<<<n_images, 2 * kCrop, 0, st>>>(d_hw, d_offsets, n_images, max_taps, transform, w.plan,
                                                       w.status, w.lut, w.img_info, w.counters, w.items_heavy,
                                                       w.items_normal);
  IRP_CUDA_OK(cudaGetLastError());
  const bool lanczos = transform_filter(transform) != 0;  // Lanczos / bicubic: negative weights, signed arithmetic
  FusedParams fp;
  fp.pixels = d_pixels;
  fp.plan = w.plan;
  fp.img_info = w.img_info;
  fp.status = w.status;
  fp.counters = w.counters;
  fp.items_heavy = w.items_heavy;
  fp.items_normal = w.items_normal;
  fp.lut = w.lut;
  fp.out = d_out;
  fp.max_taps = max_taps;
  fp.out_size = transform_out_size(transform);
  if (out_layout == IRP_LAYOUT_U8_HWC)
    return lanczos ? launch_fused<IRP_LAYOUT_U8_HWC, true>(fp, d_offsets, d_hw, n_images, st)
                   : launch_fused<IRP_LAYOUT_U8_HWC, false>(fp, d_offsets, d_hw, n_images, st);
  if (out_layout == IRP_LAYOUT_NHWC4P)
    return lanczos ? launch_fused<IRP_LAYOUT_NHWC4P, true>(fp, d_offsets, d_hw, n_images, st)
                   : launch_fused<IRP_LAYOUT_NHWC4P, false>(fp, d_offsets, d_hw, n_images, st);
  return lanczos ? launch_fused<IRP_LAYOUT_NCHW, true>(fp, d_offsets, d_hw, n_images, st)
                 : launch_fused<IRP_LAYOUT_NCHW, false>(fp, d_offsets, d_hw, n_images, st);
}

int irp_preprocess_status(const void* d_workspace, int n_images, int max_taps, void* stream) {
  IRP_REQUIRE(d_workspace != nullptr && n_images > 0 && max_taps > 0, "preprocess_status: bad argument");
  const PreWs w = pre_layout(const_cast<void*>(d_workspace), n_images, max_taps);
  int32_t h = 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  IRP_CUDA_OK(cudaMemcpyAsync(&h, w.status, sizeof(h), cudaMemcpyDeviceToHost, st));
  IRP_CUDA_OK(cudaStreamSynchronize(st));
  IRP_REQUIRE(h == 0, "preprocess: max_taps %d is too small for at least one image of the batch (its filter was "
                      "truncated; bf16 outputs were written as NaN)", max_taps);
  return IRP_OK;
}

/* Host-side view of the resize/crop geometry (used by the Python mirror to size max_taps and by tests). */
int irp_preprocess_geometry(int h, int w, int* out_h, int* out_w, int* top, int* left, int* taps) {
  return irp_preprocess_geometry_ex(h, w, IRP_TRANSFORM_WEIGHTS_DEFAULT, out_h, out_w, top, left, taps);
}

int irp_preprocess_geometry_ex(int h, int w, int transform, int* out_h, int* out_w, int* top, int* left, int* taps) {
  IRP_REQUIRE(h > 0 && w > 0, "geometry: bad size %dx%d", h, w);
  IRP_REQUIRE(transform == IRP_TRANSFORM_WEIGHTS_DEFAULT || transform == IRP_TRANSFORM_VAL_256 ||
                  transform == IRP_TRANSFORM_WDS_LANCZOS || transform == IRP_TRANSFORM_HASH_64,
              "geometry: unknown transform %d", transform);
  const Geometry g = compute_geometry(h, w, transform);
  if (out_h) *out_h = g.out_h;
  if (out_w) *out_w = g.out_w;
  if (top) *top = g.top;
  if (left) *left = g.left;
  if (taps) {
    const double sx = static_cast<double>(w) / g.out_w, sy = static_cast<double>(h) / g.out_h;
    double s = sx > sy ? sx : sy;
    if (s < 1.0) s = 1.0;
    if (transform == IRP_TRANSFORM_WDS_LANCZOS) s *= 3.0;  // Lanczos support
    if (transform == IRP_TRANSFORM_HASH_64) s *= 2.0;      // bicubic support
    int c = static_cast<int>(s);
    if (static_cast<double>(c) < s) ++c;
    *taps = 2 * c + 1;
  }
  return IRP_OK;
}

}  // extern "C"
