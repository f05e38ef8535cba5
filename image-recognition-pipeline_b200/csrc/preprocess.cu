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
constexpr int kFastTaps = 6;         // tap bound of the shared-memory-staged fast path (downscale <= 2.5x)

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

__global__ void resample_plan_kernel(const int32_t* __restrict__ hw, int n_images, int max_taps, int transform,
                                     int32_t* __restrict__ plan, int32_t* __restrict__ status,
                                     int32_t* __restrict__ img_taps, __nv_bfloat16* __restrict__ lut) {
  const int img = blockIdx.x;
  const int j = threadIdx.x;
  if (img == 0 && lut != nullptr) {
    // normalisation table for the two-pass path: lut[c*256 + v] = bf16(((v / 255) - mean_c) / std_c)
    for (int i = j; i < 768; i += blockDim.x) {
      const int c = i >> 8, v = i & 255;
      const float mean = c == 0 ? 0.485f : (c == 1 ? 0.456f : 0.406f);
      const float stdv = c == 0 ? 0.229f : (c == 1 ? 0.224f : 0.225f);
      const float f = __fdiv_rn(static_cast<float>(v), 255.0f);
      lut[i] = __float2bfloat16_rn(__fdiv_rn(__fsub_rn(f, mean), stdv));
    }
  }
  if (img >= n_images || j >= 2 * kCrop) return;
  const int axis = j / kCrop;  // 0: horizontal (x), 1: vertical (y)
  const int o = j % kCrop;
  const int h = hw[2 * img], w = hw[2 * img + 1];
  const Geometry g = compute_geometry(h, w, transform);
  const int in_size = axis == 0 ? w : h;
  const int out_size = axis == 0 ? g.out_w : g.out_h;
  const int xx = o + (axis == 0 ? g.left : g.top);

  int32_t* base = plan + (static_cast<size_t>(img) * 2 + axis) * plan_ints_per_axis(max_taps);
  int32_t* first = base;
  int32_t* count = base + kCrop;
  int32_t* coef = base + 2 * kCrop + static_cast<size_t>(o) * max_taps;

  if (in_size == out_size) {  // Pillow skips the pass: identity tap
    first[o] = xx;
    count[o] = 1;
    coef[0] = 1 << kPrecisionBits;
    atomicMax(&img_taps[img], 1);
    return;
  }
  // Pillow precompute_coeffs (bilinear: support 1.0, Lanczos: support 3.0), evaluated in fp64 without FMA
  // contraction.
  const bool lanczos = transform == IRP_TRANSFORM_WDS_LANCZOS;
  const double scale = __ddiv_rn(static_cast<double>(in_size), static_cast<double>(out_size));
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = lanczos ? __dmul_rn(3.0, filterscale) : filterscale;
  const double center = __dmul_rn(__dadd_rn(static_cast<double>(xx), 0.5), scale);
  const double ss = __ddiv_rn(1.0, filterscale);
  int xmin = static_cast<int>(__dadd_rn(__dsub_rn(center, support), 0.5));
  if (xmin < 0) xmin = 0;
  int xmax = static_cast<int>(__dadd_rn(__dadd_rn(center, support), 0.5));
  if (xmax > in_size) xmax = in_size;
  int n = xmax - xmin;
  if (n > max_taps) {
    atomicExch(status, 1);
    n = max_taps;
  }
  double ww = 0.0;
  for (int x = 0; x < n; ++x) {
    const double a = __dmul_rn(__dadd_rn(__dsub_rn(static_cast<double>(x + xmin), center), 0.5), ss);
    ww = __dadd_rn(ww, lanczos ? lanczos3_weight(a) : triangle_weight(a));
  }
  for (int x = 0; x < n; ++x) {
    const double a = __dmul_rn(__dadd_rn(__dsub_rn(static_cast<double>(x + xmin), center), 0.5), ss);
    double wgt = lanczos ? lanczos3_weight(a) : triangle_weight(a);
    if (ww != 0.0) wgt = __ddiv_rn(wgt, ww);
    // normalize_coeffs_8bpc: round half away from zero into 22-bit fixed point
    const double scaled = __dmul_rn(wgt, static_cast<double>(1 << kPrecisionBits));
    coef[x] = wgt < 0.0 ? static_cast<int>(__dadd_rn(-0.5, scaled)) : static_cast<int>(__dadd_rn(0.5, scaled));
  }
  first[o] = xmin;
  count[o] = n;
  atomicMax(&img_taps[img], n);
}

__device__ __forceinline__ int clip8_fixed(int acc) {
  const int v = acc >> kPrecisionBits;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

constexpr int kTwoPassMaxRows = 640;  // source rows of the crop window an image may need to take this path
constexpr int kTwoPassTaps = 16;      // and its largest tap count (<= kFastTaps: unrolled register-resident weights)

// Images whose crop window needs at most kTwoPassMaxRows source rows and kTwoPassTaps taps take the two-pass
// path; anything beyond that (downscales by more than ~2.7x with the Lanczos filter, ~2.8x bilinear) the generic
// band kernel.
__device__ __forceinline__ bool two_pass_image(const int32_t* __restrict__ plan_v, int taps) {
  const int span = plan_v[kCrop - 1] + plan_v[2 * kCrop - 1] - plan_v[0];
  return taps <= kTwoPassTaps && span <= kTwoPassMaxRows;
}

// dynamic smem: int32 hfirst[224], hcount[224], hcoef[224*T], vfirst[8], vcount[8], vcoef[8*T];
//               bf16 lut[768]; uint8 hbuf[kChunkRows*672]; bf16 obuf[...]
template <int LAYOUT>
__global__ void __launch_bounds__(kThreads) resample_kernel(const uint8_t* __restrict__ pixels,
                                                            const int64_t* __restrict__ offsets,
                                                            const int32_t* __restrict__ hw, int max_taps,
                                                            const int32_t* __restrict__ plan,
                                                            const int32_t* __restrict__ img_taps,
                                                            __nv_bfloat16* __restrict__ out, int two_pass) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int T = max_taps;
  {  // images the fast kernels of the selected mode handle are skipped here
    const int taps = img_taps[blockIdx.y];
    const int32_t* pv = plan + (static_cast<size_t>(blockIdx.y) * 2 + 1) * plan_ints_per_axis(T);
    if (two_pass ? two_pass_image(pv, taps) : taps <= kFastTaps) return;
  }
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
    uint8_t* o8 = reinterpret_cast<uint8_t*>(out) + (static_cast<size_t>(img) * kCrop + y0) * kRowElems;
#pragma unroll
    for (int y = 0; y < kBandRows; ++y)
#pragma unroll
      for (int k = 0; k < kElemsPerThread; ++k) {
        const int e = tid + k * kThreads;
        if (e < kRowElems) o8[static_cast<size_t>(y) * kRowElems + e] = static_cast<uint8_t>(clip8_fixed(acc[y][k]));
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


// ------------------------------------------------------------------------------------------------------------
// Fast path (<= kFastTaps taps per output, i.e. downscale factors up to 2.5): the source rows a band needs are
// first staged in shared memory with 16-byte coalesced loads (rows are realigned to 16 bytes, the per-row byte
// shift is remembered), the horizontal pass then reads bytes from shared memory with its tap weights held in
// registers, the vertical pass reads the uint8 intermediate rows from shared memory.
// dynamic smem: int32 vfirst[TH], vcount[TH], vcoef[TH*8], rshift[kFastRows]; bf16 lut[768];
//               uint8 hbuf[kFastRows*672]; uint8 inbuf[kFastInBytes + 64]; bf16 obuf[...]
// ------------------------------------------------------------------------------------------------------------
constexpr int kFastRows = 24;             // source rows staged per pass
constexpr int kFastInBytes = 28 * 1024;   // staging budget for those rows

// NT = tap count every output of this image is padded to (zero weights beyond its own count), so the inner
// loops are fully unrolled without predication.
template <int LAYOUT, int TH, int NT>
__device__ __forceinline__ void resample_fast_body(uint8_t* smem, const uint8_t* __restrict__ pixels,
                                                   const int64_t* __restrict__ offsets,
                                                   const int32_t* __restrict__ hw, int max_taps,
                                                   const int32_t* __restrict__ plan,
                                                   __nv_bfloat16* __restrict__ out) {
  const int img = blockIdx.y;
  const int T = max_taps;
  int32_t* vfirst = reinterpret_cast<int32_t*>(smem);
  int32_t* vcount = vfirst + TH;
  int32_t* vcoef = vcount + TH;                 // [TH][kFastTaps] (first NT used)
  int32_t* rshift = vcoef + TH * kFastTaps;     // [kFastRows]
  __nv_bfloat16* lut = reinterpret_cast<__nv_bfloat16*>(rshift + kFastRows);
  uint8_t* hbuf = reinterpret_cast<uint8_t*>(lut + 768);
  uint8_t* inbuf = hbuf + kFastRows * kRowElems;
  __nv_bfloat16* obuf = reinterpret_cast<__nv_bfloat16*>(inbuf + kFastInBytes + 64);

  const int band = blockIdx.x;
  const int y0 = band * TH;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int w = hw[2 * img + 1];
  const uint8_t* src = pixels + offsets[img];
  const size_t row_bytes = static_cast<size_t>(w) * 3;
  const int32_t* plan_h = plan + (static_cast<size_t>(img) * 2 + 0) * plan_ints_per_axis(T);
  const int32_t* plan_v = plan + (static_cast<size_t>(img) * 2 + 1) * plan_ints_per_axis(T);

  // ---- vertical tables, LUT ----
  if (tid < TH) {
    vfirst[tid] = plan_v[y0 + tid];
    vcount[tid] = plan_v[kCrop + y0 + tid];
  }
  for (int i = tid; i < TH * kFastTaps; i += kThreads) {
    const int y = i / kFastTaps, t = i % kFastTaps;
    // weights past the output's own tap count are not written by the plan kernel: pad with zeros
    vcoef[i] = t < plan_v[kCrop + y0 + y] ? plan_v[2 * kCrop + static_cast<size_t>(y0 + y) * T + t] : 0;
  }
  for (int i = tid; i < 768; i += kThreads) {
    const int c = i >> 8, v = i & 255;
    const float mean = c == 0 ? 0.485f : (c == 1 ? 0.456f : 0.406f);
    const float stdv = c == 0 ? 0.229f : (c == 1 ? 0.224f : 0.225f);
    const float f = __fdiv_rn(static_cast<float>(v), 255.0f);
    lut[i] = __float2bfloat16_rn(__fdiv_rn(__fsub_rn(f, mean), stdv));
  }
  // ---- this thread's horizontal taps (fixed (x, c) elements for every row) ----
  const int col_first = plan_h[0] * 3;                                      // first source byte any output needs
  const int col_last = (plan_h[kCrop - 1] + plan_h[2 * kCrop - 1]) * 3;     // one past the last
  const int width_bytes = col_last - col_first;
  const int pitch = ((width_bytes + 15) & ~15) + 16;
  int hoff[kElemsPerThread], hcf[kElemsPerThread][NT];
#pragma unroll
  for (int k = 0; k < kElemsPerThread; ++k) {
    const int e = min(tid + k * kThreads, kRowElems - 1);  // surplus threads shadow the last element
    const int x = e / 3, c = e - x * 3;
    const int n = plan_h[kCrop + x];
    hoff[k] = plan_h[x] * 3 + c - col_first;
#pragma unroll
    for (int t = 0; t < NT; ++t) hcf[k][t] = t < n ? plan_h[2 * kCrop + static_cast<size_t>(x) * T + t] : 0;
  }
  __syncthreads();

  int row_lo = vfirst[0], row_hi = vfirst[0] + vcount[0];
#pragma unroll
  for (int y = 1; y < TH; ++y) {
    row_lo = min(row_lo, vfirst[y]);
    row_hi = max(row_hi, vfirst[y] + vcount[y]);
  }
  int rows_cap = kFastInBytes / pitch;
  if (rows_cap > kFastRows) rows_cap = kFastRows;

  int acc[TH][kElemsPerThread];
#pragma unroll
  for (int y = 0; y < TH; ++y)
#pragma unroll
    for (int k = 0; k < kElemsPerThread; ++k) acc[y][k] = 1 << (kPrecisionBits - 1);

  const int vec_per_row = pitch >> 4;
  for (int chunk = row_lo; chunk < row_hi; chunk += rows_cap) {
    const int rows = min(rows_cap, row_hi - chunk);
    // ---- stage source rows: 16-byte loads from the 16-byte-aligned address at or before the first byte ----
    for (int r = warp; r < rows; r += kThreads / 32) {
      const uintptr_t a = reinterpret_cast<uintptr_t>(src + static_cast<size_t>(chunk + r) * row_bytes + col_first);
      const uintptr_t al = a & ~static_cast<uintptr_t>(15);
      const int shift = static_cast<int>(a - al);
      if (lane == 0) rshift[r] = shift;
      const int nvec = (shift + width_bytes + 15) >> 4;
      const uint4* gp = reinterpret_cast<const uint4*>(al);
      uint4* sp = reinterpret_cast<uint4*>(inbuf + static_cast<size_t>(r) * pitch);
      for (int v = lane; v < nvec && v < vec_per_row; v += 32) sp[v] = __ldg(gp + v);
    }
    __syncthreads();
    // ---- horizontal pass from shared memory ----
    for (int r = 0; r < rows; ++r) {
      const uint8_t* rp = inbuf + static_cast<size_t>(r) * pitch + rshift[r];
#pragma unroll
      for (int k = 0; k < kElemsPerThread; ++k) {
        const int e = tid + k * kThreads;
        const uint8_t* sp = rp + hoff[k];
        int a = 1 << (kPrecisionBits - 1);
#pragma unroll
        for (int t = 0; t < NT; ++t) a += static_cast<int>(sp[3 * t]) * hcf[k][t];  // padded taps read slack, weight 0
        if (e < kRowElems) hbuf[r * kRowElems + e] = static_cast<uint8_t>(clip8_fixed(a));
      }
    }
    __syncthreads();
    // ---- vertical pass: taps that fall in this chunk ----
#pragma unroll
    for (int y = 0; y < TH; ++y) {
      const int vf = vfirst[y];
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const int rr = vf + t - chunk;  // staged row of this tap; weight is 0 beyond the output's own count
        const int cf = (rr >= 0 && rr < rows) ? vcoef[y * kFastTaps + t] : 0;
        const uint8_t* hp = hbuf + min(max(rr, 0), kFastRows - 1) * kRowElems;
#pragma unroll
        for (int k = 0; k < kElemsPerThread; ++k) {
          const int e = min(tid + k * kThreads, kRowElems - 1);
          acc[y][k] += static_cast<int>(hp[e]) * cf;
        }
      }
    }
    __syncthreads();
  }

  // ---- normalise + stage + 16-byte stores ----
  if (LAYOUT == IRP_LAYOUT_NHWC4P) {
    uint32_t* z = reinterpret_cast<uint32_t*>(obuf);
    for (int i = tid; i < TH * kPad * 2; i += kThreads) z[i] = 0u;
    __syncthreads();
#pragma unroll
    for (int y = 0; y < TH; ++y)
#pragma unroll
      for (int k = 0; k < kElemsPerThread; ++k) {
        const int e = tid + k * kThreads;
        if (e < kRowElems) {
          const int x = e / 3, c = e - x * 3;
          obuf[(y * kPad + 3 + x) * 4 + c] = lut[c * 256 + clip8_fixed(acc[y][k])];
        }
      }
    __syncthreads();
    constexpr int kVecPerRow = kPad * 8 / 16;
    uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<size_t>(img) * kPad + 3 + y0) * kPad * 4);
    const uint4* sv = reinterpret_cast<const uint4*>(obuf);
    for (int i = tid; i < TH * kVecPerRow; i += kThreads) dst[i] = sv[i];
    if (band == 0) {
      uint4* top = reinterpret_cast<uint4*>(out + static_cast<size_t>(img) * kPad * kPad * 4);
      for (int i = tid; i < 3 * kVecPerRow; i += kThreads) top[i] = make_uint4(0, 0, 0, 0);
    }
    if (band == gridDim.x - 1) {
      uint4* bot = reinterpret_cast<uint4*>(out + (static_cast<size_t>(img) * kPad + 3 + kCrop) * kPad * 4);
      for (int i = tid; i < 3 * kVecPerRow; i += kThreads) bot[i] = make_uint4(0, 0, 0, 0);
    }
  } else {
#pragma unroll
    for (int y = 0; y < TH; ++y)
#pragma unroll
      for (int k = 0; k < kElemsPerThread; ++k) {
        const int e = tid + k * kThreads;
        if (e < kRowElems) {
          const int x = e / 3, c = e - x * 3;
          obuf[(c * TH + y) * kCrop + x] = lut[c * 256 + clip8_fixed(acc[y][k])];
        }
      }
    __syncthreads();
    constexpr int kVecPerRow = kCrop * 2 / 16;  // 28
    const uint4* sv = reinterpret_cast<const uint4*>(obuf);
    for (int i = tid; i < 3 * TH * kVecPerRow; i += kThreads) {
      const int v = i % kVecPerRow;
      const int y = (i / kVecPerRow) % TH;
      const int c = i / (kVecPerRow * TH);
      uint4* dst = reinterpret_cast<uint4*>(out + ((static_cast<size_t>(img) * 3 + c) * kCrop + y0 + y) * kCrop);
      dst[v] = sv[i];
    }
  }
}

template <int LAYOUT, int TH>
__global__ void __launch_bounds__(kThreads, (TH <= 8 ? 3 : (TH <= 16 ? 2 : 1))) resample_fast_kernel(const uint8_t* __restrict__ pixels,
                                                                    const int64_t* __restrict__ offsets,
                                                                    const int32_t* __restrict__ hw, int max_taps,
                                                                    const int32_t* __restrict__ plan,
                                                                    const int32_t* __restrict__ img_taps,
                                                                    __nv_bfloat16* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int nt = img_taps[blockIdx.y];
  if (nt > kFastTaps) return;  // handled by the generic kernel
  if (nt <= 2) resample_fast_body<LAYOUT, TH, 2>(smem, pixels, offsets, hw, max_taps, plan, out);
  else if (nt == 3) resample_fast_body<LAYOUT, TH, 3>(smem, pixels, offsets, hw, max_taps, plan, out);
  else if (nt == 4) resample_fast_body<LAYOUT, TH, 4>(smem, pixels, offsets, hw, max_taps, plan, out);
  else if (nt == 5) resample_fast_body<LAYOUT, TH, 5>(smem, pixels, offsets, hw, max_taps, plan, out);
  else resample_fast_body<LAYOUT, TH, 6>(smem, pixels, offsets, hw, max_taps, plan, out);
}

// ------------------------------------------------------------------------------------------------------------
// Two-pass path (default for images with <= kFastTaps taps): no shared memory, no barriers, full occupancy.
//   hpass_kernel : one thread per (source row inside the crop window, output column): the horizontal filter is
//                  applied ONCE per source row (the band kernels recompute the rows shared by adjacent bands) and
//                  its rounded uint8 result goes to an intermediate [image][row][224*3] that stays in L2;
//   vpass_kernel : one thread per padded output pixel: vertical filter over the intermediate rows, Pillow's
//                  rounding, LUT normalisation, and one 8-byte store of the (R, G, B, 0) bf16 pixel (NHWC4P) or
//                  three 2-byte stores (NCHW); the threads of the 3-pixel border write zeros.
// The arithmetic is the band kernels' (and Pillow's): acc = 2^21 + sum src * coef, clip8(acc >> 22), twice.
// ------------------------------------------------------------------------------------------------------------
constexpr int kHRowsPerCta = 16;
constexpr int kVRowsPerCta = 8;  // (padded) output rows per CTA of the vertical pass

__global__ void __launch_bounds__(kCrop) hpass_kernel(const uint8_t* __restrict__ pixels,
                                                     const int64_t* __restrict__ offsets,
                                                     const int32_t* __restrict__ hw, int max_taps,
                                                     const int32_t* __restrict__ plan,
                                                     const int32_t* __restrict__ img_taps,
                                                     uint8_t* __restrict__ inter) {
  const int img = blockIdx.y;
  const int nt = img_taps[img];
  const int T = max_taps;
  const int32_t* plan_h = plan + (static_cast<size_t>(img) * 2 + 0) * plan_ints_per_axis(T);
  const int32_t* plan_v = plan + (static_cast<size_t>(img) * 2 + 1) * plan_ints_per_axis(T);
  if (!two_pass_image(plan_v, nt)) return;  // the generic kernel handles this image
  const int row_lo = plan_v[0];
  const int row_hi = plan_v[kCrop - 1] + plan_v[2 * kCrop - 1];  // first + count of the last output row
  const int r0 = row_lo + blockIdx.x * kHRowsPerCta;
  if (r0 >= row_hi) return;
  const int x = threadIdx.x;
  const int first = plan_h[x], n = plan_h[kCrop + x];
  const int w = hw[2 * img + 1];
  const size_t row_bytes = static_cast<size_t>(w) * 3;
  const uint8_t* src = pixels + offsets[img] + static_cast<size_t>(first) * 3;
  uint8_t* dst = inter + (static_cast<size_t>(img) * kTwoPassMaxRows) * kRowElems + x * 3;
  if (nt > kFastTaps) {
    // 7..kTwoPassTaps taps (Lanczos, or bilinear downscales by 2.5-2.8x): weights stay in the (L1-cached) plan
    const int32_t* cfp = plan_h + 2 * kCrop + static_cast<size_t>(x) * T;
    for (int rr = 0; rr < kHRowsPerCta; ++rr) {
      const int r = r0 + rr;
      if (r >= row_hi) break;
      const uint8_t* sp = src + static_cast<size_t>(r) * row_bytes;
      int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
#pragma unroll 4
      for (int t = 0; t < n; ++t) {
        const int c = __ldg(cfp + t);
        a0 += static_cast<int>(__ldg(sp + 3 * t + 0)) * c;
        a1 += static_cast<int>(__ldg(sp + 3 * t + 1)) * c;
        a2 += static_cast<int>(__ldg(sp + 3 * t + 2)) * c;
      }
      uint8_t* d = dst + static_cast<size_t>(r - row_lo) * kRowElems;
      d[0] = static_cast<uint8_t>(clip8_fixed(a0));
      d[1] = static_cast<uint8_t>(clip8_fixed(a1));
      d[2] = static_cast<uint8_t>(clip8_fixed(a2));
    }
    return;
  }
  int cf[kFastTaps];
#pragma unroll
  for (int t = 0; t < kFastTaps; ++t) cf[t] = t < n ? plan_h[2 * kCrop + static_cast<size_t>(x) * T + t] : 0;
#pragma unroll 4
  for (int rr = 0; rr < kHRowsPerCta; ++rr) {
    const int r = r0 + rr;
    if (r >= row_hi) break;
    const uint8_t* sp = src + static_cast<size_t>(r) * row_bytes;
    int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
#pragma unroll
    for (int t = 0; t < kFastTaps; ++t) {
      if (t < n) {  // taps past this column's own count would read past its window: skipped, not zero-weighted
        a0 += static_cast<int>(__ldg(sp + 3 * t + 0)) * cf[t];
        a1 += static_cast<int>(__ldg(sp + 3 * t + 1)) * cf[t];
        a2 += static_cast<int>(__ldg(sp + 3 * t + 2)) * cf[t];
      }
    }
    uint8_t* d = dst + static_cast<size_t>(r - row_lo) * kRowElems;
    d[0] = static_cast<uint8_t>(clip8_fixed(a0));
    d[1] = static_cast<uint8_t>(clip8_fixed(a1));
    d[2] = static_cast<uint8_t>(clip8_fixed(a2));
  }
}

template <int LAYOUT>
__global__ void __launch_bounds__(256) vpass_kernel(const int32_t* __restrict__ plan, int max_taps,
                                                    const int32_t* __restrict__ img_taps,
                                                    const uint8_t* __restrict__ inter,
                                                    const __nv_bfloat16* __restrict__ lut,
                                                    __nv_bfloat16* __restrict__ out) {
  const int img = blockIdx.y;
  const int T = max_taps;
  const int32_t* plan_v = plan + (static_cast<size_t>(img) * 2 + 1) * plan_ints_per_axis(T);
  if (!two_pass_image(plan_v, img_taps[img])) return;
  constexpr int kSide = LAYOUT == IRP_LAYOUT_NHWC4P ? kPad : kCrop;  // rows / columns this launch covers
  constexpr int kBorder = LAYOUT == IRP_LAYOUT_NHWC4P ? 3 : 0;
  const int px = threadIdx.x;          // (padded) output column
  if (px >= kSide) return;
  const int x = px - kBorder;
  const int row_lo = plan_v[0];
  const uint16_t* l16 = reinterpret_cast<const uint16_t*>(lut);
#pragma unroll 2
  for (int py = blockIdx.x * kVRowsPerCta; py < min(kSide, (static_cast<int>(blockIdx.x) + 1) * kVRowsPerCta); ++py) {
  const int y = py - kBorder;          // (padded) output row py
  const bool inside = y >= 0 && y < kCrop && x >= 0 && x < kCrop;
  uint32_t rgb[3] = {0u, 0u, 0u};  // bf16 bit patterns (uint8 values for IRP_LAYOUT_U8_HWC)
  if (inside) {
    const int vf = plan_v[y], vn = plan_v[kCrop + y];
    const int32_t* vc = plan_v + 2 * kCrop + static_cast<size_t>(y) * T;
    const uint8_t* ip = inter + (static_cast<size_t>(img) * kTwoPassMaxRows + (vf - row_lo)) * kRowElems + x * 3;
    int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
    for (int t = 0; t < vn; ++t) {
      const int c = vc[t];
      const uint8_t* rp = ip + static_cast<size_t>(t) * kRowElems;
      a0 += static_cast<int>(rp[0]) * c;
      a1 += static_cast<int>(rp[1]) * c;
      a2 += static_cast<int>(rp[2]) * c;
    }
    if (LAYOUT == IRP_LAYOUT_U8_HWC) {
      rgb[0] = clip8_fixed(a0);
      rgb[1] = clip8_fixed(a1);
      rgb[2] = clip8_fixed(a2);
    } else {
      rgb[0] = __ldg(l16 + clip8_fixed(a0));
      rgb[1] = __ldg(l16 + 256 + clip8_fixed(a1));
      rgb[2] = __ldg(l16 + 512 + clip8_fixed(a2));
    }
  }
  if (LAYOUT == IRP_LAYOUT_U8_HWC) {
    uint8_t* o8 = reinterpret_cast<uint8_t*>(out) + ((static_cast<size_t>(img) * kCrop + py) * kCrop + px) * 3;
    o8[0] = static_cast<uint8_t>(rgb[0]);
    o8[1] = static_cast<uint8_t>(rgb[1]);
    o8[2] = static_cast<uint8_t>(rgb[2]);
  } else if (LAYOUT == IRP_LAYOUT_NHWC4P) {
    uint2 o;
    o.x = rgb[0] | (rgb[1] << 16);
    o.y = rgb[2];
    reinterpret_cast<uint2*>(out)[(static_cast<size_t>(img) * kPad + py) * kPad + px] = o;
  } else {
    uint16_t* o16 = reinterpret_cast<uint16_t*>(out);
#pragma unroll
    for (int c = 0; c < 3; ++c)
      o16[((static_cast<size_t>(img) * 3 + c) * kCrop + py) * kCrop + px] = static_cast<uint16_t>(rgb[c]);
  }
  }
}

template <int TH>
static size_t fast_smem_bytes(int layout) {
  size_t b = (static_cast<size_t>(TH) * (2 + kFastTaps) + kFastRows) * 4 + 768 * 2;
  b += static_cast<size_t>(kFastRows) * kRowElems + kFastInBytes + 64;
  b += layout == IRP_LAYOUT_NHWC4P ? static_cast<size_t>(TH) * kPad * 4 * 2 : static_cast<size_t>(3) * TH * kCrop * 2;
  return b;
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

template <int LAYOUT, int TH>
static int launch_fast(const uint8_t* d_pixels, const int64_t* d_offsets, const int32_t* d_hw, int n_images,
                       int max_taps, const int32_t* plan, const int32_t* img_taps, __nv_bfloat16* out,
                       cudaStream_t st) {
  const size_t smem = fast_smem_bytes<TH>(LAYOUT);
  auto k = resample_fast_kernel<LAYOUT, TH>;
  static bool cfg = false;
  if (!cfg) {
    IRP_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    cfg = true;
  }
  dim3 grid(kCrop / TH, n_images);
  k<<<grid, kThreads, smem, st>>>(d_pixels, d_offsets, d_hw, max_taps, plan, img_taps, out);
  IRP_CUDA_OK(cudaGetLastError());
  return IRP_OK;
}

// output rows per CTA of the fast path (IRP_PRE_TH = 8 | 16 | 32)
static int fast_band_rows() {
  static int v = 0;
  if (v == 0) {
    const char* e = getenv("IRP_PRE_TH");
    v = e ? atoi(e) : 8;
    if (v != 8 && v != 16 && v != 32) v = 8;
  }
  return v;
}

// IRP_PRE_BANDS=1 selects the single-kernel band path (A/B runs)
static bool two_pass_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("IRP_PRE_BANDS");
    v = (e && atoi(e) != 0) ? 0 : 1;
  }
  return v == 1;
}

template <int LAYOUT>
static int launch_resample(const uint8_t* d_pixels, const int64_t* d_offsets, const int32_t* d_hw, int n_images,
                           int max_taps, const int32_t* plan, const int32_t* img_taps, __nv_bfloat16* out,
                           cudaStream_t st, uint8_t* inter, const __nv_bfloat16* lut) {
  if (two_pass_enabled() || LAYOUT == IRP_LAYOUT_U8_HWC) {
    dim3 hgrid((kTwoPassMaxRows + kHRowsPerCta - 1) / kHRowsPerCta, n_images);
    hpass_kernel<<<hgrid, kCrop, 0, st>>>(d_pixels, d_offsets, d_hw, max_taps, plan, img_taps, inter);
    constexpr int kSide = LAYOUT == IRP_LAYOUT_NHWC4P ? kPad : kCrop;
    dim3 vgrid((kSide + kVRowsPerCta - 1) / kVRowsPerCta, n_images);
    vpass_kernel<LAYOUT><<<vgrid, 256, 0, st>>>(plan, max_taps, img_taps, inter, lut, out);
    IRP_CUDA_OK(cudaGetLastError());
    return IRP_OK;
  }
  constexpr int BL = LAYOUT == IRP_LAYOUT_U8_HWC ? IRP_LAYOUT_NCHW : LAYOUT;  // band kernels: bf16 layouts only
  switch (fast_band_rows()) {
    default: IRP_TRY((launch_fast<BL, 8>(d_pixels, d_offsets, d_hw, n_images, max_taps, plan, img_taps, out, st))); break;
    case 32: IRP_TRY((launch_fast<BL, 32>(d_pixels, d_offsets, d_hw, n_images, max_taps, plan, img_taps, out, st))); break;
    case 16: IRP_TRY((launch_fast<BL, 16>(d_pixels, d_offsets, d_hw, n_images, max_taps, plan, img_taps, out, st))); break;
  }
  return IRP_OK;
}

// The rare images that need more than kFastTaps taps (downscale > 2.5x) run through the generic kernel: a few
// long CTAs per image.  After the fast path they would be a pure tail (112 us for one 1200x1200 image against
// 158 us for the other 255 images of a batch), so they run CONCURRENTLY on a side stream: fork after the plan
// kernel, join after the fast path.  The two kernels write disjoint images.
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
    // highest priority: its few long CTAs must get SM slots while the fast path's thousands of CTAs are queued
    int least = 0, greatest = 0;
    IRP_CUDA_OK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    IRP_CUDA_OK(cudaStreamCreateWithPriority(&s.stream, cudaStreamNonBlocking, greatest));
    IRP_CUDA_OK(cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming));
    IRP_CUDA_OK(cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming));
  }
  *out = &s;
  return IRP_OK;
}

template <int LAYOUT>
static int launch_generic(const uint8_t* d_pixels, const int64_t* d_offsets, const int32_t* d_hw, int n_images,
                          int max_taps, const int32_t* plan, const int32_t* img_taps, __nv_bfloat16* out,
                          cudaStream_t st) {
  const size_t smem = resample_smem_bytes(max_taps, LAYOUT);
  IRP_REQUIRE(smem <= 227 * 1024, "preprocess: max_taps %d needs %zu bytes of shared memory", max_taps, smem);
  auto k = resample_kernel<LAYOUT>;
  static size_t cfg = 0;
  if (smem > 48 * 1024 && smem > cfg) {
    IRP_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    cfg = smem;
  }
  dim3 grid(kCrop / kBandRows, n_images);
  k<<<grid, kThreads, smem, st>>>(d_pixels, d_offsets, d_hw, max_taps, plan, img_taps, out,
                                  (two_pass_enabled() || LAYOUT == IRP_LAYOUT_U8_HWC) ? 1 : 0);
  IRP_CUDA_OK(cudaGetLastError());
  return IRP_OK;
}

template <int LAYOUT>
static int launch_both(const uint8_t* d_pixels, const int64_t* d_offsets, const int32_t* d_hw, int n_images,
                       int max_taps, const int32_t* plan, const int32_t* img_taps, __nv_bfloat16* out,
                       cudaStream_t st, uint8_t* inter, const __nv_bfloat16* lut) {
  if (max_taps <= kFastTaps)  // no image can need the generic path
    return launch_resample<LAYOUT>(d_pixels, d_offsets, d_hw, n_images, max_taps, plan, img_taps, out, st, inter, lut);
  SideStream* side = nullptr;
  IRP_TRY(side_stream(&side));
  IRP_CUDA_OK(cudaEventRecord(side->fork, st));  // the plan kernel (and the caller's earlier work) precede both
  IRP_CUDA_OK(cudaStreamWaitEvent(side->stream, side->fork, 0));
  IRP_TRY(launch_generic<LAYOUT>(d_pixels, d_offsets, d_hw, n_images, max_taps, plan, img_taps, out, side->stream));
  IRP_CUDA_OK(cudaEventRecord(side->join, side->stream));
  IRP_TRY(launch_resample<LAYOUT>(d_pixels, d_offsets, d_hw, n_images, max_taps, plan, img_taps, out, st, inter, lut));
  IRP_CUDA_OK(cudaStreamWaitEvent(st, side->join, 0));
  return IRP_OK;
}

extern "C" {

size_t irp_preprocess_workspace_bytes(int n_images, int max_taps) {
  if (n_images <= 0 || max_taps <= 0) return 0;
  size_t b = static_cast<size_t>(n_images) * 2 * plan_ints_per_axis(max_taps) * sizeof(int32_t) + 16 +
             static_cast<size_t>(n_images) * sizeof(int32_t);
  b = (b + 255) & ~static_cast<size_t>(255);
  b += 2048;                                                              // normalisation LUT (768 bf16)
  b += static_cast<size_t>(n_images) * kTwoPassMaxRows * kRowElems;       // horizontally filtered rows (uint8)
  return b;
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
                  transform == IRP_TRANSFORM_WDS_LANCZOS,
              "preprocess: unknown transform %d", transform);
  IRP_REQUIRE(n_images > 0, "preprocess: n_images %d", n_images);
  IRP_REQUIRE(max_taps >= 3 && max_taps <= 513, "preprocess: max_taps %d out of range", max_taps);
  IRP_REQUIRE(out_layout == IRP_LAYOUT_NCHW || out_layout == IRP_LAYOUT_NHWC4P || out_layout == IRP_LAYOUT_U8_HWC,
              "preprocess: bad layout %d", out_layout);
  IRP_REQUIRE(workspace_bytes >= irp_preprocess_workspace_bytes(n_images, max_taps),
              "preprocess: workspace %zu < %zu bytes", workspace_bytes,
              irp_preprocess_workspace_bytes(n_images, max_taps));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // workspace: plan tables | status word (16 bytes) | per-image max tap count
  int32_t* plan = static_cast<int32_t*>(d_workspace);
  const size_t plan_bytes = static_cast<size_t>(n_images) * 2 * plan_ints_per_axis(max_taps) * sizeof(int32_t);
  int32_t* status = reinterpret_cast<int32_t*>(static_cast<uint8_t*>(d_workspace) + plan_bytes);
  int32_t* img_taps = status + 4;
  size_t head = plan_bytes + 16 + static_cast<size_t>(n_images) * sizeof(int32_t);
  head = (head + 255) & ~static_cast<size_t>(255);
  __nv_bfloat16* lut = reinterpret_cast<__nv_bfloat16*>(static_cast<uint8_t*>(d_workspace) + head);
  uint8_t* inter = static_cast<uint8_t*>(d_workspace) + head + 2048;
  IRP_CUDA_OK(cudaMemsetAsync(status, 0, 16 + static_cast<size_t>(n_images) * sizeof(int32_t), st));
  resample_plan_kernel<<<n_images, 2 * kCrop, 0, st>>>(d_hw, n_images, max_taps, transform, plan, status, img_taps,
                                                       lut);
  IRP_CUDA_OK(cudaGetLastError());
  if (out_layout == IRP_LAYOUT_U8_HWC)
    return launch_both<IRP_LAYOUT_U8_HWC>(d_pixels, d_offsets, d_hw, n_images, max_taps, plan, img_taps,
                                          static_cast<__nv_bfloat16*>(d_out), st, inter, lut);
  if (out_layout == IRP_LAYOUT_NHWC4P)
    return launch_both<IRP_LAYOUT_NHWC4P>(d_pixels, d_offsets, d_hw, n_images, max_taps, plan, img_taps,
                                          static_cast<__nv_bfloat16*>(d_out), st, inter, lut);
  return launch_both<IRP_LAYOUT_NCHW>(d_pixels, d_offsets, d_hw, n_images, max_taps, plan, img_taps,
                                      static_cast<__nv_bfloat16*>(d_out), st, inter, lut);
}

/* Host-side view of the resize/crop geometry (used by the Python mirror to size max_taps and by tests). */
int irp_preprocess_geometry(int h, int w, int* out_h, int* out_w, int* top, int* left, int* taps) {
  return irp_preprocess_geometry_ex(h, w, IRP_TRANSFORM_WEIGHTS_DEFAULT, out_h, out_w, top, left, taps);
}

int irp_preprocess_geometry_ex(int h, int w, int transform, int* out_h, int* out_w, int* top, int* left, int* taps) {
  IRP_REQUIRE(h > 0 && w > 0, "geometry: bad size %dx%d", h, w);
  IRP_REQUIRE(transform == IRP_TRANSFORM_WEIGHTS_DEFAULT || transform == IRP_TRANSFORM_VAL_256 ||
                  transform == IRP_TRANSFORM_WDS_LANCZOS,
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
    int c = static_cast<int>(s);
    if (static_cast<double>(c) < s) ++c;
    *taps = 2 * c + 1;
  }
  return IRP_OK;
}

}  // extern "C"
