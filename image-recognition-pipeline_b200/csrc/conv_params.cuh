// Implicit-GEMM convolution for the ResNet-50 trunk on tcgen05 / TMEM, operands fed by TMA.
//
//   out[n,ho,wo,co] = act( sum_{r,s,ci} x[n, ho*stride+r-pad, wo*stride+s-pad, ci] * w[co,r,s,ci]
//                          + bias[co] (+ residual[n,ho,wo,co]) )
//
// GEMM view (torchvision/models/resnet.py:108-160 Bottleneck, BN folded into w/bias):
//   M = output pixels, N = Cout, K = taps * Cin.  Activations are NHWC bf16, weights [Cout][tap][Cin] bf16.
//   One CTA tile = 128 output pixels (a (bw,bh,bn) box in (wo,ho,n)) x BN output channels.  For each filter
//   tap the A operand is ONE shifted TMA box of the input (out-of-bounds rows/cols are zero-filled by TMA,
//   which is exactly the conv's zero padding); stride-2 convs read one of four "parity" views of the input,
//   each a plain strided tensor map.  Accumulators live in TMEM (double buffered) so the epilogue of tile i
//   overlaps the main loop of tile i+1.
//
// Epilogue: TMEM -> registers -> (+bias, +residual, ReLU, bf16) -> swizzled shared-memory staging tile -> TMA
// store with the same box geometry as the A loads (TMA clips rows outside the tensor, so ragged tiles need no
// predication).  The residual tile is prefetched by TMA into the staging buffer it will be overwritten in.
//
// This header holds the parameter block shared by the trunk's convolution kernels (conv_gemm2.cuh: CTA-pair
// implicit GEMM, conv3x3_c64.cuh: patch-resident 3x3).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdint>

#include "ptx.cuh"

namespace irp {

constexpr int kMaxTaps = 9;
constexpr int kTileM = 128;
constexpr int kStgChunkBytes = kTileM * 128;  // one 64-channel column chunk of the output tile (128 B rows)

struct alignas(64) ConvParams {
  CUtensorMap tmA[4];  // activation views (index = parity for stride 2; [0] only for stride 1 / stem)
  CUtensorMap tmB;     // weights [Cout][K] bf16, K-major
  CUtensorMap tmOut;   // output  (Cout, Wo, Ho, B) / flat (Cout, M, 1, 1), box (64, bw, bh, bn)
  CUtensorMap tmRes;   // residual, same geometry as tmOut (valid only for the RES instances)
  CUtensorMap tmOutW;  // flat (1x1, stride 1) convs only: the output with a (64, 32) box -- one epilogue warp's 32 rows
  // M tiling: box (bw,bh,bn) in output coordinates
  int bw, bh, bn;
  int tiles_w, tiles_h, tiles_n;
  int n_tiles_n;  // Cout / BN
  int num_tiles;
  int a_box_bytes;    // bytes one A TMA box delivers (rows_in_box * BK * 2)
  int out_box_bytes;  // bytes one 64-channel output/residual box moves (rows_in_box * 128)
  // problem
  int ntaps, kc_blocks;  // K loop = ntaps * kc_blocks blocks of BK
  int cin;               // channels per tap in the weight matrix (K offset of tap t = t*cin)
  int n_base;            // stem only: first image of this micro-batch inside the input tensor map
  int8_t tap_map[kMaxTaps], tap_dw[kMaxTaps], tap_dh[kMaxTaps];
  // epilogue
  const float* bias;  // [Cout]
  int cout;           // total output channels (length of bias)
  int relu;
};

}  // namespace irp
