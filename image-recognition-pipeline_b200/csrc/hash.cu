// N3: the duplicate-detection hash of the dataset-cleaning pass,
// functions/data_curation.py:283-292 compute_image_hash = md5(img.resize((64, 64)).convert("RGB").tobytes()).
// The resize is irp_preprocess_ex(IRP_TRANSFORM_HASH_64) (bit-exact Pillow bicubic); this file is the digest:
// RFC 1321 MD5 of every row of a [n, row_bytes] uint8 matrix, one thread per row (the 64-byte blocks of one message
// are a serial chain, the rows are independent).  oracle: hashlib.md5 (tests/test_oracle.py, tests/golden/hash.npz).
#include <cstdint>

#include "common.h"

namespace irp {

__constant__ uint32_t kMd5K[64] = {
    0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501,
    0x698098d8, 0x8b44f7af, 0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821,
    0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8,
    0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a,
    0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70,
    0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665,
    0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1,
    0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};

__device__ __forceinline__ uint32_t rotl32(uint32_t x, int s) { return __funnelshift_l(x, x, s); }

__device__ __forceinline__ void md5_block(uint32_t (&st)[4], const uint32_t (&m)[16]) {
  constexpr int kS[4][4] = {{7, 12, 17, 22}, {5, 9, 14, 20}, {4, 11, 16, 23}, {6, 10, 15, 21}};
  uint32_t a = st[0], b = st[1], c = st[2], d = st[3];
#pragma unroll
  for (int i = 0; i < 64; ++i) {
    uint32_t f;
    int g;
    if (i < 16) {
      f = (b & c) | (~b & d);
      g = i;
    } else if (i < 32) {
      f = (d & b) | (~d & c);
      g = (5 * i + 1) & 15;
    } else if (i < 48) {
      f = b ^ c ^ d;
      g = (3 * i + 5) & 15;
    } else {
      f = c ^ (b | ~d);
      g = (7 * i) & 15;
    }
    const uint32_t t = d;
    d = c;
    c = b;
    b = b + rotl32(a + f + kMd5K[i] + m[g], kS[i >> 4][i & 3]);
    a = t;
  }
  st[0] += a;
  st[1] += b;
  st[2] += c;
  st[3] += d;
}

// message byte i of row `row` (zero past the end): reads stay inside the row
__global__ void __launch_bounds__(128) md5_rows_kernel(const uint8_t* __restrict__ data, int n_rows,
                                                       long long row_bytes, uint8_t* __restrict__ digest) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_rows) return;
  const uint8_t* p = data + static_cast<size_t>(row) * row_bytes;
  uint32_t st[4] = {0x67452301u, 0xefcdab89u, 0x98badcfeu, 0x10325476u};
  const long long full = row_bytes / 64;
  const bool aligned = (reinterpret_cast<uintptr_t>(p) & 3u) == 0;
  for (long long blk = 0; blk < full; ++blk) {
    uint32_t m[16];
    if (aligned) {
      const uint32_t* w = reinterpret_cast<const uint32_t*>(p + blk * 64);
#pragma unroll
      for (int i = 0; i < 16; ++i) m[i] = w[i];  // little-endian words, as MD5 wants them
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const uint8_t* q = p + blk * 64 + 4 * i;
        m[i] = static_cast<uint32_t>(q[0]) | (static_cast<uint32_t>(q[1]) << 8) | (static_cast<uint32_t>(q[2]) << 16) |
               (static_cast<uint32_t>(q[3]) << 24);
      }
    }
    md5_block(st, m);
  }
  // tail: remaining bytes, the 0x80 marker, zero fill, the message length in bits (one or two blocks)
  const int rem = static_cast<int>(row_bytes - full * 64);
  const unsigned long long bits = static_cast<unsigned long long>(row_bytes) * 8ull;
  for (int pass = 0; pass < 2; ++pass) {
    uint32_t m[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      uint32_t w = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int idx = pass * 64 + 4 * i + j;  // byte index inside the tail
        uint32_t byte = 0;
        if (idx < rem) byte = p[full * 64 + idx];
        else if (idx == rem) byte = 0x80u;
        w |= byte << (8 * j);
      }
      m[i] = w;
    }
    const bool last = pass == 1 || rem < 56;
    if (last) {
      m[14] = static_cast<uint32_t>(bits);
      m[15] = static_cast<uint32_t>(bits >> 32);
    }
    md5_block(st, m);
    if (last) break;
  }
  uint32_t* out = reinterpret_cast<uint32_t*>(digest + static_cast<size_t>(row) * 16);
#pragma unroll
  for (int i = 0; i < 4; ++i) out[i] = st[i];  // digest bytes = the four state words, little-endian
}

}  // namespace irp

using namespace irp;

extern "C" int irp_md5_rows(const uint8_t* d_data, int n_rows, int64_t row_bytes, uint8_t* d_digest, void* stream) {
  IRP_REQUIRE((d_data || row_bytes == 0) && d_digest, "md5_rows: null argument");
  IRP_REQUIRE(n_rows > 0 && row_bytes >= 0, "md5_rows: n_rows %d, row_bytes %lld", n_rows,
              static_cast<long long>(row_bytes));
  IRP_REQUIRE((reinterpret_cast<uintptr_t>(d_digest) & 3u) == 0, "md5_rows: d_digest must be 4-byte aligned");
  md5_rows_kernel<<<(n_rows + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(d_data, n_rows, row_bytes,
                                                                                      d_digest);
  IRP_CUDA_OK(cudaGetLastError());
  return IRP_OK;
}
