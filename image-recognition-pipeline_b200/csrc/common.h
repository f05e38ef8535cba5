// Host-side plumbing shared by every translation unit of libirp_b200.so:
// status codes, the thread-local last-error string, CUDA error checking that never throws across the
// C ABI, and lazy lookup of the driver's tensor-map encoder (no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/irp_b200.h"

namespace irp {

void set_last_error(const char* fmt, ...);

#define IRP_CUDA_OK(expr)                                                                       \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      ::irp::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return IRP_ERR_CUDA;                                                                      \
    }                                                                                           \
  } while (0)

#define IRP_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      ::irp::set_last_error(__VA_ARGS__); \
      return IRP_ERR_INVALID;             \
    }                                     \
  } while (0)

#define IRP_TRY(expr)            \
  do {                           \
    int _s = (expr);             \
    if (_s != IRP_OK) return _s; \
  } while (0)

// cuTensorMapEncodeTiled resolved through cudaGetDriverEntryPoint; nullptr (+ last error) if unavailable.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_tiled();

// Encode a bf16 tiled tensor map. dims/strides/box are innermost-first; strides are in BYTES for dims 1..rank-1.
int encode_bf16_map(CUtensorMap* map, void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, int swizzle_bytes);

// SM count of the CURRENT device (cached per device).
int num_sms();

// Opt the kernel `func` in to `bytes` of dynamic shared memory on the CURRENT device (no-op up to 32 KB; cached per
// (device, kernel), so a second GPU in the same process gets its own opt-in).
int ensure_smem(const void* func, size_t bytes);
template <typename K>
inline int ensure_smem(K* kernel, size_t bytes) {
  return ensure_smem(reinterpret_cast<const void*>(kernel), bytes);
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace irp
