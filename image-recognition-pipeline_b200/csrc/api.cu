// Library-wide plumbing: last-error string, device check, tensor-map encoder lookup.
#include <cstring>
#include <map>
#include <mutex>
#include <utility>

#include "common.h"

namespace irp {

static thread_local char g_last_error[1024] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  if (!fn) set_last_error("cuTensorMapEncodeTiled is not available from the installed driver");
  return fn;
}

int encode_bf16_map(CUtensorMap* map, void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, int swizzle_bytes) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return IRP_ERR_CUDA;
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), base, gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error(
        "cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu %llu %llu %llu %llu] box [%u %u %u %u %u] "
        "stride0 %llu swizzle %d base %p",
        static_cast<int>(r), rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
        (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
        (unsigned long long)(rank > 4 ? dims[4] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0,
        rank > 3 ? box[3] : 0, rank > 4 ? box[4] : 0, (unsigned long long)(rank > 1 ? strides_bytes[0] : 0),
        swizzle_bytes, base);
    return IRP_ERR_CUDA;
  }
  return IRP_OK;
}

constexpr int kMaxDevices = 64;

int num_sms() {
  static int n[kMaxDevices] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 148;
  if (n[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    n[dev] = v;
  }
  return n[dev];
}

int ensure_smem(const void* func, size_t bytes) {
  // kernel attributes are per device: remember the limit already granted per (device, kernel)
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, size_t> granted;
  if (bytes <= 32 * 1024) return IRP_OK;  // static shared memory counts against the 48 KB default limit too
  int dev = 0;
  IRP_CUDA_OK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  size_t& g = granted[std::make_pair(dev, func)];
  if (bytes > g) {
    IRP_CUDA_OK(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
    g = bytes;
  }
  return IRP_OK;
}

}  // namespace irp

extern "C" {

int irp_abi_version(void) { return IRP_B200_ABI_VERSION; }

#ifndef IRP_BUILD_ID
#define IRP_BUILD_ID "unknown"
#endif
const char* irp_build_id(void) { return IRP_BUILD_ID; }

const char* irp_last_error(void) { return irp::g_last_error; }

int irp_init(int device) {
  // validates `device` without changing the caller's current device
  int major = 0, minor = 0;
  IRP_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  IRP_CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
  if (major != 10) {
    irp::set_last_error("device %d is sm_%d%d; libirp_b200 carries sm_100a code only", device, major, minor);
    return IRP_ERR_DEVICE;
  }
  if (!irp::get_encode_tiled()) return IRP_ERR_CUDA;
  return IRP_OK;
}

}  // extern "C"
