// Host-side plumbing of the C ABI: error reporting, device checks, TMA descriptor encoding.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>

#include <stdlib.h>

#include "common.cuh"
#include "../../include/l4p_b200.h"

namespace l4p {

static thread_local char g_err[512] = "";

int host_set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int host_check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return L4P_OK;
  return host_set_error(L4P_ERR_CUDA, "CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;
static std::once_flag g_encode_once;

static void resolve_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) g_encode = (PFN_encodeTiled)fn;
}

int host_make_tmap_16b(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                       const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  std::call_once(g_encode_once, resolve_encode);
  if (!g_encode) return host_set_error(L4P_ERR_DRIVER, "cuTensorMapEncodeTiled not available");
  if (rank < 1 || rank > 5) return host_set_error(L4P_ERR_ARG, "tensor map rank %d", rank);
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0)
    return host_set_error(L4P_ERR_ARG, "TMA base %p not 16-byte aligned", base);
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (box[i] == 0 || box[i] > 256) return host_set_error(L4P_ERR_ARG, "TMA box[%d]=%u", i, box[i]);
  }
  for (int i = 0; i + 1 < rank; ++i) {
    gstr[i] = strides_bytes[i];
    if (gstr[i] % 16 != 0)
      return host_set_error(L4P_ERR_ARG, "TMA stride[%d]=%llu not multiple of 16", i,
                            (unsigned long long)gstr[i]);
  }
  CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_NONE;
  if (swizzle_bytes == 32) sw = CU_TENSOR_MAP_SWIZZLE_32B;
  else if (swizzle_bytes == 64) sw = CU_TENSOR_MAP_SWIZZLE_64B;
  else if (swizzle_bytes == 128) sw = CU_TENSOR_MAP_SWIZZLE_128B;
  else if (swizzle_bytes != 0) return host_set_error(L4P_ERR_ARG, "swizzle %d", swizzle_bytes);
  if (swizzle_bytes && box[0] * 2u > (uint32_t)swizzle_bytes)
    return host_set_error(L4P_ERR_ARG, "TMA inner box %u B exceeds swizzle span %d", box[0] * 2u, swizzle_bytes);
  // BFLOAT16 and FLOAT16 are both opaque 2-byte moves for tiled loads; zero OOB fill.
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim,
                        gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return host_set_error(L4P_ERR_DRIVER, "cuTensorMapEncodeTiled failed: %d", (int)r);
  return L4P_OK;
}

static int g_num_sms = 0;
int host_num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

bool host_pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("L4P_NO_PDL");
    on = (e != nullptr && e[0] == '1') ? 0 : 1;
  }
  return on != 0;
}

}  // namespace l4p

extern "C" {

const char* l4p_last_error(void) { return l4p::g_err; }

int l4p_version(void) { return 100; }

int l4p_init(int device, int arch_check) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0)
    return l4p::host_set_error(L4P_ERR_CUDA, "no CUDA device available (%s)", cudaGetErrorString(e));
  if (device < 0 || device >= n) return l4p::host_set_error(L4P_ERR_ARG, "device %d out of range (%d)", device, n);
  L4P_CHECK_CUDA(cudaSetDevice(device));
  if (arch_check) {
    int major = 0, minor = 0;
    L4P_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    L4P_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
    if (major != 10)
      return l4p::host_set_error(L4P_ERR_ARCH, "device %d is sm_%d%d; this library is built for sm_100a only",
                                 device, major, minor);
  }
  l4p::g_num_sms = 0;
  (void)l4p::host_num_sms();
  return L4P_OK;
}

}  // extern "C"
