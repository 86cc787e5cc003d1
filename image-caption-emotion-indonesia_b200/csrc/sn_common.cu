#include <stdlib.h>

#include "sn_common.cuh"

#include <cuda.h>

#include <mutex>
#include <string.h>

namespace sn {

char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int32_t fail(int32_t code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

int32_t check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail((int32_t)e, "%s: %s", what, cudaGetErrorString(e));
  return 0;
}

static DevInfo g_info[64];
static std::once_flag g_once[64];

const DevInfo& dev_info() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  std::call_once(g_once[dev], [dev]() {
    DevInfo& d = g_info[dev];
    cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&d.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&d.cc_minor, cudaDevAttrComputeCapabilityMinor, dev);
  });
  return g_info[dev];
}

}  // namespace sn

namespace sn {
bool recur_cooperative() {
  static const bool coop = [] {
    const char* e = getenv("SN_RECUR_COOP");
    return e && e[0] == '1';
  }();
  return coop;
}
}  // namespace sn

extern "C" {

int32_t sn_version(void) { return SN_VERSION; }
const char* sn_last_error(void) { return sn::err_buf(); }

int32_t sn_device_info(int32_t* sm_count, int32_t* smem_optin, int32_t* cc_major, int32_t* cc_minor) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) return sn::fail(e ? (int32_t)e : -2, "no CUDA device (libsn100 has no CPU fallback)");
  const sn::DevInfo& d = sn::dev_info();
  if (sm_count) *sm_count = d.sm_count;
  if (smem_optin) *smem_optin = d.smem_optin;
  if (cc_major) *cc_major = d.cc_major;
  if (cc_minor) *cc_minor = d.cc_minor;
  if (d.cc_major != 10) return sn::fail(-3, "device is sm_%d%d; libsn100 is built for sm_100a only", d.cc_major, d.cc_minor);
  return 0;
}

// ---- CUDA IPC helpers for the peer-memory data-parallel exchange (sn_dp_adam_fused) ----------------------
// The exporting rank describes a device pointer as (64-byte handle of its cudaMalloc allocation, byte offset);
// the importing rank opens it with ITS device current, so the mapping is a peer mapping usable by its kernels.
int32_t sn_ipc_export(const void* ptr, uint8_t* handle64, int64_t* offset) {
  SN_REQUIRE(ptr && handle64 && offset, "sn_ipc_export: null argument");
  CUdeviceptr base = 0;
  size_t size = 0;
  // resolved at run time: the library must load (and export its symbols) on a box without libcuda.so.1
  typedef CUresult (*RangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
  static RangeFn range_fn = nullptr;
  if (!range_fn) {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fp, cudaEnableDefault, &q) != cudaSuccess || !fp)
      return sn::fail(-6, "cuMemGetAddressRange entry point not available");
    range_fn = (RangeFn)fp;
  }
  CUresult r = range_fn(&base, &size, (CUdeviceptr)ptr);
  if (r != CUDA_SUCCESS) return sn::fail(-6, "cuMemGetAddressRange failed (%d)", (int)r);
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, (void*)base);
  if (e != cudaSuccess) return sn::fail((int32_t)e, "cudaIpcGetMemHandle: %s (expandable segments are not IPC-exportable)", cudaGetErrorString(e));
  memcpy(handle64, &h, sizeof(h));
  *offset = (int64_t)((CUdeviceptr)ptr - base);
  return 0;
}

int32_t sn_ipc_open(const uint8_t* handle64, void** base_out) {
  SN_REQUIRE(handle64 && base_out, "sn_ipc_open: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  cudaError_t e = cudaIpcOpenMemHandle(base_out, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return sn::fail((int32_t)e, "cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
  return 0;
}

int32_t sn_ipc_close(void* base) {
  cudaError_t e = cudaIpcCloseMemHandle(base);
  if (e != cudaSuccess) return sn::fail((int32_t)e, "cudaIpcCloseMemHandle: %s", cudaGetErrorString(e));
  return 0;
}

}  // extern "C"
