#include "sn_common.cuh"

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

}  // extern "C"
