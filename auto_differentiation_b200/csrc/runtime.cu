// Library-wide state: last-error string (thread local), launch counter, device probe.
#include "common.cuh"
#include <atomic>
#include <string.h>

namespace b200ad {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count() {
  static const int n = [] {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        v <= 0) {
      cudaGetLastError();
      return B200AD_NUM_SMS;
    }
    return v;
  }();
  return n;
}

}  // namespace b200ad

extern "C" {

int b200ad_abi_version(void) { return B200AD_ABI_VERSION; }

const char* b200ad_last_error(void) { return b200ad::g_err; }

int64_t b200ad_launch_count(void) { return b200ad::g_launches.load(); }

int b200ad_device_info(char* name, int cap, int* cc_major, int* cc_minor) {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return b200ad::set_error(B200AD_ERR_CUDA, "no CUDA device: %s",
                             e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceProp p;
  e = cudaGetDeviceProperties(&p, dev);
  if (e != cudaSuccess) return b200ad::set_error(B200AD_ERR_CUDA, "%s", cudaGetErrorString(e));
  if (name && cap > 0) {
    strncpy(name, p.name, cap - 1);
    name[cap - 1] = 0;
  }
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return p.multiProcessorCount;
}
}
