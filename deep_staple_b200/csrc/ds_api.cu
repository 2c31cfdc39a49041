// Library-level entry points of include/deep_staple_b200.h: version, error string, device check,
// launch accounting.  No kernels here.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "ds_common.cuh"

namespace ds {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();  // clear the (non-sticky) launch error so the next call starts clean
    return fail((int)e, "%s: %s", what, cudaGetErrorString(e));
  }
  return 0;
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int sm_count() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
  return n;
}

int env_int(const char* name, int fallback) {
  const char* s = getenv(name);
  if (!s || !*s) return fallback;
  return atoi(s);
}

}  // namespace ds

extern "C" {

int ds_version(void) { return DS_VERSION; }

const char* ds_last_error_string(void) { return ds::g_err; }

int ds_device_check(void) {
  int dev = 0, major = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return ds::fail(DS_ERR_DEVICE, "no usable CUDA device (%s); this library has no CPU fallback",
                    cudaGetErrorString(e));
  }
  if (major != 10)
    return ds::fail(DS_ERR_DEVICE, "device %d is compute capability %d.x; kernels are built for sm_100a only", dev,
                    major);
  return DS_OK;
}

uint64_t ds_launch_count(void) { return ds::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
