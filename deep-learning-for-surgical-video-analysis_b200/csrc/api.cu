// Error plumbing and device queries shared by every translation unit of libsurgvid.
#include <mutex>

#include "common.cuh"

namespace sv {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

int device_sm_count() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  static std::mutex mu;
  static int cached_dev = -1, cached = 0;
  std::lock_guard<std::mutex> lock(mu);
  if (cached_dev == dev) return cached;
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  cached_dev = dev;
  cached = n;
  return n;
}

}  // namespace sv

extern "C" {
const char* sv_last_error(void) { return sv::g_last_error.c_str(); }
int sv_abi_version(void) { return SV_ABI_VERSION; }
}
