// Error plumbing and device queries shared by every translation unit of libsurgvid.
#include <map>
#include <mutex>
#include <utility>

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

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: remember which (device, kernel, bytes) were configured and
// never cache a failure, so a handle created on a second GPU of the same process opts its kernels in again.
int ensure_dynamic_smem(const void* fn, int bytes) {
  if (bytes <= 48 * 1024) return SV_OK;
  int dev = 0;
  SV_CUDA_OK(cudaGetDevice(&dev));
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, int> configured;
  std::lock_guard<std::mutex> lock(mu);
  auto key = std::make_pair(dev, fn);
  auto it = configured.find(key);
  if (it != configured.end() && it->second >= bytes) return SV_OK;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return fail(SV_ERR_CUDA, std::string("cudaFuncSetAttribute(MaxDynamicSharedMemorySize): ") + cudaGetErrorString(e));
  configured[key] = bytes;
  return SV_OK;
}

}  // namespace sv

extern "C" {
const char* sv_last_error(void) { return sv::g_last_error.c_str(); }
int sv_abi_version(void) { return SV_ABI_VERSION; }
}
