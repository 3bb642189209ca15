// api.cu — error plumbing, version / device checks, launch accounting of libmcd_sm100.
#include "common.cuh"
#include <atomic>

namespace mcd {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace mcd

extern "C" {

const char* mcd_last_error(void) { return mcd::g_err; }

int mcd_version(void) { return MCD_ABI_VERSION; }

int64_t mcd_launch_count(void) { return mcd::g_launches.load(std::memory_order_relaxed); }

int mcd_check_device(int device) {
  cudaDeviceProp p;
  cudaError_t e = cudaGetDeviceProperties(&p, device);
  if (e != cudaSuccess) {
    mcd::set_error("cudaGetDeviceProperties(%d): %s", device, cudaGetErrorString(e));
    return MCD_E_CUDA;
  }
  if (p.major != 10) {
    mcd::set_error("device %d is sm_%d%d; libmcd_sm100 is built for sm_100a only", device, p.major,
                   p.minor);
    return MCD_E_ARCH;
  }
  return MCD_OK;
}

}  // extern "C"
