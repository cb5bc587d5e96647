// Library-level plumbing of the C-ABI: error string, launch counter, device watchdog word.
#include "common.cuh"
#include <string.h>

namespace uavdet {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// One word per device context; kernels OR a code into it when a bounded pipeline wait
// expires (blind-development safety net: a wrong barrier must never hang the GPU box).
__device__ unsigned int g_watchdog = 0;

unsigned int* watchdog_word() {
  unsigned int* p = nullptr;
  cudaGetSymbolAddress((void**)&p, g_watchdog);
  return p;
}

}  // namespace uavdet

extern "C" const char* uavdet_last_error(void) { return uavdet::g_err; }
extern "C" int uavdet_version(void) { return 100; }
extern "C" uint64_t uavdet_launch_count(void) { return uavdet::g_launches.load(); }

extern "C" int uavdet_check_device(void* stream, int* flag_host) {
  unsigned int v = 0;
  UAVDET_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  UAVDET_CUDA(cudaMemcpyFromSymbol(&v, uavdet::g_watchdog, sizeof(v)));
  if (v) {
    unsigned int z = 0;
    UAVDET_CUDA(cudaMemcpyToSymbol(uavdet::g_watchdog, &z, sizeof(z)));
  }
  if (flag_host) *flag_host = (int)v;
  if (v) {
    uavdet::set_error("device watchdog tripped: code 0x%x", v);
    return UAVDET_ERR_DEVICE;
  }
  return UAVDET_OK;
}
