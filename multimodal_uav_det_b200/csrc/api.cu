// Library-level plumbing of the C-ABI: error string, launch counter, device watchdog word.
#include "common.cuh"
#include <string.h>
#include <stdlib.h>

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

int pdl_mask() {
  // off by default: measured on the BaselineModel step (B200, same box, 20 steps each): no attribute 26.63 / 26.69 ms,
  // implicit GEMM only 26.61, BatchNorm forward only 26.82, BatchNorm backward (reduce + apply) 27.43, all four 27.85
  static const int mask = getenv("UAVDET_PDL_MASK") ? atoi(getenv("UAVDET_PDL_MASK")) : 0;
  return mask;
}

static std::atomic<int> g_sm_margin{0};
static std::atomic<int> g_sm_margin_launches{-1};    // < 0: until reset; >= 0: persistent-kernel launches it still covers
int sm_budget() {
  const int m = g_sm_margin.load(std::memory_order_relaxed);
  if (m == 0) return kNumSMs;
  const int left = g_sm_margin_launches.load(std::memory_order_relaxed);
  if (left == 0) return kNumSMs;
  if (left > 0) g_sm_margin_launches.fetch_sub(1, std::memory_order_relaxed);
  const int b = kNumSMs - m;
  return b < 1 ? 1 : b;
}

__global__ void stamp_kernel(unsigned long long* slot) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  *slot = t;
}

}  // namespace uavdet

extern "C" int uavdet_timestamp(unsigned long long* slot_dev, void* stream) {
  UAVDET_CHECK_ARG(slot_dev, "timestamp: null slot");
  uavdet::stamp_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(slot_dev);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { uavdet::set_error("timestamp launch failed: %s", cudaGetErrorString(e)); return UAVDET_ERR_CUDA; }
  return UAVDET_OK;   // not counted in uavdet_launch_count: a measurement aid, not part of the path
}

extern "C" int uavdet_set_sm_margin(int margin, int launches) {
  if (margin < 0) margin = 0;
  if (margin > uavdet::kNumSMs - 1) margin = uavdet::kNumSMs - 1;
  uavdet::g_sm_margin_launches.store(launches < 0 ? -1 : launches);
  return uavdet::g_sm_margin.exchange(margin);
}

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
