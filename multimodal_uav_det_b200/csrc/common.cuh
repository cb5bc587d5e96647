// Shared host/device helpers for libuavdet_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>
#include <mutex>

#include "../../include/uavdet_b200.h"

namespace uavdet {

// ---- error plumbing -------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;
inline void count_launch(uint64_t n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }
// device-side watchdog word (igemm/wgrad pipeline waits) — defined in api.cu
unsigned int* watchdog_word();

#define UAVDET_CHECK_ARG(cond, ...)            \
  do {                                         \
    if (!(cond)) {                             \
      uavdet::set_error(__VA_ARGS__);          \
      return UAVDET_ERR_ARG;                   \
    }                                          \
  } while (0)

#define UAVDET_CUDA(expr)                                                               \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      uavdet::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                        __LINE__);                                                      \
      return UAVDET_ERR_CUDA;                                                           \
    }                                                                                   \
  } while (0)

#define UAVDET_LAUNCH_CHECK()                                                            \
  do {                                                                                   \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess) {                                                             \
      uavdet::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),      \
                        __FILE__, __LINE__);                                             \
      return UAVDET_ERR_CUDA;                                                            \
    }                                                                                    \
    uavdet::count_launch();                                                              \
  } while (0)

constexpr int kNumSMs = 148;  // B200

// SMs the persistent tensor-core kernels may fill: kNumSMs minus the margin a data-parallel trainer reserves for the
// NCCL all-reduce kernels that run beside backward (uavdet_set_sm_margin; a persistent CTA that cannot be placed
// because a collective's CTA holds the SM would otherwise start a second wave).
int sm_budget();

// "first launch on this device" latch for per-device function attributes (cudaFuncSetAttribute is per device).
struct PerDeviceOnce {
  std::mutex mu;
  unsigned long long done = 0;
  template <typename F>
  cudaError_t run(F&& f) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    std::lock_guard<std::mutex> lock(mu);
    if (done & bit) return cudaSuccess;
    e = f();
    if (e == cudaSuccess) done |= bit;
    return e;
  }
};

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- device helpers -------------------------------------------------------------------
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

// sigmoid(z) = 0.5 * tanh(z / 2) + 0.5 with the hardware tanh: ONE special-function instruction (tanh.approx.f32,
// relative error ~2^-11, far below the bf16 rounding of every consumer) where 1 / (1 + exp(-z)) costs an exp2, a
// reciprocal and an IEEE division sequence.  The SiLU BatchNorm passes of DySOEM_SimFPN / DyYOLO ran at 3.5-4.2 TB/s
// against 5.4-6.1 TB/s for the LeakyReLU ones: two special-function results per element at the HBM rate is 76 % of the
// SM's special-function throughput (16 / clk) before anything else issues.
__device__ __forceinline__ float fast_sigmoid(float z) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * z));
  return fmaf(0.5f, t, 0.5f);
}

template <int ACT>
__device__ __forceinline__ float act_fwd(float z) {
  if (ACT == UAVDET_ACT_LEAKY) return z > 0.f ? z : 0.1f * z;
  if (ACT == UAVDET_ACT_SILU) return z * fast_sigmoid(z);
  if (ACT == UAVDET_ACT_RELU) return z > 0.f ? z : 0.f;
  if (ACT == UAVDET_ACT_GELU) return 0.5f * z * (1.f + erff(z * 0.70710678118654752f));
  return z;
}
__device__ __forceinline__ float act_fwd_rt(int act, float z) {
  switch (act) {
    case UAVDET_ACT_LEAKY: return act_fwd<UAVDET_ACT_LEAKY>(z);
    case UAVDET_ACT_SILU: return act_fwd<UAVDET_ACT_SILU>(z);
    case UAVDET_ACT_RELU: return act_fwd<UAVDET_ACT_RELU>(z);
    case UAVDET_ACT_GELU: return act_fwd<UAVDET_ACT_GELU>(z);
    default: return z;
  }
}
// derivative of the activation w.r.t. its pre-activation z
__device__ __forceinline__ float act_grad_rt(int act, float z) {
  switch (act) {
    case UAVDET_ACT_LEAKY: return z > 0.f ? 1.f : 0.1f;
    case UAVDET_ACT_SILU: {
      float s = fast_sigmoid(z);
      return s * (1.f + z * (1.f - s));
    }
    case UAVDET_ACT_RELU: return z > 0.f ? 1.f : 0.f;
    case UAVDET_ACT_GELU: {
      float cdf = 0.5f * (1.f + erff(z * 0.70710678118654752f));
      float pdf = 0.3989422804014327f * __expf(-0.5f * z * z);
      return cdf + z * pdf;
    }
    default: return 1.f;
  }
}

}  // namespace uavdet
