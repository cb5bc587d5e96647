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

// ---- programmatic dependent launch (PDL): an experiment that stayed a switch ----------------------------------------
// A training step is ~560 dependent launches of 10-700 us.  With cudaLaunchAttributeProgrammaticStreamSerialization a
// kernel that calls pdl_launch_dependents() lets its successor be launched as soon as every CTA of the grid has started
// (the successor's CTAs take the SMs this grid's CTAs leave and run their prologue), and the successor's pdl_wait()
// returns once the predecessor grid has COMPLETED and its writes are visible, so the data flow is unchanged.  Both
// instructions are no-ops for a kernel launched without the attribute.  Measured (UAVDET_PDL_MASK, see pdl_mask()):
// neutral for the implicit GEMM launches and 0.2-1.2 ms SLOWER per step for the BatchNorm kernels — their early-resident
// blocks (parked in pdl_wait) hold the registers the weight-gradient kernel of the side stream needs to get onto the SM,
// i.e. they trade the overlap that matters (tensor-core work under the HBM-bound passes) for launch latency that the
// CUDA graph had already made small.  Default: off.
// which launches carry the attribute (UAVDET_PDL_MASK, A/B): 1 implicit GEMM, 2 BatchNorm forward, 4 BatchNorm backward
// reduce, 8 BatchNorm backward apply
enum { kPdlIgemm = 1, kPdlBnFwd = 2, kPdlBnReduce = 4, kPdlBnApply = 8 };
int pdl_mask();
inline bool pdl_enabled(int who) { return (pdl_mask() & who) != 0; }
// fills attrs[*n] with the PDL attribute when enabled
inline void pdl_attr(cudaLaunchAttribute* attrs, unsigned* n, int who) {
  if (!pdl_enabled(who)) return;
  attrs[*n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[*n].val.programmaticStreamSerializationAllowed = 1;
  ++*n;
}
// <<<grid, block, smem, stream>>> with the PDL attribute
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(int who, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  unsigned n = 0;
  pdl_attr(attr, &n, who);
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- device helpers -------------------------------------------------------------------
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
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

// Standard normal CDF for the exact (erf) GELU of nn.GELU() (RTMUAVDet.py:155 channel MLP): Abramowitz & Stegun 26.2.17,
//   Q(|z|) = phi(|z|) (b1 t + ... + b5 t^5),  t = 1 / (1 + 0.2316419 |z|),  |error| < 7.5e-8,
// with 1 / sqrt(2 pi) folded into the coefficients and the hardware reciprocal / exp2 (two special-function
// instructions, 12 others).  Measured against erf in fp32 emulation over [-8, 8]: |gelu error| <= 6.2e-7, three orders
// below the bf16 rounding of the result.  erff() is ~40 instructions with a branch; in the implicit-GEMM epilogue of the
// 192 -> 192 and 384 -> 384 channel MLPs (629 M / 315 M elements at batch 128) it made those two launches 1.6 / 0.8 ms
// against 0.39 / 0.19 ms of HBM time.
__device__ __forceinline__ float gelu_cdf(float z) {
  const float a = fabsf(z);
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.2316419f, a, 1.f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-0.72134752f * a * a));
  float p = fmaf(t, 0.530702714f, -0.726576013f);
  p = fmaf(p, t, 0.710706871f);
  p = fmaf(p, t, -0.142248368f);
  p = fmaf(p, t, 0.127414796f);
  const float q = p * t * e;
  return z >= 0.f ? 1.f - q : q;
}

// Two GELUs per instruction stream (fma.rn.f32x2 / mul / add on register pairs, sm_100): the same operations per half as
// gelu_cdf, 13 packed + 4 special-function instructions per pair instead of 2 x 15 — the GELU epilogue of the channel-MLP
// GEMM is bound by instruction issue on its 8 epilogue warps.
__device__ __forceinline__ void gelu_pair(float& x0, float& x1) {
  const float2 a = make_float2(fabsf(x0), fabsf(x1));
  const float2 d = __ffma2_rn(make_float2(0.2316419f, 0.2316419f), a, make_float2(1.f, 1.f));
  const float2 ea = __fmul2_rn(__fmul2_rn(make_float2(-0.72134752f, -0.72134752f), a), a);
  float2 t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(d.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(d.y));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(ea.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(ea.y));
  float2 p = __ffma2_rn(t, make_float2(0.530702714f, 0.530702714f), make_float2(-0.726576013f, -0.726576013f));
  p = __ffma2_rn(p, t, make_float2(0.710706871f, 0.710706871f));
  p = __ffma2_rn(p, t, make_float2(-0.142248368f, -0.142248368f));
  p = __ffma2_rn(p, t, make_float2(0.127414796f, 0.127414796f));
  const float2 q = __fmul2_rn(__fmul2_rn(p, t), e);
  const float2 u = __fadd2_rn(make_float2(1.f, 1.f), make_float2(-q.x, -q.y));
  const float2 r = __fmul2_rn(make_float2(x0, x1), make_float2(x0 >= 0.f ? u.x : q.x, x1 >= 0.f ? u.y : q.y));
  x0 = r.x; x1 = r.y;
}

// Two SiLUs (z * fast_sigmoid(z)) on packed instructions: 3 packed + 2 special-function instructions per pair.
__device__ __forceinline__ void silu_pair(float& x0, float& x1) {
  const float2 z = make_float2(x0, x1);
  const float2 h = __fmul2_rn(make_float2(0.5f, 0.5f), z);
  float2 t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(h.x));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(h.y));
  const float2 r = __fmul2_rn(z, __ffma2_rn(make_float2(0.5f, 0.5f), t, make_float2(0.5f, 0.5f)));
  x0 = r.x; x1 = r.y;
}

template <int ACT>
__device__ __forceinline__ float act_fwd(float z) {
  if (ACT == UAVDET_ACT_LEAKY) return z > 0.f ? z : 0.1f * z;
  if (ACT == UAVDET_ACT_SILU) return z * fast_sigmoid(z);
  if (ACT == UAVDET_ACT_RELU) return z > 0.f ? z : 0.f;
  if (ACT == UAVDET_ACT_GELU) return z * gelu_cdf(z);
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
      float cdf = gelu_cdf(z);
      float pdf = 0.3989422804014327f * __expf(-0.5f * z * z);
      return cdf + z * pdf;
    }
    default: return 1.f;
  }
}

}  // namespace uavdet
