// K5 and friends — HBM-bound kernels: BN finalize / normalise+activation (fwd, bwd), nearest
// upsample, adds, layout conversion, weight (un)packing, SGD.  All NHWC bf16, 128-bit accesses,
// grid-stride loops sized to a multiple of the SM count.
// Reference sites: BaselineModel.py:14-22 (BN+LeakyReLU), _base.py:19-20,50,75 (BN+SiLU/ReLU),
// BaselineModel.py:43 (residual), BaselineModel.py:86,120-122 (Upsample+cat), _base.py:292 (SGD).
#include "common.cuh"
#include <stdlib.h>

namespace uavdet {

static inline int ew_grid(long long work_items, int threads) {
  long long blocks = (work_items + threads - 1) / threads;
  long long cap = (long long)kNumSMs * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

struct View {
  __nv_bfloat16* p;
  long long npix;  // n*h*w
  int c, ld;
};
static inline View mkview(const uavdet_act* a) {
  return View{(__nv_bfloat16*)a->ptr, (long long)a->n * a->h * a->w, a->c, a->ld};
}
static int check_view(const uavdet_act* a, const char* what) {
  UAVDET_CHECK_ARG(a && a->ptr, "%s: null view", what);
  UAVDET_CHECK_ARG(a->c % 8 == 0 && a->ld % 8 == 0 && ((uintptr_t)a->ptr & 15) == 0,
                   "%s: view must be 8-channel / 16-byte aligned (c=%d ld=%d)", what, a->c, a->ld);
  return UAVDET_OK;
}
static int same_shape(const uavdet_act* a, const uavdet_act* b, const char* what) {
  UAVDET_CHECK_ARG(a->n == b->n && a->h == b->h && a->w == b->w && a->c == b->c,
                   "%s: shape mismatch (%d,%d,%d,%d) vs (%d,%d,%d,%d)", what, a->n, a->h, a->w, a->c, b->n,
                   b->h, b->w, b->c);
  return UAVDET_OK;
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  return u;
}
__device__ __forceinline__ void load8f(const float* p, float (&f)[8]) {
  float4 a = __ldg(reinterpret_cast<const float4*>(p));
  float4 b = __ldg(reinterpret_cast<const float4*>(p + 4));
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// ---- BN finalize --------------------------------------------------------------------------
__global__ void bn_finalize_kernel(const float* sum, const float* sumsq, int c, double count, float eps,
                                   float momentum, const float* gamma, const float* beta,
                                   float* running_mean, float* running_var, float* mean, float* invstd,
                                   float* scale, float* shift) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  double m = (double)sum[i] / count;
  double var = (double)sumsq[i] / count - m * m;
  if (var < 0.0) var = 0.0;
  float is = (float)(1.0 / sqrt(var + (double)eps));
  float g = gamma ? gamma[i] : 1.f, b = beta ? beta[i] : 0.f;
  if (mean) mean[i] = (float)m;
  if (invstd) invstd[i] = is;
  scale[i] = g * is;
  shift[i] = b - (float)m * g * is;
  if (running_mean) {
    double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[i] = (1.f - momentum) * running_mean[i] + momentum * (float)m;
    running_var[i] = (1.f - momentum) * running_var[i] + momentum * (float)unbiased;
  }
}

// ---- streaming skeleton ---------------------------------------------------------------------
// blockDim 256 = Gb channel groups (8 channels = 16 B each) x PL pixel lanes; a thread keeps its
// channel group for the whole kernel, so per-channel parameters are loaded once, and walks pixels
// with a 4x unrolled grid-stride loop (4 independent 16 B loads per operand in flight).
struct PixLane {
  int c;            // first channel of this thread's group
  long long px0;    // first pixel
  long long step;   // pixel stride of the loop
  bool active;
};
__device__ __forceinline__ PixLane pix_lane(int channels) {
  const int G = channels >> 3;
  const int Gb = G < 32 ? G : 32;
  const int PL = 256 / Gb;
  const int g = threadIdx.x % Gb + blockIdx.y * Gb;
  const int pl = threadIdx.x / Gb;
  PixLane L;
  L.c = g << 3;
  L.px0 = (long long)blockIdx.x * PL + pl;
  L.step = (long long)gridDim.x * PL;
  L.active = g < G && pl < PL;
  return L;
}
static dim3 stream_grid(const uavdet_act* v, int unroll, int blocks_per_sm = 16) {
  int G = v->c / 8, Gb = G < 32 ? G : 32, PL = 256 / Gb;
  long long npix = (long long)v->n * v->h * v->w;
  long long bx = (npix + (long long)PL * unroll - 1) / ((long long)PL * unroll);
  long long cap = ((long long)kNumSMs * blocks_per_sm) / ceil_div(G, Gb);
  if (cap < kNumSMs) cap = kNumSMs;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  return dim3((unsigned)bx, (unsigned)ceil_div(G, Gb), 1);
}

// ---- y = act(raw*scale+shift) (+res) -------------------------------------------------------
template <bool HAS_RES>
__global__ void __launch_bounds__(256)
bn_act_fwd_kernel(View raw, const float* __restrict__ scale, const float* __restrict__ shift, int act,
                  const __nv_bfloat16* __restrict__ res, int res_ld, int reverse, View y) {
  pdl_launch_dependents();
  pdl_wait();
  const PixLane L = pix_lane(raw.c);
  if (!L.active) return;
  float s[8], t[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s[j] = 1.f; t[j] = 0.f; }
  if (scale) load8f(scale + L.c, s);
  if (shift) load8f(shift + L.c, t);
  auto body = [&](const uint4& in, const uint4& rin, long long px) {
    float v[8];
    unpack8(in, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = act_fwd_rt(act, fmaf(v[j], s[j], t[j]));
    if (HAS_RES) {
      float r[8];
      unpack8(rin, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += r[j];
    }
    *reinterpret_cast<uint4*>(y.p + px * y.ld + L.c) = pack8(v);
  };
  // reverse: pixels from the end — the convolution that produced `raw` wrote it front to back, so its tail is what the
  // L2 still holds, and the head of y, written last here, is what the next convolution reads first
  const long long last = raw.npix - 1;
  long long px = L.px0;
  for (; px + 3 * L.step < raw.npix; px += 4 * L.step) {
    uint4 a[4], r[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long q = reverse ? last - (px + u * L.step) : px + u * L.step;
      a[u] = __ldcs(reinterpret_cast<const uint4*>(raw.p + q * raw.ld + L.c));
      if (HAS_RES) r[u] = __ldg(reinterpret_cast<const uint4*>(res + q * res_ld + L.c));
      else r[u] = make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) body(a[u], r[u], reverse ? last - (px + u * L.step) : px + u * L.step);
  }
  for (; px < raw.npix; px += L.step) {
    const long long q = reverse ? last - px : px;
    uint4 a = __ldcs(reinterpret_cast<const uint4*>(raw.p + q * raw.ld + L.c));
    uint4 r = make_uint4(0, 0, 0, 0);
    if (HAS_RES) r = __ldg(reinterpret_cast<const uint4*>(res + q * res_ld + L.c));
    body(a, r, q);
  }
}

// ---- train-mode BN in one pass over the activations: finalize folded in --------------------------------
// The first (channels of the block) threads each derive scale/shift of ONE channel from the batch sums and pass them
// through shared memory (the fp64 part — E[x^2] - E[x]^2 cancels in fp32 — is three operations per block and channel,
// not 24 per thread); block column 0 also publishes mean / invstd / scale / shift (backward needs them) and updates
// the running statistics (momentum, unbiased variance — nn.BatchNorm2d, BaselineModel.py:14).
template <bool HAS_RES>
__global__ void __launch_bounds__(256)
bn_train_fwd_kernel(View raw, const float* __restrict__ sum, const float* __restrict__ sumsq, double count, float eps,
                    float momentum, const float* __restrict__ gamma, const float* __restrict__ beta,
                    float* __restrict__ running_mean, float* __restrict__ running_var, float* __restrict__ mean_out,
                    float* __restrict__ invstd_out, float* __restrict__ scale_out, float* __restrict__ shift_out, int act,
                    const __nv_bfloat16* __restrict__ res, int res_ld, View y) {
  __shared__ float sh_s[256], sh_t[256];
  pdl_launch_dependents();
  pdl_wait();
  const PixLane L = pix_lane(raw.c);
  {
    const int G = raw.c >> 3;
    const int Gb = G < 32 ? G : 32;
    const int c = blockIdx.y * Gb * 8 + threadIdx.x;
    if ((int)threadIdx.x < Gb * 8 && c < raw.c) {
      const double inv_count = 1.0 / count;
      const double m = (double)__ldg(sum + c) * inv_count;
      double var = fma(-m, m, (double)__ldg(sumsq + c) * inv_count);
      if (var < 0.0) var = 0.0;
      const float is = rsqrtf((float)var + eps);
      const float g = gamma ? __ldg(gamma + c) : 1.f, b = beta ? __ldg(beta + c) : 0.f;
      const float sc = g * is, sf = b - (float)m * g * is;
      sh_s[threadIdx.x] = sc;
      sh_t[threadIdx.x] = sf;
      if (blockIdx.x == 0) {
        if (mean_out) mean_out[c] = (float)m;
        if (invstd_out) invstd_out[c] = is;
        scale_out[c] = sc;
        shift_out[c] = sf;
        if (running_mean) {
          const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
          running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
          running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
        }
      }
    }
  }
  __syncthreads();
  if (!L.active) return;
  float s[8], t[8];
  {
    const int G = raw.c >> 3;
    const int Gb = G < 32 ? G : 32;
    const int o = L.c - blockIdx.y * Gb * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] = sh_s[o + j]; t[j] = sh_t[o + j]; }
  }
  auto body = [&](const uint4& in, const uint4& rin, long long px) {
    float v[8];
    unpack8(in, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = act_fwd_rt(act, fmaf(v[j], s[j], t[j]));
    if (HAS_RES) {
      float r[8];
      unpack8(rin, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += r[j];
    }
    *reinterpret_cast<uint4*>(y.p + px * y.ld + L.c) = pack8(v);
  };
  long long px = L.px0;
  for (; px + 3 * L.step < raw.npix; px += 4 * L.step) {
    uint4 a[4], r[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      a[u] = __ldcs(reinterpret_cast<const uint4*>(raw.p + (px + u * L.step) * raw.ld + L.c));
      if (HAS_RES) r[u] = __ldg(reinterpret_cast<const uint4*>(res + (px + u * L.step) * res_ld + L.c));
      else r[u] = make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) body(a[u], r[u], px + u * L.step);
  }
  for (; px < raw.npix; px += L.step) {
    uint4 a = __ldcs(reinterpret_cast<const uint4*>(raw.p + px * raw.ld + L.c));
    uint4 r = make_uint4(0, 0, 0, 0);
    if (HAS_RES) r = __ldg(reinterpret_cast<const uint4*>(res + px * res_ld + L.c));
    body(a, r, px);
  }
}

// ---- BN backward, phase 1: per-channel sums of dz and dz*raw ---------------------------------
// At most 85 registers: the kernel runs beside the weight-gradient kernel of the previous layer (192 threads x 96
// registers and 176 KB of shared memory per SM), and its two blocks per SM only fit next to it below 92 registers — at
// 98 the second block of every SM waited for the first, i.e. the pass ran in two waves whenever a weight gradient was
// in flight.
template <int MINB>
__global__ void __launch_bounds__(256, MINB)
bn_bwd_reduce_kernel(View dy, View raw, const float* __restrict__ scale, const float* __restrict__ shift, int act,
                     float* __restrict__ sum_dz, float* __restrict__ sum_dzr) {
  __shared__ float red[256 * 16];
  pdl_launch_dependents();
  pdl_wait();
  const PixLane L = pix_lane(dy.c);
  float a_dz[8], a_dzr[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a_dz[j] = 0.f; a_dzr[j] = 0.f; }
  if (L.active) {
    float s[8], t[8];
    load8f(scale + L.c, s);
    load8f(shift + L.c, t);
    auto body = [&](const uint4& din, const uint4& rin) {
      float d[8], r[8];
      unpack8(din, d);
      unpack8(rin, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float dz = d[j] * act_grad_rt(act, fmaf(r[j], s[j], t[j]));
        a_dz[j] += dz;
        a_dzr[j] = fmaf(dz, r[j], a_dzr[j]);
      }
    };
    long long px = L.px0;
    for (; px + 3 * L.step < dy.npix; px += 4 * L.step) {
      uint4 a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a[u] = __ldg(reinterpret_cast<const uint4*>(dy.p + (px + u * L.step) * dy.ld + L.c));
        b[u] = __ldg(reinterpret_cast<const uint4*>(raw.p + (px + u * L.step) * raw.ld + L.c));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) body(a[u], b[u]);
    }
    for (; px < dy.npix; px += L.step)
      body(__ldg(reinterpret_cast<const uint4*>(dy.p + px * dy.ld + L.c)),
           __ldg(reinterpret_cast<const uint4*>(raw.p + px * raw.ld + L.c)));
  }
  // block reduce over pixel lanes
  const int G = dy.c >> 3;
  const int Gb = G < 32 ? G : 32;
  const int PL = 256 / Gb;
  float* mine = red + threadIdx.x * 16;
#pragma unroll
  for (int j = 0; j < 8; ++j) { mine[j] = a_dz[j]; mine[8 + j] = a_dzr[j]; }
  __syncthreads();
  for (int t = threadIdx.x; t < Gb * 16; t += 256) {
    const int gg = t >> 4, slot = t & 15;
    float acc = 0.f;
    for (int l = 0; l < PL; ++l) acc += red[(l * Gb + gg) * 16 + slot];
    const int gch = gg + blockIdx.y * Gb;
    if (gch < G) atomicAdd((slot < 8 ? sum_dz : sum_dzr) + (gch << 3) + (slot & 7), acc);
  }
}

// per channel: dgamma, dbeta and the folded coefficients of phase 2
//   d_raw = scale*dz + k1*raw + k0,  k1 = -scale*invstd*dgamma/M,  k0 = -scale*dbeta/M - k1*mean
__global__ void bn_bwd_finalize_kernel(const float* sum_dz, const float* sum_dzr, const float* mean,
                                       const float* invstd, const float* scale, int c, float inv_count,
                                       float* dgamma, float* dbeta, float* k1, float* k0) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  const float sd = sum_dz[i];
  const float dg = invstd[i] * (sum_dzr[i] - mean[i] * sd);
  dgamma[i] = dg;
  dbeta[i] = sd;
  const float a1 = -scale[i] * invstd[i] * dg * inv_count;
  k1[i] = a1;
  k0[i] = -scale[i] * sd * inv_count - a1 * mean[i];
}

// ---- BN backward, phase 2 ----------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(View dy, View raw, const float* __restrict__ scale, const float* __restrict__ shift,
                    const float* __restrict__ k1, const float* __restrict__ k0, int act, View dr) {
  const PixLane L = pix_lane(dy.c);
  if (!L.active) return;
  float s[8], t[8], a1[8], a0[8];
  load8f(scale + L.c, s);
  load8f(shift + L.c, t);
  load8f(k1 + L.c, a1);
  load8f(k0 + L.c, a0);
  auto body = [&](const uint4& din, const uint4& rin, long long px) {
    float d[8], r[8];
    unpack8(din, d);
    unpack8(rin, r);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float dz = d[j] * act_grad_rt(act, fmaf(r[j], s[j], t[j]));
      d[j] = fmaf(s[j], dz, fmaf(a1[j], r[j], a0[j]));
    }
    *reinterpret_cast<uint4*>(dr.p + px * dr.ld + L.c) = pack8(d);
  };
  long long px = L.px0;
  for (; px + 3 * L.step < dy.npix; px += 4 * L.step) {
    uint4 a[4], b[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      a[u] = __ldg(reinterpret_cast<const uint4*>(dy.p + (px + u * L.step) * dy.ld + L.c));
      b[u] = __ldg(reinterpret_cast<const uint4*>(raw.p + (px + u * L.step) * raw.ld + L.c));
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) body(a[u], b[u], px + u * L.step);
  }
  for (; px < dy.npix; px += L.step)
    body(__ldg(reinterpret_cast<const uint4*>(dy.p + px * dy.ld + L.c)),
         __ldg(reinterpret_cast<const uint4*>(raw.p + px * raw.ld + L.c)), px);
}

// ---- BN backward, phase 2 with the per-channel finalize folded in -------------------------------------------------
template <int U, int MINB>
__global__ void __launch_bounds__(256, MINB)
bn_bwd_apply_fused_kernel(View dy, View raw, const float* __restrict__ scale, const float* __restrict__ shift,
                          const float* __restrict__ sum_dz, const float* __restrict__ sum_dzr,
                          const float* __restrict__ mean, const float* __restrict__ invstd, float inv_count, int act,
                          float* __restrict__ dgamma, float* __restrict__ dbeta, int accumulate, int reverse, View dr) {
  pdl_launch_dependents();
  pdl_wait();
  const PixLane L = pix_lane(dy.c);
  if (!L.active) return;
  float s[8], t[8], a1[8], a0[8];
  load8f(scale + L.c, s);
  load8f(shift + L.c, t);
  {
    float sd[8], sdr[8], mu[8], is[8];
    load8f(sum_dz + L.c, sd);
    load8f(sum_dzr + L.c, sdr);
    load8f(mean + L.c, mu);
    load8f(invstd + L.c, is);
    const bool publish = blockIdx.x == 0 && threadIdx.x < (dy.c >> 3 < 32 ? dy.c >> 3 : 32);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float dg = is[j] * (sdr[j] - mu[j] * sd[j]);
      a1[j] = -s[j] * is[j] * dg * inv_count;
      a0[j] = -s[j] * sd[j] * inv_count - a1[j] * mu[j];
      // one thread per channel publishes; `accumulate` adds into a live gradient buffer (param.grad) instead
      if (publish) {
        dgamma[L.c + j] = accumulate ? dgamma[L.c + j] + dg : dg;
        dbeta[L.c + j] = accumulate ? dbeta[L.c + j] + sd[j] : sd[j];
      }
    }
  }
  auto body = [&](const uint4& din, const uint4& rin, long long px) {
    float d[8], r[8];
    unpack8(din, d);
    unpack8(rin, r);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float dz = d[j] * act_grad_rt(act, fmaf(r[j], s[j], t[j]));
      d[j] = fmaf(s[j], dz, fmaf(a1[j], r[j], a0[j]));
    }
    *reinterpret_cast<uint4*>(dr.p + px * dr.ld + L.c) = pack8(d);
  };
  // reverse: walk the pixels from the end.  The reduction pass that ran just before read both tensors front to back,
  // so what is still in the 126 MB L2 is their tail — and what this pass writes last (the head of d_raw) is what the
  // data-gradient kernel that follows reads first.
  const long long last = dy.npix - 1;
  long long px = L.px0;
  for (; px + (U - 1) * L.step < dy.npix; px += U * L.step) {
    uint4 a[U], b[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long q = reverse ? last - (px + u * L.step) : px + u * L.step;
      a[u] = __ldcs(reinterpret_cast<const uint4*>(dy.p + q * dy.ld + L.c));
      b[u] = __ldcs(reinterpret_cast<const uint4*>(raw.p + q * raw.ld + L.c));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) body(a[u], b[u], reverse ? last - (px + u * L.step) : px + u * L.step);
  }
  for (; px < dy.npix; px += L.step) {
    const long long q = reverse ? last - px : px;
    body(__ldcs(reinterpret_cast<const uint4*>(dy.p + q * dy.ld + L.c)),
         __ldcs(reinterpret_cast<const uint4*>(raw.p + q * raw.ld + L.c)), q);
  }
}

// ---- BN backward, phase 2, coefficients in shared memory -----------------------------------------------------------
// Same math as bn_bwd_apply_fused_kernel.  The four per-channel coefficient vectors (scale, shift, k1, k0: 32 registers
// per thread there) live in shared memory instead — derived once per block by its first threads, one channel each —
// and are read back with 16-byte loads at the point of use, which frees the registers for U pixel rows x two operands
// of loads in flight per thread at <= 85 registers.  That pass shares the SM with the weight-gradient kernel of the
// previous layer (18 K registers, 176 KB of shared memory): two 256-thread blocks fit beside it either way, and with
// U = 4 they keep 64 KB of loads in flight instead of 32 KB; beside that kernel HBM ran at 4.1 TB/s, i.e. the pass was
// latency-bound, not bandwidth-bound.
template <int U, int MINB>
__global__ void __launch_bounds__(256, MINB)
bn_bwd_apply_smem_kernel(View dy, View raw, const float* __restrict__ scale, const float* __restrict__ shift,
                         const float* __restrict__ sum_dz, const float* __restrict__ sum_dzr,
                         const float* __restrict__ mean, const float* __restrict__ invstd, float inv_count, int act,
                         float* __restrict__ dgamma, float* __restrict__ dbeta, int accumulate, View dr) {
  __shared__ __align__(16) float cs[4][256];            // scale, shift, k1, k0 of the block's (<= 256) channels
  pdl_launch_dependents();
  pdl_wait();
  const PixLane L = pix_lane(dy.c);
  const int G = dy.c >> 3;
  const int Gb = G < 32 ? G : 32;
  const int cbase = blockIdx.y * Gb * 8;
  {
    const int c = cbase + threadIdx.x;
    if ((int)threadIdx.x < Gb * 8 && c < dy.c) {
      const float sc = __ldg(scale + c), is = __ldg(invstd + c), sd = __ldg(sum_dz + c), mu = __ldg(mean + c);
      const float dg = is * (__ldg(sum_dzr + c) - mu * sd);
      const float a1 = -sc * is * dg * inv_count;
      cs[0][threadIdx.x] = sc;
      cs[1][threadIdx.x] = __ldg(shift + c);
      cs[2][threadIdx.x] = a1;
      cs[3][threadIdx.x] = -sc * sd * inv_count - a1 * mu;
      if (blockIdx.x == 0) {
        dgamma[c] = accumulate ? dgamma[c] + dg : dg;
        dbeta[c] = accumulate ? dbeta[c] + sd : sd;
      }
    }
  }
  __syncthreads();
  if (!L.active) return;
  const int o = L.c - cbase;
  const float4* c4 = reinterpret_cast<const float4*>(&cs[0][0]);
  auto body = [&](const uint4& din, const uint4& rin, long long px) {
    float d[8], r[8];
    unpack8(din, d);
    unpack8(rin, r);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float4 s4 = c4[(0 * 256 + o) / 4 + h], t4 = c4[(1 * 256 + o) / 4 + h];
      const float4 a4 = c4[(2 * 256 + o) / 4 + h], b4 = c4[(3 * 256 + o) / 4 + h];
      const float sv[4] = {s4.x, s4.y, s4.z, s4.w}, tv[4] = {t4.x, t4.y, t4.z, t4.w};
      const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int q = 4 * h + j;
        const float dz = d[q] * act_grad_rt(act, fmaf(r[q], sv[j], tv[j]));
        d[q] = fmaf(sv[j], dz, fmaf(av[j], r[q], bv[j]));
      }
    }
    *reinterpret_cast<uint4*>(dr.p + px * dr.ld + L.c) = pack8(d);
  };
  long long px = L.px0;
  for (; px + (U - 1) * L.step < dy.npix; px += U * L.step) {
    uint4 a[U], b[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long q = px + u * L.step;
      a[u] = __ldcs(reinterpret_cast<const uint4*>(dy.p + q * dy.ld + L.c));
      b[u] = __ldcs(reinterpret_cast<const uint4*>(raw.p + q * raw.ld + L.c));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) body(a[u], b[u], px + u * L.step);
  }
  for (; px < dy.npix; px += L.step)
    body(__ldcs(reinterpret_cast<const uint4*>(dy.p + px * dy.ld + L.c)),
         __ldcs(reinterpret_cast<const uint4*>(raw.p + px * raw.ld + L.c)), px);
}

__global__ void act_bwd_kernel(View dy, View raw, const float* __restrict__ scale,
                               const float* __restrict__ shift, int act, View dx) {
  const int c8 = dy.c >> 3;
  const long long total = dy.npix * c8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long px = i / c8;
    const int c = (int)(i - px * c8) << 3;
    float d[8], r[8], s[8], t[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(dy.p + px * dy.ld + c)), d);
    unpack8(__ldg(reinterpret_cast<const uint4*>(raw.p + px * raw.ld + c)), r);
    if (scale) load8f(scale + c, s);
    if (shift) load8f(shift + c, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float sc = scale ? s[j] : 1.f;
      float z = r[j] * sc + (shift ? t[j] : 0.f);
      d[j] = d[j] * act_grad_rt(act, z) * sc;
    }
    *reinterpret_cast<uint4*>(dx.p + px * dx.ld + c) = pack8(d);
  }
}

// ---- upsample / add ---------------------------------------------------------------------------
__global__ void upsample2x_fwd_kernel(const __nv_bfloat16* __restrict__ x, int x_ld, int n, int h, int w, int c,
                                      __nv_bfloat16* __restrict__ y, int y_ld) {
  const int c8 = c >> 3;
  const int H = 2 * h, W = 2 * w;
  const long long total = (long long)n * H * W * c8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long px = i / c8;
    const int cc = (int)(i - px * c8) << 3;
    const int ox = (int)(px % W); long long t = px / W;
    const int oy = (int)(t % H);
    const int b = (int)(t / H);
    const long long src = ((long long)b * h + (oy >> 1)) * w + (ox >> 1);
    *reinterpret_cast<uint4*>(y + px * y_ld + cc) = __ldg(reinterpret_cast<const uint4*>(x + src * x_ld + cc));
  }
}

__global__ void upsample2x_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int dy_ld, int n, int h, int w, int c,
                                      __nv_bfloat16* __restrict__ dx, int dx_ld, int accumulate) {
  // dx is (n,h,w,c); dy is (n,2h,2w,c)
  const int c8 = c >> 3;
  const int W = 2 * w;
  const long long total = (long long)n * h * w * c8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long px = i / c8;
    const int cc = (int)(i - px * c8) << 3;
    const int x0 = (int)(px % w); long long t = px / w;
    const int y0 = (int)(t % h);
    const int b = (int)(t / h);
    const long long base = ((long long)b * 2 * h + 2 * y0) * W + 2 * x0;
    float acc[8], v[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(dy + base * dy_ld + cc)), acc);
    unpack8(__ldg(reinterpret_cast<const uint4*>(dy + (base + 1) * dy_ld + cc)), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += v[j];
    unpack8(__ldg(reinterpret_cast<const uint4*>(dy + (base + W) * dy_ld + cc)), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += v[j];
    unpack8(__ldg(reinterpret_cast<const uint4*>(dy + (base + W + 1) * dy_ld + cc)), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += v[j];
    if (accumulate) {
      unpack8(*reinterpret_cast<const uint4*>(dx + px * dx_ld + cc), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
    *reinterpret_cast<uint4*>(dx + px * dx_ld + cc) = pack8(acc);
  }
}

__global__ void upsample2x_add_kernel(const __nv_bfloat16* __restrict__ b, int b_ld, int n, int h, int w, int c,
                                      const __nv_bfloat16* __restrict__ a, int a_ld, float a_mult,
                                      __nv_bfloat16* __restrict__ y, int y_ld) {
  // (n, 2h, 2w, c) output; b is (n, h, w, c)
  const int c8 = c >> 3;
  const int H = 2 * h, W = 2 * w;
  const long long total = (long long)n * H * W * c8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long px = i / c8;
    const int cc = (int)(i - px * c8) << 3;
    const int ox = (int)(px % W); long long t = px / W;
    const int oy = (int)(t % H);
    const int img = (int)(t / H);
    const long long src = ((long long)img * h + (oy >> 1)) * w + (ox >> 1);
    float va[8], vb[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(a + px * a_ld + cc)), va);
    unpack8(__ldg(reinterpret_cast<const uint4*>(b + src * b_ld + cc)), vb);
#pragma unroll
    for (int j = 0; j < 8; ++j) va[j] = fmaf(a_mult, va[j], vb[j]);
    *reinterpret_cast<uint4*>(y + px * y_ld + cc) = pack8(va);
  }
}

__global__ void add_kernel(View a, const __nv_bfloat16* __restrict__ b, int b_ld, View y) {
  const int c8 = a.c >> 3;
  const long long total = a.npix * c8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long px = i / c8;
    const int c = (int)(i - px * c8) << 3;
    uint4 u = __ldg(reinterpret_cast<const uint4*>(a.p + px * a.ld + c));
    if (b) {
      float va[8], vb[8];
      unpack8(u, va);
      unpack8(__ldg(reinterpret_cast<const uint4*>(b + px * b_ld + c)), vb);
#pragma unroll
      for (int j = 0; j < 8; ++j) va[j] += vb[j];
      u = pack8(va);
    }
    *reinterpret_cast<uint4*>(y.p + px * y.ld + c) = u;
  }
}

// ---- layout conversion (API edge) ----------------------------------------------------------------
__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ x, int ld, int n, int hw, int c,
                                    float* __restrict__ y) {
  // tile transpose 32 (pixels) x 32 (channels) through shared memory
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int p = p0 + r, cc = c0 + threadIdx.x;
    tile[r][threadIdx.x] = (p < hw && cc < c) ? __bfloat162float(x[((long long)b * hw + p) * ld + cc]) : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int cc = c0 + r, p = p0 + threadIdx.x;
    if (p < hw && cc < c) y[((long long)b * c + cc) * hw + p] = tile[threadIdx.x][r];
  }
}
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, int n, int hw, int c,
                                    __nv_bfloat16* __restrict__ y, int ld) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int cc = c0 + r, p = p0 + threadIdx.x;
    tile[r][threadIdx.x] = (p < hw && cc < c) ? x[((long long)b * c + cc) * hw + p] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int p = p0 + r, cc = c0 + threadIdx.x;
    if (p < hw && cc < c) y[((long long)b * hw + p) * ld + cc] = __float2bfloat16(tile[threadIdx.x][r]);
  }
}

// ---- weights ---------------------------------------------------------------------------------------
// out[o][tap][i] (transposed: out[i][tap][o]) = w[o][i][tap]   (src_nhwc: the source is stored [o][tap][i])
__device__ __forceinline__ void pack_weight_elems(const float* __restrict__ w, int O, int I, int kk, int transposed,
                                                  int src_nhwc, __nv_bfloat16* __restrict__ out, long long first,
                                                  long long stride) {
  const long long total = (long long)O * I * kk;
  for (long long idx = first; idx < total; idx += stride) {
    int o, i, t;
    if (!transposed) { i = (int)(idx % I); long long r = idx / I; t = (int)(r % kk); o = (int)(r / kk); }
    else { o = (int)(idx % O); long long r = idx / O; t = (int)(r % kk); i = (int)(r / kk); }
    const long long src = src_nhwc ? ((long long)o * kk + t) * I + i : ((long long)o * I + i) * kk + t;
    out[idx] = __float2bfloat16(__ldg(w + src));
  }
}
__global__ void pack_weight_kernel(const float* __restrict__ w, int O, int I, int kk, int transposed, int src_nhwc,
                                   __nv_bfloat16* __restrict__ out) {
  pack_weight_elems(w, O, I, kk, transposed, src_nhwc, out, blockIdx.x * (long long)blockDim.x + threadIdx.x,
                    (long long)gridDim.x * blockDim.x);
}
// every conv weight of a model (both orientations) in ONE launch: a block = one 4096-element chunk of one job
constexpr int kPackChunk = 4096;
__global__ void __launch_bounds__(256)
pack_weights_batched_kernel(const uavdet_pack_job* __restrict__ jobs, int n_jobs) {
  int lo = 0, hi = n_jobs - 1;                       // last job whose first chunk is <= blockIdx.x
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].chunk0 <= (long long)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const uavdet_pack_job j = jobs[lo];
  const int kk = j.k * j.k, transposed = j.flags & 1, src_nhwc = (j.flags >> 1) & 1;
  const long long total = (long long)j.O * j.I * kk;
  const long long base = ((long long)blockIdx.x - j.chunk0) * kPackChunk;
  __nv_bfloat16* out = (__nv_bfloat16*)j.dst;
  const int n_here = (int)((total - base) < kPackChunk ? (total - base) : kPackChunk);
  if (!transposed && (src_nhwc || kk == 1)) {
    // same element order in source and destination: a straight fp32 -> bf16 conversion, 8 elements per thread
    const float* src = j.src + base;
    if ((((uintptr_t)src | (uintptr_t)(out + base)) & 15) == 0) {
      for (int e = threadIdx.x * 8; e + 8 <= n_here; e += 256 * 8) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(src + e));
        const float4 b = __ldg(reinterpret_cast<const float4*>(src + e + 4));
        uint4 o;
        o.x = pack_bf16x2(a.x, a.y); o.y = pack_bf16x2(a.z, a.w);
        o.z = pack_bf16x2(b.x, b.y); o.w = pack_bf16x2(b.z, b.w);
        *reinterpret_cast<uint4*>(out + base + e) = o;
      }
      for (int e = (n_here & ~7) + threadIdx.x; e < n_here; e += 256) out[base + e] = __float2bfloat16(__ldg(src + e));
      return;
    }
  }
  const uint32_t I = (uint32_t)j.I, O = (uint32_t)j.O, ukk = (uint32_t)kk;
  if (transposed && (src_nhwc || kk == 1) && (O & 31u) == 0 && (I & 31u) == 0) {
    // per tap a 2-D transpose [O][I] -> [I][O]: 32 x 32 tiles through shared memory, reads contiguous along I,
    // writes contiguous along O; this block's 4096 output elements are four tiles
    __shared__ float tile[32][33];
    const uint32_t tiles_o = O >> 5, tiles_i = I >> 5;
    const uint32_t n_tiles = ukk * tiles_i * tiles_o;
    const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (uint32_t q = 0; q < kPackChunk / 1024; ++q) {
      const uint32_t tile_id = (uint32_t)(base >> 10) + q;
      if (tile_id >= n_tiles) break;
      const uint32_t to = tile_id % tiles_o, r = tile_id / tiles_o, ti = r % tiles_i, t = r / tiles_i;
      const uint32_t o0 = to << 5, i0 = ti << 5;
#pragma unroll
      for (uint32_t rr = ty; rr < 32; rr += 8)
        tile[rr][tx] = __ldg(j.src + ((size_t)(o0 + rr) * ukk + t) * I + i0 + tx);
      __syncthreads();
#pragma unroll
      for (uint32_t rr = ty; rr < 32; rr += 8)
        out[((size_t)(i0 + rr) * ukk + t) * O + o0 + tx] = __float2bfloat16(tile[tx][rr]);
      __syncthreads();
    }
    return;
  }
  // general case; a weight tensor has < 2^31 elements, so the index arithmetic stays in 32 bits
  for (int e = threadIdx.x; e < n_here; e += 256) {
    const uint32_t idx = (uint32_t)base + (uint32_t)e;
    uint32_t o, i, t;
    if (!transposed) { i = idx % I; const uint32_t r = idx / I; t = r % ukk; o = r / ukk; }
    else { o = idx % O; const uint32_t r = idx / O; t = r % ukk; i = r / ukk; }
    const uint32_t src = src_nhwc ? (o * ukk + t) * I + i : (o * I + i) * ukk + t;
    out[idx] = __float2bfloat16(__ldg(j.src + src));
  }
}
__global__ void unpack_wgrad_kernel(const float* __restrict__ dwp, int O, int I, int kk, float* __restrict__ g,
                                    int accumulate) {
  const long long total = (long long)O * I * kk;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    // idx enumerates the OIHW destination; read the packed [o][tap][i] source
    int t = (int)(idx % kk); long long r = idx / kk; int i = (int)(r % I); int o = (int)(r / I);
    float v = __ldg(dwp + ((long long)o * kk + t) * I + i);
    g[idx] = accumulate ? g[idx] + v : v;
  }
}

// out[b][o][tap][i] = sum_k attn[b][k] * bank[k][o][i][tap]   (transposed: out[b][i][tap][o])
__global__ void dyn_aggregate_kernel(const float* __restrict__ attn, int n, int K, const float* __restrict__ bank,
                                     int O, int I, int kk, int transposed, __nv_bfloat16* __restrict__ out) {
  const long long per = (long long)O * I * kk;
  const long long total = per * n;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(idx / per);
    const long long e = idx - (long long)b * per;
    int o, i, t;
    if (!transposed) { i = (int)(e % I); long long r = e / I; t = (int)(r % kk); o = (int)(r / kk); }
    else { o = (int)(e % O); long long r = e / O; t = (int)(r % kk); i = (int)(r / kk); }
    const long long src = ((long long)o * I + i) * kk + t;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc += __ldg(attn + b * K + k) * __ldg(bank + (long long)k * per + src);
    out[idx] = __float2bfloat16(acc);
  }
}
// Tiled form for k*k <= 9: a block stages the K experts of a 16 (cout) x 16 (cin) x k*k tile of the bank in shared
// memory with coalesced reads — ONCE, not once per sample — and then writes that tile of every sample's aggregated
// kernel in 32-byte runs of the packed layout (cin fastest, or cout fastest for the transposed pack).  The one-thread-
// per-output kernel above re-read the bank n times with a 36-byte stride (2 ms per DySOEM step).
constexpr int kAggT = 16;
__global__ void __launch_bounds__(256)
dyn_aggregate_tiled_kernel(const float* __restrict__ attn, int n, int K, const float* __restrict__ bank, int O, int I,
                           int kk, int transposed, __nv_bfloat16* __restrict__ out) {
  extern __shared__ float agg_sm[];
  const int row = kAggT * kk + 1;                    // floats per (k, oo) row of the tile, padded against bank conflicts
  float* tile = agg_sm;                              // [K][16 oo][16 ii][kk] (+1 pad per oo)
  float* sattn = agg_sm + (size_t)K * kAggT * row;   // [n][K]
  const int o0 = blockIdx.y * kAggT, i0 = blockIdx.x * kAggT;
  const long long per = (long long)O * I * kk;
  for (int idx = threadIdx.x; idx < K * kAggT * kAggT * kk; idx += 256) {
    const int k = idx / (kAggT * kAggT * kk), r = idx - k * (kAggT * kAggT * kk);
    const int oo = r / (kAggT * kk), rr = r - oo * (kAggT * kk);
    const int ii = rr / kk;
    float v = 0.f;
    if (o0 + oo < O && i0 + ii < I) v = __ldg(bank + (long long)k * per + ((long long)(o0 + oo) * I + i0) * kk + rr);
    tile[(k * kAggT + oo) * row + rr] = v;
  }
  for (int idx = threadIdx.x; idx < n * K; idx += 256) sattn[idx] = attn[idx];
  __syncthreads();
  // this thread's (at most 9) elements of the tile: shared-memory offset and offset inside one sample's output
  int soff[9];
  long long doff[9];
#pragma unroll
  for (int j = 0; j < 9; ++j) {
    const int e = threadIdx.x + 256 * j;
    soff[j] = -1;
    doff[j] = 0;
    if (e < kAggT * kAggT * kk) {
      int oo, ii, t;
      if (!transposed) { ii = e % kAggT; t = (e / kAggT) % kk; oo = e / (kAggT * kk); }
      else { oo = e % kAggT; t = (e / kAggT) % kk; ii = e / (kAggT * kk); }
      if (o0 + oo < O && i0 + ii < I) {
        soff[j] = oo * row + ii * kk + t;
        doff[j] = !transposed ? ((long long)(o0 + oo) * kk + t) * I + i0 + ii : ((long long)(i0 + ii) * kk + t) * O + o0 + oo;
      }
    }
  }
  for (int b = 0; b < n; ++b) {
    __nv_bfloat16* ob = out + (long long)b * per;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      if (soff[j] >= 0) {
        float acc = 0.f;
        for (int k = 0; k < K; ++k) acc += sattn[b * K + k] * tile[k * kAggT * row + soff[j]];
        ob[doff[j]] = __float2bfloat16(acc);
      }
    }
  }
}

// cin <= 3 stem sites that run as im2col + 1x1 GEMM: out[b][o][32] = sum_k attn[b][k] * bank[k][o][j] for the
// j < I*kk taps of the OIHW-flattened kernel (the im2col channel order), zero for the padding columns.
__global__ void dyn_aggregate_stem_kernel(const float* __restrict__ attn, int n, int K, const float* __restrict__ bank,
                                          int O, int taps, __nv_bfloat16* __restrict__ out) {
  const long long total = (long long)n * O * 32;
  const long long per = (long long)O * taps;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx & 31);
    const long long bo = idx >> 5;
    const int o = (int)(bo % O), b = (int)(bo / O);
    float acc = 0.f;
    if (j < taps)
      for (int k = 0; k < K; ++k) acc += __ldg(attn + b * K + k) * __ldg(bank + (long long)k * per + (long long)o * taps + j);
    out[idx] = __float2bfloat16(acc);
  }
}
__global__ void dyn_bias_kernel(const float* __restrict__ attn, int n, int K, const float* __restrict__ bias_bank,
                                int O, float* __restrict__ bias_out) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * O) return;
  int b = idx / O, o = idx - b * O;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) acc += attn[b * K + k] * bias_bank[k * O + o];
  bias_out[idx] = acc;
}

// ---- backward of the per-sample kernel aggregation ---------------------------------------------------
// thread <-> up to E kernel elements (OIHW index e); loops over the batch once: d_bank accumulates in
// registers, the d_attn partial products are folded per warp (shuffles) and parked in shared memory
// [warp][b][k]; one global reduction per (b, k) per block at the end.
constexpr int kDbcThreads = 256;
constexpr int kDbcE = 4;
constexpr int kDbcMaxK = 8;
__global__ void __launch_bounds__(kDbcThreads)
dyn_bwd_contract_kernel(const float* __restrict__ dwb, int n, int K, const float* __restrict__ attn,
                        const float* __restrict__ bank, int O, int I, int kk, int packed,
                        float* __restrict__ d_bank, float* __restrict__ d_attn) {
  extern __shared__ float sda[];   // [warps][n][K]
  const long long per = (long long)O * I * kk;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long e[kDbcE], pidx[kDbcE];
  float bk[kDbcE][kDbcMaxK], acc[kDbcE][kDbcMaxK];
#pragma unroll
  for (int j = 0; j < kDbcE; ++j) {
    e[j] = ((long long)blockIdx.x * kDbcE + j) * kDbcThreads + threadIdx.x;
    const bool live = e[j] < per;
    pidx[j] = -1;
    if (live) {
      if (packed) {
        const int t = (int)(e[j] % kk);
        const long long r = e[j] / kk;
        const int i = (int)(r % I);
        const long long o = r / I;
        pidx[j] = (o * kk + t) * I + i;
      } else {
        pidx[j] = e[j];
      }
    }
#pragma unroll
    for (int k = 0; k < kDbcMaxK; ++k) {
      bk[j][k] = (live && k < K) ? __ldg(bank + (long long)k * per + e[j]) : 0.f;
      acc[j][k] = 0.f;
    }
  }
  for (int b = 0; b < n; ++b) {
    float g[kDbcE];
#pragma unroll
    for (int j = 0; j < kDbcE; ++j) g[j] = pidx[j] >= 0 ? __ldg(dwb + (long long)b * per + pidx[j]) : 0.f;
#pragma unroll
    for (int k = 0; k < kDbcMaxK; ++k) {
      if (k < K) {
        const float a = __ldg(attn + b * K + k);
        float da = 0.f;
#pragma unroll
        for (int j = 0; j < kDbcE; ++j) {
          acc[j][k] = fmaf(a, g[j], acc[j][k]);
          da = fmaf(g[j], bk[j][k], da);
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) da += __shfl_xor_sync(0xffffffffu, da, off);
        if (lane == 0) sda[(warp * n + b) * K + k] = da;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < kDbcE; ++j) {
    if (pidx[j] >= 0) {
#pragma unroll
      for (int k = 0; k < kDbcMaxK; ++k)
        if (k < K) d_bank[(long long)k * per + e[j]] += acc[j][k];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n * K; i += kDbcThreads) {
    float s = 0.f;
    for (int w = 0; w < kDbcThreads / 32; ++w) s += sda[w * n * K + i];
    atomicAdd(d_attn + i, s);
  }
}

// ---- global average pool -------------------------------------------------------------------------
// out[b][q*C + c] += sum over pixels of parity class q (s2d) / all pixels (q = 0), scaled by inv_count
__global__ void gap_kernel(const __nv_bfloat16* __restrict__ x, int ld, int h, int w, int c, int s2d,
                           float inv_count, float* __restrict__ out) {
  __shared__ float red[256 * 8];
  const int G = c >> 3;
  const int Gb = G < 32 ? G : 32;
  const int PL = 256 / Gb;
  const int g = threadIdx.x % Gb + blockIdx.y * Gb;
  const int pl = threadIdx.x / Gb;
  const int b = blockIdx.z;
  const int nq = s2d ? 4 : 1;
  const long long hw = (long long)h * w;
  for (int q = 0; q < nq; ++q) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    if (g < G && pl < PL) {
      // pixels of class q, four independent 16-byte loads in flight per thread (the first version kept one load in
      // flight and walked ALL pixels once per class with a 64-bit divide each: 38 % / 15 % of the HBM rate)
      const __nv_bfloat16* xb_ = x + (long long)b * hw * ld + (g << 3);
      const int w2 = w >> 1;
      const int total = s2d ? (h >> 1) * w2 : (int)hw;
      const int qy = q >> 1, qx = q & 1;
      auto pixel = [&](int i) -> long long {
        if (!s2d) return i;
        const int yy = i / w2, xx = i - yy * w2;
        return (long long)(2 * yy + qy) * w + 2 * xx + qx;
      };
      int i = blockIdx.x * PL + pl;
      const int istep = gridDim.x * PL;
      for (; i + 3 * istep < total; i += 4 * istep) {
        uint4 u[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) u[k] = __ldcs(reinterpret_cast<const uint4*>(xb_ + pixel(i + k * istep) * ld));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float v[8];
          unpack8(u[k], v);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += v[j];
        }
      }
      for (; i < total; i += istep) {
        float v[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(xb_ + pixel(i) * ld)), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[j];
      }
    }
    float* mine = red + threadIdx.x * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) mine[j] = acc[j];
    __syncthreads();
    for (int t = threadIdx.x; t < Gb * 8; t += 256) {
      const int gg = t >> 3, slot = t & 7;
      float s = 0.f;
      for (int l = 0; l < PL; ++l) s += red[(l * Gb + gg) * 8 + slot];
      const int gch = gg + blockIdx.y * Gb;
      if (gch < G) atomicAdd(out + (long long)b * nq * c + q * c + (gch << 3) + slot, s * inv_count);
    }
    __syncthreads();
  }
}
__global__ void gap_nchw_kernel(const float* __restrict__ x, int hw, float* __restrict__ out) {
  // blockIdx.x = (n, c) plane, blockIdx.y = slice of the plane (a plane per block left 96 blocks for a 3-channel batch
  // of 32: 0.7 ms for 157 MB); float4 loads, partial sums combined with one atomic per block (out is zeroed first)
  const float* p = x + (long long)blockIdx.x * hw;
  float s = 0.f;
  const int hw4 = (hw & 3) == 0 ? hw >> 2 : 0;
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < hw4; i += gridDim.y * blockDim.x) {
    const float4 v = __ldcs(reinterpret_cast<const float4*>(p) + i);
    s += (v.x + v.y) + (v.z + v.w);
  }
  for (int i = 4 * hw4 + blockIdx.y * blockDim.x + threadIdx.x; i < hw; i += gridDim.y * blockDim.x) s += p[i];
  __shared__ float red[32];
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) atomicAdd(out + blockIdx.x, s / (float)hw);
  }
}

// ---- attention MLP + softmax (one block per sample) ---------------------------------------------------
__global__ void attn_mlp_softmax_kernel(const float* __restrict__ pooled, int c, const float* __restrict__ w1,
                                        const float* __restrict__ b1, int hid, const float* __restrict__ w2,
                                        const float* __restrict__ b2, int K, float inv_t, float* __restrict__ attn,
                                        float* __restrict__ hidden) {
  extern __shared__ float sh[];  // [c] pooled, [hid] hidden, [K] logits
  float* sp = sh; float* shid = sh + c; float* slog = shid + hid;
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < c; i += blockDim.x) sp[i] = pooled[(long long)b * c + i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int j = warp; j < hid; j += nw) {
    float s = 0.f;
    for (int i = lane; i < c; i += 32) s += sp[i] * __ldg(w1 + (long long)j * c + i);
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
      s += b1 ? b1[j] : 0.f;
      s = s > 0.f ? s : 0.f;
      shid[j] = s;
      if (hidden) hidden[(long long)b * hid + j] = s;
    }
  }
  __syncthreads();
  for (int k = warp; k < K; k += nw) {
    float s = 0.f;
    for (int i = lane; i < hid; i += 32) s += shid[i] * __ldg(w2 + (long long)k * hid + i);
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) slog[k] = (s + (b2 ? b2[k] : 0.f)) * inv_t;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float mx = -INFINITY;
    for (int k = 0; k < K; ++k) mx = fmaxf(mx, slog[k]);
    float den = 0.f;
    for (int k = 0; k < K; ++k) { float e = expf(slog[k] - mx); slog[k] = e; den += e; }
    for (int k = 0; k < K; ++k) attn[(long long)b * K + k] = slog[k] / den;
  }
}

// ---- detection-head gradients -> the A operand of the head's data / weight gradient GEMMs ------------------------------
// d_obj (n,A,H,W,1), d_bbox (n,A,H,W,4) fp32 (the layouts YOLOHead returns, _base.py:91-94,112-115) -> dyh (n,H,W,32)
// bf16 with channels [A obj | 4A bbox | zero pad], and the bias gradients sum_p d_obj / d_bbox (fp32, accumulated).
// One thread per pixel: 5A coalesced fp32 reads, one 64-byte row written.
__global__ void __launch_bounds__(256)
head_grad_pack_kernel(const float* __restrict__ d_obj, const float4* __restrict__ d_bbox, int n, int A, long long hw,
                      __nv_bfloat16* __restrict__ dyh, long long dyh_ld, float* __restrict__ gb_obj,
                      float* __restrict__ gb_bbox) {
  __shared__ float red[8][16];
  const long long total = (long long)n * hw;
  float bs[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) bs[j] = 0.f;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
    const long long img = p / hw, px = p - img * hw;
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      if (a < A) {
        if (d_obj) v[a] = __ldcs(d_obj + (img * A + a) * hw + px);
        if (d_bbox) {
          const float4 t = __ldcs(d_bbox + (img * A + a) * hw + px);
          v[A + 4 * a + 0] = t.x; v[A + 4 * a + 1] = t.y; v[A + 4 * a + 2] = t.z; v[A + 4 * a + 3] = t.w;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) bs[j] += v[j];
    uint4* dst = reinterpret_cast<uint4*>(dyh + p * dyh_ld);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 o;
      o.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]); o.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
      o.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]); o.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
      dst[j] = o;
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float s = bs[j];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) red[warp][j] = s;
  }
  __syncthreads();
  if (threadIdx.x < 5 * A) {
    float s = 0.f;
    for (int wi = 0; wi < 8; ++wi) s += red[wi][threadIdx.x];
    if (threadIdx.x < A) { if (gb_obj) atomicAdd(gb_obj + threadIdx.x, s); }
    else if (gb_bbox) atomicAdd(gb_bbox + (threadIdx.x - A), s);
  }
}

// ---- backward of the attention MLP + softmax(./T) ------------------------------------------------------------------
// step 1, one block per sample: g = a * (d_a - <a, d_a>) / T (softmax backward), dh = (g @ w2) * [hidden > 0]
__global__ void attn_bwd_dh_kernel(const float* __restrict__ attn, const float* __restrict__ d_attn,
                                   const float* __restrict__ hidden, const float* __restrict__ w2, int hid, int K,
                                   float inv_t, float* __restrict__ g_out, float* __restrict__ dh_out) {
  __shared__ float sg[32];
  const int b = blockIdx.x;
  if (threadIdx.x == 0) {
    float dot = 0.f;
    for (int k = 0; k < K; ++k) dot += attn[b * K + k] * d_attn[b * K + k];
    for (int k = 0; k < K; ++k) {
      const float g = attn[b * K + k] * (d_attn[b * K + k] - dot) * inv_t;
      sg[k] = g;
      g_out[b * K + k] = g;
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < hid; j += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < K; ++k) s += sg[k] * __ldg(w2 + (long long)k * hid + j);
    dh_out[(long long)b * hid + j] = hidden[(long long)b * hid + j] > 0.f ? s : 0.f;
  }
}
// step 2: dW1[j][c] += sum_b dh[b][j] pooled[b][c];  d_pooled[b][c] = out_scale * sum_j dh[b][j] w1[j][c];
//         block 0 also: dW2[k][j] += sum_b g[b][k] hidden[b][j], db2[k] += sum_b g[b][k], db1[j] += sum_b dh[b][j]
__global__ void attn_bwd_params_kernel(const float* __restrict__ g, const float* __restrict__ dh,
                                       const float* __restrict__ hidden, const float* __restrict__ pooled,
                                       const float* __restrict__ w1, int n, int c, int hid, int K, float out_scale,
                                       float* __restrict__ dw1, float* __restrict__ db1, float* __restrict__ dw2,
                                       float* __restrict__ db2, float* __restrict__ d_pooled) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long t0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (dw1) {
    for (long long e = t0; e < (long long)hid * c; e += stride) {
      const int j = (int)(e / c), cc = (int)(e - (long long)j * c);
      float acc = 0.f;
      for (int b = 0; b < n; ++b) acc += __ldg(dh + (long long)b * hid + j) * __ldg(pooled + (long long)b * c + cc);
      dw1[e] += acc;
    }
  }
  if (d_pooled) {
    for (long long e = t0; e < (long long)n * c; e += stride) {
      const int b = (int)(e / c), cc = (int)(e - (long long)b * c);
      float acc = 0.f;
      for (int j = 0; j < hid; ++j) acc += __ldg(dh + (long long)b * hid + j) * __ldg(w1 + (long long)j * c + cc);
      d_pooled[e] = acc * out_scale;
    }
  }
  if (blockIdx.x == 0) {
    for (int e = threadIdx.x; e < K * hid; e += blockDim.x) {
      const int k = e / hid, j = e - k * hid;
      float acc = 0.f;
      for (int b = 0; b < n; ++b) acc += g[b * K + k] * hidden[(long long)b * hid + j];
      if (dw2) dw2[e] += acc;
    }
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      float acc = 0.f;
      for (int b = 0; b < n; ++b) acc += g[b * K + k];
      if (db2) db2[k] += acc;
    }
    if (db1)
      for (int j = threadIdx.x; j < hid; j += blockDim.x) {
        float acc = 0.f;
        for (int b = 0; b < n; ++b) acc += dh[(long long)b * hid + j];
        db1[j] += acc;
      }
  }
}

// ---- SGD ------------------------------------------------------------------------------------------
__global__ void sgd_momentum_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                    long long count, float lr, float momentum, float grad_scale, int first,
                                    const float* __restrict__ hyper) {
  if (hyper) {      // learning rate / momentum / gradient scale live on the device (CUDA-graph replay, schedulers)
    lr = __ldg(hyper);
    momentum = __ldg(hyper + 1);
    grad_scale = __ldg(hyper + 2);
  }
  const long long n4 = count >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 gp = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 pp = reinterpret_cast<float4*>(p)[i];
    float4 mm = first ? make_float4(0, 0, 0, 0) : reinterpret_cast<float4*>(m)[i];
    gp.x *= grad_scale; gp.y *= grad_scale; gp.z *= grad_scale; gp.w *= grad_scale;
    mm.x = first ? gp.x : momentum * mm.x + gp.x; mm.y = first ? gp.y : momentum * mm.y + gp.y;
    mm.z = first ? gp.z : momentum * mm.z + gp.z; mm.w = first ? gp.w : momentum * mm.w + gp.w;
    pp.x -= lr * mm.x; pp.y -= lr * mm.y; pp.z -= lr * mm.z; pp.w -= lr * mm.w;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(p)[i] = pp;
  }
  // tail
  for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x) {
    float gg = g[i] * grad_scale;
    float mm = first ? gg : momentum * m[i] + gg;
    m[i] = mm;
    p[i] -= lr * mm;
  }
}

// ---- bias bank of a dynamic conv (DynamicSOEM, DySOEM_SimFPN.py:56-60): bias[b] = attn[b] @ bias_bank ---------------
// d_bias_bank[kk][o] = sum_b attn[b][kk] * g[b][o];  d_attn[b][kk] += sum_o g[b][o] * bias_bank[kk][o], where
// g = scale * pooled_grad is the per-sample channel sum of the output gradient (pool kernel x pixel count).  n <= 128,
// K <= 8, O <= 1024: a few hundred thousand MACs, one thread per output value (the two torch matmuls this replaces were
// the last cuBLAS launches of the DySOEM_SimFPN step).
__global__ void dyn_bias_bwd_kernel(const float* __restrict__ g, float scale, int n, int K, int O,
                                    const float* __restrict__ attn, const float* __restrict__ bias_bank,
                                    float* __restrict__ d_bias_bank, float* __restrict__ d_attn) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < K * O) {
    const int kk = i / O, o = i - kk * O;
    float acc = 0.f;
    for (int b = 0; b < n; ++b) acc = fmaf(attn[b * K + kk], g[(size_t)b * O + o], acc);
    d_bias_bank[i] = acc * scale;
  } else if (i < K * O + n * K) {
    const int j = i - K * O;
    const int b = j / K, kk = j - b * K;
    float acc = 0.f;
    for (int o = 0; o < O; ++o) acc = fmaf(g[(size_t)b * O + o], bias_bank[(size_t)kk * O + o], acc);
    d_attn[j] += acc * scale;
  }
}

}  // namespace uavdet

using namespace uavdet;
#define ST ((cudaStream_t)stream)

extern "C" int uavdet_bn_finalize(const float* sum, const float* sumsq, int c, double count, float eps,
                                  float momentum, const float* gamma, const float* beta, float* running_mean,
                                  float* running_var, float* mean, float* invstd, float* scale, float* shift,
                                  void* stream) {
  UAVDET_CHECK_ARG(sum && sumsq && scale && shift && c > 0 && count > 0, "bn_finalize: bad arguments");
  bn_finalize_kernel<<<ceil_div(c, 128), 128, 0, ST>>>(sum, sumsq, c, count, eps, momentum, gamma, beta,
                                                       running_mean, running_var, mean, invstd, scale, shift);
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}

extern "C" int uavdet_bn_act_fwd(const uavdet_act* raw, const float* scale, const float* shift, int act,
                                 const uavdet_act* res, const uavdet_act* y, void* stream) {
  int rc;
  if ((rc = check_view(raw, "bn_act_fwd raw")) || (rc = check_view(y, "bn_act_fwd y")) ||
      (rc = same_shape(raw, y, "bn_act_fwd")))
    return rc;
  if (res && ((rc = check_view(res, "bn_act_fwd res")) || (rc = same_shape(raw, res, "bn_act_fwd res")))) return rc;
  View r = mkview(raw), o = mkview(y);
  if (r.npix == 0) return UAVDET_OK;
  dim3 grid = stream_grid(raw, 8);
  static const int reverse = getenv("UAVDET_BN_FWD_REVERSE") ? atoi(getenv("UAVDET_BN_FWD_REVERSE")) : 0;
  if (res)
    launch_pdl(kPdlBnFwd, bn_act_fwd_kernel<true>, grid, dim3(256), 0, ST, r, scale, shift, act, (const __nv_bfloat16*)res->ptr, res->ld, reverse, o);
  else
    launch_pdl(kPdlBnFwd, bn_act_fwd_kernel<false>, grid, dim3(256), 0, ST, r, scale, shift, act, (const __nv_bfloat16*)nullptr, 0, reverse, o);
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}


// The BatchNorm-backward kernels run concurrently with the weight-gradient kernel of the previous layer (another
// stream).  Two kernels share an SM only if they agree on its L1 / shared-memory split, so these streaming kernels
// ask for the maximum-shared-memory carveout the tensor-core kernels use (they do not need the L1).
static void prefer_max_smem_carveout_once() {
  static PerDeviceOnce once;
  once.run([] {
    cudaFuncSetAttribute(bn_bwd_reduce_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(bn_bwd_reduce_kernel<3>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(bn_bwd_apply_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(bn_bwd_apply_fused_kernel<4, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(bn_bwd_apply_fused_kernel<4, 3>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(bn_bwd_apply_fused_kernel<2, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(bn_bwd_apply_smem_kernel<4, 3>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(bn_bwd_apply_smem_kernel<8, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    return cudaFuncSetAttribute(bn_bwd_apply_fused_kernel<2, 3>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                cudaSharedmemCarveoutMaxShared);
  });
}

extern "C" int uavdet_bn_act_bwd_reduce(const uavdet_act* dy, const uavdet_act* raw, const float* scale,
                                        const float* shift, int act, float* sum_dz, float* sum_dzr,
                                        void* stream) {
  int rc;
  if ((rc = check_view(dy, "bn_bwd_reduce dy")) || (rc = check_view(raw, "bn_bwd_reduce raw")) ||
      (rc = same_shape(dy, raw, "bn_bwd_reduce")))
    return rc;
  UAVDET_CHECK_ARG(scale && shift && sum_dz && sum_dzr, "bn_bwd_reduce: null stats");
  prefer_max_smem_carveout_once();
  // few long-lived blocks: every block ends with one atomic per channel sum, and thousands of blocks hammering
  // the same 2*c addresses serialise in L2 (the kernel ran at 1.5 TB/s with 1,600 blocks)
  static const int bps = getenv("UAVDET_BN_REDUCE_BPS") ? atoi(getenv("UAVDET_BN_REDUCE_BPS")) : 2;
  static const int minb = getenv("UAVDET_BN_REDUCE_MINB") ? atoi(getenv("UAVDET_BN_REDUCE_MINB")) : 3;   // A/B switch
  if (minb >= 3)
    launch_pdl(kPdlBnReduce, bn_bwd_reduce_kernel<3>, stream_grid(dy, 16, bps), dim3(256), 0, ST, mkview(dy), mkview(raw), scale, shift, act,
               sum_dz, sum_dzr);
  else
    launch_pdl(kPdlBnReduce, bn_bwd_reduce_kernel<1>, stream_grid(dy, 16, bps), dim3(256), 0, ST, mkview(dy), mkview(raw), scale, shift, act,
               sum_dz, sum_dzr);
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}

extern "C" int uavdet_bn_bwd_finalize(const float* sum_dz, const float* sum_dzr, const float* mean,
                                      const float* invstd, const float* scale, int c, double count, float* dgamma,
                                      float* dbeta, float* k1, float* k0, void* stream) {
  UAVDET_CHECK_ARG(sum_dz && sum_dzr && mean && invstd && scale && dgamma && dbeta && k1 && k0 && c > 0 && count > 0,
                   "bn_bwd_finalize: bad arguments");
  bn_bwd_finalize_kernel<<<ceil_div(c, 128), 128, 0, ST>>>(sum_dz, sum_dzr, mean, invstd, scale, c,
                                                           (float)(1.0 / count), dgamma, dbeta, k1, k0);
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}

extern "C" int uavdet_bn_act_bwd_apply(const uavdet_act* dy, const uavdet_act* raw, const float* scale,
                                       const float* shift, const float* k1, const float* k0, int act,
                                       const uavdet_act* d_raw, void* stream) {
  int rc;
  if ((rc = check_view(dy, "bn_bwd_apply dy")) || (rc = check_view(raw, "bn_bwd_apply raw")) ||
      (rc = check_view(d_raw, "bn_bwd_apply d_raw")) || (rc = same_shape(dy, raw, "bn_bwd_apply")) ||
      (rc = same_shape(dy, d_raw, "bn_bwd_apply")))
    return rc;
  UAVDET_CHECK_ARG(scale && shift && k1 && k0, "bn_bwd_apply: null coefficients");
  bn_bwd_apply_kernel<<<stream_grid(dy, 8), 256, 0, ST>>>(mkview(dy), mkview(raw), scale, shift, k1, k0, act,
                                                          mkview(d_raw));
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}

extern "C" int uavdet_bn_train_fwd(const uavdet_act* raw, const float* sum, const float* sumsq, double count, float eps,
                                   float momentum, const float* gamma, const float* beta, float* running_mean,
                                   float* running_var, float* mean, float* invstd, float* scale, float* shift, int act,
                                   const uavdet_act* res, const uavdet_act* y, void* stream) {
  int rc;
  if ((rc = check_view(raw, "bn_train_fwd raw")) || (rc = check_view(y, "bn_train_fwd y")) ||
      (rc = same_shape(raw, y, "bn_train_fwd")))
    return rc;
  UAVDET_CHECK_ARG(sum && sumsq && scale && shift && count > 0, "bn_train_fwd: sums / outputs missing");
  if ((long long)raw->n * raw->h * raw->w == 0) return UAVDET_OK;
  dim3 grid = stream_grid(raw, 8);
  if (res) {
    if ((rc = check_view(res, "bn_train_fwd res")) || (rc = same_shape(raw, res, "bn_train_fwd res"))) return rc;
    launch_pdl(kPdlBnFwd, bn_train_fwd_kernel<true>, grid, dim3(256), 0, ST, mkview(raw), sum, sumsq, count, eps, momentum, gamma, beta,
               running_mean, running_var, mean, invstd, scale, shift, act, (const __nv_bfloat16*)res->ptr, res->ld,
               mkview(y));
  } else {
    launch_pdl(kPdlBnFwd, bn_train_fwd_kernel<false>, grid, dim3(256), 0, ST, mkview(raw), sum, sumsq, count, eps, momentum, gamma, beta,
               running_mean, running_var, mean, invstd, scale, shift, act, (const __nv_bfloat16*)nullptr, 0, mkview(y));
  }
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}

extern "C" int uavdet_bn_act_bwd_apply_fused(const uavdet_act* dy, const uavdet_act* raw, const float* scale,
                                             const float* shift, const float* sum_dz, const float* sum_dzr,
                                             const float* mean, const float* invstd, double count, int act,
                                             float* dgamma, float* dbeta, int accumulate, const uavdet_act* d_raw,
                                             void* stream) {
  int rc;
  if ((rc = check_view(dy, "bn_bwd_apply_fused dy")) || (rc = check_view(raw, "bn_bwd_apply_fused raw")) ||
      (rc = check_view(d_raw, "bn_bwd_apply_fused d_raw")) || (rc = same_shape(dy, raw, "bn_bwd_apply_fused")) ||
      (rc = same_shape(dy, d_raw, "bn_bwd_apply_fused")))
    return rc;
  UAVDET_CHECK_ARG(scale && shift && sum_dz && sum_dzr && mean && invstd && dgamma && dbeta && count > 0,
                   "bn_bwd_apply_fused: null argument");
  // pixels per thread: every thread first derives the coefficients of its 8 channels from six per-channel vectors,
  // so very short threads spend more on that prologue than on their data
  static const int ppt = getenv("UAVDET_BN_APPLY_PPT") ? atoi(getenv("UAVDET_BN_APPLY_PPT")) : 4;
  // register budget / unroll variants (occupancy against loads in flight per thread): A/B switch
  static const int variant = getenv("UAVDET_BN_APPLY_VARIANT") ? atoi(getenv("UAVDET_BN_APPLY_VARIANT")) : 3;
  static const int reverse = getenv("UAVDET_BN_APPLY_REVERSE") ? atoi(getenv("UAVDET_BN_APPLY_REVERSE")) : 0;
  prefer_max_smem_carveout_once();
#define UAVDET_BN_APPLY(U, MINB, BPS)                                                                                       \
  launch_pdl(kPdlBnApply, bn_bwd_apply_fused_kernel<U, MINB>, stream_grid(dy, ppt, BPS), dim3(256), 0, ST,                                 \
             mkview(dy), mkview(raw), scale, shift, sum_dz, sum_dzr, mean, invstd, (float)(1.0 / count), act, dgamma, dbeta, \
             accumulate, reverse, mkview(d_raw))
  if (variant == 5 || variant == 6) {
    // coefficients in shared memory (see bn_bwd_apply_smem_kernel)
    const dim3 grid = stream_grid(dy, ppt * 2, 24);
    if (variant == 5)
      launch_pdl(kPdlBnApply, bn_bwd_apply_smem_kernel<4, 3>, grid, dim3(256), 0, ST, mkview(dy), mkview(raw), scale, shift, sum_dz,
                 sum_dzr, mean, invstd, (float)(1.0 / count), act, dgamma, dbeta, accumulate, mkview(d_raw));
    else
      launch_pdl(kPdlBnApply, bn_bwd_apply_smem_kernel<8, 2>, grid, dim3(256), 0, ST, mkview(dy), mkview(raw), scale, shift, sum_dz,
                 sum_dzr, mean, invstd, (float)(1.0 / count), act, dgamma, dbeta, accumulate, mkview(d_raw));
    UAVDET_LAUNCH_CHECK();
    return UAVDET_OK;
  }
  switch (variant) {
    case 1: UAVDET_BN_APPLY(4, 3, 24); break;
    case 2: UAVDET_BN_APPLY(2, 4, 32); break;
    case 3: UAVDET_BN_APPLY(2, 3, 24); break;
    default: UAVDET_BN_APPLY(4, 1, 16); break;
  }
#undef UAVDET_BN_APPLY
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}

extern "C" int uavdet_act_bwd(const uavdet_act* dy, const uavdet_act* raw, const float* scale,
                              const float* shift, int act, const uavdet_act* dx, void* stream) {
  int rc;
  if ((rc = check_view(dy, "act_bwd dy")) || (rc = check_view(raw, "act_bwd raw")) ||
      (rc = check_view(dx, "act_bwd dx")) || (rc = same_shape(dy, raw, "act_bwd")) ||
      (rc = same_shape(dy, dx, "act_bwd")))
    return rc;
  View d = mkview(dy);
  act_bwd_kernel<<<ew_grid(d.npix * (d.c / 8), 256), 256, 0, ST>>>(d, mkview(raw), scale, shift, act, mkview(dx));
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}

extern "C" int uavdet_upsample2x_fwd(const uavdet_act* x, const uavdet_act* y, void* stream) {
  int rc;
  if ((rc = check_view(x, "upsample2x_fwd x")) || (rc = check_view(y, "upsample2x_fwd y"))) return rc;
  UAVDET_CHECK_ARG(y->n == x->n && y->h == 2 * x->h && y->w == 2 * x->w && y->c == x->c, "upsample2x_fwd: shapes");
  long long total = (long long)y->n * y->h * y->w * (y->c / 8);
  upsample2x_fwd_kernel<<<ew_grid(total, 256), 256, 0, ST>>>((const __nv_bfloat16*)x->ptr, x->ld, x->n, x->h, x->w,
                                                            x->c, (__nv_bfloat16*)y->ptr, y->ld);
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}

extern "C" int uavdet_upsample2x_bwd(const uavdet_act* dy, const uavdet_act* dx, int accumulate, void* stream) {
  int rc;
  if ((rc = check_view(dy, "upsample2x_bwd dy")) || (rc = check_view(dx, "upsample2x_bwd dx"))) return rc;
  UAVDET_CHECK_ARG(dy->n == dx->n && dy->h == 2 * dx->h && dy->w == 2 * dx->w && dy->c == dx->c, "upsample2x_bwd: shapes");
  long long total = (long long)dx->n * dx->h * dx->w * (dx->c / 8);
  upsample2x_bwd_kernel<<<ew_grid(total, 256), 256, 0, ST>>>((const __nv_bfloat16*)dy->ptr, dy->ld, dx->n, dx->h,
                                                            dx->w, dx->c, (__nv_bfloat16*)dx->ptr, dx->ld, accumulate);
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}

extern "C" int uavdet_upsample2x_add(const uavdet_act* b_low, const uavdet_act* a, float a_mult,
                                     const uavdet_act* y, void* stream) {
  int rc;
  if ((rc = check_view(b_low, "upsample2x_add b")) || (rc = check_view(a, "upsample2x_add a")) ||
      (rc = check_view(y, "upsample2x_add y")) || (rc = same_shape(a, y, "upsample2x_add")))
    return rc;
  UAVDET_CHECK_ARG(a->n == b_low->n && a->h == 2 * b_low->h && a->w == 2 * b_low->w && a->c == b_low->c,
                   "upsample2x_add: shapes");
  long long total = (long long)a->n * a->h * a->w * (a->c / 8);
  upsample2x_add_kernel<<<ew_grid(total, 256), 256, 0, ST>>>((const __nv_bfloat16*)b_low->ptr, b_low->ld, b_low->n,
                                                            b_low->h, b_low->w, b_low->c, (const __nv_bfloat16*)a->ptr,
                                                            a->ld, a_mult, (__nv_bfloat16*)y->ptr, y->ld);
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}

extern "C" int uavdet_add(const uavdet_act* a, const uavdet_act* b, const uavdet_act* y, void* stream) {
  int rc;
  if ((rc = check_view(a, "add a")) || (rc = check_view(y, "add y")) || (rc = same_shape(a, y, "add"))) return rc;
  if (b && ((rc = check_view(b, "add b")) || (rc = same_shape(a, b, "add b")))) return rc;
  View va = mkview(a);
  add_kernel<<<ew_grid(va.npix * (va.c / 8), 256), 256, 0, ST>>>(va, b ? (const __nv_bfloat16*)b->ptr : nullptr,
                                                                 b ? b->ld : 0, mkview(y));
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}

extern "C" int uavdet_nhwc_to_nchw_f32(const uavdet_act* x, float* y_nchw, void* stream) {
  UAVDET_CHECK_ARG(x && x->ptr && y_nchw, "nhwc_to_nchw: null");
  int hw = x->h * x->w;
  dim3 grid(ceil_div(hw, 32), ceil_div(x->c, 32), x->n), block(32, 8);
  nhwc_to_nchw_kernel<<<grid, block, 0, ST>>>((const __nv_bfloat16*)x->ptr, x->ld, x->n, hw, x->c, y_nchw);
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}
extern "C" int uavdet_nchw_f32_to_nhwc(const float* x_nchw, const uavdet_act* y, void* stream) {
  UAVDET_CHECK_ARG(y && y->ptr && x_nchw, "nchw_to_nhwc: null");
  int hw = y->h * y->w;
  dim3 grid(ceil_div(hw, 32), ceil_div(y->c, 32), y->n), block(32, 8);
  nchw_to_nhwc_kernel<<<grid, block, 0, ST>>>(x_nchw, y->n, hw, y->c, (__nv_bfloat16*)y->ptr, y->ld);
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}

extern "C" int uavdet_pack_weight(const float* w_oihw, int O, int I, int k, int flags, void* out_bf16,
                                  void* stream) {
  UAVDET_CHECK_ARG(w_oihw && out_bf16 && O > 0 && I > 0 && k > 0, "pack_weight: bad arguments");
  pack_weight_kernel<<<ew_grid((long long)O * I * k * k, 256), 256, 0, ST>>>(w_oihw, O, I, k * k, flags & 1,
                                                                          (flags >> 1) & 1, (__nv_bfloat16*)out_bf16);
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}
extern "C" int uavdet_pack_weights_batched(const uavdet_pack_job* jobs_dev, int n_jobs, long long total_chunks,
                                           void* stream) {
  UAVDET_CHECK_ARG(jobs_dev && n_jobs >= 0 && total_chunks >= 0 && total_chunks < (1ll << 31),
                   "pack_weights_batched: bad arguments");
  if (n_jobs == 0 || total_chunks == 0) return UAVDET_OK;
  pack_weights_batched_kernel<<<(unsigned)total_chunks, 256, 0, ST>>>(jobs_dev, n_jobs);
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}
extern "C" int uavdet_unpack_wgrad(const float* dw_packed, int O, int I, int k, float* grad_oihw, int accumulate,
                                   void* stream) {
  UAVDET_CHECK_ARG(dw_packed && grad_oihw, "unpack_wgrad: null");
  long long total = (long long)O * I * k * k;
  unpack_wgrad_kernel<<<ew_grid(total, 256), 256, 0, ST>>>(dw_packed, O, I, k * k, grad_oihw, accumulate);
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}

extern "C" int uavdet_dyn_aggregate(const float* attn, int n, int K, const float* bank, int O, int I, int k,
                                    int transposed, void* out_bf16, const float* bias_bank, float* bias_out,
                                    void* stream) {
  UAVDET_CHECK_ARG(attn && bank && out_bf16 && n > 0 && K > 0, "dyn_aggregate: bad arguments");
  if (transposed == 2) {      // stem layout: OIHW-flat rows zero-padded to the 32 im2col channels
    UAVDET_CHECK_ARG(I * k * k <= 32, "dyn_aggregate: the stem layout holds at most 32 taps");
    dyn_aggregate_stem_kernel<<<ew_grid((long long)n * O * 32, 256), 256, 0, ST>>>(attn, n, K, bank, O, I * k * k,
                                                                                   (__nv_bfloat16*)out_bf16);
    UAVDET_LAUNCH_CHECK();
    return UAVDET_OK;
  }
  long long total = (long long)n * O * I * k * k;
  const size_t agg_smem = sizeof(float) * ((size_t)K * kAggT * (kAggT * k * k + 1) + (size_t)n * K);
  if (k * k <= 9 && agg_smem <= 48 * 1024) {
    dim3 grid((unsigned)ceil_div(I, kAggT), (unsigned)ceil_div(O, kAggT));
    dyn_aggregate_tiled_kernel<<<grid, 256, agg_smem, ST>>>(attn, n, K, bank, O, I, k * k, transposed ? 1 : 0,
                                                           (__nv_bfloat16*)out_bf16);
  } else {
    dyn_aggregate_kernel<<<ew_grid(total, 256), 256, 0, ST>>>(attn, n, K, bank, O, I, k * k, transposed,
                                                              (__nv_bfloat16*)out_bf16);
  }
  UAVDET_LAUNCH_CHECK();
  if (bias_bank && bias_out) {
    dyn_bias_kernel<<<ceil_div(n * O, 256), 256, 0, ST>>>(attn, n, K, bias_bank, O, bias_out);
    UAVDET_LAUNCH_CHECK();
  }
  return UAVDET_OK;
}

extern "C" int uavdet_dyn_bwd_contract(const float* dwb, int n, int K, const float* attn, const float* bank, int O,
                                       int I, int k, int packed, float* d_bank, float* d_attn, void* stream) {
  UAVDET_CHECK_ARG(dwb && attn && bank && d_bank && d_attn && n > 0 && K > 0 && K <= kDbcMaxK,
                   "dyn_bwd_contract: bad arguments (K <= %d)", kDbcMaxK);
  const long long per = (long long)O * I * k * k;
  const size_t sh = sizeof(float) * (size_t)(kDbcThreads / 32) * n * K;
  UAVDET_CHECK_ARG(sh <= 48 * 1024, "dyn_bwd_contract: batch too large for the shared accumulator");
  const long long blocks = (per + (long long)kDbcThreads * kDbcE - 1) / ((long long)kDbcThreads * kDbcE);
  dyn_bwd_contract_kernel<<<(unsigned)blocks, kDbcThreads, sh, ST>>>(dwb, n, K, attn, bank, O, I, k * k, packed, d_bank,
                                                                    d_attn);
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}

extern "C" int uavdet_dyn_bias_bwd(const float* pooled_grad, float scale, int n, int K, int O, const float* attn,
                                   const float* bias_bank, float* d_bias_bank, float* d_attn, void* stream) {
  UAVDET_CHECK_ARG(pooled_grad && attn && bias_bank && d_bias_bank && d_attn && n > 0 && K > 0 && O > 0,
                   "dyn_bias_bwd: bad arguments");
  const int work = K * O + n * K;
  dyn_bias_bwd_kernel<<<ceil_div(work, 128), 128, 0, ST>>>(pooled_grad, scale, n, K, O, attn, bias_bank, d_bias_bank, d_attn);
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}

extern "C" int uavdet_gap(const uavdet_act* x, int s2d, float* out, void* stream) {
  int rc;
  if ((rc = check_view(x, "gap x"))) return rc;
  UAVDET_CHECK_ARG(out, "gap: null out");
  const int nq = s2d ? 4 : 1;
  UAVDET_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)x->n * nq * x->c, ST));
  int G = x->c / 8, Gb = G < 32 ? G : 32, PL = 256 / Gb;
  long long hw = (long long)x->h * x->w;
  long long bx = (hw + (long long)PL * 8 - 1) / ((long long)PL * 8);
  long long cap = (kNumSMs * 8) / ((long long)(x->n > 0 ? x->n : 1) * ceil_div(G, Gb)) + 1;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  dim3 grid((unsigned)bx, (unsigned)ceil_div(G, Gb), (unsigned)x->n);
  gap_kernel<<<grid, 256, 0, ST>>>((const __nv_bfloat16*)x->ptr, x->ld, x->h, x->w, x->c, s2d,
                                   (float)(nq / (double)hw), out);
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}
extern "C" int uavdet_gap_nchw(const float* x_nchw, int n, int c, int hw, float* out, void* stream) {
  UAVDET_CHECK_ARG(x_nchw && out && n > 0 && c > 0 && hw > 0, "gap_nchw: bad arguments");
  UAVDET_CHECK_ARG(((uintptr_t)x_nchw & 15) == 0, "gap_nchw: input must be 16-byte aligned");
  UAVDET_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)n * c, ST));
  int slices = (kNumSMs * 8) / (n * c) + 1;
  const int max_slices = (hw / 4 + 255) / 256;
  if (slices > max_slices) slices = max_slices;
  if (slices < 1) slices = 1;
  gap_nchw_kernel<<<dim3((unsigned)(n * c), (unsigned)slices), 256, 0, ST>>>(x_nchw, hw, out);
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}

extern "C" int uavdet_attn_mlp_softmax(const float* pooled, int n, int c, const float* w1, const float* b1, int hid,
                                       const float* w2, const float* b2, int K, float temperature, float* attn,
                                       float* hidden, void* stream) {
  UAVDET_CHECK_ARG(pooled && w1 && w2 && attn && n > 0 && c > 0 && hid > 0 && K > 0, "attn_mlp_softmax: bad arguments");
  size_t sh = sizeof(float) * (size_t)(c + hid + K);
  UAVDET_CHECK_ARG(sh <= 48 * 1024, "attn_mlp_softmax: c+hid+K too large");
  attn_mlp_softmax_kernel<<<n, 256, sh, ST>>>(pooled, c, w1, b1, hid, w2, b2, K, 1.f / temperature, attn, hidden);
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}

extern "C" int uavdet_head_grad_pack(const float* d_obj, const float* d_bbox, int n, int anchors, int h, int w,
                                     const uavdet_act* dyh, float* bias_obj_grad, float* bias_bbox_grad, void* stream) {
  int rc;
  if ((rc = check_view(dyh, "head_grad_pack dyh"))) return rc;
  UAVDET_CHECK_ARG((d_obj || d_bbox) && n > 0 && anchors >= 1 && anchors <= 3 && h > 0 && w > 0,
                   "head_grad_pack: bad arguments (anchors <= 3)");
  UAVDET_CHECK_ARG(dyh->n == n && dyh->h == h && dyh->w == w && dyh->c == 32, "head_grad_pack: dyh must be (n,h,w,32)");
  UAVDET_CHECK_ARG(((uintptr_t)d_bbox & 15) == 0, "head_grad_pack: d_bbox must be 16-byte aligned");
  const long long hw = (long long)h * w, total = hw * n;
  head_grad_pack_kernel<<<ew_grid(total, 256), 256, 0, ST>>>(d_obj, (const float4*)d_bbox, n, anchors, hw,
                                                            (__nv_bfloat16*)dyh->ptr, dyh->ld, bias_obj_grad, bias_bbox_grad);
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}

extern "C" int uavdet_attn_mlp_bwd(const float* attn, const float* d_attn, const float* hidden, const float* pooled, int n,
                                   int c, const float* w1, int hid, const float* w2, int K, float temperature,
                                   float out_scale, float* workspace, float* dw1, float* db1, float* dw2, float* db2,
                                   float* d_pooled, void* stream) {
  UAVDET_CHECK_ARG(attn && d_attn && hidden && pooled && w1 && w2 && workspace && n > 0 && c > 0 && hid > 0 && K > 0 &&
                       K <= 32 && temperature != 0.f,
                   "attn_mlp_bwd: bad arguments (K <= 32)");
  float* g = workspace;
  float* dh = workspace + (size_t)n * K;
  attn_bwd_dh_kernel<<<n, 128, 0, ST>>>(attn, d_attn, hidden, w2, hid, K, 1.f / temperature, g, dh);
  UAVDET_LAUNCH_CHECK();
  const long long work = (long long)(hid > n ? hid : n) * c;
  attn_bwd_params_kernel<<<ew_grid(work, 256), 256, 0, ST>>>(g, dh, hidden, pooled, w1, n, c, hid, K, out_scale, dw1, db1,
                                                            dw2, db2, d_pooled);
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}

extern "C" int uavdet_sgd_momentum(float* param, const float* grad, float* momentum_buf, int64_t count, float lr,
                                   float momentum, float grad_scale, int first_step, void* stream) {
  UAVDET_CHECK_ARG(param && grad && momentum_buf && count >= 0, "sgd: bad arguments");
  UAVDET_CHECK_ARG((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)momentum_buf) & 15) == 0, "sgd: 16-byte alignment");
  if (count == 0) return UAVDET_OK;
  sgd_momentum_kernel<<<ew_grid(count / 4 + 1, 256), 256, 0, ST>>>(param, grad, momentum_buf, count, lr, momentum,
                                                                   grad_scale, first_step, nullptr);
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}

extern "C" int uavdet_sgd_momentum_dev(float* param, const float* grad, float* momentum_buf, int64_t count,
                                       const float* hyper_dev, int first_step, void* stream) {
  UAVDET_CHECK_ARG(param && grad && momentum_buf && hyper_dev && count >= 0, "sgd_dev: bad arguments");
  UAVDET_CHECK_ARG((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)momentum_buf) & 15) == 0, "sgd_dev: 16-byte alignment");
  if (count == 0) return UAVDET_OK;
  sgd_momentum_kernel<<<ew_grid(count / 4 + 1, 256), 256, 0, ST>>>(param, grad, momentum_buf, count, 0.f, 0.f, 1.f,
                                                                   first_step, hyper_dev);
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}
