// K3 / K6 / K7 — RTMUAVDet's memory-bound kernels (reference model/RTMUAVDet.py):
//   * per-sample depthwise dynamic convolution + residual  (MDyConv.forward :80-98; K <= 25 taps,
//     far below the tensor-core ridge: one thread owns 8 channels of one pixel, 128-bit accesses)
//   * GroupNorm(num_groups=1) as per-sample statistics + one normalise pass, with the residual add
//     of MDyEncoder (:171-174) folded into both passes
//   * bilinear x2 upsampling, align_corners=False (MFDFEncoderModule :193)
//   * tiny fully-connected layers of the attention branch (:54-62) and the sigmoid + box decode of
//     RTMHead (:234,253,274-291).
#include "common.cuh"

namespace uavdet {

__device__ __forceinline__ void unpack8r(const uint4& u, float (&f)[8]) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8r(const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  return u;
}
static inline int rtm_grid(long long items, int threads) {
  long long b = (items + threads - 1) / threads, cap = (long long)kNumSMs * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// Thread layout of the streaming kernels below: blockDim 256 = Gb channel groups (8 channels = 16 B) x PL pixel
// lanes; a thread keeps its channel group, so per-channel parameters are loaded once and no index needs a divide.
struct Lanes { int g, pl, Gb, PL; bool active; };
__device__ __forceinline__ Lanes lanes_of(int c) {
  const int G = c >> 3;
  Lanes L;
  L.Gb = G < 32 ? G : 32;
  L.PL = 256 / L.Gb;
  L.g = threadIdx.x % L.Gb + blockIdx.y * L.Gb;
  L.pl = threadIdx.x / L.Gb;
  L.active = L.g < G && L.pl < L.PL;
  return L;
}

// out[b,y,x,c] = x[b,y,x,c] + channel_w[b,c] * sum_t kernel_w[b,t] * x[b,y+dy,x+dx,c]
// A thread walks a vertical strip of `rows` output pixels of one column (8 channels).  Every INPUT row is loaded and
// converted to fp32 once and scattered into the KS output rows it contributes to: a ring of KS accumulators whose slot
// indices are compile-time constants because the row loop is unrolled by KS.  (The first version gathered a KS x KS
// window per output row: KS x more bf16 -> fp32 conversions and a register shuffle of the window per row made it
// instruction-bound — 18 % of the HBM rate for k = 5, 50 % for k = 3.)
// kFold (the MDyEncoder site, RTMUAVDet.py:171-174): the encoder's residual `res` (a channel slice of its input) is
// added to the result, and the per-sample sum / sum of squares of the bf16-ROUNDED outputs — the values the 1x1
// convolution behind GroupNorm(1 group) is going to read — are accumulated into s1 / s2: GroupNorm(cat + residual) then
// needs neither a statistics pass nor a normalise pass (its affine map is folded into that convolution).
template <int KS, bool kFold>
__device__ __forceinline__ void dwdynconv_body(const __nv_bfloat16* __restrict__ x, int x_ld, int h, int w, int c,
                 const float* __restrict__ channel_w, const float* __restrict__ kernel_w, int xblocks, int rows,
                 __nv_bfloat16* __restrict__ y, int y_ld, const __nv_bfloat16* __restrict__ res, int r_ld,
                 float& s1, float& s2) {
  const Lanes L = lanes_of(c);
  if (!L.active) return;
  const int b = blockIdx.z;
  const int xb = blockIdx.x % xblocks, strip = blockIdx.x / xblocks;
  const int ox = xb * L.PL + L.pl;
  if (ox >= w) return;
  const int cc = L.g << 3;
  const int y0 = strip * rows;
  const int nrows = min(h, y0 + rows) - y0;          // output rows of this strip
  constexpr int PAD = KS / 2;
  float kw[KS * KS], cw[8];
#pragma unroll
  for (int t = 0; t < KS * KS; ++t) kw[t] = __ldg(kernel_w + b * KS * KS + t);
#pragma unroll
  for (int j = 0; j < 8; ++j) cw[j] = __ldg(channel_w + (long long)b * c + cc + j);
  const __nv_bfloat16* xb_ = x + (long long)b * h * w * x_ld + cc;
  __nv_bfloat16* yb_ = y + (long long)b * h * w * y_ld + cc;
  const __nv_bfloat16* rb_ = kFold ? res + (long long)b * h * w * r_ld + cc : nullptr;
  if (KS == 1) {
    // 1x1: y = x * (1 + channel_w * kernel_w) — a pure stream, four loads in flight
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = fmaf(cw[j], kw[0], 1.f);
    for (int r0 = 0; r0 < nrows; r0 += 4) {
      uint4 raw[4], rr[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (r0 + u < nrows) {
          raw[u] = __ldcs(reinterpret_cast<const uint4*>(xb_ + ((long long)(y0 + r0 + u) * w + ox) * x_ld));
          if (kFold) rr[u] = __ldcs(reinterpret_cast<const uint4*>(rb_ + ((long long)(y0 + r0 + u) * w + ox) * r_ld));
        }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (r0 + u < nrows) {
          float v[8], o[8];
          unpack8r(raw[u], v);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = fmaf(fmaf(kw[0], v[j], 0.f), cw[j], v[j]);
          if (kFold) {
            // same association as the unfused pair of kernels would have in fp32: (x + cw * dw) + res
            float r[8];
            unpack8r(rr[u], r);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] += r[j];
            const uint4 out = pack8r(o);
            unpack8r(out, o);
#pragma unroll
            for (int j = 0; j < 8; ++j) { s1 += o[j]; s2 = fmaf(o[j], o[j], s2); }
            *reinterpret_cast<uint4*>(yb_ + ((long long)(y0 + r0 + u) * w + ox) * y_ld) = out;
          } else {
            *reinterpret_cast<uint4*>(yb_ + ((long long)(y0 + r0 + u) * w + ox) * y_ld) = pack8r(o);
          }
        }
    }
    (void)f;
    return;
  }
  float acc[KS][8];
#pragma unroll
  for (int s_ = 0; s_ < KS; ++s_)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[s_][j] = 0.f;
  // local input row li <-> image row y0 - PAD + li; it feeds output rows y0 + li - r (r = filter row), ring slot
  // (li - r) mod KS; output row y0 + li - (KS - 1) is complete after input row li.  The raw row li + 1 is requested
  // before row li is consumed, so its latency hides behind the KS*KS*8 FMAs of the current row.
  const int n_in = nrows + KS - 1;
  // Addresses are running pointers (one 64-bit add per row and pointer; computing ((long long)iy * w + ix) * ld for every
  // access was a third of the kernel's instructions, and the kernel is bound by instruction issue): `fp` walks the input
  // rows at the centre column, `cp` / `yp` / `rp` the completed output rows.  Pointers outside the image are never
  // dereferenced.
  const long long in_stride = (long long)w * x_ld;
  const __nv_bfloat16* fp = xb_ + ((long long)(y0 - PAD) * w + ox) * x_ld;
  const __nv_bfloat16* cp = xb_ + ((long long)y0 * w + ox) * x_ld;
  __nv_bfloat16* yp = yb_ + ((long long)y0 * w + ox) * y_ld;
  const __nv_bfloat16* rp = kFold ? rb_ + ((long long)y0 * w + ox) * r_ld : nullptr;
  bool col_ok[KS];
  int col_off[KS];
#pragma unroll
  for (int kx = 0; kx < KS; ++kx) {
    const int ix = ox + kx - PAD;
    col_ok[kx] = ix >= 0 && ix < w;
    col_off[kx] = (kx - PAD) * x_ld;
  }
  int fy = y0 - PAD;                                     // image row `fp` points at
  auto fetch = [&](uint4 (&raw)[KS]) {
    const bool row_ok = fy >= 0 && fy < h && fy < y0 - PAD + n_in;
#pragma unroll
    for (int kx = 0; kx < KS; ++kx)
      raw[kx] = (row_ok && col_ok[kx]) ? __ldg(reinterpret_cast<const uint4*>(fp + col_off[kx])) : make_uint4(0u, 0u, 0u, 0u);
    fp += in_stride;
    ++fy;
  };
  uint4 buf[2][KS];                                       // ping-pong: the loop is unrolled 2 * KS rows, no copies
  fetch(buf[0]);
  for (int base = 0; base < n_in; base += 2 * KS) {
#pragma unroll
    for (int u2 = 0; u2 < 2 * KS; ++u2) {
      const int li = base + u2;
      const int u = u2 % KS;
      if (li < n_in) {
        fetch(buf[(u2 + 1) & 1]);
        float v[KS][8];
#pragma unroll
        for (int kx = 0; kx < KS; ++kx) unpack8r(buf[u2 & 1][kx], v[kx]);
#pragma unroll
        for (int r = 0; r < KS; ++r) {
          const int slot = ((u - r) % KS + KS) % KS;            // compile-time after unrolling
#pragma unroll
          for (int kx = 0; kx < KS; ++kx) {
            // two channels per instruction (fma.rn.f32x2: each half rounds like fmaf)
            const float2 w2 = make_float2(kw[r * KS + kx], kw[r * KS + kx]);
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
              const float2 a2 = __ffma2_rn(w2, make_float2(v[kx][j], v[kx][j + 1]), make_float2(acc[slot][j], acc[slot][j + 1]));
              acc[slot][j] = a2.x; acc[slot][j + 1] = a2.y;
            }
          }
        }
        const int dslot = (u + 1) % KS;
        if (li >= KS - 1) {                                        // local output row li - (KS - 1) is complete
          float centre[8], o[8];
          unpack8r(__ldg(reinterpret_cast<const uint4*>(cp)), centre);
          cp += in_stride;
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = fmaf(cw[j], acc[dslot][j], centre[j]);
          if (kFold) {
            float r[8];
            unpack8r(__ldcs(reinterpret_cast<const uint4*>(rp)), r);
            rp += (long long)w * r_ld;
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] += r[j];
            const uint4 out = pack8r(o);
            unpack8r(out, o);
#pragma unroll
            for (int j = 0; j < 8; ++j) { s1 += o[j]; s2 = fmaf(o[j], o[j], s2); }
            *reinterpret_cast<uint4*>(yp) = out;
          } else {
            *reinterpret_cast<uint4*>(yp) = pack8r(o);
          }
          yp += (long long)w * y_ld;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[dslot][j] = 0.f;          // the slot now belongs to output row done + KS
      }
    }
  }
}

// block-wide sum of (s1, s2) -> two atomics per block
__device__ __forceinline__ void block_sum2_atomic(float s1, float s2, float* dst) {
  __shared__ float red[2][8];
  for (int off = 16; off; off >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, off); s2 += __shfl_xor_sync(0xffffffffu, s2, off); }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s1; red[1][threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x < 32) {
    s1 = threadIdx.x < 8 ? red[0][threadIdx.x] : 0.f;
    s2 = threadIdx.x < 8 ? red[1][threadIdx.x] : 0.f;
    for (int off = 4; off; off >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, off); s2 += __shfl_xor_sync(0xffffffffu, s2, off); }
    if (threadIdx.x == 0) { atomicAdd(dst, s1); atomicAdd(dst + 1, s2); }
  }
}

template <int KS>
__global__ void __launch_bounds__(256, KS <= 3 ? 3 : (KS <= 5 ? 2 : 1))
dwdynconv_kernel(const __nv_bfloat16* __restrict__ x, int x_ld, int h, int w, int c,
                 const float* __restrict__ channel_w, const float* __restrict__ kernel_w, int xblocks, int rows,
                 __nv_bfloat16* __restrict__ y, int y_ld) {
  float s1 = 0.f, s2 = 0.f;
  dwdynconv_body<KS, false>(x, x_ld, h, w, c, channel_w, kernel_w, xblocks, rows, y, y_ld, nullptr, 0, s1, s2);
}

// + residual + per-sample statistics of the result (stats[2 * b], stats[2 * b + 1], caller-zeroed)
template <int KS>
__global__ void __launch_bounds__(256, KS <= 3 ? 3 : (KS <= 5 ? 2 : 1))
dwdynconv_res_stats_kernel(const __nv_bfloat16* __restrict__ x, int x_ld, int h, int w, int c,
                           const float* __restrict__ channel_w, const float* __restrict__ kernel_w, int xblocks, int rows,
                           __nv_bfloat16* __restrict__ y, int y_ld, const __nv_bfloat16* __restrict__ res, int r_ld,
                           float* __restrict__ stats) {
  float s1 = 0.f, s2 = 0.f;
  dwdynconv_body<KS, true>(x, x_ld, h, w, c, channel_w, kernel_w, xblocks, rows, y, y_ld, res, r_ld, s1, s2);
  block_sum2_atomic(s1, s2, stats + 2 * blockIdx.z);
}

// GroupNorm(1 group) folded into the 1x1 convolution behind it.  With z = (v - mean_n) * rstd_n * gamma + beta and
// W' = W * diag(gamma):  conv(z)[o] = rstd_n * (W' v)[o] + (W beta)[o] - mean_n * rstd_n * sum_c W'[o][c], so the GEMM runs on
// the un-normalised v and its epilogue needs two numbers per image: (rstd_n, mean_n * rstd_n).
__global__ void gn_fold_kernel(const float* __restrict__ stats, int n, double inv_count, float eps, float* __restrict__ out) {
  const int img = blockIdx.x * blockDim.x + threadIdx.x;
  if (img >= n) return;
  // E[x^2] - E[x]^2 in double (cancellation), one fp32 rsqrt — as gn_apply_kernel
  const double m = (double)stats[2 * img] * inv_count;
  double var = fma(-m, m, (double)stats[2 * img + 1] * inv_count);
  if (var < 0) var = 0;
  const float rstd = rsqrtf((float)var + eps);
  out[2 * img] = rstd;
  out[2 * img + 1] = (float)m * rstd;
}

// out[r][o] = act(sum_c in[r][c] * W[o][c] + b[o]); one warp per output element
__global__ void linear_kernel(const float* __restrict__ in, int rows, int c, const float* __restrict__ W,
                              const float* __restrict__ bias, int o_dim, int act, float* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows * o_dim) return;
  const int r = warp / o_dim, o = warp - r * o_dim;
  float s = 0.f;
  for (int i = lane; i < c; i += 32) s = fmaf(__ldg(in + (long long)r * c + i), __ldg(W + (long long)o * c + i), s);
  for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) out[warp] = act_fwd_rt(act, s + (bias ? bias[o] : 0.f));
}

// per-sample sum / sum of squares of (a + b) over h*w*c
__global__ void __launch_bounds__(256)
gn_stats_kernel(const __nv_bfloat16* __restrict__ a, int a_ld, const __nv_bfloat16* __restrict__ b2, int b_ld, int hw,
                int c, float* __restrict__ stats) {
  const Lanes L = lanes_of(c);
  const int img = blockIdx.z;
  float s1 = 0.f, s2 = 0.f;
  if (L.active) {
    const int cc = L.g << 3;
    const __nv_bfloat16* ap = a + (long long)img * hw * a_ld + cc;
    const __nv_bfloat16* bp = b2 ? b2 + (long long)img * hw * b_ld + cc : nullptr;
    const int step = gridDim.x * L.PL;
    auto body = [&](const uint4& ua, const uint4& ub) {
      float v[8], u[8];
      unpack8r(ua, v);
      unpack8r(ub, u);
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float t = v[j] + u[j]; s1 += t; s2 = fmaf(t, t, s2); }
    };
    int px = blockIdx.x * L.PL + L.pl;
    for (; px + 3 * step < hw; px += 4 * step) {
      uint4 ua[4], ub[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        ua[q] = __ldg(reinterpret_cast<const uint4*>(ap + (long long)(px + q * step) * a_ld));
        ub[q] = bp ? __ldg(reinterpret_cast<const uint4*>(bp + (long long)(px + q * step) * b_ld)) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) body(ua[q], ub[q]);
    }
    for (; px < hw; px += step)
      body(__ldg(reinterpret_cast<const uint4*>(ap + (long long)px * a_ld)),
           bp ? __ldg(reinterpret_cast<const uint4*>(bp + (long long)px * b_ld)) : make_uint4(0u, 0u, 0u, 0u));
  }
  __shared__ float red[2][8];
  for (int off = 16; off; off >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, off); s2 += __shfl_xor_sync(0xffffffffu, s2, off); }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s1; red[1][threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x < 32) {
    s1 = threadIdx.x < 8 ? red[0][threadIdx.x] : 0.f;
    s2 = threadIdx.x < 8 ? red[1][threadIdx.x] : 0.f;
    for (int off = 4; off; off >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, off); s2 += __shfl_xor_sync(0xffffffffu, s2, off); }
    if (threadIdx.x == 0) { atomicAdd(stats + 2 * img, s1); atomicAdd(stats + 2 * img + 1, s2); }
  }
}

__global__ void __launch_bounds__(256)
gn_apply_kernel(const __nv_bfloat16* __restrict__ a, int a_ld, const __nv_bfloat16* __restrict__ b2, int b_ld, int hw,
                int c, const float* __restrict__ stats, float eps, const float* __restrict__ gamma,
                const float* __restrict__ beta, __nv_bfloat16* __restrict__ y, int y_ld) {
  const Lanes L = lanes_of(c);
  if (!L.active) return;
  const int img = blockIdx.z;
  const int cc = L.g << 3;
  // E[x^2] - E[x]^2 in double (cancellation), one fp32 rsqrt
  const double inv_cnt = 1.0 / ((double)hw * c);
  const double m = (double)stats[2 * img] * inv_cnt;
  double var = fma(-m, m, (double)stats[2 * img + 1] * inv_cnt);
  if (var < 0) var = 0;
  const float mean = (float)m, invstd = rsqrtf((float)var + eps);
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = invstd * __ldg(gamma + cc + j);
    sh[j] = fmaf(-mean, sc[j], __ldg(beta + cc + j));
  }
  const __nv_bfloat16* ap = a + (long long)img * hw * a_ld + cc;
  const __nv_bfloat16* bp = b2 ? b2 + (long long)img * hw * b_ld + cc : nullptr;
  __nv_bfloat16* yp = y + (long long)img * hw * y_ld + cc;
  const int step = gridDim.x * L.PL;
  auto body = [&](const uint4& ua, const uint4& ub, int px) {
    float v[8], u[8];
    unpack8r(ua, v);
    unpack8r(ub, u);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j] + u[j], sc[j], sh[j]);
    *reinterpret_cast<uint4*>(yp + (long long)px * y_ld) = pack8r(v);
  };
  int px = blockIdx.x * L.PL + L.pl;
  for (; px + 3 * step < hw; px += 4 * step) {
    uint4 ua[4], ub[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      ua[q] = __ldg(reinterpret_cast<const uint4*>(ap + (long long)(px + q * step) * a_ld));
      ub[q] = bp ? __ldg(reinterpret_cast<const uint4*>(bp + (long long)(px + q * step) * b_ld)) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) body(ua[q], ub[q], px + q * step);
  }
  for (; px < hw; px += step)
    body(__ldg(reinterpret_cast<const uint4*>(ap + (long long)px * a_ld)),
         bp ? __ldg(reinterpret_cast<const uint4*>(bp + (long long)px * b_ld)) : make_uint4(0u, 0u, 0u, 0u), px);
}

// nn.Upsample(scale_factor=2, mode='bilinear') (align_corners=False).
// A thread owns 8 channels of a run of kBilinearRun input columns of one input row i and writes the 2 x (2 * run)
// output pixels that row produces.  The 3 x 3 input neighbourhood rolls along the run in registers, so an input pixel
// is fetched 3 x (run + 2) / run times per thread instead of 4 times per OUTPUT pixel (the first version was bound by
// the L2 -> SM traffic of those re-reads: 6.7 GB for 1.6 GB written).  Arithmetic is the textbook expression
// hy*(hx*v00 + lx*v01) + ly*(hx*v10 + lx*v11) with torch's source-index rule, unchanged.
constexpr int kBilinearRun = 8;
__global__ void __launch_bounds__(256)
bilinear2x_kernel(const __nv_bfloat16* __restrict__ x, int x_ld, int h, int w, int c, int gy,
                  __nv_bfloat16* __restrict__ y, int y_ld) {
  const int G = c >> 3;
  const int Gb = G < 32 ? G : 32;
  const int PL = 256 / Gb;
  const int i = blockIdx.y / gy;                              // input row
  const int g = threadIdx.x % Gb + (blockIdx.y - i * gy) * Gb;
  const int pl = threadIdx.x / Gb;
  const int j0 = (blockIdx.x * PL + pl) * kBilinearRun;       // first input column of the run
  if (g >= G || pl >= PL || j0 >= w) return;
  const int b = blockIdx.z;
  const int cc = g << 3;
  const int W = 2 * w;
  const __nv_bfloat16* xb_ = x + (long long)b * h * w * x_ld + cc;
  __nv_bfloat16* yb_ = y + (long long)b * (2 * h) * W * y_ld + cc;
  const int ym = i > 0 ? i - 1 : 0, yp = i < h - 1 ? i + 1 : i;
  // vertical weights of the two output rows 2i and 2i+1 (torch: src = max((dst + 0.5) / 2 - 0.5, 0))
  float ly[2], hy[2];
  int top[2];                                                 // window row holding y0 (0 = row ym, 1 = row i)
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const float sy = fmaxf((2 * i + r + 0.5f) * 0.5f - 0.5f, 0.f);
    const int y0 = (int)sy;
    ly[r] = sy - y0;
    hy[r] = 1.f - ly[r];
    top[r] = (r == 0) ? 0 : 1;
  }
  float win[3][3][8];                                         // [row ym, i, yp][column j-1, j, j+1][channel]
  auto load_col = [&](int col, int slot) {
    const int jc = col < 0 ? 0 : (col > w - 1 ? w - 1 : col);
    unpack8r(__ldg(reinterpret_cast<const uint4*>(xb_ + ((long long)ym * w + jc) * x_ld)), win[0][slot]);
    unpack8r(__ldg(reinterpret_cast<const uint4*>(xb_ + ((long long)i * w + jc) * x_ld)), win[1][slot]);
    unpack8r(__ldg(reinterpret_cast<const uint4*>(xb_ + ((long long)yp * w + jc) * x_ld)), win[2][slot]);
  };
  load_col(j0 - 1, 0);
  load_col(j0, 1);
  const int j_end = min(w, j0 + kBilinearRun);
  for (int j = j0; j < j_end; ++j) {
    load_col(j + 1, 2);
#pragma unroll
    for (int q = 0; q < 2; ++q) {                              // output column 2j + q
      const float sx = fmaxf((2 * j + q + 0.5f) * 0.5f - 0.5f, 0.f);
      const int x0 = (int)sx;
      const float lx = sx - x0, hx = 1.f - lx;
      const int left = q == 0 ? 0 : 1;                        // window column holding x0
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e)
          o[e] = hy[r] * (hx * win[top[r]][left][e] + lx * win[top[r]][left + 1][e]) +
                 ly[r] * (hx * win[top[r] + 1][left][e] + lx * win[top[r] + 1][left + 1][e]);
        *reinterpret_cast<uint4*>(yb_ + ((long long)(2 * i + r) * W + 2 * j + q) * y_ld) = pack8r(o);
      }
    }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int e = 0; e < 8; ++e) { win[r][0][e] = win[r][1][e]; win[r][1][e] = win[r][2][e]; }
  }
}

struct RtmAnchors { float w[8]; float h[8]; };
// RTMHead: sigmoid on both heads, then px = 2s-0.5+gx, pw = (2s)^2*anchor_w (anchors NOT stride-scaled)
__global__ void rtm_head_post_kernel(const float4* __restrict__ bbox_logits, const float* __restrict__ obj_logits,
                                     int batch, int A, int Sh, int Sw, RtmAnchors anc, float4* __restrict__ bbox_out,
                                     float* __restrict__ obj_out) {
  const int per_img = A * Sh * Sw;
  const long long total = (long long)batch * per_img;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i % per_img);
    const int a = r / (Sh * Sw);
    const int yx = r - a * (Sh * Sw);
    const int gy = yx / Sw, gx = yx - gy * Sw;
    const float4 t = __ldg(&bbox_logits[i]);
    const float s0 = 1.f / (1.f + expf(-t.x)), s1 = 1.f / (1.f + expf(-t.y));
    const float s2 = 1.f / (1.f + expf(-t.z)), s3 = 1.f / (1.f + expf(-t.w));
    const float px = __fadd_rn(__fsub_rn(__fmul_rn(s0, 2.f), 0.5f), (float)gx);
    const float py = __fadd_rn(__fsub_rn(__fmul_rn(s1, 2.f), 0.5f), (float)gy);
    const float bw = __fmul_rn(s2, 2.f), bh = __fmul_rn(s3, 2.f);
    bbox_out[i] = make_float4(px, py, __fmul_rn(__fmul_rn(bw, bw), anc.w[a]), __fmul_rn(__fmul_rn(bh, bh), anc.h[a]));
    obj_out[i] = 1.f / (1.f + expf(-__ldg(&obj_logits[i])));
  }
}

}  // namespace uavdet

using namespace uavdet;
#define ST ((cudaStream_t)stream)

static int chk(const uavdet_act* a, const char* what) {
  UAVDET_CHECK_ARG(a && a->ptr, "%s: null view", what);
  UAVDET_CHECK_ARG(a->c % 8 == 0 && a->ld % 8 == 0 && ((uintptr_t)a->ptr & 15) == 0, "%s: alignment (c=%d ld=%d)", what, a->c, a->ld);
  return UAVDET_OK;
}

extern "C" int uavdet_dwdynconv_fwd(const uavdet_act* x, const float* channel_w, const float* kernel_w, int k, int pad,
                                    const uavdet_act* y, void* stream) {
  int rc;
  if ((rc = chk(x, "dwdynconv x")) || (rc = chk(y, "dwdynconv y"))) return rc;
  UAVDET_CHECK_ARG(channel_w && kernel_w && k >= 1 && k <= 7 && 2 * pad == k - 1, "dwdynconv: bad arguments (k=%d pad=%d)", k, pad);
  UAVDET_CHECK_ARG(x->n == y->n && x->h == y->h && x->w == y->w && x->c == y->c, "dwdynconv: shape mismatch");
  const int G = x->c / 8, Gb = G < 32 ? G : 32, PL = 256 / Gb;
  const int xblocks = ceil_div(x->w, PL);
  int rows = 32;
  while (rows > 4 && (long long)xblocks * ceil_div(x->h, rows) * ceil_div(G, Gb) * x->n < 8 * kNumSMs) rows >>= 1;
  dim3 grid((unsigned)(xblocks * ceil_div(x->h, rows)), (unsigned)ceil_div(G, Gb), (unsigned)x->n);
  const __nv_bfloat16* xp = (const __nv_bfloat16*)x->ptr;
  __nv_bfloat16* yp = (__nv_bfloat16*)y->ptr;
#define UAVDET_DW(KS) dwdynconv_kernel<KS><<<grid, 256, 0, ST>>>(xp, x->ld, x->h, x->w, x->c, channel_w, kernel_w, xblocks, rows, yp, y->ld)
  switch (k) {
    case 1: UAVDET_DW(1); break;
    case 3: UAVDET_DW(3); break;
    case 5: UAVDET_DW(5); break;
    default: UAVDET_DW(7); break;
  }
#undef UAVDET_DW
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}

extern "C" int uavdet_dwdynconv_res_stats_fwd(const uavdet_act* x, const float* channel_w, const float* kernel_w, int k, int pad,
                                              const uavdet_act* res, float* stats, const uavdet_act* y, void* stream) {
  int rc;
  if ((rc = chk(x, "dwdynconv x")) || (rc = chk(y, "dwdynconv y")) || (rc = chk(res, "dwdynconv res"))) return rc;
  UAVDET_CHECK_ARG(channel_w && kernel_w && stats && k >= 1 && k <= 7 && 2 * pad == k - 1, "dwdynconv: bad arguments (k=%d pad=%d)", k, pad);
  UAVDET_CHECK_ARG(x->n == y->n && x->h == y->h && x->w == y->w && x->c == y->c, "dwdynconv: shape mismatch");
  UAVDET_CHECK_ARG(x->n == res->n && x->h == res->h && x->w == res->w && x->c == res->c, "dwdynconv: residual shape mismatch");
  const int G = x->c / 8, Gb = G < 32 ? G : 32, PL = 256 / Gb;
  const int xblocks = ceil_div(x->w, PL);
  int rows = 32;
  while (rows > 4 && (long long)xblocks * ceil_div(x->h, rows) * ceil_div(G, Gb) * x->n < 8 * kNumSMs) rows >>= 1;
  dim3 grid((unsigned)(xblocks * ceil_div(x->h, rows)), (unsigned)ceil_div(G, Gb), (unsigned)x->n);
  const __nv_bfloat16* xp = (const __nv_bfloat16*)x->ptr;
  __nv_bfloat16* yp = (__nv_bfloat16*)y->ptr;
#define UAVDET_DW(KS) dwdynconv_res_stats_kernel<KS><<<grid, 256, 0, ST>>>(xp, x->ld, x->h, x->w, x->c, channel_w, kernel_w, xblocks, \
                                                                          rows, yp, y->ld, (const __nv_bfloat16*)res->ptr, res->ld, stats)
  switch (k) {
    case 1: UAVDET_DW(1); break;
    case 3: UAVDET_DW(3); break;
    case 5: UAVDET_DW(5); break;
    default: UAVDET_DW(7); break;
  }
#undef UAVDET_DW
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}

extern "C" int uavdet_linear(const float* in, int rows, int c, const float* w, const float* bias, int out_dim, int act,
                             float* out, void* stream) {
  UAVDET_CHECK_ARG(in && w && out && rows > 0 && c > 0 && out_dim > 0, "linear: bad arguments");
  long long warps = (long long)rows * out_dim;
  linear_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, ST>>>(in, rows, c, w, bias, out_dim, act, out);
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}

static int gn_grid(const uavdet_act* a, dim3* grid, int* hw_out) {
  const long long hw64 = (long long)a->h * a->w;
  UAVDET_CHECK_ARG(hw64 < (1ll << 30), "groupnorm: map too large");
  const int hw = (int)hw64;
  const int G = a->c / 8, Gb = G < 32 ? G : 32, PL = 256 / Gb, gy = ceil_div(G, Gb);
  long long bx = ceil_div(hw, PL * 4);
  long long cap = ((long long)kNumSMs * 16) / ((long long)a->n * gy) + 1;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  *grid = dim3((unsigned)bx, (unsigned)gy, (unsigned)a->n);
  *hw_out = hw;
  return UAVDET_OK;
}

extern "C" int uavdet_groupnorm1_stats(const uavdet_act* a, const uavdet_act* b, float* stats, void* stream) {
  int rc;
  if ((rc = chk(a, "groupnorm a"))) return rc;
  if (b && (rc = chk(b, "groupnorm b"))) return rc;
  UAVDET_CHECK_ARG(stats, "groupnorm: null pointer");
  if (b) UAVDET_CHECK_ARG(a->n == b->n && a->h == b->h && a->w == b->w && a->c == b->c, "groupnorm: residual shape mismatch");
  UAVDET_CUDA(cudaMemsetAsync(stats, 0, sizeof(float) * 2 * a->n, ST));
  dim3 grid;
  int hw;
  if ((rc = gn_grid(a, &grid, &hw))) return rc;
  gn_stats_kernel<<<grid, 256, 0, ST>>>((const __nv_bfloat16*)a->ptr, a->ld, b ? (const __nv_bfloat16*)b->ptr : nullptr,
                                        b ? b->ld : 0, hw, a->c, stats);
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}

extern "C" int uavdet_groupnorm1_fold(const float* stats, int n, double count, float eps, float* sample_affine, void* stream) {
  UAVDET_CHECK_ARG(stats && sample_affine && n > 0 && count > 0, "groupnorm fold: bad arguments");
  gn_fold_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ST>>>(stats, n, 1.0 / count, eps, sample_affine);
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}

extern "C" int uavdet_groupnorm1(const uavdet_act* a, const uavdet_act* b, const float* gamma, const float* beta,
                                 float eps, float* stats_ws, const uavdet_act* y, void* stream) {
  int rc;
  if ((rc = chk(a, "groupnorm a")) || (rc = chk(y, "groupnorm y"))) return rc;
  UAVDET_CHECK_ARG(gamma && beta && stats_ws, "groupnorm: null pointer");
  UAVDET_CHECK_ARG(a->n == y->n && a->h == y->h && a->w == y->w && a->c == y->c, "groupnorm: shape mismatch");
  if ((rc = uavdet_groupnorm1_stats(a, b, stats_ws, stream))) return rc;
  dim3 grid;
  int hw;
  if ((rc = gn_grid(a, &grid, &hw))) return rc;
  gn_apply_kernel<<<grid, 256, 0, ST>>>((const __nv_bfloat16*)a->ptr, a->ld, b ? (const __nv_bfloat16*)b->ptr : nullptr,
                                        b ? b->ld : 0, hw, a->c, stats_ws, eps, gamma, beta, (__nv_bfloat16*)y->ptr, y->ld);
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}

extern "C" int uavdet_bilinear2x_fwd(const uavdet_act* x, const uavdet_act* y, void* stream) {
  int rc;
  if ((rc = chk(x, "bilinear2x x")) || (rc = chk(y, "bilinear2x y"))) return rc;
  UAVDET_CHECK_ARG(y->n == x->n && y->h == 2 * x->h && y->w == 2 * x->w && y->c == x->c, "bilinear2x: shapes");
  const int G = x->c / 8, Gb = G < 32 ? G : 32, PL = 256 / Gb, gy = ceil_div(G, Gb);
  UAVDET_CHECK_ARG((long long)x->h * gy <= 65535, "bilinear2x: map too tall");
  dim3 grid((unsigned)ceil_div(ceil_div(x->w, kBilinearRun), PL), (unsigned)(x->h * gy), (unsigned)x->n);
  bilinear2x_kernel<<<grid, 256, 0, ST>>>((const __nv_bfloat16*)x->ptr, x->ld, x->h, x->w, x->c, gy,
                                          (__nv_bfloat16*)y->ptr, y->ld);
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}

extern "C" int uavdet_rtm_head_post(const float* bbox_logits, const float* obj_logits, int batch, int A, int S_h, int S_w,
                                    const float* anchors_host, float* bbox_out, float* obj_out, void* stream) {
  UAVDET_CHECK_ARG(A > 0 && A <= 8 && bbox_logits && obj_logits && bbox_out && obj_out && anchors_host, "rtm_head_post: bad arguments");
  UAVDET_CHECK_ARG((((uintptr_t)bbox_logits | (uintptr_t)bbox_out) & 15) == 0, "rtm_head_post: alignment");
  if (batch == 0) return UAVDET_OK;
  RtmAnchors anc{};
  for (int a = 0; a < A; ++a) { anc.w[a] = anchors_host[2 * a]; anc.h[a] = anchors_host[2 * a + 1]; }
  long long total = (long long)batch * A * S_h * S_w;
  rtm_head_post_kernel<<<rtm_grid(total, 256), 256, 0, ST>>>((const float4*)bbox_logits, obj_logits, batch, A, S_h, S_w,
                                                            anc, (float4*)bbox_out, obj_out);
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}
