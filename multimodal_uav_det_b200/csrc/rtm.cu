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

// out[b,y,x,c] = x[b,y,x,c] + channel_w[b,c] * sum_t kernel_w[b,t] * x[b,y+dy,x+dx,c]
__global__ void dwdynconv_kernel(const __nv_bfloat16* __restrict__ x, int x_ld, int n, int h, int w, int c,
                                 const float* __restrict__ channel_w, const float* __restrict__ kernel_w, int k,
                                 int pad, __nv_bfloat16* __restrict__ y, int y_ld) {
  const int c8 = c >> 3;
  const long long total = (long long)n * h * w * c8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long px = i / c8;
    const int cc = (int)(i - px * c8) << 3;
    const int ox = (int)(px % w); long long t = px / w;
    const int oy = (int)(t % h);
    const int b = (int)(t / h);
    float acc[8], centre[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    unpack8r(__ldg(reinterpret_cast<const uint4*>(x + px * x_ld + cc)), centre);
    const float* kw = kernel_w + (long long)b * k * k;
    for (int kh = 0; kh < k; ++kh) {
      const int iy = oy + kh - pad;
      if (iy < 0 || iy >= h) continue;
      for (int kx = 0; kx < k; ++kx) {
        const int ix = ox + kx - pad;
        if (ix < 0 || ix >= w) continue;
        const float wt = __ldg(kw + kh * k + kx);
        float v[8];
        unpack8r(__ldg(reinterpret_cast<const uint4*>(x + (((long long)b * h + iy) * w + ix) * x_ld + cc)), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(wt, v[j], acc[j]);
      }
    }
    const float* cw = channel_w + (long long)b * c + cc;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = fmaf(__ldg(cw + j), acc[j], centre[j]);
    *reinterpret_cast<uint4*>(y + px * y_ld + cc) = pack8r(acc);
  }
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
__global__ void gn_stats_kernel(const __nv_bfloat16* __restrict__ a, int a_ld, const __nv_bfloat16* __restrict__ b2,
                                int b_ld, long long hw, int c, float* __restrict__ stats) {
  const int img = blockIdx.y;
  const int c8 = c >> 3;
  const long long total = hw * c8;
  float s1 = 0.f, s2 = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long px = i / c8 + (long long)img * hw;
    const int cc = (int)(i % c8) << 3;
    float v[8];
    unpack8r(__ldg(reinterpret_cast<const uint4*>(a + px * a_ld + cc)), v);
    if (b2) {
      float u[8];
      unpack8r(__ldg(reinterpret_cast<const uint4*>(b2 + px * b_ld + cc)), u);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += u[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { s1 += v[j]; s2 = fmaf(v[j], v[j], s2); }
  }
  __shared__ float red[2][32];
  for (int off = 16; off; off >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, off); s2 += __shfl_xor_sync(0xffffffffu, s2, off); }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s1; red[1][threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x < 32) {
    s1 = threadIdx.x < (blockDim.x >> 5) ? red[0][threadIdx.x] : 0.f;
    s2 = threadIdx.x < (blockDim.x >> 5) ? red[1][threadIdx.x] : 0.f;
    for (int off = 16; off; off >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, off); s2 += __shfl_xor_sync(0xffffffffu, s2, off); }
    if (threadIdx.x == 0) { atomicAdd(stats + 2 * img, s1); atomicAdd(stats + 2 * img + 1, s2); }
  }
}

__global__ void gn_apply_kernel(const __nv_bfloat16* __restrict__ a, int a_ld, const __nv_bfloat16* __restrict__ b2,
                                int b_ld, long long hw, int c, const float* __restrict__ stats, float eps,
                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                __nv_bfloat16* __restrict__ y, int y_ld) {
  const int img = blockIdx.y;
  const int c8 = c >> 3;
  const long long total = hw * c8;
  const double cnt = (double)hw * c;
  const double m = (double)stats[2 * img] / cnt;
  double var = (double)stats[2 * img + 1] / cnt - m * m;
  if (var < 0) var = 0;
  const float mean = (float)m, invstd = (float)(1.0 / sqrt(var + (double)eps));
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long px = i / c8 + (long long)img * hw;
    const int cc = (int)(i % c8) << 3;
    float v[8];
    unpack8r(__ldg(reinterpret_cast<const uint4*>(a + px * a_ld + cc)), v);
    if (b2) {
      float u[8];
      unpack8r(__ldg(reinterpret_cast<const uint4*>(b2 + px * b_ld + cc)), u);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += u[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (v[j] - mean) * invstd * __ldg(gamma + cc + j) + __ldg(beta + cc + j);
    *reinterpret_cast<uint4*>(y + px * y_ld + cc) = pack8r(v);
  }
}

// nn.Upsample(scale_factor=2, mode='bilinear') (align_corners=False)
__global__ void bilinear2x_kernel(const __nv_bfloat16* __restrict__ x, int x_ld, int n, int h, int w, int c,
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
    const float sy = fmaxf((oy + 0.5f) * 0.5f - 0.5f, 0.f), sx = fmaxf((ox + 0.5f) * 0.5f - 0.5f, 0.f);
    const int y0 = (int)sy, x0 = (int)sx;
    const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
    const float ly = sy - y0, lx = sx - x0, hy = 1.f - ly, hx = 1.f - lx;
    const long long base = (long long)b * h * w;
    float v00[8], v01[8], v10[8], v11[8], o[8];
    unpack8r(__ldg(reinterpret_cast<const uint4*>(x + (base + (long long)y0 * w + x0) * x_ld + cc)), v00);
    unpack8r(__ldg(reinterpret_cast<const uint4*>(x + (base + (long long)y0 * w + x1) * x_ld + cc)), v01);
    unpack8r(__ldg(reinterpret_cast<const uint4*>(x + (base + (long long)y1 * w + x0) * x_ld + cc)), v10);
    unpack8r(__ldg(reinterpret_cast<const uint4*>(x + (base + (long long)y1 * w + x1) * x_ld + cc)), v11);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = hy * (hx * v00[j] + lx * v01[j]) + ly * (hx * v10[j] + lx * v11[j]);
    *reinterpret_cast<uint4*>(y + px * y_ld + cc) = pack8r(o);
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
  long long total = (long long)x->n * x->h * x->w * (x->c / 8);
  dwdynconv_kernel<<<rtm_grid(total, 256), 256, 0, ST>>>((const __nv_bfloat16*)x->ptr, x->ld, x->n, x->h, x->w, x->c,
                                                        channel_w, kernel_w, k, pad, (__nv_bfloat16*)y->ptr, y->ld);
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

extern "C" int uavdet_groupnorm1(const uavdet_act* a, const uavdet_act* b, const float* gamma, const float* beta,
                                 float eps, float* stats_ws, const uavdet_act* y, void* stream) {
  int rc;
  if ((rc = chk(a, "groupnorm a")) || (rc = chk(y, "groupnorm y"))) return rc;
  if (b && (rc = chk(b, "groupnorm b"))) return rc;
  UAVDET_CHECK_ARG(gamma && beta && stats_ws, "groupnorm: null pointer");
  UAVDET_CHECK_ARG(a->n == y->n && a->h == y->h && a->w == y->w && a->c == y->c, "groupnorm: shape mismatch");
  if (b) UAVDET_CHECK_ARG(a->n == b->n && a->h == b->h && a->w == b->w && a->c == b->c, "groupnorm: residual shape mismatch");
  UAVDET_CUDA(cudaMemsetAsync(stats_ws, 0, sizeof(float) * 2 * a->n, ST));
  const long long hw = (long long)a->h * a->w;
  long long bx = (hw * (a->c / 8) + 256 * 8 - 1) / (256 * 8);
  long long cap = (kNumSMs * 8) / (a->n > 0 ? a->n : 1) + 1;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  dim3 grid((unsigned)bx, (unsigned)a->n);
  const __nv_bfloat16* bp = b ? (const __nv_bfloat16*)b->ptr : nullptr;
  gn_stats_kernel<<<grid, 256, 0, ST>>>((const __nv_bfloat16*)a->ptr, a->ld, bp, b ? b->ld : 0, hw, a->c, stats_ws);
  UAVDET_LAUNCH_CHECK();
  gn_apply_kernel<<<grid, 256, 0, ST>>>((const __nv_bfloat16*)a->ptr, a->ld, bp, b ? b->ld : 0, hw, a->c, stats_ws, eps,
                                        gamma, beta, (__nv_bfloat16*)y->ptr, y->ld);
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}

extern "C" int uavdet_bilinear2x_fwd(const uavdet_act* x, const uavdet_act* y, void* stream) {
  int rc;
  if ((rc = chk(x, "bilinear2x x")) || (rc = chk(y, "bilinear2x y"))) return rc;
  UAVDET_CHECK_ARG(y->n == x->n && y->h == 2 * x->h && y->w == 2 * x->w && y->c == x->c, "bilinear2x: shapes");
  long long total = (long long)y->n * y->h * y->w * (y->c / 8);
  bilinear2x_kernel<<<rtm_grid(total, 256), 256, 0, ST>>>((const __nv_bfloat16*)x->ptr, x->ld, x->n, x->h, x->w, x->c,
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
