// Stem convolution for cin in {1,3}: direct CUDA-core kernel reading the NCHW fp32 network
// input and writing NHWC bf16.  K = cin*k*k <= 75 is far below the tensor-core ridge
// (SURVEY.md §7 "thin-channel early layers"): the layer is bound by the 64 B/pixel output
// write, so one thread owns one output pixel and all 32 output channels.
// Reference: first layer of BaselineModel.py:89-97 / DyYOLO.py:89-100 (via per-sample weights),
// DySOEM_SimFPN.py:30 (1x1), RTMUAVDet.py:31 (5x5 s2 p1).
#include "common.cuh"

namespace uavdet {

constexpr int kStemCout = 32;
constexpr int kStemMaxK = 75;  // 3 * 5 * 5

struct StemParams {
  const float* x; int n, cin, h, w;
  const float* wgt;        // [w_batch][32][cin][k][k] fp32
  int w_batch;
  int k, stride, pad, ho, wo;
  int yh, yw;              // output buffer grid (== ho,wo, or ho+1,wo+1: zero-padded last row/col)
  __nv_bfloat16* y; long long y_ld;
  int epi, act;
  const float* scale; const float* shift;
  float* sum; float* sumsq;
};

__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int step = 16; step >= 1; step >>= 1) {
    const bool upper = (lane & step) != 0;
#pragma unroll
    for (int i = 0; i < step; ++i) {
      float send = upper ? v[i] : v[i + step];
      float keep = upper ? v[i + step] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, step);
    }
  }
  return v[0];
}

// grid.y = image; each block handles 256 consecutive output pixels of that image
__global__ void __launch_bounds__(256) stem_fwd_kernel(StemParams P) {
  __shared__ float sw[kStemMaxK * kStemCout];  // [tap*cin][32] (cout fastest -> broadcast-free reads)
  __shared__ float ssum[2][kStemCout];
  const int img = blockIdx.y;
  const int K = P.cin * P.k * P.k;
  const float* wsrc = P.wgt + (P.w_batch > 1 ? (long long)img * kStemCout * K : 0);
  for (int i = threadIdx.x; i < K * kStemCout; i += blockDim.x) {
    const int co = i % kStemCout, kk = i / kStemCout;  // kk = (ci*k + kh)*k + kw
    sw[kk * kStemCout + co] = wsrc[co * K + kk];
  }
  if (threadIdx.x < 2 * kStemCout) ssum[threadIdx.x / kStemCout][threadIdx.x % kStemCout] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long hw_out = (long long)P.yh * P.yw;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_buf = p < hw_out;
  const bool valid = in_buf && (int)(p / P.yw) < P.ho && (int)(p % P.yw) < P.wo;
  float acc[kStemCout];
#pragma unroll
  for (int c = 0; c < kStemCout; ++c) acc[c] = 0.f;
  if (valid) {
    const int oy = (int)(p / P.yw), ox = (int)(p - (long long)oy * P.yw);
    const float* xin = P.x + (long long)img * P.cin * P.h * P.w;
    for (int ci = 0; ci < P.cin; ++ci)
      for (int kh = 0; kh < P.k; ++kh) {
        const int iy = oy * P.stride + kh - P.pad;
        if (iy < 0 || iy >= P.h) continue;
        for (int kw = 0; kw < P.k; ++kw) {
          const int ix = ox * P.stride + kw - P.pad;
          if (ix < 0 || ix >= P.w) continue;
          const float xv = __ldg(xin + ((long long)ci * P.h + iy) * P.w + ix);
          const float* wr = sw + ((ci * P.k + kh) * P.k + kw) * kStemCout;
#pragma unroll
          for (int c = 0; c < kStemCout; ++c) acc[c] = fmaf(xv, wr[c], acc[c]);
        }
      }
  }
  __nv_bfloat16* yp = P.y + ((long long)img * hw_out + p) * P.y_ld;
  if (P.epi == UAVDET_EPI_STATS) {
    if (valid) {
#pragma unroll
      for (int c = 0; c < kStemCout; c += 8) {
        uint4 o;
        o.x = pack_bf16x2(acc[c], acc[c + 1]); o.y = pack_bf16x2(acc[c + 2], acc[c + 3]);
        o.z = pack_bf16x2(acc[c + 4], acc[c + 5]); o.w = pack_bf16x2(acc[c + 6], acc[c + 7]);
        *reinterpret_cast<uint4*>(yp + c) = o;
      }
    }
    float sq[kStemCout];
#pragma unroll
    for (int c = 0; c < kStemCout; ++c) sq[c] = acc[c] * acc[c];  // invalid lanes hold zeros
    const float s1 = warp_colsum32(acc, lane);
    const float s2 = warp_colsum32(sq, lane);
    atomicAdd(&ssum[0][lane], s1);
    atomicAdd(&ssum[1][lane], s2);
    __syncthreads();
    if (threadIdx.x < kStemCout) {
      atomicAdd(P.sum + threadIdx.x, ssum[0][threadIdx.x]);
      atomicAdd(P.sumsq + threadIdx.x, ssum[1][threadIdx.x]);
    }
  } else if (in_buf) {
#pragma unroll
    for (int c = 0; c < kStemCout; ++c) {
      float z = acc[c];
      if (P.scale) z *= __ldg(P.scale + c);
      if (P.shift) z += __ldg(P.shift + c);
      acc[c] = valid ? act_fwd_rt(P.act, z) : 0.f;   // padded last row / column stays zero
    }
#pragma unroll
    for (int c = 0; c < kStemCout; c += 8) {
      uint4 o;
      o.x = pack_bf16x2(acc[c], acc[c + 1]); o.y = pack_bf16x2(acc[c + 2], acc[c + 3]);
      o.z = pack_bf16x2(acc[c + 4], acc[c + 5]); o.w = pack_bf16x2(acc[c + 6], acc[c + 7]);
      *reinterpret_cast<uint4*>(yp + c) = o;
    }
  }
}

// dW[co][ci][kh][kw] += sum_p dy[p][co] * x[p + tap][ci]   (fp32 CUDA cores: K <= 75, the layer is
// bound by streaming dy once).  Persistent blocks walk 32x8 output-pixel tiles; the fp32 input patch
// (with halo) and the bf16 dy tile are staged in shared memory; thread (warp w, lane l) owns the
// output-channel pair 2*(l%16) and the taps {16*j + 2*w + l/16}: dy reads are conflict-free, x reads
// are broadcasts, and the 2 x TPG accumulators stay in registers across all tiles of the block.
constexpr int kSwTileW = 32, kSwTileH = 8, kSwThreads = 256, kSwMaxTpg = 5;

template <int TPG>
__global__ void __launch_bounds__(kSwThreads) stem_wgrad_kernel(const float* __restrict__ x, int n, int cin, int h,
                                                                int w, const __nv_bfloat16* __restrict__ dy,
                                                                long long dy_ld, int k, int stride, int pad, int ho,
                                                                int wo, float* __restrict__ grad, int per_sample) {
  extern __shared__ float sw_smem[];
  const int K = cin * k * k;
  const int pw = (kSwTileW - 1) * stride + k, ph = (kSwTileH - 1) * stride + k;  // input patch
  float* xs = sw_smem;                                                            // [cin][ph][pw]
  uint32_t* dys = reinterpret_cast<uint32_t*>(sw_smem + ((cin * ph * pw + 3) & ~3));  // [256 px][16 co pairs], 16 B aligned
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cp = lane & 15;                    // output-channel pair
  const int slot = 2 * warp + (lane >> 4);     // 0..15
  // per-thread taps and their patch offsets
  int tap_off[TPG];
  bool tap_ok[TPG];
#pragma unroll
  for (int j = 0; j < TPG; ++j) {
    const int t = 16 * j + slot;               // tap id = (ci*k + kh)*k + kw
    tap_ok[j] = t < K;
    const int tt = tap_ok[j] ? t : 0;
    const int ci = tt / (k * k), r = tt - ci * k * k, kh = r / k, kw = r - kh * k;
    tap_off[j] = (ci * ph + kh) * pw + kw;
  }
  float acc0[TPG], acc1[TPG];
#pragma unroll
  for (int j = 0; j < TPG; ++j) { acc0[j] = 0.f; acc1[j] = 0.f; }

  const int tiles_w = (wo + kSwTileW - 1) / kSwTileW, tiles_h = (ho + kSwTileH - 1) / kSwTileH;
  const int tiles_img = tiles_w * tiles_h;
  const int img_lo = per_sample ? blockIdx.y : 0, img_hi = per_sample ? blockIdx.y + 1 : n;
  const long long total = (long long)(img_hi - img_lo) * tiles_img;
  for (long long tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int img = img_lo + (int)(tile / tiles_img);
    const int tr = (int)(tile % tiles_img);
    const int ty = tr / tiles_w, tx = tr - ty * tiles_w;
    const int oy0 = ty * kSwTileH, ox0 = tx * kSwTileW;
    const int iy0 = oy0 * stride - pad, ix0 = ox0 * stride - pad;
    __syncthreads();  // previous tile fully consumed
    const float* xin = x + (long long)img * cin * h * w;
    for (int i = threadIdx.x; i < cin * ph * pw; i += kSwThreads) {
      const int ci = i / (ph * pw), r = i - ci * ph * pw, yy = r / pw, xx = r - yy * pw;
      const int iy = iy0 + yy, ix = ix0 + xx;
      xs[i] = (iy >= 0 && iy < h && ix >= 0 && ix < w) ? __ldg(xin + ((long long)ci * h + iy) * w + ix) : 0.f;
    }
    // dy tile: 256 pixels x 32 channels bf16 = 64 B per pixel = 4 x 16 B
    for (int i = threadIdx.x; i < kSwTileW * kSwTileH * 4; i += kSwThreads) {
      const int p = i >> 2, q = i & 3;
      const int oy = oy0 + p / kSwTileW, ox = ox0 + p % kSwTileW;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (oy < ho && ox < wo)
        v = __ldg(reinterpret_cast<const uint4*>(dy + (((long long)img * ho + oy) * wo + ox) * dy_ld) + q);
      reinterpret_cast<uint4*>(dys)[i] = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int p = 0; p < kSwTileW * kSwTileH; ++p) {
      const uint32_t g2 = dys[p * 16 + cp];
      const float g0 = bf16_lo(g2), g1 = bf16_hi(g2);
      const int base = ((p / kSwTileW) * stride) * pw + (p % kSwTileW) * stride;
#pragma unroll
      for (int j = 0; j < TPG; ++j) {
        const float xv = xs[base + tap_off[j]];
        acc0[j] = fmaf(g0, xv, acc0[j]);
        acc1[j] = fmaf(g1, xv, acc1[j]);
      }
    }
  }
  float* gdst = grad + (per_sample ? (long long)blockIdx.y * kStemCout * K : 0);
#pragma unroll
  for (int j = 0; j < TPG; ++j) {
    if (tap_ok[j]) {
      const int t = 16 * j + slot;
      atomicAdd(gdst + (2 * cp) * K + t, acc0[j]);
      atomicAdd(gdst + (2 * cp + 1) * K + t, acc1[j]);
    }
  }
}

// ---- 1x1 stems (DySOEM_SimFPN.py:14-33: 1- or 3-channel input -> 32 channels) --------------------------------------
// Pure streaming layers (12 B in, 64 B out per pixel).  Four lanes share a pixel, each owning 8 output channels, so a
// warp writes (or, for the weight gradient, reads) 512 contiguous bytes per instruction; a thread walks pixels with a
// grid-stride loop and keeps its statistics / gradient partial sums in registers until the end.  (The generic kernels
// above — a pixel and all 32 channels per thread, a full warp-wide butterfly of the statistics per pixel; a 32x8 tile
// with 16 tap slots of which a 1x1 kernel fills 3 — ran this layer at 25 % / 10 % of the HBM rate.)
template <int CIN, bool STATS>
__global__ void __launch_bounds__(256, 3) stem1x1_fwd_kernel(StemParams P) {
  __shared__ float red[2][8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cg = (lane & 3) << 3;                    // first of this thread's 8 output channels
  // one instance per epilogue: the statistics sums and the affine coefficients never live in registers together
  // (one kernel for both sat at 125 registers, two blocks per SM)
  float wr[8][CIN], sc[STATS ? 1 : 8], sh[STATS ? 1 : 8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) wr[j][ci] = __ldg(P.wgt + (cg + j) * CIN + ci);
    if constexpr (!STATS) {
      sc[j] = P.scale ? __ldg(P.scale + cg + j) : 1.f;
      sh[j] = P.shift ? __ldg(P.shift + cg + j) : 0.f;
    }
  }
  float s1[STATS ? 8 : 1], s2[STATS ? 8 : 1];
  if constexpr (STATS) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
  }
  const long long hw = (long long)P.h * P.w, total = hw * P.n;
  // A warp takes 32 consecutive pixels per iteration: lane l loads the CIN inputs of pixel l (one coalesced 128-byte
  // request per channel plane), shuffles hand pixel 8u + (lane >> 2) to its four lanes, and the warp issues four
  // 512-byte stores.  The inputs of the NEXT chunk are requested before the current one is computed: with one pixel
  // per thread and iteration (the first version) the ~80-register kernel had 768 threads x 16 bytes in flight per SM
  // and ran at 1.9 TB/s.  (image, pixel) of the chunk base advance incrementally: no 64-bit divide in the loop.
  const long long step = (long long)gridDim.x * 256;                 // pixels per grid sweep (8 warps x 32 per block)
  long long p0 = ((long long)blockIdx.x * 8 + warp) * 32;            // first pixel of this warp's chunk
  long long img = p0 / hw, px = p0 - img * hw;
  const long long step_img = step / hw, step_px = step - step_img * hw;
  auto load_chunk = [&](long long base, long long bimg, long long bpx, float (&xv)[CIN]) {
    long long li = bimg, lp = bpx + lane;
    while (lp >= hw) { lp -= hw; ++li; }
    const bool ok = base + lane < total;
    const float* xin = P.x + li * CIN * hw + lp;
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) xv[ci] = ok ? __ldcs(xin + ci * hw) : 0.f;
  };
  float xcur[CIN], xnext[CIN];
  if (p0 < total) load_chunk(p0, img, px, xcur);
  for (; p0 < total; p0 += step) {
    long long nimg = img + step_img, npx = px + step_px;
    if (npx >= hw) { npx -= hw; ++nimg; }
    if (p0 + step < total) load_chunk(p0 + step, nimg, npx, xnext);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int src = 8 * u + (lane >> 2);
      float xv[CIN];
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) xv[ci] = __shfl_sync(0xffffffffu, xcur[ci], src);
      const long long p = p0 + src;
      if (p < total) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float a = 0.f;
#pragma unroll
          for (int ci = 0; ci < CIN; ++ci) a = fmaf(xv[ci], wr[j][ci], a);
          v[j] = a;
        }
        if constexpr (STATS) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { s1[j] += v[j]; s2[j] = fmaf(v[j], v[j], s2[j]); }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = act_fwd_rt(P.act, fmaf(v[j], sc[j], sh[j]));
        }
        uint4 o;
        o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
        o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
        *reinterpret_cast<uint4*>(P.y + p * P.y_ld + cg) = o;
      }
    }
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) xcur[ci] = xnext[ci];
    img = nimg; px = npx;
  }
  if constexpr (STATS) {
    // lanes with equal (lane & 3) hold the same channels: fold the 8 pixel lanes, then the 8 warps
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a = s1[j], b = s2[j];
#pragma unroll
      for (int off = 4; off < 32; off <<= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, off);
        b += __shfl_xor_sync(0xffffffffu, b, off);
      }
      if (lane < 4) { red[0][warp][cg + j] = a; red[1][warp][cg + j] = b; }
    }
    __syncthreads();
    if (threadIdx.x < 64) {
      const int which = threadIdx.x >> 5, ch = threadIdx.x & 31;
      float a = 0.f;
      for (int wi = 0; wi < 8; ++wi) a += red[which][wi][ch];
      atomicAdd((which ? P.sumsq : P.sum) + ch, a);
    }
  }
}

// dW[co][ci] += sum_p dy[p][co] * x[p][ci] for a 1x1 stem, same thread layout.
template <int CIN>
__global__ void __launch_bounds__(256)
stem1x1_wgrad_kernel(const float* __restrict__ x, int n, long long hw, const __nv_bfloat16* __restrict__ dy,
                     long long dy_ld, float* __restrict__ grad) {
  __shared__ float red[8][32][CIN];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cg = (lane & 3) << 3;
  float acc[8][CIN];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) acc[j][ci] = 0.f;
  const long long total = hw * n;
  const long long step = ((long long)gridDim.x * 256) >> 2;
  long long p = ((long long)blockIdx.x * 256 + threadIdx.x) >> 2;
  long long img = p / hw, px = p - img * hw;
  const long long step_img = step / hw, step_px = step - step_img * hw;
  for (; p < total; p += step, img += step_img, px += step_px) {
    if (px >= hw) { px -= hw; ++img; }
    const float* xin = x + img * CIN * hw + px;
    float xv[CIN];
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) xv[ci] = __ldcs(xin + ci * hw);
    const uint4 g = __ldcs(reinterpret_cast<const uint4*>(dy + p * dy_ld + cg));
    const float gv[8] = {bf16_lo(g.x), bf16_hi(g.x), bf16_lo(g.y), bf16_hi(g.y),
                         bf16_lo(g.z), bf16_hi(g.z), bf16_lo(g.w), bf16_hi(g.w)};
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) acc[j][ci] = fmaf(gv[j], xv[ci], acc[j][ci]);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) {
      float a = acc[j][ci];
#pragma unroll
      for (int off = 4; off < 32; off <<= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
      if (lane < 4) red[warp][cg + j][ci] = a;
    }
  __syncthreads();
  if (threadIdx.x < 32 * CIN) {
    const int ch = threadIdx.x / CIN, ci = threadIdx.x - ch * CIN;
    float a = 0.f;
    for (int wi = 0; wi < 8; ++wi) a += red[wi][ch][ci];
    atomicAdd(grad + ch * CIN + ci, a);
  }
}

// ---- 5x5 stride-2 stem (RTMUAVDet.py:28-36: 3 -> 32, pad 1), inference epilogue ---------------------------------------
// Two horizontally adjacent output pixels per thread: the 7 input columns they need are loaded once per filter row, and
// every 16-byte read of the filter from shared memory feeds 8 FMAs (the one-pixel kernel above issued one 4-byte shared
// load per FMA and ran at a quarter of the fp32 rate: 3.5 ms of the 24 ms batch-128 forward).
__global__ void __launch_bounds__(256) stem5x5s2_fwd_kernel(StemParams P) {
  __shared__ float4 sw4[75 * 8];          // [tap = (ci*5 + kh)*5 + kw][32 output channels]
  const int img = blockIdx.y;
  float* sw = reinterpret_cast<float*>(sw4);
  for (int i = threadIdx.x; i < 75 * kStemCout; i += blockDim.x) {
    const int co = i % kStemCout, kk = i / kStemCout;
    sw[kk * kStemCout + co] = P.wgt[co * 75 + kk];
  }
  __syncthreads();
  const int half_w = P.yw >> 1;
  const long long pairs = (long long)P.yh * half_w;
  const long long pp = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pp >= pairs) return;
  const int oy = (int)(pp / half_w), ox = 2 * (int)(pp - (long long)oy * half_w);
  float acc[2][kStemCout];
#pragma unroll
  for (int c = 0; c < kStemCout; ++c) { acc[0][c] = 0.f; acc[1][c] = 0.f; }
  const float* xin = P.x + (long long)img * 3 * P.h * P.w;
  if (oy < P.ho) {
    for (int ci = 0; ci < 3; ++ci)
      for (int kh = 0; kh < 5; ++kh) {
        const int iy = 2 * oy + kh - P.pad;
        if (iy < 0 || iy >= P.h) continue;
        const float* xrow = xin + ((long long)ci * P.h + iy) * P.w;
        float col[7];
#pragma unroll
        for (int c = 0; c < 7; ++c) {
          const int ix = 2 * ox + c - P.pad;
          col[c] = (ix >= 0 && ix < P.w) ? __ldg(xrow + ix) : 0.f;
        }
        const float4* wrow = sw4 + ((ci * 5 + kh) * 5) * 8;
#pragma unroll
        for (int kw = 0; kw < 5; ++kw) {
          const float x0 = col[kw], x1 = col[kw + 2];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 w4 = wrow[kw * 8 + q];
            acc[0][4 * q + 0] = fmaf(x0, w4.x, acc[0][4 * q + 0]); acc[1][4 * q + 0] = fmaf(x1, w4.x, acc[1][4 * q + 0]);
            acc[0][4 * q + 1] = fmaf(x0, w4.y, acc[0][4 * q + 1]); acc[1][4 * q + 1] = fmaf(x1, w4.y, acc[1][4 * q + 1]);
            acc[0][4 * q + 2] = fmaf(x0, w4.z, acc[0][4 * q + 2]); acc[1][4 * q + 2] = fmaf(x1, w4.z, acc[1][4 * q + 2]);
            acc[0][4 * q + 3] = fmaf(x0, w4.w, acc[0][4 * q + 3]); acc[1][4 * q + 3] = fmaf(x1, w4.w, acc[1][4 * q + 3]);
          }
        }
      }
  }
#pragma unroll
  for (int px = 0; px < 2; ++px) {
    const bool valid = oy < P.ho && ox + px < P.wo;     // the padded last row / column stays zero
    __nv_bfloat16* yp = P.y + (((long long)img * P.yh + oy) * P.yw + ox + px) * P.y_ld;
#pragma unroll
    for (int c = 0; c < kStemCout; c += 8) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float z = acc[px][c + j];
        if (P.scale) z *= __ldg(P.scale + c + j);
        if (P.shift) z += __ldg(P.shift + c + j);
        v[j] = valid ? act_fwd_rt(P.act, z) : 0.f;
      }
      uint4 o;
      o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
      o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
      *reinterpret_cast<uint4*>(yp + c) = o;
    }
  }
}

// im2col of the cin<=3 network input: y[n][oy][ox][(ci*k + kh)*k + kw] = x[n][ci][oy*s+kh-p][ox*s+kw-p] (zero outside
// the image and for the pad channels K..31), bf16 NHWC with 32 channels = 64 B per pixel.  With it the stem
// convolution, its per-sample dynamic variant and their weight gradients run on the same tcgen05 kernels as
// every other layer (a 1x1 conv over 32 "channels") instead of CUDA-core direct kernels.
template <int CIN, int KS>
__global__ void __launch_bounds__(256)
im2col_stem_kernel(const float* __restrict__ x, int h, int w, int stride, int pad, int ho, int wo,
                   __nv_bfloat16* __restrict__ y, long long y_ld) {
  const int img = blockIdx.y;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= (long long)ho * wo) return;
  const int oy = (int)(p / wo), ox = (int)(p - (long long)oy * wo);
  const float* xin = x + (long long)img * CIN * h * w;
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = 0.f;
#pragma unroll
  for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
    for (int kh = 0; kh < KS; ++kh) {
      const int iy = oy * stride + kh - pad;
      const bool row_ok = iy >= 0 && iy < h;
#pragma unroll
      for (int kw = 0; kw < KS; ++kw) {
        const int ix = ox * stride + kw - pad;
        if (row_ok && ix >= 0 && ix < w) v[(ci * KS + kh) * KS + kw] = __ldg(xin + ((long long)ci * h + iy) * w + ix);
      }
    }
  __nv_bfloat16* yp = y + ((long long)img * ho * wo + p) * y_ld;
#pragma unroll
  for (int c = 0; c < 32; c += 8) {
    uint4 o;
    o.x = pack_bf16x2(v[c], v[c + 1]); o.y = pack_bf16x2(v[c + 2], v[c + 3]);
    o.z = pack_bf16x2(v[c + 4], v[c + 5]); o.w = pack_bf16x2(v[c + 6], v[c + 7]);
    *reinterpret_cast<uint4*>(yp + c) = o;
  }
}

// Space-to-depth(2) of the 3-channel NCHW fp32 network input into NHWC bf16 with 32 channels:
//   y[n][by][bx][(py*2 + px)*3 + ci] = x[n][ci][2*by + py][2*bx + px]   (channels 12..31 zero).
// A k x k stride-2 stem becomes a ceil(k/2)+1-tap stride-1 convolution over this map (RTMUAVDet's 5x5 s2 p1 stem =
// a 3x3 pad-1 convolution with 12 live input channels), which runs on the implicit-GEMM kernel.
__global__ void __launch_bounds__(256)
stem_s2d_pack_kernel(const float* __restrict__ x, int h, int w, __nv_bfloat16* __restrict__ y, long long y_ld) {
  const int img = blockIdx.y;
  const int hs = h >> 1, ws = w >> 1;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= (long long)hs * ws) return;
  const int by = (int)(p / ws), bx = (int)(p - (long long)by * ws);
  const float* xin = x + (long long)img * 3 * h * w;
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = 0.f;
#pragma unroll
  for (int ci = 0; ci < 3; ++ci)
#pragma unroll
    for (int py = 0; py < 2; ++py) {
      const float2 t = __ldcs(reinterpret_cast<const float2*>(xin + ((long long)ci * h + 2 * by + py) * w + 2 * bx));
      v[(py * 2 + 0) * 3 + ci] = t.x;
      v[(py * 2 + 1) * 3 + ci] = t.y;
    }
  uint4* dst = reinterpret_cast<uint4*>(y + ((long long)img * hs * ws + p) * y_ld);
  uint4 o;
  o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]); o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
  dst[0] = o;
  o.x = pack_bf16x2(v[8], v[9]); o.y = pack_bf16x2(v[10], v[11]); o.z = 0u; o.w = 0u;
  dst[1] = o;
  dst[2] = make_uint4(0u, 0u, 0u, 0u);
  dst[3] = make_uint4(0u, 0u, 0u, 0u);
}

}  // namespace uavdet

using namespace uavdet;

extern "C" int uavdet_stem_s2d_pack(const float* x_nchw, int n, int h, int w, const uavdet_act* y, void* stream) {
  UAVDET_CHECK_ARG(x_nchw && y && y->ptr && n > 0 && h > 0 && w > 0 && h % 2 == 0 && w % 2 == 0,
                   "stem_s2d_pack: bad arguments (even H, W)");
  UAVDET_CHECK_ARG(y->n == n && y->h == h / 2 && y->w == w / 2 && y->c == 32, "stem_s2d_pack: output must be (n,H/2,W/2,32)");
  UAVDET_CHECK_ARG(y->ld % 8 == 0 && ((uintptr_t)y->ptr & 15) == 0 && ((uintptr_t)x_nchw & 7) == 0, "stem_s2d_pack: alignment");
  dim3 grid((unsigned)ceil_div64((long long)(h / 2) * (w / 2), 256), (unsigned)n);
  stem_s2d_pack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x_nchw, h, w, (__nv_bfloat16*)y->ptr, y->ld);
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}

extern "C" int uavdet_stem_fwd(const float* x_nchw, int n, int cin, int h, int w, const float* w_oihw, int cout,
                               int k, int stride, int pad, const uavdet_act* y, const uavdet_epilogue* epi,
                               void* stream) {
  UAVDET_CHECK_ARG(x_nchw && w_oihw && y && y->ptr, "stem_fwd: null pointer");
  UAVDET_CHECK_ARG(cout == kStemCout, "stem_fwd: cout must be %d (got %d)", kStemCout, cout);
  UAVDET_CHECK_ARG(cin >= 1 && cin <= 3 && cin * k * k <= kStemMaxK, "stem_fwd: cin=%d k=%d unsupported", cin, k);
  StemParams P{};
  P.x = x_nchw; P.n = n; P.cin = cin; P.h = h; P.w = w; P.wgt = w_oihw;
  P.w_batch = (epi && epi->head_anchors < 0) ? n : 1;  // head_anchors = -1 flags per-sample stem weights
  P.k = k; P.stride = stride; P.pad = pad;
  P.ho = (h + 2 * pad - k) / stride + 1;
  P.wo = (w + 2 * pad - k) / stride + 1;
  const bool padded = (y->h == P.ho + 1 && y->w == P.wo + 1);
  UAVDET_CHECK_ARG(y->n == n && y->c == cout && ((y->h == P.ho && y->w == P.wo) || padded),
                   "stem_fwd: output view (%d,%d) != conv output (%d,%d) [or +1 zero-padded]", y->h, y->w, P.ho, P.wo);
  P.yh = y->h; P.yw = y->w;
  UAVDET_CHECK_ARG(y->ld % 8 == 0 && ((uintptr_t)y->ptr & 15) == 0, "stem_fwd: output alignment");
  P.y = (__nv_bfloat16*)y->ptr; P.y_ld = y->ld;
  P.epi = epi ? epi->epi : UAVDET_EPI_AFFINE;
  P.act = epi ? epi->act : UAVDET_ACT_NONE;
  P.scale = epi ? epi->scale : nullptr; P.shift = epi ? epi->shift : nullptr;
  P.sum = epi ? epi->sum : nullptr; P.sumsq = epi ? epi->sumsq : nullptr;
  if (P.epi == UAVDET_EPI_STATS) UAVDET_CHECK_ARG(P.sum && P.sumsq, "stem_fwd: STATS needs sum/sumsq");
  UAVDET_CHECK_ARG(P.epi != UAVDET_EPI_HEAD, "stem_fwd: HEAD epilogue unsupported");
  UAVDET_CHECK_ARG(!(padded && P.epi == UAVDET_EPI_STATS), "stem_fwd: zero-padded output needs the AFFINE epilogue");
  cudaStream_t st = (cudaStream_t)stream;
  if (k == 1 && stride == 1 && pad == 0 && P.w_batch == 1 && !padded && (cin == 1 || cin == 3)) {
    // streaming 1x1 stem: 4 lanes per pixel, grid-stride
    long long blocks = ceil_div64((long long)n * h * w, 256 * 4);
    if (blocks > (long long)kNumSMs * 8) blocks = (long long)kNumSMs * 8;
    if (blocks < 1) blocks = 1;
    const bool stats = P.epi == UAVDET_EPI_STATS;
    if (cin == 1 && stats) stem1x1_fwd_kernel<1, true><<<(unsigned)blocks, 256, 0, st>>>(P);
    else if (cin == 1) stem1x1_fwd_kernel<1, false><<<(unsigned)blocks, 256, 0, st>>>(P);
    else if (stats) stem1x1_fwd_kernel<3, true><<<(unsigned)blocks, 256, 0, st>>>(P);
    else stem1x1_fwd_kernel<3, false><<<(unsigned)blocks, 256, 0, st>>>(P);
    UAVDET_LAUNCH_CHECK();
    return UAVDET_OK;
  }
  if (k == 5 && stride == 2 && cin == 3 && P.w_batch == 1 && P.epi == UAVDET_EPI_AFFINE && (P.yw & 1) == 0) {
    dim3 grid2((unsigned)ceil_div64((long long)P.yh * (P.yw / 2), 256), (unsigned)n);
    stem5x5s2_fwd_kernel<<<grid2, 256, 0, st>>>(P);
    UAVDET_LAUNCH_CHECK();
    return UAVDET_OK;
  }
  dim3 grid((unsigned)ceil_div64((long long)P.yh * P.yw, 256), (unsigned)n);
  stem_fwd_kernel<<<grid, 256, 0, st>>>(P);
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}

extern "C" int uavdet_stem_wgrad(const float* x_nchw, int n, int cin, int h, int w, const uavdet_act* dy, int k,
                                 int stride, int pad, float* grad_oihw, void* stream) {
  UAVDET_CHECK_ARG(x_nchw && dy && dy->ptr && grad_oihw, "stem_wgrad: null pointer");
  UAVDET_CHECK_ARG(dy->c == kStemCout && cin >= 1 && cin <= 3 && cin * k * k <= kStemMaxK, "stem_wgrad: unsupported shape");
  const int per_sample = n < 0 ? 1 : 0;  // n < 0 flags per-sample gradients ([|n|][32][K])
  if (n < 0) n = -n;
  const int K = cin * k * k;
  const int tpg = (K + 15) / 16;
  UAVDET_CHECK_ARG(tpg >= 1 && tpg <= kSwMaxTpg, "stem_wgrad: K=%d unsupported", K);
  UAVDET_CHECK_ARG(dy->ld % 8 == 0 && ((uintptr_t)dy->ptr & 15) == 0, "stem_wgrad: dy alignment");
  if (k == 1 && stride == 1 && pad == 0 && !per_sample && (cin == 1 || cin == 3) && dy->h == h && dy->w == w) {
    long long blocks = ceil_div64((long long)n * h * w * 4, 256 * 8);
    if (blocks > (long long)kNumSMs * 8) blocks = (long long)kNumSMs * 8;
    if (blocks < 1) blocks = 1;
    const __nv_bfloat16* dyq = (const __nv_bfloat16*)dy->ptr;
    if (cin == 1)
      stem1x1_wgrad_kernel<1><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x_nchw, n, (long long)h * w, dyq, dy->ld, grad_oihw);
    else
      stem1x1_wgrad_kernel<3><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x_nchw, n, (long long)h * w, dyq, dy->ld, grad_oihw);
    UAVDET_LAUNCH_CHECK();
    return UAVDET_OK;
  }
  const int pw = (kSwTileW - 1) * stride + k, ph = (kSwTileH - 1) * stride + k;
  const size_t smem = sizeof(float) * (size_t)((cin * ph * pw + 3) & ~3) + (size_t)kSwTileW * kSwTileH * 64;
  UAVDET_CHECK_ARG(smem <= 48 * 1024, "stem_wgrad: patch does not fit shared memory");
  const long long tiles = (long long)ceil_div(dy->w, kSwTileW) * ceil_div(dy->h, kSwTileH) * (per_sample ? 1 : n);
  int bx = (int)(tiles < (long long)kNumSMs * 4 ? tiles : (long long)kNumSMs * 4);
  if (per_sample) bx = (int)(tiles < 16 ? tiles : 16);
  if (bx < 1) bx = 1;
  dim3 grid((unsigned)bx, (unsigned)(per_sample ? n : 1));
  cudaStream_t st = (cudaStream_t)stream;
  const __nv_bfloat16* dyp = (const __nv_bfloat16*)dy->ptr;
#define UAVDET_SW_LAUNCH(T)                                                                                   \
  stem_wgrad_kernel<T><<<grid, kSwThreads, smem, st>>>(x_nchw, n, cin, h, w, dyp, dy->ld, k, stride, pad, dy->h, \
                                                       dy->w, grad_oihw, per_sample)
  switch (tpg) {
    case 1: UAVDET_SW_LAUNCH(1); break;
    case 2: UAVDET_SW_LAUNCH(2); break;
    case 3: UAVDET_SW_LAUNCH(3); break;
    case 4: UAVDET_SW_LAUNCH(4); break;
    default: UAVDET_SW_LAUNCH(5); break;
  }
#undef UAVDET_SW_LAUNCH
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}

extern "C" int uavdet_im2col_stem(const float* x_nchw, int n, int cin, int h, int w, int k, int stride, int pad,
                                  const uavdet_act* y, void* stream) {
  UAVDET_CHECK_ARG(x_nchw && y && y->ptr, "im2col_stem: null pointer");
  UAVDET_CHECK_ARG(cin >= 1 && cin * k * k <= 32 && (stride == 1 || stride == 2), "im2col_stem: cin*k*k must be <= 32");
  const int ho = (h + 2 * pad - k) / stride + 1, wo = (w + 2 * pad - k) / stride + 1;
  UAVDET_CHECK_ARG(y->n == n && y->h == ho && y->w == wo && y->c == 32, "im2col_stem: output view must be (n,%d,%d,32)", ho, wo);
  UAVDET_CHECK_ARG(y->ld % 8 == 0 && ((uintptr_t)y->ptr & 15) == 0, "im2col_stem: output alignment");
  dim3 grid((unsigned)ceil_div64((long long)ho * wo, 256), (unsigned)n);
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* yp = (__nv_bfloat16*)y->ptr;
#define UAVDET_I2C(CI, KS) im2col_stem_kernel<CI, KS><<<grid, 256, 0, st>>>(x_nchw, h, w, stride, pad, ho, wo, yp, y->ld)
  if (cin == 3 && k == 3) UAVDET_I2C(3, 3);
  else if (cin == 1 && k == 3) UAVDET_I2C(1, 3);
  else if (cin == 3 && k == 1) UAVDET_I2C(3, 1);
  else if (cin == 1 && k == 1) UAVDET_I2C(1, 1);
  else if (cin == 1 && k == 5) UAVDET_I2C(1, 5);
  else if (cin == 2 && k == 3) UAVDET_I2C(2, 3);
  else { UAVDET_CHECK_ARG(false, "im2col_stem: (cin=%d, k=%d) not instantiated", cin, k); }
#undef UAVDET_I2C
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}
