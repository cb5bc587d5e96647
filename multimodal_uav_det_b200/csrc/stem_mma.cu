// K1s — the cin <= 3 stem convolution (3x3, K = cin*k*k <= 32) on tcgen05 without a patch tensor in HBM.
//
// Replaces the first nn.Conv2d of BaselineModel.py:89-97 / DyYOLO.py:89-100 (per-sample kernels) and its weight
// gradient.  The layer moves 12 B in and 64 B out per pixel and has 27 MACs per output value: far too thin for the
// generic implicit GEMM (its A operand needs >= 32 bf16 channels per TMA box), so round 1 materialised an im2col patch
// tensor (n, h, w, 32) bf16 — 839 MB written and read twice per step at batch 32 / 640x640, 2.7x the layer's
// algorithmic bytes.  Here the patch tile is built in shared memory:
//   forward:  thread r of a 128-thread CTA gathers the 27 fp32 taps of output pixel r straight from the NCHW input
//             (coalesced along x, the shifted re-reads hit L1), writes them as one 64-byte bf16 row of a K-major
//             SWIZZLE_64B tile, one thread issues two M128 x N32 x K16 MMAs against the 2 KB weight tile, every warp
//             drains its 32 TMEM lanes, keeps per-thread partial sums of the batch statistics in registers across all
//             of the CTA's tiles (one butterfly + 64 atomics per warp at the very end), stages the bf16 rows and one
//             thread sends the 8 KB tile to global memory with a single bulk copy (128 consecutive pixels x 32
//             channels are contiguous in NHWC).
//   wgrad:    dW[co][kk] = sum_p dy[p][co] * patch[p][kk]: the same patch tile is the MN-major B operand, the dy tile
//             (copied as is, 64-byte rows) the MN-major A operand; K = 128 pixels per tile = 8 MMAs, accumulated in
//             TMEM over all tiles of the CTA, one 32x32 atomic flush per CTA (per image for per-sample gradients).
// No role specialisation: a CTA works in lock step and 3-4 CTAs share an SM (32 TMEM columns, < 42 KB of shared
// memory each), so one CTA's gather overlaps another's MMA / epilogue.
#include "common.cuh"
#include "sm100.cuh"

namespace uavdet {
using namespace sm100;

constexpr int kSmThreads = 128;
constexpr int kSmCtasPerSm = 4;

struct StemMmaParams {
  const float* x; int n, h, w;
  const __nv_bfloat16* wgt; int w_batch;   // forward: [w_batch][32][32] bf16, row = cout, K-major, zero padded
  int stride, pad, ho, wo;
  int tiles_per_img, total_tiles;
  __nv_bfloat16* y; long long y_ld;        // forward output / wgrad dy
  int act;
  const float* scale; const float* shift;
  float* sum; float* sumsq;
  float* dw; int per_sample;               // wgrad: [per_sample ? n : 1][32][32] fp32, accumulated
  unsigned int* watchdog;
};

// Column sums across the 32 lanes of a warp: lane c returns sum_lanes v[c] (31 shuffles instead of 32 x 5).
__device__ __forceinline__ float colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int step = 16; step >= 1; step >>= 1) {
    const bool upper = (lane & step) != 0;
#pragma unroll
    for (int i = 0; i < step; ++i) {
      const float send = upper ? v[i] : v[i + step];
      const float keep = upper ? v[i + step] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, step);
    }
  }
  return v[0];
}

// The CIN*KS*KS taps of output pixel (oy, ox), bf16, in w.flatten(1) order ((ci*KS + kh)*KS + kw), zero padded to 32.
template <int CIN, int KS>
__device__ __forceinline__ void gather_patch(const float* __restrict__ xin, int h, int w, int oy, int ox, int stride,
                                             int pad, bool valid, uint32_t (&a)[16]) {
  constexpr int K = CIN * KS * KS;
  static_assert(K <= 32, "stem_mma: cin*k*k must fit the 32-channel patch row");
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = 0.f;
  const int iy0 = oy * stride - pad, ix0 = ox * stride - pad;
  bool okx[KS];
#pragma unroll
  for (int kw = 0; kw < KS; ++kw) okx[kw] = (unsigned)(ix0 + kw) < (unsigned)w;
#pragma unroll
  for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
    for (int kh = 0; kh < KS; ++kh) {
      const int iy = iy0 + kh;
      const bool oky = valid && (unsigned)iy < (unsigned)h;
      const float* row = xin + ((long long)ci * h + (oky ? iy : 0)) * w + ix0;
#pragma unroll
      for (int kw = 0; kw < KS; ++kw)
        if (oky && okx[kw]) v[(ci * KS + kh) * KS + kw] = __ldg(row + kw);
    }
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = (2 * i < K) ? pack_bf16x2(v[2 * i], v[2 * i + 1]) : 0u;
}

// One 64-byte row (4 x 16 B) of a SWIZZLE_64B tile: 16-byte chunk j of row r lives at chunk j ^ ((r >> 1) & 3).
__device__ __forceinline__ void store_row_sw64(uint8_t* tile, int r, const uint32_t (&a)[16]) {
  const int sw = (r >> 1) & 3;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    *reinterpret_cast<uint4*>(tile + r * 64 + ((j ^ sw) << 4)) = make_uint4(a[4 * j], a[4 * j + 1], a[4 * j + 2], a[4 * j + 3]);
}

__device__ __forceinline__ uint8_t* align1024(uint8_t* p) {
  const uint32_t a = smem_u32(p);
  return p + (((a + 1023u) & ~1023u) - a);
}

__device__ __forceinline__ void bulk_store_1d(void* gdst, uint32_t ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}

// ------------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------------
template <int CIN, int KS, bool kStats>
__global__ void __launch_bounds__(kSmThreads, kSmCtasPerSm) stem_mma_fwd_kernel(const StemMmaParams P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ uint32_t dead_flag;
  __shared__ float s_sc[32], s_sh[32];
  __shared__ float s_stat[2][32];
  uint8_t* sA = align1024(smem_raw);        // [128 pixels][64 B]  K-major SWIZZLE_64B
  uint8_t* sOut = sA + 8192;                // [128 pixels][64 B]  dense (bulk-copied to global memory)
  uint8_t* sB = sOut + 8192;                // [32 cout][64 B]     K-major SWIZZLE_64B
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  volatile uint32_t* dead = &dead_flag;

  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    dead_flag = 0;
    fence_barrier_init();
  }
  if (tid < 32) {
    s_sc[tid] = P.scale ? P.scale[tid] : 1.f;
    s_sh[tid] = P.shift ? P.shift[tid] : 0.f;
    s_stat[0][tid] = 0.f;
    s_stat[1][tid] = 0.f;
  }
  if (warp == 0) { tmem_alloc(smem_u32(&tmem_slot), 32); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = make_idesc_bf16(32, 0, 0, 128);
  const uint64_t adesc = make_smem_desc(smem_u32(sA), 16, 512, 4u);
  const uint64_t bdesc = make_smem_desc(smem_u32(sB), 16, 512, 4u);
  const long long hw = (long long)P.ho * P.wo;
  const bool dense = P.y_ld == 32;

  float s[32], q[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) { s[i] = 0.f; q[i] = 0.f; }
  uint32_t ph = 0;
  int cur_wimg = -1;
  for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
    const int img = tile / P.tiles_per_img;
    const int t = tile - img * P.tiles_per_img;
    const long long p = (long long)t * 128 + tid;
    const bool valid = p < hw;
    const int oy = valid ? (int)(p / P.wo) : 0;
    const int ox = valid ? (int)(p - (long long)oy * P.wo) : 0;
    uint32_t a[16];
    gather_patch<CIN, KS>(P.x + (long long)img * CIN * P.h * P.w, P.h, P.w, oy, ox, P.stride, P.pad, valid, a);
    store_row_sw64(sA, tid, a);
    const int wimg = P.w_batch > 1 ? img : 0;
    if (wimg != cur_wimg) {            // CTA-uniform; the previous tile's MMAs have completed (awaited below)
      cur_wimg = wimg;
      const int row = tid >> 2, ch = tid & 3;
      const uint4 wv = __ldg(reinterpret_cast<const uint4*>(P.wgt + (long long)wimg * 1024 + row * 32 + ch * 8));
      *reinterpret_cast<uint4*>(sB + row * 64 + ((ch ^ ((row >> 1) & 3)) << 4)) = wv;
    }
    fence_proxy_async();
    if (tid == 0) tma_store_wait_read<0>();       // the previous tile's bulk copy has finished reading sOut
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      tc_mma_bf16(tmem, adesc, bdesc, idesc, 0u);
      tc_mma_bf16(tmem, adesc + 2, bdesc + 2, idesc, 1u);     // +32 bytes along K inside the swizzled row
      tc_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), ph, dead, P.watchdog, 0x100u);
    ph ^= 1u;
    tc_fence_after();
    uint32_t r[32];
    tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16), r);
    tmem_ld_wait();
    tc_fence_before();
    uint32_t o[16];
    if (kStats) {
      // rows past the image hold zeros (zero patch rows): they add nothing to the sums
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        o[i] = pack_bf16x2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
        const float lo = bf16_lo(o[i]), hi = bf16_hi(o[i]);       // the values BatchNorm will normalise
        s[2 * i] += lo; q[2 * i] = fmaf(lo, lo, q[2 * i]);
        s[2 * i + 1] += hi; q[2 * i + 1] = fmaf(hi, hi, q[2 * i + 1]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float z0 = fmaf(__uint_as_float(r[2 * i]), s_sc[2 * i], s_sh[2 * i]);
        const float z1 = fmaf(__uint_as_float(r[2 * i + 1]), s_sc[2 * i + 1], s_sh[2 * i + 1]);
        o[i] = pack_bf16x2(act_fwd_rt(P.act, z0), act_fwd_rt(P.act, z1));
      }
    }
    __nv_bfloat16* ytile = P.y + ((long long)img * hw + (long long)t * 128) * P.y_ld;
    if (dense) {
      // Stage the row so that a quarter-warp's 16-byte stores hit 8 different bank groups: in iteration j thread r
      // writes chunk (j + (r >> 1)) & 3 of its row (rows are 64 B apart, i.e. only 2 rows per 128-byte bank line).
      uint4 c4[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) c4[j] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
      const int rot = (tid >> 1) & 3;
      if (rot & 1) { const uint4 t0 = c4[0]; c4[0] = c4[1]; c4[1] = c4[2]; c4[2] = c4[3]; c4[3] = t0; }
      if (rot & 2) { uint4 t0 = c4[0]; c4[0] = c4[2]; c4[2] = t0; t0 = c4[1]; c4[1] = c4[3]; c4[3] = t0; }
#pragma unroll
      for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(sOut + tid * 64 + (((j + rot) & 3) << 4)) = c4[j];
      fence_proxy_async();
      __syncthreads();
      if (tid == 0) {
        const long long rows = hw - (long long)t * 128 < 128 ? hw - (long long)t * 128 : 128;
        bulk_store_1d(ytile, smem_u32(sOut), (uint32_t)(rows * 64));
        tma_store_commit();
      }
    } else if (valid) {
      uint4* dst = reinterpret_cast<uint4*>(ytile + (long long)tid * P.y_ld);
#pragma unroll
      for (int j = 0; j < 4; ++j) dst[j] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
    }
  }
  if (kStats) {
    const float s1 = colsum32(s, lane);
    const float s2 = colsum32(q, lane);
    atomicAdd(&s_stat[0][lane], s1);
    atomicAdd(&s_stat[1][lane], s2);
    __syncthreads();
    if (tid < 32) {
      atomicAdd(P.sum + tid, s_stat[0][tid]);
      atomicAdd(P.sumsq + tid, s_stat[1][tid]);
    }
  }
  if (tid == 0) tma_store_wait_all();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 32); }
}

// ------------------------------------------------------------------------------------------------------------------
// weight gradient
// ------------------------------------------------------------------------------------------------------------------
template <int CIN, int KS>
__global__ void __launch_bounds__(kSmThreads, kSmCtasPerSm) stem_mma_wgrad_kernel(const StemMmaParams P) {
  constexpr int K = CIN * KS * KS;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ uint32_t dead_flag;
  // A operand (dy^T, M = cout): MN-major SWIZZLE_64B, M = 128 = four 32-channel blocks 8 KB apart of which only the
  // first one exists — the other three stay zero (accumulator rows 32..127 are never read).
  uint8_t* sDY = align1024(smem_raw);       // 4 x [128 pixels][64 B]
  uint8_t* sP = sDY + 4 * 8192;             // [128 pixels][64 B] patch rows (B operand, N = 32 taps)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  volatile uint32_t* dead = &dead_flag;

  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    dead_flag = 0;
    fence_barrier_init();
  }
  for (int i = tid; i < 3 * 8192 / 16; i += kSmThreads) reinterpret_cast<uint4*>(sDY + 8192)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (warp == 0) { tmem_alloc(smem_u32(&tmem_slot), 32); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = make_idesc_bf16(32, 1, 1, 128);                  // both operands MN-major
  const uint64_t adesc = make_smem_desc(smem_u32(sDY), 8192, 512, 4u);    // LBO: next 32-channel block; SBO: 8 pixel rows
  const uint64_t bdesc = make_smem_desc(smem_u32(sP), 8192, 512, 4u);
  const long long hw = (long long)P.ho * P.wo;

  // contiguous chunk of the (image, tile) sequence: a CTA crosses an image boundary at most a few times
  const int per_cta = (P.total_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
  const int t_begin = (int)blockIdx.x * per_cta;
  const int t_end = min(t_begin + per_cta, P.total_tiles);
  uint32_t ph = 0;
  bool pending = false;      // MMAs of the previous tile not yet awaited
  bool fresh = true;         // the next MMA starts a new accumulation
  int cur_seg = -1;

  auto flush = [&](int seg) {
    // accumulator rows 0..31 (cout) x columns 0..K-1 (taps) -> dw[seg][co][kk]
    if (warp == 0) {
      uint32_t r[32];
      tmem_ld_32x32(tmem, r);
      tmem_ld_wait();
      float* dst = P.dw + (long long)seg * 1024 + lane * 32;
#pragma unroll
      for (int i = 0; i < K; ++i) atomicAdd(dst + i, __uint_as_float(r[i]));
    }
    tc_fence_before();
  };

  for (int tile = t_begin; tile < t_end; ++tile) {
    const int img = tile / P.tiles_per_img;
    const int t = tile - img * P.tiles_per_img;
    const long long p = (long long)t * 128 + tid;
    const bool valid = p < hw;
    const int oy = valid ? (int)(p / P.wo) : 0;
    const int ox = valid ? (int)(p - (long long)oy * P.wo) : 0;
    uint32_t a[16];
    gather_patch<CIN, KS>(P.x + (long long)img * CIN * P.h * P.w, P.h, P.w, oy, ox, P.stride, P.pad, valid, a);
    // dy rows: instruction i covers rows 32*i + tid/4, 16-byte chunk tid%4 (a warp reads 512 contiguous bytes)
    uint4 g[4];
    const int ch = tid & 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long pp = (long long)t * 128 + 32 * i + (tid >> 2);
      g[i] = make_uint4(0u, 0u, 0u, 0u);
      if (pp < hw) g[i] = __ldg(reinterpret_cast<const uint4*>(P.y + ((long long)img * hw + pp) * P.y_ld) + ch);
    }
    if (pending) {
      mbar_wait(smem_u32(&bar), ph, dead, P.watchdog, 0x200u);
      ph ^= 1u;
      tc_fence_after();
      pending = false;
    }
    const int seg = P.per_sample ? img : 0;
    if (seg != cur_seg) {
      if (cur_seg >= 0) { flush(cur_seg); fresh = true; }
      cur_seg = seg;
    }
    store_row_sw64(sP, tid, a);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = 32 * i + (tid >> 2);
      *reinterpret_cast<uint4*>(sDY + row * 64 + ((ch ^ ((row >> 1) & 3)) << 4)) = g[i];
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)      // 16 pixel rows = 1024 bytes per K step
        tc_mma_bf16(tmem, adesc + (uint64_t)(64 * ks), bdesc + (uint64_t)(64 * ks), idesc, (fresh && ks == 0) ? 0u : 1u);
      tc_commit(smem_u32(&bar));
    }
    fresh = false;
    pending = true;
  }
  if (pending) {
    mbar_wait(smem_u32(&bar), ph, dead, P.watchdog, 0x200u);
    tc_fence_after();
  }
  if (cur_seg >= 0) flush(cur_seg);
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 32); }
}

static int stem_mma_fill(StemMmaParams& P, const float* x_nchw, int n, int cin, int h, int w, int k, int stride, int pad,
                         const uavdet_act* y) {
  UAVDET_CHECK_ARG(x_nchw && y && y->ptr && n > 0 && h > 0 && w > 0, "stem_mma: null pointer / empty input");
  UAVDET_CHECK_ARG((stride == 1 || stride == 2) && pad >= 0, "stem_mma: stride=%d pad=%d unsupported", stride, pad);
  UAVDET_CHECK_ARG(h + 2 * pad >= k && w + 2 * pad >= k, "stem_mma: kernel larger than the padded input");
  P.x = x_nchw; P.n = n; P.h = h; P.w = w;
  P.stride = stride; P.pad = pad;
  P.ho = (h + 2 * pad - k) / stride + 1;
  P.wo = (w + 2 * pad - k) / stride + 1;
  UAVDET_CHECK_ARG(y->n == n && y->h == P.ho && y->w == P.wo && y->c == 32,
                   "stem_mma: NHWC view must be (%d,%d,%d,32), got (%d,%d,%d,%d)", n, P.ho, P.wo, y->n, y->h, y->w, y->c);
  UAVDET_CHECK_ARG(y->ld % 8 == 0 && ((uintptr_t)y->ptr & 15) == 0, "stem_mma: NHWC view must be 16-byte aligned");
  const long long tiles_per_img = ceil_div64((long long)P.ho * P.wo, 128);
  UAVDET_CHECK_ARG(tiles_per_img * n < (1ll << 31), "stem_mma: too many tiles");
  P.tiles_per_img = (int)tiles_per_img;
  P.total_tiles = (int)(tiles_per_img * n);
  P.y = (__nv_bfloat16*)y->ptr; P.y_ld = y->ld;
  P.watchdog = watchdog_word();
  (void)cin;
  return UAVDET_OK;
}

static bool stem_mma_shape_ok(int cin, int k) { return (k == 3 && cin >= 1 && cin <= 3); }

}  // namespace uavdet

using namespace uavdet;

extern "C" int uavdet_stem_mma_supported(int cin, int cout, int k) {
  return (cout == 32 && stem_mma_shape_ok(cin, k)) ? 1 : 0;
}

extern "C" int uavdet_stem_mma_fwd(const float* x_nchw, int n, int cin, int h, int w, const void* w_bf16, int w_batch, int k,
                                   int stride, int pad, const uavdet_act* y, const uavdet_epilogue* epi, void* stream) {
  UAVDET_CHECK_ARG(w_bf16 && ((uintptr_t)w_bf16 & 15) == 0, "stem_mma_fwd: weights must be 16-byte aligned");
  UAVDET_CHECK_ARG(stem_mma_shape_ok(cin, k), "stem_mma_fwd: (cin=%d, k=%d) not instantiated", cin, k);
  UAVDET_CHECK_ARG(w_batch == 1 || w_batch == n, "stem_mma_fwd: w_batch must be 1 or n");
  StemMmaParams P{};
  int rc = stem_mma_fill(P, x_nchw, n, cin, h, w, k, stride, pad, y);
  if (rc) return rc;
  P.wgt = (const __nv_bfloat16*)w_bf16; P.w_batch = w_batch;
  const int e = epi ? epi->epi : UAVDET_EPI_AFFINE;
  UAVDET_CHECK_ARG(e == UAVDET_EPI_AFFINE || e == UAVDET_EPI_STATS, "stem_mma_fwd: epilogue must be AFFINE or STATS");
  P.act = epi ? epi->act : UAVDET_ACT_NONE;
  P.scale = epi ? epi->scale : nullptr; P.shift = epi ? epi->shift : nullptr;
  P.sum = epi ? epi->sum : nullptr; P.sumsq = epi ? epi->sumsq : nullptr;
  const bool stats = e == UAVDET_EPI_STATS;
  if (stats) UAVDET_CHECK_ARG(P.sum && P.sumsq, "stem_mma_fwd: STATS needs sum/sumsq");
  UAVDET_CHECK_ARG(!(epi && epi->res), "stem_mma_fwd: no residual operand");
  const int smem = 8192 + 8192 + 2048 + 1024;
  int grid = kNumSMs * kSmCtasPerSm;
  if (grid > P.total_tiles) grid = P.total_tiles;
  cudaStream_t st = (cudaStream_t)stream;
#define UAVDET_SMF(CI)                                                                            \
  do {                                                                                            \
    if (stats) stem_mma_fwd_kernel<CI, 3, true><<<grid, kSmThreads, smem, st>>>(P);               \
    else stem_mma_fwd_kernel<CI, 3, false><<<grid, kSmThreads, smem, st>>>(P);                    \
  } while (0)
  if (cin == 3) UAVDET_SMF(3);
  else if (cin == 2) UAVDET_SMF(2);
  else UAVDET_SMF(1);
#undef UAVDET_SMF
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}

extern "C" int uavdet_stem_mma_wgrad(const float* x_nchw, int n, int cin, int h, int w, const uavdet_act* dy, int k,
                                     int stride, int pad, float* dw_o32, int per_sample, void* stream) {
  UAVDET_CHECK_ARG(dw_o32, "stem_mma_wgrad: null gradient buffer");
  UAVDET_CHECK_ARG(stem_mma_shape_ok(cin, k), "stem_mma_wgrad: (cin=%d, k=%d) not instantiated", cin, k);
  StemMmaParams P{};
  int rc = stem_mma_fill(P, x_nchw, n, cin, h, w, k, stride, pad, dy);
  if (rc) return rc;
  P.dw = dw_o32; P.per_sample = per_sample ? 1 : 0;
  const int smem = 4 * 8192 + 8192 + 1024;
  static PerDeviceOnce attr_once[3];
  int grid = kNumSMs * kSmCtasPerSm;
  if (grid > P.total_tiles) grid = P.total_tiles;
  cudaStream_t st = (cudaStream_t)stream;
#define UAVDET_SMW(CI)                                                                                                \
  do {                                                                                                                \
    UAVDET_CUDA(attr_once[CI - 1].run([=] {                                                                           \
      return cudaFuncSetAttribute(stem_mma_wgrad_kernel<CI, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);   \
    }));                                                                                                              \
    stem_mma_wgrad_kernel<CI, 3><<<grid, kSmThreads, smem, st>>>(P);                                                  \
  } while (0)
  if (cin == 3) UAVDET_SMW(3);
  else if (cin == 2) UAVDET_SMW(2);
  else UAVDET_SMW(1);
#undef UAVDET_SMW
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}
