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
#include "igemm.h"
#include <stdlib.h>

namespace uavdet {
using namespace sm100;

constexpr int kSmThreads = 128;
constexpr int kSmCtasPerSm = 6;      // weight gradient: 83 registers
constexpr int kSmFwdCtasPerSm = 6;      // forward, affine + activation epilogue: 80 registers
constexpr int kSmFwdStatsCtasPerSm = 8; // forward, statistics epilogue: 62 registers

struct StemMmaParams {
  const float* x; int n, h, w;
  const __nv_bfloat16* wgt; int w_batch;   // forward: [w_batch][32][32] bf16, row = cout, K-major, zero padded
  int stride, pad, ho, wo;
  int tiles_per_img, total_tiles;
  FastDiv fd_wo, fd_tpi;                   // n / wo and n / tiles_per_img without a hardware divide (a 64-bit division per
                                           // tile and thread was a third of the forward kernel's instructions)
  __nv_bfloat16* y; long long y_ld;        // forward output / wgrad dy
  int act;
  const float* scale; const float* shift;
  float* sum; float* sumsq;
  float* dw; int per_sample;               // wgrad: [per_sample ? n : 1][32][32] fp32, accumulated
  unsigned int* watchdog;
  int dbg;                                 // UAVDET_STEM_DBG bits (timing experiments only): 1 no global loads, 2 no MMA, 4 no store, 8 no commit / wait, 16 no TMEM load
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
__device__ __forceinline__ void load_patch(const float* __restrict__ xin, int h, int w, int oy, int ox, int stride,
                                           int pad, bool valid, float (&v)[CIN * KS * KS]) {
  constexpr int K = CIN * KS * KS;
  static_assert(K <= 32, "stem_mma: cin*k*k must fit the 32-channel patch row");
#pragma unroll
  for (int i = 0; i < K; ++i) v[i] = 0.f;
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
}
template <int K>
__device__ __forceinline__ void pack_patch(const float (&v)[K], uint32_t (&a)[16]) {
#pragma unroll
  for (int i = 0; i < 16; ++i)
    a[i] = (2 * i + 1 < K) ? pack_bf16x2(v[2 * i], v[2 * i + 1]) : (2 * i < K ? pack_bf16x2(v[2 * i], 0.f) : 0u);
}
template <int CIN, int KS>
__device__ __forceinline__ void gather_patch(const float* __restrict__ xin, int h, int w, int oy, int ox, int stride,
                                             int pad, bool valid, uint32_t (&a)[16]) {
  float v[CIN * KS * KS];
  load_patch<CIN, KS>(xin, h, w, oy, ox, stride, pad, valid, v);
  pack_patch<CIN * KS * KS>(v, a);
}

__device__ __forceinline__ int fdiv(int n, const FastDiv& f) {
  return f.d == 1 ? n : (int)(__umulhi((uint32_t)n, f.mul) >> f.shr);
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
__global__ void __launch_bounds__(kSmThreads, kStats ? kSmFwdStatsCtasPerSm : kSmFwdCtasPerSm) stem_mma_fwd_kernel(const StemMmaParams P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ uint32_t dead_flag;
  __shared__ float s_sc[32], s_sh[32];
  __shared__ float s_stat[2][32];
  // The kernel is bound by instruction latency, not by a memory or tensor unit (with its loads, MMAs and stores
  // switched off it ran 14 % faster, UAVDET_STEM_DBG): a CTA works through a tile in lock step, so the only latency
  // hiding is other CTAs.  Hence 80 registers and 6 CTAs per SM: the batch statistics are column sums of the staged
  // output tile (4 accumulators per thread) instead of 64 per-thread partial sums, and nothing is prefetched across
  // tiles (a two-deep software pipeline at 152 registers / 3 CTAs per SM measured the same 389 us as the plain loop).
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
  const int hw = P.ho * P.wo;
  const bool dense = P.y_ld == 32;

  float st_s0 = 0.f, st_s1 = 0.f, st_q0 = 0.f, st_q1 = 0.f;   // statistics of columns 2*(lane%16), +1 over rows of parity lane/16
  uint32_t ph = 0;
  int cur_wimg = -1;
  for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
    const int img = fdiv(tile, P.fd_tpi);
    const int t = tile - img * P.tiles_per_img;
    const int p = t * 128 + tid;
    const bool valid = p < hw;
    const int oy = valid ? fdiv(p, P.fd_wo) : 0;
    const int ox = valid ? p - oy * P.wo : 0;
    {
      uint32_t a[16];
      gather_patch<CIN, KS>(P.x + (long long)img * CIN * P.h * P.w, P.h, P.w, oy, ox, P.stride, P.pad, valid && !(P.dbg & 1), a);
      store_row_sw64(sA, tid, a);
    }
    const int wimg = P.w_batch > 1 ? img : 0;
    if (wimg != cur_wimg) {            // CTA-uniform; the previous tile's MMAs have completed (awaited below)
      cur_wimg = wimg;
      const int row = tid >> 2, ch = tid & 3;
      const uint4 wv = __ldg(reinterpret_cast<const uint4*>(P.wgt + (long long)wimg * 1024 + row * 32 + ch * 8));
      *reinterpret_cast<uint4*>(sB + row * 64 + ((ch ^ ((row >> 1) & 3)) << 4)) = wv;
    }
    if (!(P.dbg & 32)) fence_proxy_async();
    if (tid == 0) tma_store_wait_read<0>();       // the previous tile's bulk copy has finished reading sOut
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      if (!(P.dbg & 2)) {
        tc_mma_bf16(tmem, adesc, bdesc, idesc, 0u);
        tc_mma_bf16(tmem, adesc + 2, bdesc + 2, idesc, 1u);     // +32 bytes along K inside the swizzled row
      }
      if (!(P.dbg & 8)) tc_commit(smem_u32(&bar));
    }
    if (!(P.dbg & 8)) {
      mbar_wait(smem_u32(&bar), ph, dead, P.watchdog, 0x100u);
      ph ^= 1u;
    }
    tc_fence_after();
    uint32_t o[16];
    {
      uint32_t r[32];
      if (!(P.dbg & 16)) {
        tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16), r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = (uint32_t)(tid + i);
      }
      tc_fence_before();
      if (kStats) {
        // rows past the image hold zeros (zero patch rows): they add nothing to the sums
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = pack_bf16x2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
      } else {
        float z[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) z[i] = fmaf(__uint_as_float(r[i]), s_sc[i], s_sh[i]);
        switch (P.act) {       // one branch per tile, not one per element
          case UAVDET_ACT_LEAKY:
#pragma unroll
            for (int i = 0; i < 32; ++i) z[i] = act_fwd<UAVDET_ACT_LEAKY>(z[i]);
            break;
          case UAVDET_ACT_SILU:
#pragma unroll
            for (int i = 0; i < 32; ++i) z[i] = act_fwd<UAVDET_ACT_SILU>(z[i]);
            break;
          case UAVDET_ACT_RELU:
#pragma unroll
            for (int i = 0; i < 32; ++i) z[i] = act_fwd<UAVDET_ACT_RELU>(z[i]);
            break;
          case UAVDET_ACT_GELU:
#pragma unroll
            for (int i = 0; i < 32; ++i) z[i] = act_fwd<UAVDET_ACT_GELU>(z[i]);
            break;
          default: break;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = pack_bf16x2(z[2 * i], z[2 * i + 1]);
      }
    }
    __nv_bfloat16* ytile = P.y + ((long long)img * hw + (long long)t * 128) * P.y_ld;
    if (dense || kStats) {
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
      if (dense && !(P.dbg & 32)) fence_proxy_async();
      if (!(P.dbg & 64)) __syncthreads();
      if (dense && tid == 0) {
        const int rows = hw - t * 128 < 128 ? hw - t * 128 : 128;
        if (!(P.dbg & 4)) bulk_store_1d(ytile, smem_u32(sOut), (uint32_t)(rows * 64));
        tma_store_commit();
      }
      if (kStats) {
        // column sums of the staged (bf16-rounded) rows — exactly the values BatchNorm will normalise.  Warp w owns
        // rows 32w..32w+31; lane -> column pair lane % 16 of the rows of parity lane / 16: one conflict-free 128-byte
        // shared-memory line per load.
        const uint8_t* base = sOut + (warp * 32 + (lane >> 4)) * 64 + (lane & 15) * 4;
#pragma unroll 8
        for (int i = 0; i < 16; ++i) {
          const uint32_t u = *reinterpret_cast<const uint32_t*>(base + i * 128);
          const float lo = bf16_lo(u), hi = bf16_hi(u);
          st_s0 += lo; st_q0 = fmaf(lo, lo, st_q0);
          st_s1 += hi; st_q1 = fmaf(hi, hi, st_q1);
        }
        if (!dense) __syncthreads();      // nothing else orders the next tile's staging writes after these reads
      }
    }
    if (!dense && valid) {
      uint4* dst = reinterpret_cast<uint4*>(ytile + (long long)tid * P.y_ld);
#pragma unroll
      for (int j = 0; j < 4; ++j) dst[j] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
    }
  }
  if (kStats) {
    st_s0 += __shfl_xor_sync(0xffffffffu, st_s0, 16); st_s1 += __shfl_xor_sync(0xffffffffu, st_s1, 16);
    st_q0 += __shfl_xor_sync(0xffffffffu, st_q0, 16); st_q1 += __shfl_xor_sync(0xffffffffu, st_q1, 16);
    if (lane < 16) {
      atomicAdd(&s_stat[0][2 * lane], st_s0); atomicAdd(&s_stat[0][2 * lane + 1], st_s1);
      atomicAdd(&s_stat[1][2 * lane], st_q0); atomicAdd(&s_stat[1][2 * lane + 1], st_q1);
    }
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
// One 64-byte half (pixel `r` of the tile) of a 128-byte pixel-PAIR row of a SWIZZLE_128B tile: pair row R = r / 2,
// 16-byte chunk (r % 2) * 4 + j at chunk index ^ (R % 8).
__device__ __forceinline__ void store_half_row_sw128(uint8_t* tile, int r, int j, const uint4& v) {
  const int R = r >> 1;
  *reinterpret_cast<uint4*>(tile + R * 128 + (((((r & 1) << 2) + j) ^ (R & 7)) << 4)) = v;
}

// dW[co][kk] = sum_p dy[p][co] * patch[p][kk].  Both tiles are kept as 64 pixel-PAIR rows of 128 bytes (two pixels x 32
// values — the bytes are the same as 128 rows of 64, only the swizzle differs): MN-major SWIZZLE_128B operands with
// M = [pixel parity][co], N = [pixel parity][kk], K = 64 pairs.  The accumulator's diagonal blocks are the sums over
// the even and over the odd pixels (added in the flush), the off-diagonal blocks are never read.  Four M128 x N64 x K16
// instructions per tile on 128-byte rows instead of eight N = 32 ones on 64-byte rows (which cost the tensor core twice
// the cycles each: see igemm.cu, halo_mode()).
template <int CIN, int KS>
__global__ void __launch_bounds__(kSmThreads, kSmCtasPerSm) stem_mma_wgrad_kernel(const StemMmaParams P) {
  constexpr int K = CIN * KS * KS;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ uint32_t dead_flag;
  // A operand (dy, M = 128 = two 64-element blocks 8 KB apart of which only the first exists: the second stays zero)
  uint8_t* sDY = align1024(smem_raw);       // 2 x [64 pairs][128 B]
  uint8_t* sP = sDY + 2 * 8192;             // [64 pairs][128 B] patch rows (B operand, N = 64)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  volatile uint32_t* dead = &dead_flag;

  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    dead_flag = 0;
    fence_barrier_init();
  }
  for (int i = tid; i < 8192 / 16; i += kSmThreads) reinterpret_cast<uint4*>(sDY + 8192)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (warp == 0) { tmem_alloc(smem_u32(&tmem_slot), 64); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = make_idesc_bf16(64, 1, 1, 128);                  // both operands MN-major
  const uint64_t adesc = make_smem_desc(smem_u32(sDY), 8192, 1024, 2u);   // LBO: next 64-element block; SBO: 8 pair rows
  const uint64_t bdesc = make_smem_desc(smem_u32(sP), 8192, 1024, 2u);
  const int hw = P.ho * P.wo;

  // contiguous chunk of the (image, tile) sequence: a CTA crosses an image boundary at most a few times
  const int per_cta = (P.total_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
  const int t_begin = (int)blockIdx.x * per_cta;
  const int t_end = min(t_begin + per_cta, P.total_tiles);
  uint32_t ph = 0;
  bool pending = false;      // MMAs of the previous tile not yet awaited
  bool fresh = true;         // the next MMA starts a new accumulation
  int cur_seg = -1;

  auto flush = [&](int seg) {
    // rows [parity][co] x columns [parity][kk]: warp 0 owns the even-pixel block, warp 1 the odd-pixel block
    if (warp < 2) {
      uint32_t r[32];
      tmem_ld_32x32(tmem + (uint32_t)(warp * 32) + ((uint32_t)(warp * 32) << 16), r);
      tmem_ld_wait();
      float* dst = P.dw + (long long)seg * 1024 + lane * 32;
#pragma unroll
      for (int i = 0; i < K; ++i) atomicAdd(dst + i, __uint_as_float(r[i]));
    }
    tc_fence_before();
  };

  for (int tile = t_begin; tile < t_end; ++tile) {
    const int img = fdiv(tile, P.fd_tpi);
    const int t = tile - img * P.tiles_per_img;
    const int p = t * 128 + tid;
    const bool valid = p < hw;
    const int oy = valid ? fdiv(p, P.fd_wo) : 0;
    const int ox = valid ? p - oy * P.wo : 0;
    uint32_t a[16];
    gather_patch<CIN, KS>(P.x + (long long)img * CIN * P.h * P.w, P.h, P.w, oy, ox, P.stride, P.pad, valid, a);
    // dy rows: instruction i covers pixels 32*i + tid/4, 16-byte chunk tid%4 (a warp reads 512 contiguous bytes)
    uint4 g[4];
    const int ch = tid & 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int pp = t * 128 + 32 * i + (tid >> 2);
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
#pragma unroll
    for (int j = 0; j < 4; ++j) store_half_row_sw128(sP, tid, j, make_uint4(a[4 * j], a[4 * j + 1], a[4 * j + 2], a[4 * j + 3]));
#pragma unroll
    for (int i = 0; i < 4; ++i) store_half_row_sw128(sDY, 32 * i + (tid >> 2), ch, g[i]);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)      // 16 pair rows = 2048 bytes per K step
        tc_mma_bf16(tmem, adesc + (uint64_t)(128 * ks), bdesc + (uint64_t)(128 * ks), idesc, (fresh && ks == 0) ? 0u : 1u);
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
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 64); }
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
  UAVDET_CHECK_ARG((long long)P.ho * P.wo < (1ll << 30), "stem_mma: image too large");
  const long long tiles_per_img = ceil_div64((long long)P.ho * P.wo, 128);
  UAVDET_CHECK_ARG(tiles_per_img * n < (1ll << 30), "stem_mma: too many tiles");
  P.fd_wo = make_fast_div(P.wo);
  P.fd_tpi = make_fast_div((int)tiles_per_img);
  P.tiles_per_img = (int)tiles_per_img;
  P.total_tiles = (int)(tiles_per_img * n);
  P.y = (__nv_bfloat16*)y->ptr; P.y_ld = y->ld;
  P.watchdog = watchdog_word();
  static const int dbg = getenv("UAVDET_STEM_DBG") ? atoi(getenv("UAVDET_STEM_DBG")) : 0;
  P.dbg = dbg;
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
  int grid = kNumSMs * (stats ? kSmFwdStatsCtasPerSm : kSmFwdCtasPerSm);
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
  const int smem = 2 * 8192 + 8192 + 1024;
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
