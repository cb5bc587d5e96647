// Weight gradient of the implicit-GEMM convolution on tcgen05:
//
//   dW[co, tap, ci] += sum_{pixels p} dY[p, co] * X[p + tap, ci]
//
// (autograd of the conv sites listed in igemm.cu; reference has no explicit kernel — ATen's
// convolution_backward.)  GEMM view: M = cout (128-row blocks), N = cin block per tap, K = pixels.
// Both operands are "MN-major" in shared memory: a TMA box {64 channels, tile_w, 1, tile_h, 1}
// lands as [pixel][64 ch] 128-byte rows, i.e. K rows x contiguous M/N — exactly the canonical
// UMMA MN-major SWIZZLE_128B atom, so no transpose is ever materialised.
//   * one CTA = (128-cout block, cin block, group of G filter taps, split of the pixel range):
//     the dY box is loaded once per pixel tile and reused by the G taps (G x cin_block <= 512
//     TMEM columns), each tap has its own shifted X box (zero padding = TMA OOB fill);
//   * split-K over pixel tiles so the grid covers all SMs; partial sums are reduced with fp32
//     atomics into the packed gradient [cout][taps*cin] (caller-zeroed).
#include "common.cuh"
#include "sm100.cuh"
#include "igemm.h"

namespace uavdet {
using namespace sm100;

constexpr int kWgradThreads = 192;

struct WgradParams {
  int n_img, ho, wo;                // dY grid
  int tile_w, tile_h, tiles_w, tiles_h, kp;  // pixel tile (kp = tile_w*tile_h, multiple of 16)
  int cout, cin_blk;                // cin_blk = N per tap (multiple of 32, <= 256)
  int a_width, b_width;             // channels per TMA box (64 | 32) for dY / X
  int num_taps, group;              // taps total, taps per CTA
  int ci_blocks, co_blocks, tap_groups, items;
  int k_splits, tiles_per_split, total_pixel_tiles;
  int per_sample;
  long long k_total;                // row length of packed dW
  long long sample_stride;          // elements between per-sample gradients
  int stages;
  float* dw;
  unsigned int* watchdog;
  ConvTap taps[kMaxTaps];
};

__global__ void __launch_bounds__(kWgradThreads, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap mapDY, const __grid_constant__ CUtensorMap mapX,
             const __grid_constant__ WgradParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- work decode -----------------------------------------------------------------------
  int item = blockIdx.x % P.items;
  const int split = blockIdx.x / P.items;
  const int tg = item % P.tap_groups; item /= P.tap_groups;
  const int cib = item % P.ci_blocks;
  const int cob = item / P.ci_blocks;
  const int tap0 = tg * P.group;
  const int ntap = min(P.group, P.num_taps - tap0);
  const int img_only = P.per_sample ? blockIdx.y : -1;
  const int pt_begin = split * P.tiles_per_split;
  const int tiles_here_total = P.per_sample ? P.tiles_h * P.tiles_w : P.total_pixel_tiles;
  const int pt_end = min(pt_begin + P.tiles_per_split, tiles_here_total);
  const int num_kb = max(pt_end - pt_begin, 0);

  const int a_row = P.a_width * 2, b_row = P.b_width * 2;        // bytes per smem row
  const int a_blocks = 128 / P.a_width;                          // 64-/32-channel boxes covering M=128
  const int b_blocks = P.cin_blk / P.b_width;
  const int a_blk_bytes = P.kp * a_row, b_blk_bytes = P.kp * b_row;
  const int a_bytes = a_blocks * a_blk_bytes;
  const int b_tap_bytes = b_blocks * b_blk_bytes;
  const int stage_bytes = a_bytes + P.group * b_tap_bytes;
  uint8_t* ctrl = smem + (size_t)P.stages * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ctrl);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* done_bar = empty_bar + kMaxStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done_bar + 1);
  volatile uint32_t* dead = tmem_ptr + 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < P.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(done_bar), 1);
    *dead = 0;
    fence_barrier_init();
    prefetch_tensormap(&mapDY);
    prefetch_tensormap(&mapX);
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_ptr), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // number of dY boxes that actually exist (cout may be < 128 or not a multiple of it)
  const int co0 = cob * 128;
  int a_live = (P.cout - co0 + P.a_width - 1) / P.a_width;
  if (a_live > a_blocks) a_live = a_blocks;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        int pt = pt_begin + kb;
        const int tw = pt % P.tiles_w; pt /= P.tiles_w;
        const int th = pt % P.tiles_h;
        const int img = P.per_sample ? img_only : pt / P.tiles_h;
        const int ow0 = tw * P.tile_w, oh0 = th * P.tile_h;
        mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u, dead, P.watchdog, 0x10u);
        const uint32_t fb = smem_u32(&full_bar[stage]);
        mbar_arrive_expect_tx(fb, (uint32_t)(a_live * a_blk_bytes + ntap * b_tap_bytes));
        uint8_t* sa = smem + (size_t)stage * stage_bytes;
        for (int ab = 0; ab < a_live; ++ab)
          tma_load_5d(smem_u32(sa + ab * a_blk_bytes), &mapDY, fb, co0 + ab * P.a_width, ow0, 0, oh0, img);
        for (int t = 0; t < ntap; ++t) {
          const ConvTap tap = P.taps[tap0 + t];
          uint8_t* sb = sa + a_bytes + t * b_tap_bytes;
          for (int bb = 0; bb < b_blocks; ++bb)
            tma_load_5d(smem_u32(sb + bb * b_blk_bytes), &mapX, fb,
                        tap.c_off + cib * P.cin_blk + bb * P.b_width, ow0 + tap.dw, tap.p, oh0 + tap.dh, img);
        }
        if (++stage == P.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16(P.cin_blk, 1, 1);  // both operands MN-major
      const uint32_t a_layout = P.a_width == 64 ? 2u : 4u, b_layout = P.b_width == 64 ? 2u : 4u;
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(smem_u32(&full_bar[stage]), phase, dead, P.watchdog, 0x20u);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
        for (int ks = 0; ks < P.kp / 16; ++ks) {
          // MN-major: LBO = stride between channel blocks, SBO = stride between 8-pixel groups
          const uint64_t a_desc = make_smem_desc(sa + ks * 16 * a_row, a_blk_bytes, 8 * a_row, a_layout);
          for (int t = 0; t < ntap; ++t) {
            const uint64_t b_desc = make_smem_desc(sa + a_bytes + t * b_tap_bytes + ks * 16 * b_row,
                                                   b_blk_bytes, 8 * b_row, b_layout);
            tc_mma_bf16(tmem_base + (uint32_t)(t * P.cin_blk), a_desc, b_desc, idesc, (kb | ks) != 0 ? 1u : 0u);
          }
        }
        tc_commit(smem_u32(&empty_bar[stage]));
        if (++stage == P.stages) { stage = 0; phase ^= 1u; }
      }
      tc_commit(smem_u32(done_bar));
    }
  } else if (num_kb > 0) {
    // ---- epilogue: TMEM -> fp32 atomics into the packed gradient ----------------------------
    // Each warp owns 32 output-channel rows; a 32x32 chunk is transposed through a per-warp smem
    // scratch so that every RED instruction covers 32 consecutive floats of ONE row (one 128-byte
    // line) instead of 32 different rows.
    const int q = warp & 3;
    float* scratch = reinterpret_cast<float*>(tmem_ptr + 4) + (warp - 2) * (32 * 33);
    mbar_wait(smem_u32(done_bar), 0, dead, P.watchdog, 0x40u);
    tc_fence_after();
    const int row0 = co0 + q * 32;
    float* base = P.dw + (P.per_sample ? (long long)img_only * P.sample_stride : 0);
    for (int t = 0; t < ntap; ++t) {
      const long long koff = P.taps[tap0 + t].w_koff + (long long)cib * P.cin_blk;
      for (int c0 = 0; c0 < P.cin_blk; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (uint32_t)(t * P.cin_blk + c0) + ((uint32_t)(q * 32) << 16), r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) scratch[lane * 33 + i] = __uint_as_float(r[i]);
        __syncwarp();
        const int nrows = min(32, P.cout - row0);
        for (int rr = 0; rr < nrows; ++rr)
          atomicAdd(base + (long long)(row0 + rr) * P.k_total + koff + c0 + lane, scratch[rr * 33 + lane]);
        __syncwarp();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// pixel tile for the K dimension: tile_w*tile_h must be an exact multiple of 16 (every smem row
// feeds the reduction) and <= 64; pick the one wasting the fewest zero-filled pixels.
static void choose_k_tile(int ho, int wo, int* tile_w, int* tile_h) {
  double best = -1.0;
  int bw = 16, bh = 1, bkp = 0;
  for (int tw = 1; tw <= (wo < 64 ? wo : 64); ++tw)
    for (int th = 1; th <= ho && tw * th <= 64; ++th) {
      const int kp = tw * th;
      if (kp % 16) continue;
      const double eff = (double)ho * wo / ((double)ceil_div(ho, th) * ceil_div(wo, tw) * kp);
      if (eff > best + 1e-9 || (eff > best - 1e-9 && kp > bkp)) { best = eff; bw = tw; bh = th; bkp = kp; }
    }
  if (best < 0) { bw = 16; bh = 1; }  // tiny maps: 16x1 with OOB fill
  *tile_w = bw;
  *tile_h = bh;
}

}  // namespace uavdet

using namespace uavdet;

extern "C" int uavdet_conv_wgrad(const uavdet_act* x, const uavdet_act* dy, int k, int stride, int pad, int s2d,
                                 float* dw_packed, int per_sample, void* stream) {
  UAVDET_CHECK_ARG(x && x->ptr && dy && dy->ptr && dw_packed, "conv_wgrad: null pointer");
  UAVDET_CHECK_ARG(k >= 1 && k <= 5 && (stride == 1 || stride == 2), "conv_wgrad: k=%d stride=%d unsupported", k, stride);
  UAVDET_CHECK_ARG(!(s2d && stride != 1), "conv_wgrad: s2d implies stride 1");
  UAVDET_CHECK_ARG(x->ld % 8 == 0 && dy->ld % 8 == 0 && (((uintptr_t)x->ptr | (uintptr_t)dy->ptr) & 15) == 0,
                   "conv_wgrad: 16-byte alignment");
  UAVDET_CHECK_ARG(x->n == dy->n, "conv_wgrad: batch mismatch");
  const int parity = (stride == 2 || s2d) ? 1 : 0;
  if (parity) UAVDET_CHECK_ARG(x->h % 2 == 0 && x->w % 2 == 0, "conv_wgrad: stride-2/s2d needs even H,W");
  const int c_blk = x->c, cin = s2d ? 4 * x->c : x->c, cout = dy->c;
  UAVDET_CHECK_ARG(c_blk % 32 == 0 && cout % 32 == 0, "conv_wgrad: channels must be multiples of 32");
  const int hin = s2d ? x->h / 2 : x->h, win = s2d ? x->w / 2 : x->w;
  UAVDET_CHECK_ARG((hin + 2 * pad - k) / stride + 1 == dy->h && (win + 2 * pad - k) / stride + 1 == dy->w,
                   "conv_wgrad: spatial sizes inconsistent");
  WgradParams P{};
  P.n_img = x->n; P.ho = dy->h; P.wo = dy->w;
  P.cout = cout;
  P.a_width = (cout % 64 == 0) ? 64 : 32;
  P.b_width = (c_blk % 64 == 0) ? 64 : 32;
  // N per tap: largest multiple-of-32 divisor of the per-tap channel count that is <= 256
  int nb = 32;
  for (int cand = 256; cand >= 32; cand -= 32)
    if (c_blk % cand == 0 && cand % P.b_width == 0) { nb = cand; break; }
  P.cin_blk = nb;
  P.ci_blocks = c_blk / nb;
  int nt = 0;
  for (int kh = 0; kh < k; ++kh)
    for (int kw = 0; kw < k; ++kw) {
      if (s2d) {
        for (int i = 0; i < 2; ++i)
          for (int j = 0; j < 2; ++j)
            P.taps[nt++] = ConvTap{j * x->ld, kw - pad, i, kh - pad, ((kh * k + kw) * 4 + (i * 2 + j)) * c_blk};
      } else if (stride == 1) {
        P.taps[nt++] = ConvTap{0, kw - pad, 0, kh - pad, (kh * k + kw) * cin};
      } else {
        const int th = kh - pad, tw = kw - pad;
        const int ph = ((th % 2) + 2) % 2, pw = ((tw % 2) + 2) % 2;
        P.taps[nt++] = ConvTap{pw * x->ld, (tw - pw) / 2, ph, (th - ph) / 2, (kh * k + kw) * cin};
      }
      UAVDET_CHECK_ARG(nt <= kMaxTaps, "conv_wgrad: too many taps");
    }
  P.num_taps = nt;
  choose_k_tile(P.ho, P.wo, &P.tile_w, &P.tile_h);
  P.kp = P.tile_w * P.tile_h;
  P.tiles_w = ceil_div(P.wo, P.tile_w);
  P.tiles_h = ceil_div(P.ho, P.tile_h);
  P.total_pixel_tiles = P.n_img * P.tiles_h * P.tiles_w;
  // taps per CTA: bounded by 512 TMEM columns and by shared memory (>= 3 stages)
  const int max_smem = 227 * 1024;
  const int ctrl_bytes = 8 * (2 * kMaxStages + 1) + 16 + 4 * 32 * 33 * 4;   // barriers + per-warp transpose scratch
  const int a_bytes = 128 * P.kp * 2;
  int group = 512 / P.cin_blk;
  if (group > nt) group = nt;
  while (group > 1 && (max_smem - 1024 - ctrl_bytes) / (a_bytes + group * P.cin_blk * P.kp * 2) < 3) --group;
  // balance groups (e.g. 9 taps, cap 8 -> 5+4 rather than 8+1)
  P.tap_groups = ceil_div(nt, group);
  group = ceil_div(nt, P.tap_groups);
  P.group = group;
  const int stage_bytes = a_bytes + group * P.cin_blk * P.kp * 2;
  int stages = (max_smem - 1024 - ctrl_bytes) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  UAVDET_CHECK_ARG(stages >= 2, "conv_wgrad: stage does not fit shared memory");
  P.stages = stages;
  P.co_blocks = ceil_div(cout, 128);
  P.items = P.co_blocks * P.ci_blocks * P.tap_groups;
  P.per_sample = per_sample ? 1 : 0;
  const int tiles_avail = per_sample ? P.tiles_h * P.tiles_w : P.total_pixel_tiles;
  const int samples = per_sample ? P.n_img : 1;
  // split-K factor: minimise (waves of CTAs) x (pixel tiles per CTA); one CTA per SM is resident
  int splits = 1;
  {
    long long best = -1;
    const int ctas_per_split = P.items * samples;
    for (int s_ = 1; s_ <= 64 && s_ <= tiles_avail; ++s_) {
      const long long waves = ceil_div(ctas_per_split * s_, kNumSMs);
      const long long cost = waves * (ceil_div(tiles_avail, s_) + 6);   // +6: per-CTA prologue/epilogue in tile units
      if (best < 0 || cost < best) { best = cost; splits = s_; }
    }
  }
  P.tiles_per_split = ceil_div(tiles_avail, splits);
  P.k_splits = ceil_div(tiles_avail, P.tiles_per_split);
  P.k_total = (long long)k * k * cin;
  P.sample_stride = (long long)cout * P.k_total;
  P.dw = dw_packed;
  P.watchdog = watchdog_word();

  CUtensorMap mapDY, mapX;
  int rc = make_act_map(&mapDY, dy, 0, P.a_width, P.tile_w, P.tile_h);
  if (rc) return rc;
  rc = make_act_map(&mapX, x, parity, P.b_width, P.tile_w, P.tile_h);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    UAVDET_CUDA(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    attr_set = true;
  }
  dim3 grid((unsigned)(P.items * P.k_splits), (unsigned)samples);
  wgrad_kernel<<<grid, kWgradThreads, max_smem, (cudaStream_t)stream>>>(mapDY, mapX, P);
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}
