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
#include <stdlib.h>

namespace uavdet {
using namespace sm100;

constexpr int kWgradThreads = 192;
constexpr int kMaxCols = 20;
constexpr int kMaxSlots = 16;   // taps per CTA (512 TMEM columns / 32)

// A "column" = the filter taps that differ only in their vertical offset (same channel block, same horizontal
// offset, same row parity).  They read the same input pixels shifted by whole tile rows, so ONE TMA box of
// (tile_h + nv - 1) x tile_w pixels feeds all nv of them: tap v starts v*tile_w smem rows further down (tile_w is a
// multiple of 8, which keeps every start on a swizzle-atom boundary).  3x3 stride 1: 3 columns of 3 taps, i.e. 3
// input boxes per pixel tile instead of 9.
struct WgCol {
  int c_off, dw, p, dh0;   // TMA coordinates of the box origin relative to the pixel tile (see ConvTap)
  int nv;                  // taps in the column
  int tap0;                // index of its first tap in WgradParams::koff
  int map;                 // which X tensor map (box height tile_h + nv - 1) loads it
  int off, blk;            // smem byte offset of its first channel block inside a stage / bytes per channel block
};

struct WgradParams {
  int n_img, ho, wo;                // dY grid
  int tile_w, tile_h, tiles_w, tiles_h;
  int kp, kp_pad;                   // pixels per K tile, rounded up to the MMA K step (16); the pad rows stay zero
  int cout, cin_blk;                // cin_blk = N per tap (multiple of 32, <= 256)
  int a_width, b_width;             // channels per TMA box (64 | 32) for dY / X
  int num_cols, num_groups, cols_per_group;
  int ci_blocks, co_blocks, items;
  int k_splits, tiles_per_split, total_pixel_tiles;
  int per_sample;
  long long k_total;                // row length of packed dW
  long long sample_stride;          // elements between per-sample gradients
  int stages, stage_bytes, a_bytes;
  float* dw;
  unsigned int* watchdog;
  WgCol cols[kMaxCols];
  int grp_x_bytes[kMaxCols];        // bytes of X one stage of column group g receives
  int koff[kMaxTaps];               // K offset of every tap inside the packed weight row, column-major order
  int tap_off[kMaxTaps];            // smem byte offset (inside a stage) of the first row each tap reads
  int tap_lbo[kMaxTaps];            // bytes between the channel blocks of that tap's box
  int merge_taps;                   // 1: one MMA per column and K step (taps as N blocks), where a tap is one channel block
};

struct WgIssue {
  uint32_t full_bar, empty_bar;     // smem addresses of the barrier arrays
  int num_kb, ksteps, stages;
  uint64_t a0;                      // dY descriptor of stage 0, K step 0
  uint32_t a_step, b_step, stage_step;   // (bytes >> 4) added per K step / per stage
  uint32_t tmem_base, cin_blk, idesc;
  volatile uint32_t* dead;
  unsigned int* watchdog;
};

// MMA issue loop of one CTA, specialised on the number of taps NS (accumulators) it owns.
template <int NS, bool kTwo>
__device__ __forceinline__ void wgrad_issue(const WgIssue& is, const uint64_t (&bd)[kMaxSlots]) {
  int stage = 0;
  uint32_t phase = 0, soff = 0;
  for (int kb = 0; kb < is.num_kb; ++kb) {
    mbar_wait(is.full_bar + 8u * (uint32_t)stage, phase, is.dead, is.watchdog, 0x20u);
    tc_fence_after();
    for (int ks = 0; ks < is.ksteps; ++ks) {
      const uint64_t ad = is.a0 + (uint64_t)(soff + ks * is.a_step);
      const uint64_t boff = (uint64_t)(soff + ks * is.b_step);
      const uint32_t accum = (kb | ks) != 0 ? 1u : 0u;
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        if (kTwo) tc_mma_bf16_2sm(is.tmem_base + (uint32_t)i * is.cin_blk, ad, bd[i] + boff, is.idesc, accum);
        else tc_mma_bf16(is.tmem_base + (uint32_t)i * is.cin_blk, ad, bd[i] + boff, is.idesc, accum);
      }
    }
    if (kTwo) tc_commit_2sm(is.empty_bar + 8u * (uint32_t)stage, 3); else tc_commit(is.empty_bar + 8u * (uint32_t)stage);
    soff += is.stage_step;
    if (++stage == is.stages) { stage = 0; phase ^= 1u; soff = 0; }
  }
}

// The same loop with the taps of a column merged into ONE instruction per K step: the taps of a column read one box at
// starts tile_w rows apart, which is exactly the layout of an MN-major operand whose N blocks (one per tap, cin_blk
// channels = one swizzle row each) lie LBO = tile_w rows apart.  The tensor core takes as long for an N = 32 instruction
// as for an N = 96 one (it is bound by re-reading the dY slice), so a 3x3 layer with <= 64 input channels issues a
// third of the instructions.  NG = number of columns (instructions per K step).
constexpr int kMaxGroups = 8;
template <int NG>
__device__ __forceinline__ void wgrad_issue_merged(const WgIssue& is, const uint64_t (&gd)[kMaxGroups],
                                                   const uint32_t (&gcol)[kMaxGroups], const uint32_t (&gidesc)[kMaxGroups]) {
  int stage = 0;
  uint32_t phase = 0, soff = 0;
  for (int kb = 0; kb < is.num_kb; ++kb) {
    mbar_wait(is.full_bar + 8u * (uint32_t)stage, phase, is.dead, is.watchdog, 0x20u);
    tc_fence_after();
    for (int ks = 0; ks < is.ksteps; ++ks) {
      const uint64_t ad = is.a0 + (uint64_t)(soff + ks * is.a_step);
      const uint64_t boff = (uint64_t)(soff + ks * is.b_step);
      const uint32_t accum = (kb | ks) != 0 ? 1u : 0u;
#pragma unroll
      for (int i = 0; i < NG; ++i) tc_mma_bf16(is.tmem_base + gcol[i], ad, gd[i] + boff, gidesc[i], accum);
    }
    tc_commit(is.empty_bar + 8u * (uint32_t)stage);
    soff += is.stage_step;
    if (++stage == is.stages) { stage = 0; phase ^= 1u; soff = 0; }
  }
}

// kTwo: CTA-pair variant (cta_group::2).  The pair owns two adjacent 128-row cout blocks (M = 256) of the same (cin
// block, tap group, pixel split): each CTA loads its own dY boxes and HALF of the channels of every X box, so the
// L2 -> SM traffic per MMA — what bounds this kernel on the 20x20 / 40x40 maps (20 KB per 384-cycle tile step, twice
// what the L2 delivers per SM) — drops by 30 %.  Accumulators, epilogue and atomics stay per CTA.
template <bool kTwo>
__global__ void __launch_bounds__(kWgradThreads, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap mapDY, const __grid_constant__ CUtensorMap mapX0,
             const __grid_constant__ CUtensorMap mapX1, const __grid_constant__ CUtensorMap mapX2,
             const __grid_constant__ WgradParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- work decode -----------------------------------------------------------------------
  const uint32_t rank = kTwo ? cluster_ctarank() : 0u;
  const int bid = kTwo ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int items = kTwo ? P.items / 2 : P.items;            // pairs of cout blocks
  int item = bid % items;
  const int split = bid / items;
  const int grp = item % P.num_groups; item /= P.num_groups;
  const int cib = item % P.ci_blocks;
  const int cob = kTwo ? 2 * (item / P.ci_blocks) + (int)rank : item / P.ci_blocks;
  const int col0 = grp * P.cols_per_group;
  const int ncol = min(P.cols_per_group, P.num_cols - col0);
  const int img_only = P.per_sample ? (int)blockIdx.y : -1;
  const int pt_begin = split * P.tiles_per_split;
  const int tiles_here_total = P.per_sample ? P.tiles_h * P.tiles_w : P.total_pixel_tiles;
  const int pt_end = min(pt_begin + P.tiles_per_split, tiles_here_total);
  const int num_kb = max(pt_end - pt_begin, 0);

  const int a_row = P.a_width * 2, b_row = P.b_width * 2;        // bytes per smem row
  const int a_blocks = 128 / P.a_width;                          // 64-/32-channel boxes covering M=128
  const int cin_here = kTwo ? P.cin_blk / 2 : P.cin_blk;         // channels of every X box this CTA stages
  const int b_blocks = cin_here / P.b_width;
  const int a_blk_bytes = P.kp_pad * a_row;
  uint8_t* ctrl = smem + (size_t)P.stages * P.stage_bytes;
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
    prefetch_tensormap(&mapX0);
  }
  if (warp == 1) {
    if (kTwo) { tmem_alloc_2sm(smem_u32(tmem_ptr), 512); tmem_relinquish_2sm(); }
    else { tmem_alloc(smem_u32(tmem_ptr), 512); tmem_relinquish(); }
  }
  // The K padding rows (and the rows past a shared box that its last tap touches) are never written by TMA:
  // zero the pipeline buffers once so they contribute exact zeros.
  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    const int n16 = P.stages * P.stage_bytes / 16;
    for (int i = threadIdx.x; i < n16; i += kWgradThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async();
  }
  tc_fence_before();
  if (kTwo) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // number of dY boxes that actually exist (cout may be < 128 or not a multiple of it)
  const int co0 = cob * 128;
  int a_live = (P.cout - co0 + P.a_width - 1) / P.a_width;
  if (a_live > a_blocks) a_live = a_blocks;

  const int tx_bytes = a_live * P.kp * a_row + P.grp_x_bytes[grp];   // bytes one stage receives

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int tw, th, img;
      {
        int pt = pt_begin;
        tw = pt % P.tiles_w; pt /= P.tiles_w;
        th = pt % P.tiles_h;
        img = P.per_sample ? img_only : pt / P.tiles_h;
      }
      for (int kb = 0; kb < num_kb; ++kb) {
        const int ow0 = tw * P.tile_w, oh0 = th * P.tile_h;
        mbar_wait<32>(smem_u32(&empty_bar[stage]), phase ^ 1u, dead, P.watchdog, 0x10u);
        // CTA pair: all bytes of both CTAs are counted on the leader's barrier, which the leader arms for both
        const uint32_t fb = kTwo ? mapa_shared(smem_u32(&full_bar[stage]), 0) : smem_u32(&full_bar[stage]);
        if (!kTwo) mbar_arrive_expect_tx(fb, (uint32_t)tx_bytes);
        else if (rank == 0) mbar_arrive_expect_tx(smem_u32(&full_bar[stage]), (uint32_t)(2 * tx_bytes));
        uint8_t* sa = smem + (size_t)stage * P.stage_bytes;
        for (int ab = 0; ab < a_live; ++ab) {
          if (kTwo) tma_load_5d_2sm(smem_u32(sa + ab * a_blk_bytes), &mapDY, fb, co0 + ab * P.a_width, ow0, 0, oh0, img);
          else tma_load_5d(smem_u32(sa + ab * a_blk_bytes), &mapDY, fb, co0 + ab * P.a_width, ow0, 0, oh0, img);
        }
        for (int c = 0; c < ncol; ++c) {
          const WgCol col = P.cols[col0 + c];
          const CUtensorMap* mx = col.map == 0 ? &mapX0 : (col.map == 1 ? &mapX1 : &mapX2);
          for (int bb = 0; bb < b_blocks; ++bb) {
            const int ch = col.c_off + cib * P.cin_blk + (int)rank * cin_here + bb * P.b_width;
            if (kTwo) tma_load_5d_2sm(smem_u32(sa + col.off + bb * col.blk), mx, fb, ch, ow0 + col.dw, col.p, oh0 + col.dh0, img);
            else tma_load_5d(smem_u32(sa + col.off + bb * col.blk), mx, fb, ch, ow0 + col.dw, col.p, oh0 + col.dh0, img);
          }
        }
        if (++tw == P.tiles_w) { tw = 0; if (++th == P.tiles_h) { th = 0; ++img; } }
        if (++stage == P.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && elect_one()) {      // a pair's MMAs are issued by its leader
      // The single issuing thread must spend < 64 cycles per MMA (M128 x N128 x K16) to keep the tensor pipe
      // busy, so everything that does not change inside the loop is hoisted: one 64-bit descriptor per tap
      // (stage 0, K step 0) lives in registers; stage / K-step offsets are added in the (address >> 4) field.
      const uint32_t idesc = make_idesc_bf16(P.cin_blk, 1, 1, kTwo ? 256 : 128);  // both operands MN-major
      const uint32_t a_layout = P.a_width == 64 ? 2u : 4u, b_layout = P.b_width == 64 ? 2u : 4u;
      const int ksteps = P.kp_pad / 16;
      const uint32_t sa0 = smem_u32(smem);
      const uint64_t a0 = make_smem_desc(sa0, a_blk_bytes, 8 * a_row, a_layout);
      const int t0 = P.cols[col0].tap0;
      int nslots = 0;
      for (int c = 0; c < ncol; ++c) nslots += P.cols[col0 + c].nv;
      uint64_t bd[kMaxSlots];
#pragma unroll
      for (int i = 0; i < kMaxSlots; ++i)
        bd[i] = i < nslots ? make_smem_desc(sa0 + (uint32_t)P.tap_off[t0 + i], (uint32_t)P.tap_lbo[t0 + i], 8 * b_row, b_layout)
                           : 0ull;
      const uint32_t a_step = (uint32_t)(16 * a_row) >> 4, b_step = (uint32_t)(16 * b_row) >> 4;
      const uint32_t stage_step = (uint32_t)P.stage_bytes >> 4;
      WgIssue is;
      is.full_bar = smem_u32(full_bar); is.empty_bar = smem_u32(empty_bar);
      is.num_kb = num_kb; is.ksteps = ksteps; is.stages = P.stages;
      is.a0 = a0; is.a_step = a_step; is.b_step = b_step; is.stage_step = stage_step;
      is.tmem_base = tmem_base; is.cin_blk = (uint32_t)P.cin_blk; is.idesc = idesc;
      is.dead = dead; is.watchdog = P.watchdog;
      // one instruction per column where every tap is a single channel block (see wgrad_issue_merged)
      bool merged = !kTwo && P.merge_taps && b_blocks == 1 && ncol <= kMaxGroups;
      for (int c = 0; c < ncol && merged; ++c) merged = P.cols[col0 + c].nv * P.cin_blk <= 256;
      if (merged) {
        uint64_t gd[kMaxGroups];
        uint32_t gcol[kMaxGroups], gidesc[kMaxGroups];
#pragma unroll
        for (int c = 0; c < kMaxGroups; ++c) {
          gd[c] = 0ull; gcol[c] = 0u; gidesc[c] = 0u;
          if (c < ncol) {
            const WgCol col = P.cols[col0 + c];
            gd[c] = make_smem_desc(sa0 + (uint32_t)P.tap_off[col.tap0], (uint32_t)(P.tile_w * b_row), 8 * b_row, b_layout);
            gcol[c] = (uint32_t)((col.tap0 - t0) * P.cin_blk);
            gidesc[c] = make_idesc_bf16(col.nv * P.cin_blk, 1, 1, 128);
          }
        }
        switch (ncol) {
          case 1: wgrad_issue_merged<1>(is, gd, gcol, gidesc); break;  case 2: wgrad_issue_merged<2>(is, gd, gcol, gidesc); break;
          case 3: wgrad_issue_merged<3>(is, gd, gcol, gidesc); break;  case 4: wgrad_issue_merged<4>(is, gd, gcol, gidesc); break;
          case 5: wgrad_issue_merged<5>(is, gd, gcol, gidesc); break;  case 6: wgrad_issue_merged<6>(is, gd, gcol, gidesc); break;
          case 7: wgrad_issue_merged<7>(is, gd, gcol, gidesc); break;  default: wgrad_issue_merged<8>(is, gd, gcol, gidesc); break;
        }
      } else
      switch (nslots) {     // one specialisation per tap count: only live MMAs in the instruction stream
        case 1: wgrad_issue<1, kTwo>(is, bd); break;    case 2: wgrad_issue<2, kTwo>(is, bd); break;
        case 3: wgrad_issue<3, kTwo>(is, bd); break;    case 4: wgrad_issue<4, kTwo>(is, bd); break;
        case 5: wgrad_issue<5, kTwo>(is, bd); break;    case 6: wgrad_issue<6, kTwo>(is, bd); break;
        case 7: wgrad_issue<7, kTwo>(is, bd); break;    case 8: wgrad_issue<8, kTwo>(is, bd); break;
        case 9: wgrad_issue<9, kTwo>(is, bd); break;    case 10: wgrad_issue<10, kTwo>(is, bd); break;
        case 11: wgrad_issue<11, kTwo>(is, bd); break;  case 12: wgrad_issue<12, kTwo>(is, bd); break;
        case 13: wgrad_issue<13, kTwo>(is, bd); break;  case 14: wgrad_issue<14, kTwo>(is, bd); break;
        case 15: wgrad_issue<15, kTwo>(is, bd); break;  default: wgrad_issue<16, kTwo>(is, bd); break;
      }
      if (kTwo) tc_commit_2sm(smem_u32(done_bar), 3); else tc_commit(smem_u32(done_bar));
    }
  } else if (num_kb > 0) {
    // ---- epilogue: TMEM -> fp32 atomics into the packed gradient ----------------------------
    // Each warp owns 32 output-channel rows; a 32x32 chunk is transposed through a per-warp smem
    // scratch so that every RED instruction covers 32 consecutive floats of ONE row (one 128-byte
    // line) instead of 32 different rows.
    const int q = warp & 3;
    float* scratch = reinterpret_cast<float*>(tmem_ptr + 4) + (warp - 2) * (32 * 33);
    mbar_wait<256>(smem_u32(done_bar), 0, dead, P.watchdog, 0x40u);
    tc_fence_after();
    const int row0 = co0 + q * 32;
    float* base = P.dw + (P.per_sample ? (long long)img_only * P.sample_stride : 0);
    int slot = 0;
    for (int c = 0; c < ncol; ++c) {
      const WgCol col = P.cols[col0 + c];
      for (int v = 0; v < col.nv; ++v, ++slot) {
        const long long koff = P.koff[col.tap0 + v] + (long long)cib * P.cin_blk;
        for (int c0 = 0; c0 < P.cin_blk; c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + (uint32_t)(slot * P.cin_blk + c0) + ((uint32_t)(q * 32) << 16), r);
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
  }

  tc_fence_before();
  if (kTwo) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (kTwo) tmem_dealloc_2sm(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

// Pixel tile of the K dimension: tile_w * tile_h <= 64 pixels, rounded up to a multiple of the MMA K step with
// zero rows.  `share`: vertical taps share one box, which needs tile_w % 8 == 0 (tap v starts v*tile_w rows down).
// Picks the tile wasting the fewest MMA rows; ties go to the larger tile.
static bool choose_k_tile(int ho, int wo, bool share, int h_limit, int* tile_w, int* tile_h) {
  double best = -1.0;
  int bw = 0, bh = 0, bkp = 0;
  const int max_w = wo < 64 ? wo : 64;
  for (int tw = share ? 8 : 1; tw <= max_w; tw += share ? 8 : 1)
    for (int th = 1; th <= ho && th <= h_limit && tw * th <= 64; ++th) {
      const int kp = tw * th, kp_pad = (kp + 15) / 16 * 16;
      const double eff = (double)ho * wo / ((double)ceil_div(ho, th) * ceil_div(wo, tw) * kp_pad);
      if (eff > best + 1e-9 || (eff > best - 1e-9 && kp > bkp)) { best = eff; bw = tw; bh = th; bkp = kp; }
    }
  if (best < 0) return false;
  *tile_w = bw;
  *tile_h = bh;
  return true;
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
  const int c_blk0 = x->c, cin = s2d ? 4 * x->c : x->c, cout = dy->c;
  UAVDET_CHECK_ARG(c_blk0 % 32 == 0 && cout % 32 == 0, "conv_wgrad: channels must be multiples of 32");
  const int hin = s2d ? x->h / 2 : x->h, win = s2d ? x->w / 2 : x->w;
  UAVDET_CHECK_ARG((hin + 2 * pad - k) / stride + 1 == dy->h && (win + 2 * pad - k) / stride + 1 == dy->w,
                   "conv_wgrad: spatial sizes inconsistent");
  WgradParams P{};
  P.n_img = x->n; P.ho = dy->h; P.wo = dy->w;
  P.cout = cout;
  P.a_width = (cout % 64 == 0) ? 64 : 32;
  // 32-channel inputs read through the parity view: both pixels of a pair as ONE 64-channel block (see the forward
  // kernel, igemm.cu) — half the MMAs, 128-byte rows.  Stride 2: the (kh, 0) tap reads channels 32..95 of the pair row;
  // the zero-filled half adds zeros to the gradient columns of (kh, 1).
  static const bool no_pair = getenv("UAVDET_IGEMM_NOPAIR") != nullptr;
  const bool pair = !no_pair && parity && x->c == 32 && x->ld == 32 && (s2d || (k == 3 && pad == 1));
  const int c_blk = pair ? 64 : c_blk0;
  P.b_width = (c_blk % 64 == 0) ? 64 : 32;

  // ---- filter taps, grouped into columns (same channel block / horizontal offset / row parity) ----
  struct Tap { int c_off, dw, p, dh, koff; };
  Tap taps[kMaxTaps];
  int nt = 0;
  for (int kh = 0; kh < k; ++kh)
    for (int kw = 0; kw < k; ++kw) {
      if (pair && s2d) {
        for (int i = 0; i < 2; ++i) {
          UAVDET_CHECK_ARG(nt < kMaxTaps, "conv_wgrad: too many taps");
          taps[nt++] = Tap{0, kw - pad, i, kh - pad, ((kh * k + kw) * 4 + i * 2) * c_blk0};
        }
      } else if (pair) {
        if (kw == 2) continue;
        const int th = kh - pad;
        const int ph = ((th % 2) + 2) % 2;
        taps[nt++] = Tap{kw == 0 ? 32 : 0, kw == 0 ? -1 : 0, ph, (th - ph) / 2, (kh * k + kw) * cin};
      } else if (s2d) {
        for (int i = 0; i < 2; ++i)
          for (int j = 0; j < 2; ++j) {
            UAVDET_CHECK_ARG(nt < kMaxTaps, "conv_wgrad: too many taps");
            taps[nt++] = Tap{j * x->ld, kw - pad, i, kh - pad, ((kh * k + kw) * 4 + (i * 2 + j)) * c_blk0};
          }
      } else if (stride == 1) {
        UAVDET_CHECK_ARG(nt < kMaxTaps, "conv_wgrad: too many taps");
        taps[nt++] = Tap{0, kw - pad, 0, kh - pad, (kh * k + kw) * cin};
      } else {
        const int th = kh - pad, tw = kw - pad;
        const int ph = ((th % 2) + 2) % 2, pw = ((tw % 2) + 2) % 2;
        UAVDET_CHECK_ARG(nt < kMaxTaps, "conv_wgrad: too many taps");
        taps[nt++] = Tap{pw * x->ld, (tw - pw) / 2, ph, (th - ph) / 2, (kh * k + kw) * cin};
      }
    }
  // sharing needs the taller box (tile_h + nv - 1 rows) to fit the (parity) view of the input
  const int nv_max = stride == 2 ? (k + 1) / 2 : k;
  const int view_h = parity ? x->h / 2 : x->h;
  static const bool no_share = getenv("UAVDET_WGRAD_NOSHARE") != nullptr;   // tuning/debug switch
  const bool share = !no_share && k > 1 && P.wo >= 8 && view_h - (nv_max - 1) >= 1;
  if (!choose_k_tile(P.ho, P.wo, share, share ? view_h - (nv_max - 1) : P.ho, &P.tile_w, &P.tile_h)) {
    UAVDET_CHECK_ARG(false, "conv_wgrad: no pixel tile for a %dx%d map", P.ho, P.wo);
  }
  P.kp = P.tile_w * P.tile_h;
  P.kp_pad = (P.kp + 15) / 16 * 16;
  const int pad_rows = P.kp_pad - P.kp;
  int ncols = 0, ntaps_flat = 0;
  bool used[kMaxTaps] = {false};
  for (int t = 0; t < nt; ++t) {
    if (used[t]) continue;
    // gather the taps of this column in increasing dh; only runs of consecutive dh can share a box
    int members[8], nm = 0;
    for (int u = t; u < nt; ++u)
      if (!used[u] && taps[u].c_off == taps[t].c_off && taps[u].dw == taps[t].dw && taps[u].p == taps[t].p) members[nm++] = u;
    for (int a = 0; a < nm; ++a)
      for (int b = a + 1; b < nm; ++b)
        if (taps[members[b]].dh < taps[members[a]].dh) { int tmp = members[a]; members[a] = members[b]; members[b] = tmp; }
    int a = 0;
    while (a < nm) {
      int b = a + 1;
      if (share)
        while (b < nm && taps[members[b]].dh == taps[members[b - 1]].dh + 1) ++b;
      UAVDET_CHECK_ARG(ncols < kMaxCols, "conv_wgrad: too many tap columns");
      WgCol& col = P.cols[ncols++];
      col.c_off = taps[members[a]].c_off; col.dw = taps[members[a]].dw; col.p = taps[members[a]].p;
      col.dh0 = taps[members[a]].dh; col.nv = b - a; col.tap0 = ntaps_flat;
      for (int m = a; m < b; ++m) { P.koff[ntaps_flat++] = taps[members[m]].koff; used[members[m]] = true; }
      a = b;
    }
  }
  P.num_cols = ncols;
  // distinct box heights -> X tensor maps
  int map_nv[3] = {0, 0, 0}, nmaps = 0;
  for (int c = 0; c < ncols; ++c) {
    int m = -1;
    for (int i = 0; i < nmaps; ++i) if (map_nv[i] == P.cols[c].nv) m = i;
    if (m < 0) {
      UAVDET_CHECK_ARG(nmaps < 3, "conv_wgrad: more than 3 distinct column heights");
      m = nmaps; map_nv[nmaps++] = P.cols[c].nv;
    }
    P.cols[c].map = m;
  }

  // ---- N per tap, columns per CTA (TMEM: taps x cin_blk <= 512 columns; smem: >= 3 pipeline stages) ----
  // The kernel asks for only ~176 KB so that HBM-bound blocks of another stream (the BatchNorm backward of the
  // next layer, which the engine runs concurrently) can share the SM: tensor work and streaming work overlap.
  static const int smem_budget_kb = getenv("UAVDET_WGRAD_SMEM_KB") ? atoi(getenv("UAVDET_WGRAD_SMEM_KB")) : 176;
  const int max_smem = smem_budget_kb * 1024;
  const int ctrl_bytes = 8 * (2 * kMaxStages + 1) + 16 + 4 * 32 * 33 * 4;   // barriers + per-warp transpose scratch
  P.a_bytes = 128 * P.kp_pad * 2;
  int max_nv = 0;
  for (int c = 0; c < ncols; ++c) max_nv = P.cols[c].nv > max_nv ? P.cols[c].nv : max_nv;
  int nb = 32;
  for (int cand = 256; cand >= 32; cand -= 32)
    if (c_blk % cand == 0 && cand % P.b_width == 0 && cand * max_nv <= 512) { nb = cand; break; }
  UAVDET_CHECK_ARG(nb * max_nv <= 512, "conv_wgrad: a tap column does not fit TMEM");
  P.cin_blk = nb;
  P.ci_blocks = c_blk / nb;
  // CTA-pair kernel: two cout blocks per pair, each CTA stages half of the channels of every X box
  static const int two_mode = getenv("UAVDET_WGRAD_2CTA") ? atoi(getenv("UAVDET_WGRAD_2CTA")) : 1;
  const bool two = two_mode != 0 && !per_sample && cout % 256 == 0 && nb >= 2 * P.b_width && (nb / 2) % P.b_width == 0 &&
                   (nb / 2) % 16 == 0;
  const int nbh = two ? nb / 2 : nb;                      // channels per X box and CTA
  auto col_bytes = [&](int c) { return (long long)nbh * 2 * ((P.tile_h + P.cols[c].nv - 1) * P.tile_w + pad_rows); };
  int cpg = 1;
  for (int cand = ncols; cand >= 1; --cand) {
    if (ncols % cand) continue;
    bool ok = true;
    for (int g0 = 0; g0 < ncols && ok; g0 += cand) {
      int tp = 0; long long bytes = P.a_bytes;
      for (int c = g0; c < g0 + cand; ++c) { tp += P.cols[c].nv; bytes += col_bytes(c); }
      if (tp * nb > 512 || tp > kMaxSlots || (max_smem - ctrl_bytes) / bytes < 3) ok = false;
    }
    if (ok) { cpg = cand; break; }
  }
  P.cols_per_group = cpg;
  P.num_groups = ncols / cpg;
  long long stage_bytes = 0;
  for (int g = 0; g < P.num_groups; ++g) {
    long long off = P.a_bytes; int xb = 0;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      const int rows = (P.tile_h + P.cols[c].nv - 1) * P.tile_w;
      P.cols[c].off = (int)off;
      P.cols[c].blk = (rows + pad_rows) * P.b_width * 2;
      for (int v = 0; v < P.cols[c].nv; ++v) {
        P.tap_off[P.cols[c].tap0 + v] = (int)off + v * P.tile_w * P.b_width * 2;
        P.tap_lbo[P.cols[c].tap0 + v] = P.cols[c].blk;
      }
      off += (long long)(nbh / P.b_width) * P.cols[c].blk;
      xb += (nbh / P.b_width) * rows * P.b_width * 2;
    }
    P.grp_x_bytes[g] = xb;
    if (off > stage_bytes) stage_bytes = off;
  }
  stage_bytes = (stage_bytes + 1023) / 1024 * 1024;
  P.stage_bytes = (int)stage_bytes;
  int stages = (int)((max_smem - ctrl_bytes) / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  UAVDET_CHECK_ARG(stages >= 2, "conv_wgrad: stage does not fit shared memory");
  P.stages = stages;

  P.tiles_w = ceil_div(P.wo, P.tile_w);
  P.tiles_h = ceil_div(P.ho, P.tile_h);
  P.total_pixel_tiles = P.n_img * P.tiles_h * P.tiles_w;
  P.co_blocks = ceil_div(cout, 128);
  P.items = P.co_blocks * P.ci_blocks * P.num_groups;
  P.per_sample = per_sample ? 1 : 0;
  const int tiles_avail = per_sample ? P.tiles_h * P.tiles_w : P.total_pixel_tiles;
  const int samples = per_sample ? P.n_img : 1;
  // split-K factor: minimise (waves of CTAs) x (pixel tiles per CTA); one CTA per SM is resident
  int splits = 1;
  {
    long long best = -1;
    const int ctas_per_split = P.items * samples;
    for (int s_ = 1; s_ <= 148 && s_ <= tiles_avail; ++s_) {
      const long long waves = ceil_div(ctas_per_split * s_, sm_budget());
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
  static const bool no_merge = getenv("UAVDET_WGRAD_NOMERGE") != nullptr;    // A/B switch
  P.merge_taps = (!no_merge && share) ? 1 : 0;      // shared boxes: the taps of a column are tile_w rows apart

  CUtensorMap mapDY, mapX[3];
  int rc = make_act_map(&mapDY, dy, 0, P.a_width, P.tile_w, P.tile_h);
  if (rc) return rc;
  for (int m = 0; m < 3; ++m) {
    const int nv = m < nmaps ? map_nv[m] : map_nv[0];
    rc = make_act_map(&mapX[m], x, parity, P.b_width, P.tile_w, P.tile_h + nv - 1);
    if (rc) return rc;
  }
  static PerDeviceOnce attr_once;     // function attributes are per device
  UAVDET_CUDA(attr_once.run([] {
    cudaError_t e = cudaFuncSetAttribute(wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    // same L1 / shared-memory split as the streaming kernels that share the SM with this one (see elementwise.cu)
    e = cudaFuncSetAttribute(wgrad_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(wgrad_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  }));
  dim3 grid((unsigned)(P.items * P.k_splits), (unsigned)samples);
  const int smem_bytes = P.stages * P.stage_bytes + ctrl_bytes;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kWgradThreads);
  cfg.dynamicSmemBytes = (size_t)smem_bytes;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute cattr[1];
  cattr[0].id = cudaLaunchAttributeClusterDimension;
  cattr[0].val.clusterDim.x = 2; cattr[0].val.clusterDim.y = 1; cattr[0].val.clusterDim.z = 1;
  cfg.attrs = cattr;
  cfg.numAttrs = two ? 1 : 0;
  if (two) UAVDET_CUDA(cudaLaunchKernelEx(&cfg, wgrad_kernel<true>, mapDY, mapX[0], mapX[1], mapX[2], P));
  else UAVDET_CUDA(cudaLaunchKernelEx(&cfg, wgrad_kernel<false>, mapDY, mapX[0], mapX[1], mapX[2], P));
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}
