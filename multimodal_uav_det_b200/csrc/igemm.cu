// K1/K2 — implicit-GEMM convolution on tcgen05 tensor cores (forward and data-gradient).
//
//   D[pixel, cout] = sum_{tap} sum_{c} A_tap[pixel, c] * W[cout, tap, c]
//
// Replaces aten::convolution behind nn.Conv2d / F.conv2d at reference
// BaselineModel.py:13, _base.py:18,72-74,85,107, DySOEM_SimFPN.py:58,103-111,
// RTMUAVDet.py:19 and their autograd data-gradients.
//
// Design (B200-first, no im2col buffer):
//   * activations NHWC bf16; the A operand of every filter tap is one TMA box
//     {BK channels, tile_w, 1, tile_h, 1} of a 5-D tensor map, shifted by the tap offset —
//     TMA zero-fills out-of-image coordinates, which *is* the zero padding;
//   * stride-2 and the DySOEM space-to-depth gather use a parity view of the same tensor
//     ([2 pixels x C, W/2, 2, H/2, N]) so they stay plain tiled TMA;
//   * weights are a K-major [cout][taps*cin] bf16 matrix (optionally one per sample for the
//     dynamic-kernel convs) fetched by a 3-D map;
//   * 128 x block_n fp32 accumulators live in TMEM, double buffered so the epilogue of tile
//     i overlaps the MMAs of tile i+1; persistent CTAs, one per SM;
//   * warp roles: warps 0-2 = TMA producers, warp 3 = MMA issuer (+TMEM alloc), warps 4-11 =
//     epilogue (tcgen05.ld -> BN-stat partial sums / affine+activation+residual -> bf16 ->
//     swizzled staging -> TMA store); one kernel instance per epilogue kind (igemm_kernel<kKind>).
#include "common.cuh"
#include "sm100.cuh"
#include "igemm.h"

namespace uavdet {
using namespace sm100;

constexpr int kProdWarps = 3;      // TMA producer warps (P.prod_warps of them active)
constexpr int kMmaWarp = 3;        // tcgen05.mma issuer (+ TMEM alloc)
constexpr int kEpiWarp0 = 4;       // first of the 8 epilogue warps (12 warps = 3 per scheduler: 168 registers/thread)
constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kIgemmThreads = (kProdWarps + 1 + kEpiWarps) * 32;   // 384
constexpr int kAccStride = 256;    // TMEM columns per accumulator buffer

// n / d for 0 <= n < 2^31 without a hardware divide (d fixed per launch, constants from the host).
__device__ __forceinline__ int fast_div(int n, const FastDiv& f) {
  return f.d == 1 ? n : (int)(__umulhi((uint32_t)n, f.mul) >> f.shr);
}

struct TileCoord { int n0, ow0, oh0, img; };
__device__ __forceinline__ TileCoord decode_tile(const IgemmParams& P, int tile) {
  int m = fast_div(tile, P.fd_n);
  TileCoord t;
  t.n0 = (tile - m * P.n_tiles) * P.block_n;
  int m2 = fast_div(m, P.fd_w);
  t.ow0 = (m - m2 * P.tiles_w) * P.tile_w;
  t.img = fast_div(m2, P.fd_h);
  t.oh0 = (m2 - t.img * P.tiles_h) * P.tile_h;
  return t;
}

// Tile of the `it`-th iteration of this CTA's persistent loop, or -1 when the loop is over.
//   one CTA per tile:  tile = blockIdx.x + it * gridDim.x
//   CTA pair (kTwo):   the pair works on two consecutive pixel tiles (2*mp + rank) of the same channel block n, so
//                      that both CTAs share one B tile (each loads half); both leave the loop in the same iteration.
//                      With an odd number of pixel tiles the last pair's second tile lies past the last image: its
//                      loads are zero-filled and its stores clipped by TMA.
template <bool kTwo>
__device__ __forceinline__ int tile_at(const IgemmParams& P, int it, uint32_t rank) {
  if (!kTwo) {
    const int tile = (int)blockIdx.x + it * (int)gridDim.x;
    return tile < P.total_tiles ? tile : -1;
  }
  const int pt = (int)(blockIdx.x >> 1) + it * (int)(gridDim.x >> 1);
  if (pt >= P.total_pairs) return -1;
  const int mp = fast_div(pt, P.fd_n);
  const int n = pt - mp * P.n_tiles;
  return (2 * mp + (int)rank) * P.n_tiles + n;
}
template <bool kTwo>
__device__ __forceinline__ int num_iters(const IgemmParams& P) {
  const int first = kTwo ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int step = kTwo ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int total = kTwo ? P.total_pairs : P.total_tiles;
  return total > first ? (total - first + step - 1) / step : 0;
}

// 16 bytes of a per-channel epilogue vector: from the CTA's staged copy in shared memory (an asm load the compiler may
// schedule freely — the copy is read-only once the epilogue warps have passed their barrier; as a generic load it had to
// stay ordered with the staging-row stores and the ALU-bound GELU layers lost 15 %) or from global memory (read-only path).
__device__ __forceinline__ float4 load_epc4(const float* p, bool staged) {
  float4 r;
  if (staged) {
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
        : "r"((uint32_t)__cvta_generic_to_shared(p)));
  } else {
    r = __ldg(reinterpret_cast<const float4*>(p));
  }
  return r;
}

// 32 accumulator columns of this lane's row -> (scale, shift, act, residual), in place.
// kFold: the GroupNorm-fold form  v * rs + u[c]  — rs = rstd of the tile's image, u = the warp's per-image additive vector
// (passed as `shift`, always in shared memory): uavdet_epilogue::sample_affine.
template <int ACT, bool kFold = false>
__device__ __forceinline__ void affine_act(float (&v)[32], const float* scale, int cg, const float* shift,
                                           const uint4 (&res)[4], bool have_res, bool staged, float rs = 1.f) {
  // The element-wise math runs on packed fp32x2 instructions (fma.rn.f32x2 & co., two columns per instruction, each half
  // rounded like the scalar form): these epilogues are bound by instruction issue on their 8 warps.
  auto fma2 = [](float& a, float& b, float2 m, float2 c) {
    const float2 r = __ffma2_rn(make_float2(a, b), m, c);
    a = r.x; b = r.y;
  };
  if (kFold) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 t = load_epc4(shift + cg + i, true);
      fma2(v[i], v[i + 1], make_float2(rs, rs), make_float2(t.x, t.y));
      fma2(v[i + 2], v[i + 3], make_float2(rs, rs), make_float2(t.z, t.w));
    }
  } else if (scale && shift) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 s = load_epc4(scale + cg + i, staged);
      const float4 t = load_epc4(shift + cg + i, staged);
      fma2(v[i], v[i + 1], make_float2(s.x, s.y), make_float2(t.x, t.y));
      fma2(v[i + 2], v[i + 3], make_float2(s.z, s.w), make_float2(t.z, t.w));
    }
  } else if (scale) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 s = load_epc4(scale + cg + i, staged);
      v[i] *= s.x; v[i + 1] *= s.y; v[i + 2] *= s.z; v[i + 3] *= s.w;
    }
  } else if (shift) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 t = load_epc4(shift + cg + i, staged);
      const float2 a = __fadd2_rn(make_float2(v[i], v[i + 1]), make_float2(t.x, t.y));
      const float2 b = __fadd2_rn(make_float2(v[i + 2], v[i + 3]), make_float2(t.z, t.w));
      v[i] = a.x; v[i + 1] = a.y; v[i + 2] = b.x; v[i + 3] = b.y;
    }
  }
  if (ACT == UAVDET_ACT_GELU) {
#pragma unroll
    for (int i = 0; i < 32; i += 2) gelu_pair(v[i], v[i + 1]);
  } else if (ACT == UAVDET_ACT_SILU) {
#pragma unroll
    for (int i = 0; i < 32; i += 2) silu_pair(v[i], v[i + 1]);
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = act_fwd<ACT>(v[i]);
  }
  if (have_res) {
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
      const uint4 rr = res[i >> 3];
      const uint32_t w4[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 a = __fadd2_rn(make_float2(v[i + 2 * j], v[i + 2 * j + 1]), make_float2(bf16_lo(w4[j]), bf16_hi(w4[j])));
        v[i + 2 * j] = a.x; v[i + 2 * j + 1] = a.y;
      }
    }
  }
}

// One 32-column chunk of one accumulator row: TMEM -> fp32 -> epilogue math -> bf16 -> swizzled staging row.
// `chunk_in_slab` = which 32-column half of the (64-wide) slab; `sw_mask` = the row's swizzle XOR.
// Deliberately NOT inlined: it is called from 13 sites and carries the 5-way activation switch; inlined, the kernel
// grew to 37,000 instructions (595 KB) and the epilogue warps thrashed the instruction cache.
// Everything it needs from the kernel parameters arrives in registers (`cfg`, `scale`): read through a reference,
// the fields became a chain of control-dependent generic loads from the parameter bank (~1,000 cycles per call).
//   cfg bit 0: statistics epilogue; bits 1-3: activation; bit 4: residual operand; bit 5: it is in the staging row;
//   bit 6: scale / shift point into the CTA's staged copy in shared memory; bit 7: GroupNorm-fold form of the affine map
__device__ __forceinline__ uint32_t chunk_cfg(const IgemmParams& P) {
  return (P.epi == UAVDET_EPI_STATS ? 1u : 0u) | ((uint32_t)P.act << 1) | (P.res ? 16u : 0u) | (P.res_tma ? 32u : 0u) |
         ((P.epi != UAVDET_EPI_STATS && P.epi != UAVDET_EPI_HEAD && P.epc_floats > 0) ? 64u : 0u) |
         (P.sample_affine ? 128u : 0u);
}
// epilogue math of one 32-column chunk (accumulator values in r) + bf16 pack + swizzled staging store
// kKind: the kernel instance (see igemm_kernel): 0 statistics epilogue, 1 affine without activation, 2 affine with
// any activation.  Each instance only carries its own epilogue code.
template <int kKind>
__device__ __forceinline__ void finish_chunk(uint32_t cfg, const float* scale, const uint32_t (&r)[32], int cg, bool valid,
                                             const float* shift, const __nv_bfloat16* res_px, uint8_t* srow,
                                             int chunk_in_slab, int sw_mask, float sa_rs) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
  if (kKind == 0) {
    if (shift) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float4 s4 = __ldg(reinterpret_cast<const float4*>(shift + cg + i));
        v[i] += s4.x; v[i + 1] += s4.y; v[i + 2] += s4.z; v[i + 3] += s4.w;
      }
    }
  } else if (valid) {
    // residual operand of these 32 columns: already in the staging row (a TMA load put the residual tile exactly
    // where the result is about to be written, see `res_tma`), or fetched from global memory
    uint4 rr[4];
    const bool have_res = (cfg & 16u) != 0u;
    if (have_res) {
      if (cfg & 32u) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          rr[j] = *reinterpret_cast<const uint4*>(srow + ((chunk_in_slab * 4 + j) ^ sw_mask) * 16);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) rr[j] = __ldg(reinterpret_cast<const uint4*>(res_px + cg + 8 * j));
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) rr[j] = make_uint4(0u, 0u, 0u, 0u);
    }
    const bool staged = (cfg & 64u) != 0u;
    if (kKind == 1) {
      affine_act<UAVDET_ACT_NONE>(v, scale, cg, shift, rr, have_res, staged);
    } else if (kKind == 4) {
      affine_act<UAVDET_ACT_GELU>(v, scale, cg, shift, rr, have_res, staged);
    } else if (kKind == 6) {
      affine_act<UAVDET_ACT_GELU, true>(v, scale, cg, shift, rr, have_res, staged, sa_rs);
    } else if (kKind == 5) {
      affine_act<UAVDET_ACT_SILU>(v, scale, cg, shift, rr, have_res, staged);
    } else {
      switch ((cfg >> 1) & 7u) {
        case UAVDET_ACT_LEAKY: affine_act<UAVDET_ACT_LEAKY>(v, scale, cg, shift, rr, have_res, staged); break;
        case UAVDET_ACT_RELU:
          if (cfg & 128u) affine_act<UAVDET_ACT_RELU, true>(v, scale, cg, shift, rr, have_res, staged, sa_rs);
          else affine_act<UAVDET_ACT_RELU>(v, scale, cg, shift, rr, have_res, staged);
          break;
        default: affine_act<UAVDET_ACT_NONE>(v, scale, cg, shift, rr, have_res, staged); break;
      }
    }
  }
  // rows outside the image / tile carry garbage or bias only: stage zeros (the clipped TMA store skips them and
  // they must not enter the statistics)
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 o;
    o.x = valid ? pack_bf16x2(v[8 * j + 0], v[8 * j + 1]) : 0u;
    o.y = valid ? pack_bf16x2(v[8 * j + 2], v[8 * j + 3]) : 0u;
    o.z = valid ? pack_bf16x2(v[8 * j + 4], v[8 * j + 5]) : 0u;
    o.w = valid ? pack_bf16x2(v[8 * j + 6], v[8 * j + 7]) : 0u;
    const int chunk = (chunk_in_slab * 4 + j) ^ sw_mask;
    *reinterpret_cast<uint4*>(srow + chunk * 16) = o;
  }
}


template <int kKind, bool kTwo>
__device__ __noinline__ void stage_chunk(uint32_t cfg, const float* scale, uint32_t taddr, int cg, bool valid,
                                            const float* shift, const __nv_bfloat16* res_px, uint8_t* srow,
                                            int chunk_in_slab, int sw_mask, bool release, uint32_t tempty, int lane,
                                            float sa_rs = 1.f) {
  uint32_t r[32];
  tmem_ld_32x32(taddr, r);
  tmem_ld_wait();
  if (release) {
    // last TMEM read of this tile by this warp: hand the accumulator buffer back to the MMA warp (of the leader CTA:
    // `tempty` is then a shared::cluster address)
    tc_fence_before();
    __syncwarp();
    if (lane == 0) { if (kTwo) mbar_arrive_cluster(tempty); else mbar_arrive(tempty); }
  }
  finish_chunk<kKind>(cfg, scale, r, cg, valid, shift, res_px, srow, chunk_in_slab, sw_mask, sa_rs);
}

// MMA issue loop of one CTA (single elected thread), KSTEPS = block_k / 16.
template <int KSTEPS, bool kTwo>
__device__ __forceinline__ void mma_issue_loop(const IgemmParams& P, uint32_t ring_base, uint32_t bres_base,
                                               uint32_t full_bar, uint32_t empty_bar, uint32_t tfull_bar,
                                               uint32_t tempty_bar, uint32_t tmem_base, int a_bytes, int b_bytes,
                                               int stage_bytes, int num_kb, volatile uint32_t* dead) {
  const uint32_t idesc = make_idesc_bf16(P.block_n, 0, 0, kTwo ? 256 : 128);
  const uint32_t layout = (KSTEPS == 4) ? 2u : 4u;            // SWIZZLE_128B : SWIZZLE_64B
  const uint32_t sbo = 8u * (uint32_t)(KSTEPS * 16) * 2u;     // 8 rows of one swizzle atom
  const bool bres = !kTwo && P.bres_bytes > 0;      // resident weights: one-CTA kernel only
  const uint64_t a0 = make_smem_desc(ring_base, 16, sbo, layout);
  // B: inside the stage (after A), or tile kb of the resident weight region
  const uint64_t b0 = make_smem_desc(bres ? bres_base : ring_base + (uint32_t)a_bytes, 16, sbo, layout);
  const uint32_t b_step = (uint32_t)b_bytes >> 4;
  const uint32_t stage_step = (uint32_t)stage_bytes >> 4;
  const uint32_t last_stage = (uint32_t)P.stages - 1u;
  uint32_t stage = 0, phase = 0, soff = 0;
  uint32_t acc = 0, acc_phase = 0;
  const int iters = num_iters<kTwo>(P);
  for (int tl = 0; tl < iters; ++tl) {
    const bool tr = P.trace && blockIdx.x == 0 && tl < P.trace_tiles;
    if (tr) P.trace[tl * 16 + 2] = clock64();
    mbar_wait(tempty_bar + 8u * acc, acc_phase ^ 1u, dead, P.watchdog, 0x2u);
    tc_fence_after();
    const uint32_t d_tmem = tmem_base + acc * (uint32_t)kAccStride;
    uint32_t accumulate = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(full_bar + 8u * stage, phase, dead, P.watchdog, 0x4u);
      tc_fence_after();
      const uint64_t ad = a0 + (uint64_t)soff, bd = b0 + (uint64_t)(bres ? (uint32_t)kb * b_step : soff);
#pragma unroll
      for (int k = 0; k < KSTEPS; ++k) {
        // advance 16 bf16 = 32 bytes along K inside the swizzle row: +2 in the >>4 field
        if (kTwo) tc_mma_bf16_2sm(d_tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, accumulate);
        else tc_mma_bf16(d_tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, accumulate);
        accumulate = 1u;
      }
      // the stage is free again in BOTH CTAs of a pair (the MMA read both shared memories)
      if (kTwo) tc_commit_2sm(empty_bar + 8u * stage, 3); else tc_commit(empty_bar + 8u * stage);
      if (stage == last_stage) { stage = 0; phase ^= 1u; soff = 0; } else { ++stage; soff += stage_step; }
    }
    if (kTwo) tc_commit_2sm(tfull_bar + 8u * acc, 3); else tc_commit(tfull_bar + 8u * acc);
    if (tr) P.trace[tl * 16 + 4] = clock64();
    acc ^= 1u;
    if (acc == 0) acc_phase ^= 1u;
  }
}

// Halo mode: the A operand of every k-block of a tile is a shifted view of the tile's halo box (P.kb_aoff), stages
// carry the weight tile only (or nothing, when the weights are resident).
template <int KSTEPS>
__device__ __forceinline__ void mma_issue_loop_halo(const IgemmParams& P, uint32_t halo_base, uint32_t ring_base,
                                                    uint32_t bres_base, uint32_t full_bar, uint32_t empty_bar,
                                                    uint32_t afull_bar, uint32_t aempty_bar, uint32_t tfull_bar,
                                                    uint32_t tempty_bar, uint32_t tmem_base, int b_bytes, int num_kb,
                                                    volatile uint32_t* dead) {
  const uint32_t idesc = make_idesc_bf16(P.block_n, 0, 0, 128);
  const uint32_t a_layout = P.halo_row_bytes == 128 ? 2u : 4u;
  const uint32_t b_layout = (KSTEPS == 4) ? 2u : 4u;
  const bool bres = P.bres_bytes > 0;
  const uint64_t a0 = make_smem_desc(halo_base, 16, (uint32_t)P.halo_sbo, a_layout);
  const uint64_t b0 = make_smem_desc(bres ? bres_base : ring_base, 16, 8u * (uint32_t)(KSTEPS * 16) * 2u, b_layout);
  const uint32_t b_step = (uint32_t)b_bytes >> 4;
  const uint32_t abuf_step = (uint32_t)P.halo_buf_bytes >> 4;
  const uint32_t last_stage = (uint32_t)P.stages - 1u, last_abuf = (uint32_t)P.halo_bufs - 1u;
  uint32_t stage = 0, phase = 0, soff = 0;
  uint32_t abuf = 0, aphase = 0, aoff = 0;
  uint32_t acc = 0, acc_phase = 0;
  const int iters = num_iters<false>(P);
  for (int tl = 0; tl < iters; ++tl) {
    const bool tr = P.trace && blockIdx.x == 0 && tl < P.trace_tiles;
    if (tr) P.trace[tl * 16 + 2] = clock64();
    mbar_wait(tempty_bar + 8u * acc, acc_phase ^ 1u, dead, P.watchdog, 0x2u);
    if (tr) P.trace[tl * 16 + 3] = clock64();
    mbar_wait(afull_bar + 8u * abuf, aphase, dead, P.watchdog, 0x400u);
    if (tr) P.trace[tl * 16 + 13] = clock64();
    tc_fence_after();
    const uint32_t d_tmem = tmem_base + acc * (uint32_t)kAccStride;
    uint32_t accumulate = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      if (!bres) {
        mbar_wait(full_bar + 8u * stage, phase, dead, P.watchdog, 0x4u);
        tc_fence_after();
      }
      const uint64_t ad = a0 + (uint64_t)(aoff + P.kb_aoff[kb]);
      const uint64_t bd = b0 + (uint64_t)(bres ? (uint32_t)kb * b_step : soff);
#pragma unroll
      for (int k = 0; k < KSTEPS; ++k) {
        tc_mma_bf16(d_tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, accumulate);
        accumulate = 1u;
      }
      if (!bres) {
        tc_commit(empty_bar + 8u * stage);
        if (stage == last_stage) { stage = 0; phase ^= 1u; soff = 0; } else { ++stage; soff += b_step; }
      }
    }
    tc_commit(aempty_bar + 8u * abuf);
    tc_commit(tfull_bar + 8u * acc);
    if (tr) P.trace[tl * 16 + 4] = clock64();
    if (abuf == last_abuf) { abuf = 0; aphase ^= 1u; aoff = 0; } else { ++abuf; aoff += abuf_step; }
    acc ^= 1u;
    if (acc == 0) acc_phase ^= 1u;
  }
}

// kKind selects the epilogue the instance is compiled with: 0 = batch statistics (training forward), 1 = affine
// without activation (data gradients), 2 = affine with any activation (fused inference epilogues), 3 = detection
// head.  One kernel holding all of them was 160 KB of code, and the step time follows the kernel's code size.
// 4 / 5 = affine with GELU / SiLU only (the inference models' activations get instances of their own: the epilogue math of
// an instance is inlined per activation, and the ALU-bound GELU epilogue lost 8 % to unrelated code in its kernel);
// 6 = GELU behind the GroupNorm-fold form of the affine map (as a runtime branch inside instance 4 it cost that layer 20 %).
// kTwo: the CTA-pair variant (cluster of two CTAs on one TPC, tcgen05.mma.cta_group::2): a 256-pixel x block_n tile per
// pair, every CTA stages its own 128 pixel rows of A and HALF of the weight tile, so the shared-memory traffic per MMA
// (what bounds the one-CTA kernel on the K >= 1152 layers: operand reads + TMA fill = 96 KB per 512-cycle k-block
// against 128 B/clk) drops by a third.  The epilogue is unchanged: every CTA drains its own 128 accumulator rows.
// barriers + TMEM pointer + flags behind the staging buffers (kernel and launch_igemm agree on this size); the staged
// epilogue constants (IgemmParams::epc_floats) follow it
constexpr int kCtrlBytes = (8 * (2 * kMaxStages + 5) + 64 + 8 * 2 * kEpiWarps + 8 * 2 * kMaxHaloBufs + 15) & ~15;

template <int kKind, bool kTwo>
__global__ void __launch_bounds__(kIgemmThreads, 1)
igemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
             const __grid_constant__ CUtensorMap mapOut, const __grid_constant__ CUtensorMap mapOutTail,
             const __grid_constant__ CUtensorMap mapRes, const __grid_constant__ CUtensorMap mapResTail,
             const __grid_constant__ IgemmParams P) {
  // 1024-byte alignment is what SWIZZLE_128B needs for TMA and UMMA; no static shared memory is used,
  // so the dynamic window starts at the (aligned) base of the CTA's shared memory.
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();     // the next kernel of the stream may be launched; its CTAs take the SMs this grid leaves

  const uint32_t rank = kTwo ? cluster_ctarank() : 0u;
  const bool halo = !kTwo && P.halo != 0;
  const int a_bytes = halo ? 0 : 128 * P.block_k * 2;  // smem reserved for A per stage (halo mode: A has its own buffers)
  const int b_bytes = (kTwo ? P.block_n / 2 : P.block_n) * P.block_k * 2;   // this CTA's part of the B tile
  // Weights small enough to stay in shared memory for the whole kernel (thin layers, 1x1 convs up to 256->128) are
  // loaded once: [resident B: one tile per k-block][stages x A]; otherwise every stage carries its B tile.
  const bool bres = !kTwo && P.bres_bytes > 0;      // resident weights: one-CTA kernel only
  const int stage_bytes = bres ? a_bytes : a_bytes + b_bytes;
  const int a_tx = P.tile_w * P.tile_h * P.block_k * 2;  // bytes the A box actually delivers
  uint8_t* ring = smem + P.bres_bytes;                   // pipeline stages (1024-aligned: tiles are multiples of 1 KB)
  uint8_t* halo_buf = ring + (size_t)P.stages * stage_bytes; // halo mode: A buffers (1024-aligned like the stages)
  uint8_t* staging = halo_buf + (halo ? (size_t)P.halo_bufs * P.halo_buf_bytes : 0);  // output staging
  uint8_t* ctrl = staging + P.staging_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ctrl);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tfull_bar = empty_bar + kMaxStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* bres_bar = tempty_bar + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bres_bar + 1);
  volatile uint32_t* dead = tmem_ptr + 1;
  // residual-tile loads of the warp-private epilogues: one barrier per (epilogue warp, staging buffer)
  uint64_t* rres_bar = reinterpret_cast<uint64_t*>(ctrl + 8 * (2 * kMaxStages + 5) + 16);
  uint64_t* afull_bar = rres_bar + 2 * kEpiWarps;          // halo mode: one full / empty pair per A buffer
  uint64_t* aempty_bar = afull_bar + kMaxHaloBufs;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2 * kEpiWarps; ++i) mbar_init(smem_u32(&rres_bar[i]), 1);
    for (int s = 0; s < P.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(bres_bar), 1);
    for (int b = 0; b < kMaxHaloBufs; ++b) {
      mbar_init(smem_u32(&afull_bar[b]), 1);
      mbar_init(smem_u32(&aempty_bar[b]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tfull_bar[a]), 1);
      // arrivals per tile: all 8 epilogue warps (of both CTAs of a pair, on the leader's barrier), or only the 4
      // that own the accumulator buffer when a tile is a single slab (warp-private modes: the two warps of a lane
      // quarter then take alternate TILES)
      mbar_init(smem_u32(&tempty_bar[a]),
                kTwo ? 2 * kEpiWarps : (kKind != 3 && P.epi_mode != 0 && P.block_n == P.slab_w) ? kEpiWarps / 2 : kEpiWarps);
    }
    *dead = 0;
    fence_barrier_init();
    prefetch_tensormap(&mapA);
    prefetch_tensormap(&mapB);
    if (kKind != 3) prefetch_tensormap(&mapOut);
    if (P.res_tma) prefetch_tensormap(&mapRes);
  }
  if (warp == kMmaWarp) {
    if (kTwo) { tmem_alloc_2sm(smem_u32(tmem_ptr), 512); tmem_relinquish_2sm(); }
    else { tmem_alloc(smem_u32(tmem_ptr), 512); tmem_relinquish(); }
  }
  tc_fence_before();
  // a pair must see each other's initialised barriers before the first remote arrive / TMA completion
  if (kTwo) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // barriers, TMEM and descriptors are ready; nothing of the predecessor kernel's output has been touched yet
  pdl_wait();
  float* epc = reinterpret_cast<float*>(ctrl + kCtrlBytes);
  if (kKind != 0 && kKind != 3 && P.epc_floats > 0 && warp >= kEpiWarp0) {
    // epilogue constants -> shared memory (scale defaults to 1, shift to 0), visible to the 8 epilogue warps only
    for (int i = (warp - kEpiWarp0) * 32 + lane; i < P.epc_floats; i += 32 * kEpiWarps) {
      epc[i] = (P.scale && i < P.cout) ? __ldg(P.scale + i) : 1.f;
      epc[P.epc_floats + i] = (P.shift && i < P.cout) ? __ldg(P.shift + i) : 0.f;
    }
    // GroupNorm fold: the per-image (rstd, mean * rstd) pairs behind the warps' additive vectors (a global load per tile
    // sat on the critical path of the epilogue-bound layers this epilogue serves)
    if ((kKind == 2 || kKind == 6) && P.sa_staged)
      for (int i = (warp - kEpiWarp0) * 32 + lane; i < 2 * P.n_img; i += 32 * kEpiWarps)
        epc[(2 + kEpiWarps) * P.epc_floats + i] = __ldg(P.sample_affine + i);
    asm volatile("bar.sync 4, 256;" ::: "memory");
  }
  const bool use_epc = kKind != 0 && kKind != 3 && P.epc_floats > 0;
  const float* ep_scale = (use_epc && P.scale) ? epc : P.scale;       // an absent vector stays absent (no multiply by 1)

  const int num_kb = P.num_taps * P.kc_per_tap;

  if (warp < kProdWarps) {
    // ============================ TMA producers ===========================
    // k-block g (counted across this CTA's tiles) belongs to producer warp g % prod_warps; `stages` is a
    // multiple of prod_warps, so a pipeline stage is always refilled by the same warp (program order keeps the
    // two uses of its empty barrier apart).
    if (halo && warp == kProdWarps - 1) {
      // halo mode: this warp loads the tiles' A boxes (one per channel sub-block), the others the weight tiles
      if (elect_one()) {
        uint32_t abuf = 0, aphase = 0;
        for (int it = 0, tile; (tile = tile_at<kTwo>(P, it, rank)) >= 0; ++it) {
          const TileCoord tc = decode_tile(P, tile);
          mbar_wait<32>(smem_u32(&aempty_bar[abuf]), aphase ^ 1u, dead, P.watchdog, 0x800u);
          const uint32_t fb = smem_u32(&afull_bar[abuf]);
          mbar_arrive_expect_tx(fb, (uint32_t)(P.halo_subs * P.halo_tx));
          uint8_t* dst = halo_buf + (size_t)abuf * P.halo_buf_bytes;
          for (int sb = 0; sb < P.halo_subs; ++sb)
            tma_load_5d(smem_u32(dst + (size_t)sb * P.halo_sub_bytes), &mapA, fb, sb * P.halo_box_c, tc.ow0 + P.halo_w0,
                        P.halo_p0, tc.oh0 + P.halo_h0, tc.img);
          if (++abuf == (uint32_t)P.halo_bufs) { abuf = 0; aphase ^= 1u; }
        }
      }
    } else if (warp < P.prod_warps && elect_one() && !(halo && bres && warp != 0)) {
      const int pw = P.prod_warps;
      const int kcpt = P.kc_per_tap;
      int it = 0, tile = tile_at<kTwo>(P, 0, rank);
      int kb = warp;                      // k-block inside the tile (may run past num_kb: normalised below)
      uint32_t stage = (uint32_t)warp, phase = 0;
      int cur_tile = -1;
      TileCoord tc{0, 0, 0, 0};
      // CTA pair: every transaction byte of both CTAs is counted on the leader's full barrier
      const uint32_t full0 = kTwo ? mapa_shared(smem_u32(&full_bar[0]), 0) : smem_u32(&full_bar[0]);
      if (bres && warp == 0) {
        // resident weights: every (tap, channel chunk) tile once (n_tiles == 1, one weight matrix for all images)
        const uint32_t bb = smem_u32(bres_bar);
        mbar_arrive_expect_tx(bb, (uint32_t)P.bres_bytes);
        for (int t = 0, kbi = 0; t < P.num_taps; ++t)
          for (int kc = 0; kc < kcpt; ++kc, ++kbi)
            tma_load_3d(smem_u32(smem + (size_t)kbi * b_bytes), &mapB, bb, P.taps[t].w_koff + kc * P.block_k, 0, 0);
      }
      while (!(halo && bres)) {        // halo mode with resident weights: nothing left to stream
        while (kb >= num_kb) { kb -= num_kb; tile = tile_at<kTwo>(P, ++it, rank); if (tile < 0) break; }
        if (tile < 0) break;
        if (tile != cur_tile) { tc = decode_tile(P, tile); cur_tile = tile; }
        const int t = kcpt == 1 ? kb : kb / kcpt;
        const int kc = kb - t * kcpt;
        const ConvTap tap = P.taps[t];
        mbar_wait<32>(smem_u32(&empty_bar[stage]), phase ^ 1u, dead, P.watchdog, 0x1u);
        const uint32_t fb = full0 + 8u * stage;
        uint8_t* sa = ring + (size_t)stage * stage_bytes;
        if (kTwo) {
          if (rank == 0) mbar_arrive_expect_tx(smem_u32(&full_bar[stage]), (uint32_t)(2 * (a_tx + b_bytes)));
          tma_load_5d_2sm(smem_u32(sa), &mapA, fb, tap.c_off + kc * P.block_k, tc.ow0 + tap.dw, tap.p, tc.oh0 + tap.dh,
                          tc.img);
          tma_load_3d_2sm(smem_u32(sa + a_bytes), &mapB, fb, tap.w_koff + kc * P.block_k,
                          tc.n0 + (int)rank * (P.block_n / 2), 0);
        } else if (halo) {
          mbar_arrive_expect_tx(fb, (uint32_t)b_bytes);
          tma_load_3d(smem_u32(sa), &mapB, fb, tap.w_koff + kc * P.block_k, tc.n0, P.w_batch > 1 ? tc.img : 0);
        } else {
          mbar_arrive_expect_tx(fb, (uint32_t)(bres ? a_tx : a_tx + b_bytes));
          tma_load_5d(smem_u32(sa), &mapA, fb, tap.c_off + kc * P.block_k, tc.ow0 + tap.dw, tap.p, tc.oh0 + tap.dh,
                      tc.img);
          if (!bres)
            tma_load_3d(smem_u32(sa + a_bytes), &mapB, fb, tap.w_koff + kc * P.block_k, tc.n0, P.w_batch > 1 ? tc.img : 0);
        }
        kb += pw;
        stage += (uint32_t)pw;
        if (stage >= (uint32_t)P.stages) { stage -= (uint32_t)P.stages; phase ^= 1u; }
      }
    }
  } else if (warp == kMmaWarp) {
    // ============================ MMA issuer ==============================
    // One thread feeds the tensor pipe; at N = 64..128 an MMA takes only 32..64 cycles, so the per-k-block
    // instruction path must be minimal: descriptors are built once (stage 0) and advanced by adding the stage /
    // K-step offset in their (address >> 4) field, barrier addresses advance incrementally, the K-step loop is
    // unrolled at compile time.
    if (rank == 0 && elect_one()) {      // a pair's MMAs are issued by its leader CTA only
      if (bres) {
        mbar_wait(smem_u32(bres_bar), 0, dead, P.watchdog, 0x80u);      // resident weights have landed
        tc_fence_after();
      }
      if (halo) {
        if (P.block_k == 64)
          mma_issue_loop_halo<4>(P, smem_u32(halo_buf), smem_u32(ring), smem_u32(smem), smem_u32(full_bar), smem_u32(empty_bar),
                                 smem_u32(afull_bar), smem_u32(aempty_bar), smem_u32(tfull_bar), smem_u32(tempty_bar),
                                 tmem_base, b_bytes, num_kb, dead);
        else
          mma_issue_loop_halo<2>(P, smem_u32(halo_buf), smem_u32(ring), smem_u32(smem), smem_u32(full_bar), smem_u32(empty_bar),
                                 smem_u32(afull_bar), smem_u32(aempty_bar), smem_u32(tfull_bar), smem_u32(tempty_bar),
                                 tmem_base, b_bytes, num_kb, dead);
      } else if (P.block_k == 64) mma_issue_loop<4, kTwo>(P, smem_u32(ring), smem_u32(smem), smem_u32(full_bar), smem_u32(empty_bar),
                                                   smem_u32(tfull_bar), smem_u32(tempty_bar), tmem_base, a_bytes, b_bytes,
                                                   stage_bytes, num_kb, dead);
      else mma_issue_loop<2, kTwo>(P, smem_u32(ring), smem_u32(smem), smem_u32(full_bar), smem_u32(empty_bar),
                                   smem_u32(tfull_bar), smem_u32(tempty_bar), tmem_base, a_bytes, b_bytes, stage_bytes,
                                   num_kb, dead);
    }
  } else {
    // ============================ epilogue (8 warps) =======================
    // Two warps per TMEM lane quarter (q = warp % 4).  bf16 outputs go TMEM -> registers -> swizzled smem ->
    // TMA store, so global writes are full 128-byte rows whatever the channel stride of the tensor is.
    const int ew = warp - kEpiWarp0;
    const int q = warp & 3;
    const int half = ew >> 2;
    const int row = q * 32 + lane;
    const int hl = row / P.tile_w, wl = row - hl * P.tile_w;
    const int kp = P.tile_w * P.tile_h;
    const int n_slabs = P.block_n / P.slab_w;
    const int row_bytes = P.slab_w * 2;
    const int sw_mask = (P.slab_w == 64) ? (row & 7) : ((row >> 1) & 3);
    int acc = 0;
    const uint32_t ccfg = chunk_cfg(P);
    uint32_t acc_phase = 0;

    if (kKind == 3) {
      for (int tl = 0, tile; (tile = tile_at<kTwo>(P, tl, rank)) >= 0; ++tl) {
        const TileCoord tc = decode_tile(P, tile);
        const int oh = tc.oh0 + hl, ow = tc.ow0 + wl;
        const bool valid = (row < kp) && (oh < P.ho) && (ow < P.wo);
        mbar_wait<64>(smem_u32(&tfull_bar[acc]), acc_phase, dead, P.watchdog, 0x8u);
        tc_fence_after();
        if (half == 0) {
          uint32_t r[16];
          tmem_ld_32x16(tmem_base + (uint32_t)(acc * kAccStride) + ((uint32_t)(q * 32) << 16), r);
          tmem_ld_wait();
          if (valid) {
            const int A = P.head_anchors;
            const size_t hw = (size_t)P.ho * P.wo;
            const size_t px = (size_t)oh * P.wo + ow;
            for (int a = 0; a < A; ++a) {
              float o = __uint_as_float(r[a]) + (P.shift ? __ldg(P.shift + a) : 0.f);
              P.head_obj[((size_t)tc.img * A + a) * hw + px] = o;
            }
            for (int a = 0; a < A; ++a) {
              float4 b;
              b.x = __uint_as_float(r[A + 4 * a + 0]);
              b.y = __uint_as_float(r[A + 4 * a + 1]);
              b.z = __uint_as_float(r[A + 4 * a + 2]);
              b.w = __uint_as_float(r[A + 4 * a + 3]);
              if (P.shift) {
                b.x += __ldg(P.shift + A + 4 * a + 0); b.y += __ldg(P.shift + A + 4 * a + 1);
                b.z += __ldg(P.shift + A + 4 * a + 2); b.w += __ldg(P.shift + A + 4 * a + 3);
              }
              reinterpret_cast<float4*>(P.head_bbox)[((size_t)tc.img * A + a) * hw + px] = b;
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[acc]));
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    } else if (P.epi_mode != 0) {
      // ---- warp-private epilogue: every warp stages and TMA-stores its own 32 accumulator rows -----------
      //   mode 1: the 32 rows are a {bw x bh} rectangle of the output tile (tile_w | 32 or 32 | tile_w);
      //   mode 2: the tile spans full image rows of a pixel-dense tensor, so the rows are a run of 32
      //           pixels of the flattened [C][H*W][N] view (a shorter tail map serves the tile's last rows).
      // No CTA-wide synchronisation: the only waits are this warp's own TMA-store reads.
      // Work units (tile, slab) alternate between the two warps of a quarter.  Batch statistics are summed
      // per warp in registers (lane <-> column pair) and flushed once per CTA / change of channel block.
      const int wbuf_bytes = 32 * row_bytes;
      uint8_t* wbuf0 = staging + (size_t)ew * P.epi_bufs * wbuf_bytes;
      const int rows_here = min(32, max(0, kp - q * 32));            // live tile rows of this warp
      const bool use_tail = rows_here > 0 && rows_here < 32;
      const int bw = P.tile_w < 32 ? P.tile_w : 32;
      const int q_ow = (q * 32) % P.tile_w, q_oh = (q * 32) / P.tile_w;   // mode 1 box origin inside the tile
      const bool single = n_slabs == 1;
      float st[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { st[j][0] = st[j][1] = st[j][2] = st[j][3] = 0.f; }
      int cur_n0 = -1;
      uint32_t ucount = 0;
      uint32_t rphase = 0;                 // bit b: parity of the next residual load into staging buffer b
      // Take the next staging buffer (its last TMA store must be done reading it) and, when the layer adds a
      // residual, start the TMA load of the residual tile for (tile tc, slab starting at channel cs) into it --
      // same map geometry as the store, so the bytes land where the result will be written.  Issued before the
      // wait for the accumulator: the load's latency hides behind the tile's main loop.
      // `early`: called in the middle of the previous unit, whose buffer is the other one -- every store committed
      // so far must then be done reading (the latest one read the buffer being taken).
      auto acquire = [&](const TileCoord& tc, int cs, bool early = false) -> uint8_t* {
        const uint32_t b = (P.epi_bufs == 2) ? (ucount & 1u) : 0u;
        uint8_t* wbuf = wbuf0 + b * wbuf_bytes;
        ++ucount;
        if (lane == 0) {
          if (P.epi_bufs == 2 && !early) tma_store_wait_read<1>(); else tma_store_wait_read<0>();
          if (P.res_tma && rows_here > 0) {
            const uint32_t bar = smem_u32(&rres_bar[ew * 2 + b]);
            if (P.epi_mode == 1) {
              mbar_arrive_expect_tx(bar, (uint32_t)(32 * row_bytes));
              const int pl = P.out_cspan ? cs / P.out_cspan : 0;
              tma_load_5d(smem_u32(wbuf), &mapRes, bar, cs - pl * P.out_cspan, tc.ow0 + q_ow, pl, tc.oh0 + q_oh, tc.img);
            } else {
              mbar_arrive_expect_tx(bar, (uint32_t)(rows_here * row_bytes));
              tma_load_3d(smem_u32(wbuf), use_tail ? &mapResTail : &mapRes, bar, cs, tc.oh0 * P.wo + q * 32, tc.img);
            }
          }
        }
        __syncwarp();
        return wbuf;
      };
      auto flush_stats = [&](int n0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int sl = single ? 0 : half + 2 * j;
          if (sl < n_slabs && (!single || j == 0)) {
            float s0 = st[j][0], s1 = st[j][1], q0 = st[j][2], q1 = st[j][3];
            int pair = lane;
            bool owner = true;
            if (P.slab_w == 32) {
              s0 += __shfl_xor_sync(0xffffffffu, s0, 16); s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
              q0 += __shfl_xor_sync(0xffffffffu, q0, 16); q1 += __shfl_xor_sync(0xffffffffu, q1, 16);
              pair = lane & 15;
              owner = lane < 16;
            }
            const int col = n0 + sl * P.slab_w + 2 * pair;
            if (owner && col < P.cout) {
              // pixel-pair GEMMs: the columns of both pixels of a pair belong to the same channels
              const int ch = P.stat_mod ? col % P.stat_mod : col;
              atomicAdd(P.sum + ch, s0); atomicAdd(P.sum + ch + 1, s1);
              atomicAdd(P.sumsq + ch, q0); atomicAdd(P.sumsq + ch + 1, q1);
            }
          }
          st[j][0] = st[j][1] = st[j][2] = st[j][3] = 0.f;
        }
      };
      // single-slab tiles: this warp owns accumulator buffer `half` and visits every second tile
      const int tstep = single ? 2 : 1;
      if (single) acc = half;
      uint8_t* wbuf_pending = nullptr;     // staging buffer (+ residual load in flight) already taken for the next tile
      for (int tl = single ? half : 0, tile; (tile = tile_at<kTwo>(P, tl, rank)) >= 0; tl += tstep) {
        const TileCoord tc = decode_tile(P, tile);
        const int oh = tc.oh0 + hl, ow = tc.ow0 + wl;
        const bool valid = (row < kp) && (oh < P.ho) && (ow < P.wo) && (tc.img < P.n_img);
        const bool tr = P.trace && blockIdx.x == 0 && tl < P.trace_tiles && (ew & 3) == 0 && lane == 0;
        if (kKind == 0 && tc.n0 != cur_n0) {
          if (cur_n0 >= 0) flush_stats(cur_n0);
          cur_n0 = tc.n0;
        }
        const __nv_bfloat16* res_px =
            P.res ? P.res + (size_t)tc.img * P.res_sn + (size_t)oh * P.res_sh + (size_t)ow * P.res_sw : nullptr;
        const float* shift = !P.shift ? nullptr : use_epc ? epc + P.epc_floats : P.shift + (size_t)tc.img * P.shift_sn;
        // GroupNorm-fold epilogue: with the per-image pair (rstd, mean * rstd) of this tile's image the warp rebuilds its
        // own additive vector u[c] = shift[c] - mean * rstd * scale[c] in shared memory; the chunk math is then one load
        // and one FMA per element (acc * rstd + u[c]) like a plain bias.  (With scale, shift and both scalars live in
        // the chunk function the register allocator serialised the GELU chains of a chunk: +26 % on that layer.)
        float sa_rs = 1.f;
        if ((kKind == 2 || kKind == 6) && P.sample_affine) {
          const int simg = min(tc.img, P.n_img - 1);
          const float2 sa = P.sa_staged ? *reinterpret_cast<const float2*>(epc + (2 + kEpiWarps) * P.epc_floats + 2 * simg)
                                        : __ldg(reinterpret_cast<const float2*>(P.sample_affine) + simg);
          sa_rs = sa.x;
          float* uw = epc + (2 + ew) * P.epc_floats;
          __syncwarp();
          for (int j = lane; j < P.epc_floats; j += 32) uw[j] = fmaf(-sa.y, epc[j], epc[P.epc_floats + j]);
          __syncwarp();
          shift = uw;
        }
        if (tr) P.trace[tl * 16 + 5] = clock64();
        // this warp's first slab of the tile: buffer + residual load before the accumulator is awaited
        // (fused parity planes: the residual always arrives by TMA — launch_igemm guarantees it — so res_px is unused)
        const int sl_first = single ? 0 : half;
        uint8_t* wbuf_next = nullptr;
        uint8_t* wbuf_first = wbuf_pending ? wbuf_pending
                                           : ((sl_first < n_slabs) ? acquire(tc, tc.n0 + sl_first * P.slab_w) : nullptr);
        wbuf_pending = nullptr;
        mbar_wait<64>(smem_u32(&tfull_bar[acc]), acc_phase, dead, P.watchdog, 0x8u);
        if (tr) P.trace[tl * 16 + 6] = clock64();
        tc_fence_after();
        const uint32_t tbase = tmem_base + (uint32_t)(acc * kAccStride) + ((uint32_t)(q * 32) << 16);
        const uint32_t tempty = kTwo ? mapa_shared(smem_u32(&tempty_bar[acc]), 0) : smem_u32(&tempty_bar[acc]);
        {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int sl = single ? 0 : half + 2 * j;
            if (sl < n_slabs && (!single || j == 0)) {
              const int cs = tc.n0 + sl * P.slab_w;                  // first global channel of the slab
              if (tr) P.trace[tl * 16 + 8] = clock64();
              uint8_t* wbuf = (j == 0) ? wbuf_first : (wbuf_next ? wbuf_next : acquire(tc, cs));
              wbuf_next = nullptr;
              if (P.res_tma && rows_here > 0) {
                const uint32_t b = (uint32_t)(wbuf != wbuf0);
                mbar_wait<0>(smem_u32(&rres_bar[ew * 2 + b]), (rphase >> b) & 1u, dead, P.watchdog, 0x10u);
                rphase ^= 1u << b;
              }
              if (tr) P.trace[tl * 16 + 9] = clock64();
              const bool last = single || (sl + 2 >= n_slabs);
              uint8_t* srow = wbuf + lane * row_bytes;
              const int c0 = sl * P.slab_w;                          // accumulator column of the slab
              if (P.slab_w == 64) {
                stage_chunk<(kKind == 3 ? 1 : kKind), kTwo>(ccfg, ep_scale, tbase + (uint32_t)c0, tc.n0 + c0, valid, shift, res_px, srow, 0, sw_mask, false,
                            tempty, lane, sa_rs);
                // the residual tile of this warp's next slab of the tile starts travelling now (other buffer)
                if (P.res_tma && P.epi_bufs == 2 && !last) wbuf_next = acquire(tc, cs + 2 * P.slab_w, true);
                stage_chunk<(kKind == 3 ? 1 : kKind), kTwo>(ccfg, ep_scale, tbase + (uint32_t)(c0 + 32), tc.n0 + c0 + 32, valid, shift, res_px, srow, 1, sw_mask,
                            last, tempty, lane, sa_rs);
              } else {
                stage_chunk<(kKind == 3 ? 1 : kKind), kTwo>(ccfg, ep_scale, tbase + (uint32_t)c0, tc.n0 + c0, valid, shift, res_px, srow, 0, sw_mask, last, tempty,
                            lane, sa_rs);
              }
              if (tr) P.trace[tl * 16 + 10] = clock64();
              fence_proxy_async();
              __syncwarp();
              if (tr) P.trace[tl * 16 + 11] = clock64();
              if (lane == 0) {
                if (P.epi_mode == 1) {
                  if (rows_here > 0) {
                    const int pl = P.out_cspan ? cs / P.out_cspan : 0;
                    tma_store_5d(&mapOut, smem_u32(wbuf), cs - pl * P.out_cspan, tc.ow0 + q_ow, pl, tc.oh0 + q_oh, tc.img);
                    tma_store_commit();
                  }
                } else if (rows_here > 0) {
                  tma_store_3d(use_tail ? &mapOutTail : &mapOut, smem_u32(wbuf), cs, tc.oh0 * P.wo + q * 32, tc.img);
                  tma_store_commit();
                }
              }
              if (tr) P.trace[tl * 16 + 12] = clock64();
              if (kKind == 0) {
                // column sums of the staged (bf16-rounded) rows: exactly the values the BatchNorm that follows
                // will normalise.  Conflict-free: a row is one 128-byte line (slab 64) / two rows are (slab 32).
                float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
                if (P.slab_w == 64) {
#pragma unroll 8
                  for (int r_ = 0; r_ < 32; ++r_) {
                    const int word = (((lane >> 2) ^ (r_ & 7)) << 2) | (lane & 3);
                    const uint32_t u = *reinterpret_cast<const uint32_t*>(wbuf + r_ * 128 + word * 4);
                    const float a = bf16_lo(u), b = bf16_hi(u);
                    s0 += a; s1 += b; q0 = fmaf(a, a, q0); q1 = fmaf(b, b, q1);
                  }
                } else {
                  const int pair = lane & 15;
#pragma unroll 8
                  for (int i = 0; i < 16; ++i) {
                    const int r_ = 2 * i + (lane >> 4);
                    const int word = (((pair >> 2) ^ ((r_ >> 1) & 3)) << 2) | (pair & 3);
                    const uint32_t u = *reinterpret_cast<const uint32_t*>(wbuf + r_ * 64 + word * 4);
                    const float a = bf16_lo(u), b = bf16_hi(u);
                    s0 += a; s1 += b; q0 = fmaf(a, a, q0); q1 = fmaf(b, b, q1);
                  }
                }
                st[j][0] += s0; st[j][1] += s1; st[j][2] += q0; st[j][3] += q1;
              }
            }
          }
        }
        if (tr) P.trace[tl * 16 + 7] = clock64();
        if (single) acc_phase ^= 1u;
        else if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        // The residual tile of this warp's first slab of its NEXT tile starts travelling now, into the other staging
        // buffer: the 1x1 data gradients are one short k-block per tile, and the load's round trip (issued only after
        // the tile had been awaited) was what their epilogue waited for.
        if (kKind != 0 && P.res_tma == 2 && P.epi_bufs == 2 && sl_first < n_slabs) {
          const int ntile = tile_at<kTwo>(P, tl + tstep, rank);
          if (ntile >= 0) {
            const TileCoord ntc = decode_tile(P, ntile);
            wbuf_pending = acquire(ntc, ntc.n0 + sl_first * P.slab_w);
          }
        }
      }
      if (kKind == 0 && cur_n0 >= 0) flush_stats(cur_n0);
      if (lane == 0) tma_store_wait_all();   // smem must stay valid until the last store has read it
    } else {
      // ---- CTA-wide epilogue (tiles whose 32-row groups are not rectangles) ------------------------------
      // The 8 warps fill one 128-row slab (two 32-column halves), one thread TMA-stores it; two slabs in
      // flight.  Statistics: shuffle-folded column sums of the staged slab, one global reduction per column.
      const int e_tid = ew * 32 + lane;
      const int slab_bytes = 128 * row_bytes;
      const int chunks_per_slab = P.slab_w / 32;              // 2 (slab 64) or 1 (slab 32)
      uint32_t slab_counter = 0;
      for (int tl = 0, tile; (tile = tile_at<kTwo>(P, tl, rank)) >= 0; ++tl) {
        const TileCoord tc = decode_tile(P, tile);
        const int oh = tc.oh0 + hl, ow = tc.ow0 + wl;
        const bool valid = (row < kp) && (oh < P.ho) && (ow < P.wo) && (tc.img < P.n_img);
        const bool tr = P.trace && blockIdx.x == 0 && tl < P.trace_tiles && ew == 0 && lane == 0;
        const __nv_bfloat16* res_px =
            P.res ? P.res + (size_t)tc.img * P.res_sn + (size_t)oh * P.res_sh + (size_t)ow * P.res_sw : nullptr;
        const float* shift = !P.shift ? nullptr : use_epc ? epc + P.epc_floats : P.shift + (size_t)tc.img * P.shift_sn;
        // GroupNorm-fold epilogue: with the per-image pair (rstd, mean * rstd) of this tile's image the warp rebuilds its
        // own additive vector u[c] = shift[c] - mean * rstd * scale[c] in shared memory; the chunk math is then one load
        // and one FMA per element (acc * rstd + u[c]) like a plain bias.  (With scale, shift and both scalars live in
        // the chunk function the register allocator serialised the GELU chains of a chunk: +26 % on that layer.)
        float sa_rs = 1.f;
        if ((kKind == 2 || kKind == 6) && P.sample_affine) {
          const int simg = min(tc.img, P.n_img - 1);
          const float2 sa = P.sa_staged ? *reinterpret_cast<const float2*>(epc + (2 + kEpiWarps) * P.epc_floats + 2 * simg)
                                        : __ldg(reinterpret_cast<const float2*>(P.sample_affine) + simg);
          sa_rs = sa.x;
          float* uw = epc + (2 + ew) * P.epc_floats;
          __syncwarp();
          for (int j = lane; j < P.epc_floats; j += 32) uw[j] = fmaf(-sa.y, epc[j], epc[P.epc_floats + j]);
          __syncwarp();
          shift = uw;
        }
        if (tr) P.trace[tl * 16 + 5] = clock64();
        mbar_wait<64>(smem_u32(&tfull_bar[acc]), acc_phase, dead, P.watchdog, 0x8u);
        if (tr) P.trace[tl * 16 + 6] = clock64();
        tc_fence_after();
        const uint32_t tbase = tmem_base + (uint32_t)(acc * kAccStride) + ((uint32_t)(q * 32) << 16);
        const uint32_t tempty = kTwo ? mapa_shared(smem_u32(&tempty_bar[acc]), 0) : smem_u32(&tempty_bar[acc]);
        for (int sl = 0; sl < n_slabs; ++sl, ++slab_counter) {
          const int cs = tc.n0 + sl * P.slab_w;
          // fused parity planes: column cs + i of the GEMM is channel (cs + i) % cspan of plane cs / cspan — shift the
          // residual base so that `res + column` lands there
          const __nv_bfloat16* res_sl =
              (res_px && P.out_cspan) ? res_px + (long long)(cs / P.out_cspan) * (P.res_sp - P.out_cspan) : res_px;
          uint8_t* sbuf = staging + (slab_counter & 1u) * slab_bytes;
          if (e_tid == 0) tma_store_wait_read<1>();           // the store that read this buffer two slabs ago
          asm volatile("bar.sync 2, 256;" ::: "memory");
          const bool last = sl == n_slabs - 1;
          if (half < chunks_per_slab) {
            const int c0 = sl * P.slab_w + half * 32;
            stage_chunk<(kKind == 3 ? 1 : kKind), kTwo>(ccfg, ep_scale, tbase + (uint32_t)c0, tc.n0 + c0, valid, shift, res_sl, sbuf + row * row_bytes, half,
                        sw_mask, last, tempty, lane, sa_rs);
            fence_proxy_async();
          } else if (last) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (kTwo) mbar_arrive_cluster(tempty); else mbar_arrive(tempty); }
          }
          asm volatile("bar.sync 3, 256;" ::: "memory");
          if (e_tid == 0) {
            const int pl = P.out_cspan ? cs / P.out_cspan : 0;
            tma_store_5d(&mapOut, smem_u32(sbuf), cs - pl * P.out_cspan, tc.ow0, pl, tc.oh0, tc.img);
            tma_store_commit();
          }
          if (kKind == 0) {
            // warp -> `ppw` column pairs, lane -> (pair, row group g) with rows g, g+groups, ... so the 32
            // lanes of a load hit 32 different banks; row groups are folded with xor-shuffles.
            const int ppw = P.slab_w >> 4;                    // 4 (slab 64) | 2 (slab 32)
            const int groups = 32 / ppw;                      // 8 | 16
            const int p_ = lane & (ppw - 1), g = lane / ppw;
            const int pr = ew * ppw + p_;                     // column pair inside the slab
            float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll 4
            for (int r_ = g; r_ < 128; r_ += groups) {
              const int m_ = (P.slab_w == 64) ? (r_ & 7) : ((r_ >> 1) & 3);
              const int word = (((pr >> 2) ^ m_) << 2) | (pr & 3);
              const uint32_t u = *reinterpret_cast<const uint32_t*>(sbuf + r_ * row_bytes + word * 4);
              const float a = bf16_lo(u), b = bf16_hi(u);
              s0 += a; s1 += b; q0 = fmaf(a, a, q0); q1 = fmaf(b, b, q1);
            }
            for (int off = ppw; off < 32; off <<= 1) {
              s0 += __shfl_xor_sync(0xffffffffu, s0, off);
              s1 += __shfl_xor_sync(0xffffffffu, s1, off);
              q0 += __shfl_xor_sync(0xffffffffu, q0, off);
              q1 += __shfl_xor_sync(0xffffffffu, q1, off);
            }
            const int col = cs + 2 * pr;
            if (lane < ppw && col < P.cout) {
              const int ch = P.stat_mod ? col % P.stat_mod : col;
              atomicAdd(P.sum + ch, s0); atomicAdd(P.sum + ch + 1, s1);
              atomicAdd(P.sumsq + ch, q0); atomicAdd(P.sumsq + ch + 1, q1);
            }
          }
        }
        if (tr) P.trace[tl * 16 + 7] = clock64();
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
      if (e_tid == 0) tma_store_wait_all();
    }
  }

  tc_fence_before();
  // a pair leaves together: the leader's MMAs read the peer's shared memory and both signal each other's barriers
  if (kTwo) cluster_sync_all(); else __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    if (kTwo) tmem_dealloc_2sm(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess || !p) {
      return nullptr;
    }
    fn = (EncodeTiledFn)p;
  }
  return fn;
}

int encode_tensor_map(CUtensorMap* m, void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return UAVDET_ERR_CUDA; }
  cuuint64_t gd[5]; cuuint64_t gs[4]; cuuint32_t bx[5]; cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                        : swizzle_bytes == 64  ? CU_TENSOR_MAP_SWIZZLE_64B
                        : swizzle_bytes == 32  ? CU_TENSOR_MAP_SWIZZLE_32B
                                               : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, base, gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank=%d dims=[%llu,%llu,%llu,%llu,%llu] box=[%u,%u,%u,%u,%u] "
              "stride0=%llu base=%p",
              (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
              (unsigned long long)(rank > 4 ? dims[4] : 0), box[0], rank > 1 ? box[1] : 0,
              rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0, rank > 4 ? box[4] : 0,
              (unsigned long long)strides_bytes[0], base);
    return UAVDET_ERR_CUDA;
  }
  return UAVDET_OK;
}

// 5-D activation map.  parity == 0: [C, W, 1, H, N];  parity == 1: [ld + C, W/2, 2, H/2, N].
int make_act_map(CUtensorMap* m, const uavdet_act* x, int parity, int box_c, int box_w, int box_h, int box_p) {
  const uint64_t eb = 2;
  uint64_t dims[5], str[4];
  uint32_t box[5] = {(uint32_t)box_c, (uint32_t)box_w, (uint32_t)box_p, (uint32_t)box_h, 1u};
  if (!parity) {
    dims[0] = x->c; dims[1] = x->w; dims[2] = 1; dims[3] = x->h; dims[4] = x->n;
    str[0] = (uint64_t)x->ld * eb;
    str[1] = (uint64_t)x->w * x->ld * eb;
    str[2] = (uint64_t)x->w * x->ld * eb;
    str[3] = (uint64_t)x->h * x->w * x->ld * eb;
  } else {
    dims[0] = (uint64_t)x->ld + x->c; dims[1] = x->w / 2; dims[2] = 2; dims[3] = x->h / 2; dims[4] = x->n;
    str[0] = (uint64_t)2 * x->ld * eb;
    str[1] = (uint64_t)x->w * x->ld * eb;
    str[2] = (uint64_t)2 * x->w * x->ld * eb;
    str[3] = (uint64_t)x->h * x->w * x->ld * eb;
  }
  return encode_tensor_map(m, x->ptr, 5, dims, str, box, box_c * 2);
}

static long long* g_trace = nullptr;
static int g_trace_tiles = 0;
void get_trace(long long** ptr, int* tiles) { *ptr = g_trace; *tiles = g_trace_tiles; }

// 5-D output map for the TMA-store epilogue: [C, Wo, 1, Ho, N] with arbitrary pixel strides (elements), so the
// same code writes plain NHWC tensors, channel slices and the parity planes of a stride-2 data gradient.
int make_out_map(CUtensorMap* m, void* ptr, int n, int ho, int wo, int c, long long sn, long long sh, long long sw,
                 int box_c, int box_w, int box_h, int planes, long long sp) {
  uint64_t dims[5] = {(uint64_t)c, (uint64_t)wo, (uint64_t)planes, (uint64_t)ho, (uint64_t)n};
  uint64_t str[4] = {(uint64_t)sw * 2, (uint64_t)(planes > 1 ? sp : sh) * 2, (uint64_t)sh * 2, (uint64_t)sn * 2};
  uint32_t box[5] = {(uint32_t)box_c, (uint32_t)box_w, 1u, (uint32_t)box_h, 1u};
  return encode_tensor_map(m, ptr, 5, dims, str, box, box_c * 2);
}

// Pick the output tile rectangle (tile_w * tile_h <= 128) and the epilogue flavour it allows.
//   mode 1: every 32-row group of the tile is a rectangle (tile_w | 32 with whole box rows, or 32 | tile_w);
//   mode 2: the tile spans full image rows of a pixel-dense tensor (32-row groups are pixel runs);
//   mode 0: anything else (CTA-wide slab epilogue).
// The warp-private modes are taken unless they waste > 10 % more MMA rows than the best rectangle; ties go to
// the squarest tile (smallest halo, best L2 reuse across the filter taps).
void choose_tile(int ho, int wo, bool dense_rows, int* tile_w, int* tile_h, int* epi_mode) {
  double best_eff[2] = {-1.0, -1.0};          // [0] any tile, [1] warp-private tiles
  int best_halo[2] = {1 << 30, 1 << 30}, bw[2] = {1, 1}, bh[2] = {1, 1}, bm[2] = {0, 0};
  const int max_w = wo < 128 ? wo : 128;
  for (int tw = 1; tw <= max_w; ++tw) {
    int th = 128 / tw;
    if (th > ho) th = ho;
    for (int pass = 0; pass < 2; ++pass) {
      int t_h = th, mode = 0;
      if (pass == 1) {
        if (tw == wo && dense_rows && (long long)ho * wo >= 32) {
          mode = 2;
        } else if (32 % tw == 0) {
          const int rows = 32 / tw;
          t_h = (th / rows) * rows;
          if (t_h == 0) continue;
          mode = 1;
        } else if (tw % 32 == 0) {
          mode = 1;
        } else {
          continue;
        }
      }
      const double tiles = (double)ceil_div(ho, t_h) * ceil_div(wo, tw);
      const double eff = (double)ho * wo / (tiles * 128.0);
      const int halo = (tw + 2) * (t_h + 2);
      if (eff > best_eff[pass] + 1e-9 || (eff > best_eff[pass] - 1e-9 && halo < best_halo[pass])) {
        best_eff[pass] = eff; best_halo[pass] = halo; bw[pass] = tw; bh[pass] = t_h; bm[pass] = mode;
      }
    }
  }
  const int pick = (best_eff[1] >= 0.9 * best_eff[0]) ? 1 : 0;
  *tile_w = bw[pick];
  *tile_h = bh[pick];
  *epi_mode = bm[pick];
}

__global__ void fill_plane_kernel(IgemmParams P) {
  // out[n, oh, ow, 0:cout] = res or 0 over one (strided) plane; 8 channels per thread
  const int c8 = P.cout / 8;
  const long long total = (long long)P.n_img * P.ho * P.wo * c8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c8) * 8;
    long long px = i / c8;
    const int ow = (int)(px % P.wo); px /= P.wo;
    const int oh = (int)(px % P.ho);
    const int n = (int)(px / P.ho);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (P.res) v = *reinterpret_cast<const uint4*>(P.res + n * P.res_sn + oh * P.res_sh + ow * P.res_sw + c);
    *reinterpret_cast<uint4*>(P.out + n * P.out_sn + oh * P.out_sh + ow * P.out_sw + c) = v;
  }
}

int fill_plane(const IgemmParams& P, cudaStream_t st) {
  const long long total = (long long)P.n_img * P.ho * P.wo * (P.cout / 8);
  if (total <= 0) return UAVDET_OK;
  long long blocks = (total + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  fill_plane_kernel<<<(int)blocks, 256, 0, st>>>(P);
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}

// UAVDET_IGEMM_RES_TMA=0 keeps the residual operand on the per-lane global loads (debug / A-B switch)
static bool res_tma_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("UAVDET_IGEMM_RES_TMA");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

// UAVDET_IGEMM_2CTA: 0 = never use the CTA-pair kernel, 1 (default) = where it applies, 2 = wherever it is legal
static int two_cta_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("UAVDET_IGEMM_2CTA");
    v = e ? atoi(e) : 1;
    if (v < 0 || v > 2) v = 1;
  }
  return v;
}

// UAVDET_IGEMM_HALO: 0 = never, 1 (default) = where it measured faster (below), 2 = the same rule for any multiple of
// 64 input channels, 3 = wherever it is legal.
// Measured on B200 (tools/trace_halo.py, batch 32 / 16 at 320x320): with the halo box the A operand is in shared memory
// ~80 cycles after the tile starts, yet a tile still takes ~2,600 cycles of MMA time on the thin layers — the tensor
// core needs ~72 cycles for every M128 x N x K16 instruction whose N is <= 128 (it re-reads the 4 KB A slice per
// instruction) and ~144 when the rows are 64 bytes (SWIZZLE_64B, 32-channel layers).  So 32 -> 64 and 64 -> 32 3x3 layers
// are bound by instruction count, not by L2 -> SM traffic, and the halo box only pays where the per-tap boxes were
// the slower side: 64 input channels, >= 64 output channels, stride 1, one weight matrix (64 -> 64 at 320x320: 228 ->
// 197 us; 64 -> 128 at 160x160 inside the training step: 141 -> 104 us; stride-2 and per-sample-weight layers lost
// 7-35 %: two boxes per tile, and weight tiles that travel alone in 4 KB stages; 128 -> 64 data gradients at 160x160
// lost 20 %, 128 -> 128 at 160x160 lost 9 %, 256 -> 256 at 80x80 gained 3 % — hence 64 input channels only).
static int halo_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("UAVDET_IGEMM_HALO");
    v = e ? atoi(e) : 1;
    if (v < 0 || v > 3) v = 1;
  }
  return v;
}

// Halo-mode geometry of a launch (see IgemmParams): decides the box, the per-k-block A offsets and forces the 8 x 16
// tile.  Returns false (P untouched except the halo fields) when the layer does not qualify.
static bool plan_halo(IgemmParams& P, const uavdet_act* a_src, int parity, int w_batch) {
  P.halo = 0;
  const int mode = halo_mode();
  if (mode == 0 || P.epi == UAVDET_EPI_HEAD || P.num_taps < 4) return false;
  if (P.wo % 8 != 0 || P.ho < 8) return false;
  const int C = a_src->c;
  if (mode == 1 && C != 64) return false;
  if (mode <= 2 && (C % 64 != 0 || parity || w_batch != 1 || P.block_n < 64)) return false;
  if (parity && a_src->ld != C) return false;           // the two pixels of a parity pair must be contiguous
  const int span = parity ? 2 * C : C;
  const int box_c = span < 64 ? span : 64;
  if (span % box_c != 0) return false;
  const int subs = span / box_c;
  const int num_kb = P.num_taps * P.kc_per_tap;
  if (num_kb > kMaxKb) return false;
  int dw0 = 1 << 30, dw1 = -(1 << 30), dh0 = 1 << 30, dh1 = -(1 << 30), p0 = 1 << 30, p1 = -(1 << 30);
  for (int t = 0; t < P.num_taps; ++t) {
    const ConvTap& tp = P.taps[t];
    dw0 = tp.dw < dw0 ? tp.dw : dw0; dw1 = tp.dw > dw1 ? tp.dw : dw1;
    dh0 = tp.dh < dh0 ? tp.dh : dh0; dh1 = tp.dh > dh1 ? tp.dh : dh1;
    p0 = tp.p < p0 ? tp.p : p0; p1 = tp.p > p1 ? tp.p : p1;
  }
  const int tile_w = 8, tile_h = 16;
  const int hw_w = tile_w + dw1 - dw0, hw_p = p1 - p0 + 1, hw_h = tile_h + dh1 - dh0;
  if (hw_w > 256 || hw_h > 256) return false;
  const int row_bytes = box_c * 2;
  const long long rows = (long long)hw_w * hw_p * hw_h;
  const long long sub_bytes = (rows * row_bytes + 1023) / 1024 * 1024;
  {
    // two A buffers, two weight stages, one staging set and the barriers must fit (launch_igemm hands out the rest)
    const long long b_tile = (long long)P.block_n * P.block_k * 2;
    const long long staging1 = 2ll * 128 * ((P.block_n % 64 == 0) ? 64 : 32) * 2;
    if (2 * sub_bytes * subs + 2 * b_tile + staging1 + 1024 > 227 * 1024) return false;
  }
  for (int t = 0; t < P.num_taps; ++t) {
    const ConvTap& tp = P.taps[t];
    for (int kc = 0; kc < P.kc_per_tap; ++kc) {
      const int cpos = tp.c_off + kc * P.block_k;
      if (cpos < 0 || cpos + P.block_k > span) return false;
      const int sub = cpos / box_c, inrow = cpos % box_c;
      if (inrow + P.block_k > box_c) return false;
      const long long row = ((long long)(tp.dh - dh0) * hw_p + (tp.p - p0)) * hw_w + (tp.dw - dw0);
      P.kb_aoff[t * P.kc_per_tap + kc] = (uint32_t)((sub * sub_bytes + row * row_bytes + inrow * 2) >> 4);
    }
  }
  P.halo = 1;
  P.halo_subs = subs;
  P.halo_sub_bytes = (int)sub_bytes;
  P.halo_buf_bytes = (int)(sub_bytes * subs);
  P.halo_tx = (int)(rows * row_bytes);
  P.halo_row_bytes = row_bytes;
  P.halo_sbo = hw_p * hw_w * row_bytes;
  P.halo_box_c = box_c;
  P.halo_w0 = dw0; P.halo_p0 = p0; P.halo_h0 = dh0;
  P.halo_bufs = 2;
  P.tile_w = tile_w; P.tile_h = tile_h; P.epi_mode = 1;
  // scratch for launch_igemm (box geometry of the A map)
  P.tiles_w = hw_w; P.tiles_h = hw_h; P.n_tiles = hw_p;
  return true;
}

static int pick_block_n(int cout) {
  if (cout <= 16) return 16;
  for (int bn = 256; bn >= 32; bn -= 32)
    if (cout % bn == 0) return bn;
  return cout < 256 ? ((cout + 31) / 32) * 32 : 256;
}

int launch_igemm(const uavdet_act* a_src, int parity, const void* w_packed, int w_rows, int k_total,
                 int w_batch, IgemmParams& P, cudaStream_t st, int w_batch_rows = 0) {
  if (w_batch_rows <= 0) w_batch_rows = w_rows;   // rows between the weight matrices of consecutive samples
  CUtensorMap mapA, mapB, mapOut, mapOutTail, mapRes, mapResTail;
  int rc;
  if (plan_halo(P, a_src, parity, w_batch)) rc = make_act_map(&mapA, a_src, parity, P.halo_box_c, P.tiles_w, P.tiles_h, P.n_tiles);
  else rc = make_act_map(&mapA, a_src, parity, P.block_k, P.tile_w, P.tile_h);
  if (rc) return rc;
  P.res_tma = 0;
  P.slab_w = (P.block_n % 64 == 0) ? 64 : 32;
  const int kp = P.tile_w * P.tile_h;
  if (P.epi == UAVDET_EPI_HEAD) {
    P.epi_mode = 0;
    mapOut = mapA;   // unused by the HEAD epilogue
    mapOutTail = mapA;
  } else if (P.epi_mode == 1) {
    const int bw = P.tile_w < 32 ? P.tile_w : 32;
    const int planes = P.out_cspan ? P.cout / P.out_cspan : 1;
    const int map_c = P.out_cspan ? P.out_cspan : P.cout;
    rc = make_out_map(&mapOut, P.out, P.n_img, P.ho, P.wo, map_c, P.out_sn, P.out_sh, P.out_sw, P.slab_w, bw, 32 / bw,
                      planes, P.out_sp);
    if (rc) return rc;
    mapOutTail = mapOut;
    if (P.res && (res_tma_enabled() || P.out_cspan)) {
      // the residual tile comes in through the same box as the result goes out
      rc = make_out_map(&mapRes, const_cast<__nv_bfloat16*>(P.res), P.n_img, P.ho, P.wo, map_c, P.res_sn, P.res_sh,
                        P.res_sw, P.slab_w, bw, 32 / bw, planes, P.res_sp);
      if (rc) return rc;
      mapResTail = mapRes;
      P.res_tma = 1;
    }
  } else if (P.epi_mode == 2) {
    uint64_t dims[3] = {(uint64_t)P.cout, (uint64_t)P.ho * P.wo, (uint64_t)P.n_img};
    uint64_t str[2] = {(uint64_t)P.out_sw * 2, (uint64_t)P.out_sn * 2};
    uint32_t box[3] = {(uint32_t)P.slab_w, 32u, 1u};
    rc = encode_tensor_map(&mapOut, P.out, 3, dims, str, box, P.slab_w * 2);
    if (rc) return rc;
    mapOutTail = mapOut;
    if (kp % 32) {
      box[1] = (uint32_t)(kp % 32);
      rc = encode_tensor_map(&mapOutTail, P.out, 3, dims, str, box, P.slab_w * 2);
      if (rc) return rc;
    }
    if (P.res && res_tma_enabled() && P.res_sh == (long long)P.wo * P.res_sw) {   // pixel-dense residual
      uint64_t rstr[2] = {(uint64_t)P.res_sw * 2, (uint64_t)P.res_sn * 2};
      box[1] = 32u;
      rc = encode_tensor_map(&mapRes, const_cast<__nv_bfloat16*>(P.res), 3, dims, rstr, box, P.slab_w * 2);
      if (rc) return rc;
      mapResTail = mapRes;
      if (kp % 32) {
        box[1] = (uint32_t)(kp % 32);
        rc = encode_tensor_map(&mapResTail, const_cast<__nv_bfloat16*>(P.res), 3, dims, rstr, box, P.slab_w * 2);
        if (rc) return rc;
      }
      P.res_tma = 1;
    }
  } else {
    rc = make_out_map(&mapOut, P.out, P.n_img, P.ho, P.wo, P.out_cspan ? P.out_cspan : P.cout, P.out_sn, P.out_sh, P.out_sw,
                      P.slab_w, P.tile_w, P.tile_h, P.out_cspan ? P.cout / P.out_cspan : 1, P.out_sp);
    if (rc) return rc;
    mapOutTail = mapOut;
  }
  P.tiles_w = ceil_div(P.wo, P.tile_w);
  P.tiles_h = ceil_div(P.ho, P.tile_h);
  P.n_tiles = ceil_div(P.cout, P.block_n);
  const int m_tiles = P.n_img * P.tiles_h * P.tiles_w;
  P.total_tiles = m_tiles * P.n_tiles;
  P.total_pairs = ceil_div(m_tiles, 2) * P.n_tiles;
  // CTA-pair kernel (cta_group::2): one weight matrix for all images, a B tile whose halves are legal MMA widths, and
  // enough K per tile for the main loop (not the epilogue) to be what is being sped up
  const int k_blocks = P.num_taps * P.kc_per_tap;
  bool two = two_cta_mode() != 0 && P.epi != UAVDET_EPI_HEAD && w_batch == 1 && P.block_n % 32 == 0 && P.block_n >= 64 &&
             P.cout % P.block_n == 0 && m_tiles >= 2 && !(P.epi_mode != 0 && P.block_n == P.slab_w);
  if (two && two_cta_mode() == 1) two = P.block_n >= 128 && k_blocks >= 9;
  if (P.halo) two = false;
  {
    uint64_t dims[3] = {(uint64_t)k_total, (uint64_t)w_rows, (uint64_t)w_batch};
    uint64_t str[2] = {(uint64_t)k_total * 2, (uint64_t)k_total * 2 * w_batch_rows};
    uint32_t box[3] = {(uint32_t)P.block_k, (uint32_t)(two ? P.block_n / 2 : P.block_n), 1u};
    rc = encode_tensor_map(&mapB, const_cast<void*>(w_packed), 3, dims, str, box, P.block_k * 2);
    if (rc) return rc;
  }
  P.fd_n = make_fast_div(P.n_tiles);
  P.fd_w = make_fast_div(P.tiles_w);
  P.fd_h = make_fast_div(P.tiles_h);
  P.w_batch = w_batch;
  // shared memory: [resident weights] [stages x (A [+ B])] [output staging] [barriers]
  const int a_stage = 128 * P.block_k * 2, b_tile = (two ? P.block_n / 2 : P.block_n) * P.block_k * 2;
  const long long b_total = (long long)P.num_taps * P.kc_per_tap * b_tile;
  P.bres_bytes = 0;
  // Per-channel scale / shift of an AFFINE epilogue (folded eval-mode BatchNorm, biases): every epilogue thread needs
  // the 32 values of its chunk, and fetching them from global memory (16 dependent-latency 16-byte loads per chunk in a
  // thread that has no registers to keep them in flight) cost ~800 of the ~1,650 cycles a chunk took in RTMUAVDet's 1x1
  // layers (tools/trace_igemm.py rtm2).  One copy per CTA in shared memory instead; per-sample shifts stay global.
  static const bool no_epc = getenv("UAVDET_IGEMM_NO_EPC") != nullptr;      // A/B switch
  P.epc_floats = (!no_epc && P.epi == UAVDET_EPI_AFFINE && (P.scale || P.shift) && P.shift_sn == 0 && P.cout <= 2048)
                     ? ((P.cout + 31) / 32) * 32 : 0;
  if (P.sample_affine) UAVDET_CHECK_ARG(P.epc_floats > 0, "conv: sample_affine needs the staged epilogue constants (cout <= 2048)");
  // barriers (+ residual-load / halo barriers) + epilogue constants (+ one additive vector per epilogue warp: GroupNorm fold)
  P.sa_staged = (P.sample_affine && P.n_img <= 2048) ? 1 : 0;
  const int ctrl_bytes = kCtrlBytes + (2 + (P.sample_affine ? kEpiWarps : 0)) * 4 * P.epc_floats +
                         (P.sa_staged ? ((8 * P.n_img + 15) & ~15) : 0);
  const int max_smem = 227 * 1024;
  const int staging1 = 2 * 128 * P.slab_w * 2;      // CTA-wide: 2 slabs; warp-private: 8 warps x 1 buffer
  if (P.halo) {
    // [resident weights | B stages] [A halo buffers] [staging] [barriers]
    const long long fixed = ctrl_bytes + staging1 + 2ll * P.halo_buf_bytes;
    if (w_batch == 1 && P.n_tiles == 1 && b_total + fixed <= max_smem && P.total_tiles > 2 * kNumSMs) P.bres_bytes = (int)b_total;
    long long left = max_smem - fixed - P.bres_bytes;
    UAVDET_CHECK_ARG(P.bres_bytes || left >= 2ll * b_tile, "igemm: halo tile does not fit shared memory");
    const long long b_min = P.bres_bytes ? 0 : 4ll * b_tile;        // weight stages kept while the rest is handed out
    P.epi_bufs = 1;
    P.staging_bytes = staging1;
    if (left - b_min >= P.halo_buf_bytes) { P.halo_bufs = 3; left -= P.halo_buf_bytes; }
    if (left - b_min >= staging1) { P.epi_bufs = 2; P.staging_bytes = 2 * staging1; left -= staging1; }
    if (P.halo_bufs == 3 && left - 2 * b_min >= P.halo_buf_bytes) { P.halo_bufs = 4; left -= P.halo_buf_bytes; }
    int stages = 1;
    P.prod_warps = 1;
    if (!P.bres_bytes) {
      stages = (int)(left / b_tile);
      if (stages > kMaxStages) stages = kMaxStages;
      if (stages >= 4) { P.prod_warps = 2; stages &= ~1; }
    }
    P.stages = stages;
  } else {
  if (!two && w_batch == 1 && P.n_tiles == 1 && b_total <= 80 * 1024 && P.total_tiles > 2 * kNumSMs) P.bres_bytes = (int)b_total;
  const int stage_bytes = P.bres_bytes ? a_stage : a_stage + b_tile;
  const int avail = max_smem - P.bres_bytes;
  P.epi_bufs = 1;
  P.staging_bytes = (P.epi == UAVDET_EPI_HEAD) ? 0 : staging1;
  if (P.epi_mode != 0 && (avail - ctrl_bytes - 2 * staging1) / stage_bytes >= 4) {
    P.epi_bufs = 2;
    P.staging_bytes = 2 * staging1;
  }
  int stages = (avail - ctrl_bytes - P.staging_bytes) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  UAVDET_CHECK_ARG(stages >= 2, "igemm: tile does not fit shared memory");
  if (stages >= 6) { stages = (stages / 3) * 3; P.prod_warps = 3; }
  else if (stages >= 4) { P.prod_warps = 2; stages &= ~1; }
  else { P.prod_warps = 1; }
  P.stages = stages;
  }
  P.watchdog = watchdog_word();
  get_trace(&P.trace, &P.trace_tiles);
  // Always request (almost) the whole shared memory so exactly one CTA is resident per SM:
  // each CTA allocates all 512 TMEM columns.
  const int smem_bytes = max_smem;
  if (P.total_tiles <= 0) return UAVDET_OK;
  const int sms = sm_budget();
  int grid = P.total_tiles < sms ? P.total_tiles : sms;
  if (two) {
    const int pairs = sms / 2;           // a CTA pair occupies the two SMs of one TPC
    grid = 2 * (P.total_pairs < pairs ? P.total_pairs : pairs);
  }
  if (!P.res_tma) { mapRes = mapOut; mapResTail = mapOutTail; }
  // res_tma == 2: the epilogue also requests the residual tile of its NEXT tile at the end of the current one
  // (UAVDET_IGEMM_RES_PREFETCH=0 switches that off: A/B)
  static const bool res_prefetch = !(getenv("UAVDET_IGEMM_RES_PREFETCH") && getenv("UAVDET_IGEMM_RES_PREFETCH")[0] == '0');
  if (P.res_tma && res_prefetch) P.res_tma = 2;
  const int kind = (P.epi == UAVDET_EPI_HEAD) ? 3 : (P.epi == UAVDET_EPI_STATS) ? 0 : P.act == UAVDET_ACT_NONE ? 1
                   : P.act == UAVDET_ACT_GELU ? (P.sample_affine ? 6 : 4) : P.act == UAVDET_ACT_SILU ? 5 : 2;
  static PerDeviceOnce attr_once[7][2];   // the dynamic-shared-memory opt-in is per device
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kIgemmThreads);
  cfg.dynamicSmemBytes = (size_t)smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute cattr[2];
  cattr[0].id = cudaLaunchAttributeClusterDimension;
  cattr[0].val.clusterDim.x = 2; cattr[0].val.clusterDim.y = 1; cattr[0].val.clusterDim.z = 1;
  cfg.attrs = two ? cattr : cattr + 1;
  cfg.numAttrs = two ? 1 : 0;
  {
    unsigned n_pdl = 0;
    pdl_attr(cattr + 1, &n_pdl, kPdlIgemm);
    cfg.numAttrs += n_pdl;
  }
#define UAVDET_LAUNCH_IGEMM(K, T)                                                                                      \
  do {                                                                                                                 \
    UAVDET_CUDA(attr_once[K][T].run([] {                                                                               \
      return cudaFuncSetAttribute(igemm_kernel<K, (T) != 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); \
    }));                                                                                                               \
    UAVDET_CUDA(cudaLaunchKernelEx(&cfg, igemm_kernel<K, (T) != 0>, mapA, mapB, mapOut, mapOutTail, mapRes, mapResTail, \
                                   P));                                                                                \
  } while (0)
  if (two) {
    switch (kind) {
      case 0: UAVDET_LAUNCH_IGEMM(0, 1); break;
      case 1: UAVDET_LAUNCH_IGEMM(1, 1); break;
      case 4: UAVDET_LAUNCH_IGEMM(4, 1); break;
      case 5: UAVDET_LAUNCH_IGEMM(5, 1); break;
      case 6: UAVDET_LAUNCH_IGEMM(6, 1); break;
      default: UAVDET_LAUNCH_IGEMM(2, 1); break;
    }
  } else {
    switch (kind) {
      case 0: UAVDET_LAUNCH_IGEMM(0, 0); break;
      case 1: UAVDET_LAUNCH_IGEMM(1, 0); break;
      case 2: UAVDET_LAUNCH_IGEMM(2, 0); break;
      case 4: UAVDET_LAUNCH_IGEMM(4, 0); break;
      case 5: UAVDET_LAUNCH_IGEMM(5, 0); break;
      case 6: UAVDET_LAUNCH_IGEMM(6, 0); break;
      default: UAVDET_LAUNCH_IGEMM(3, 0); break;
    }
  }
#undef UAVDET_LAUNCH_IGEMM
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}

static int fill_epilogue(IgemmParams& P, const uavdet_epilogue* epi, const uavdet_act* y, int cout) {
  P.epi = epi ? epi->epi : UAVDET_EPI_AFFINE;
  P.act = epi ? epi->act : UAVDET_ACT_NONE;
  P.scale = epi ? epi->scale : nullptr;
  P.shift = epi ? epi->shift : nullptr;
  P.shift_sn = (epi && (epi->shift_per_sample & 1)) ? cout : 0;
  P.sample_affine = epi ? epi->sample_affine : nullptr;
  P.res = epi ? (const __nv_bfloat16*)epi->res : nullptr;
  P.sum = epi ? epi->sum : nullptr;
  P.sumsq = epi ? epi->sumsq : nullptr;
  P.head_obj = epi ? epi->head_obj : nullptr;
  P.head_bbox = epi ? epi->head_bbox : nullptr;
  P.head_anchors = epi ? epi->head_anchors : 0;
  if (P.epi == UAVDET_EPI_HEAD) {
    UAVDET_CHECK_ARG(P.head_obj && P.head_bbox && P.head_anchors > 0 && 5 * P.head_anchors <= 16 &&
                         cout == 5 * P.head_anchors,
                     "conv: HEAD epilogue needs cout == 5*A <= 16 and output pointers");
    UAVDET_CHECK_ARG(((uintptr_t)P.head_bbox & 15) == 0, "conv: head_bbox must be 16-byte aligned");
  } else {
    UAVDET_CHECK_ARG(y && y->ptr, "conv: output view missing");
    UAVDET_CHECK_ARG(cout % 32 == 0, "conv: cout=%d must be a multiple of 32", cout);
    UAVDET_CHECK_ARG(y->ld % 8 == 0 && ((uintptr_t)y->ptr & 15) == 0, "conv: output must be 16-byte aligned");
    if (P.epi == UAVDET_EPI_STATS) UAVDET_CHECK_ARG(P.sum && P.sumsq, "conv: STATS epilogue needs sum/sumsq");
    if (P.res) UAVDET_CHECK_ARG(epi->res_ld % 8 == 0 && ((uintptr_t)P.res & 15) == 0, "conv: residual alignment");
    if (P.sample_affine)
      UAVDET_CHECK_ARG(P.epi == UAVDET_EPI_AFFINE && P.scale && P.shift && P.shift_sn == 0 &&
                           (P.act == UAVDET_ACT_RELU || P.act == UAVDET_ACT_GELU) && ((uintptr_t)P.sample_affine & 7) == 0,
                       "conv: sample_affine needs an AFFINE epilogue with shared scale and shift vectors and ReLU / GELU");
  }
  return UAVDET_OK;
}

}  // namespace uavdet

using namespace uavdet;

// debug hook (not part of the public header): device buffer receiving 8 clock64 stamps per tile of CTA 0
extern "C" void uavdet_debug_set_trace(long long* dev_buf, int tiles) { g_trace = dev_buf; g_trace_tiles = tiles; }

extern "C" int uavdet_conv_fwd(const uavdet_act* x, const void* w_packed, int w_batch, int cout, int k,
                               int stride, int pad, int s2d, const uavdet_act* y,
                               const uavdet_epilogue* epi, void* stream) {
  UAVDET_CHECK_ARG(x && x->ptr && w_packed, "conv_fwd: null input");
  UAVDET_CHECK_ARG(k >= 1 && k <= 5 && (stride == 1 || stride == 2), "conv_fwd: k=%d stride=%d unsupported", k, stride);
  UAVDET_CHECK_ARG(!(s2d && stride != 1), "conv_fwd: s2d implies stride 1 on the gathered map");
  UAVDET_CHECK_ARG(x->ld % 8 == 0 && ((uintptr_t)x->ptr & 15) == 0, "conv_fwd: input must be 16-byte aligned");
  UAVDET_CHECK_ARG(w_batch == 1 || w_batch == x->n, "conv_fwd: w_batch must be 1 or n");
  const int parity = (stride == 2 || s2d) ? 1 : 0;
  if (parity) UAVDET_CHECK_ARG(x->h % 2 == 0 && x->w % 2 == 0, "conv_fwd: stride-2/s2d needs even H,W (got %dx%d)", x->h, x->w);
  const int c_blk = x->c;                   // channels per tap block
  const int cin = s2d ? 4 * x->c : x->c;
  UAVDET_CHECK_ARG(c_blk % 32 == 0, "conv_fwd: cin=%d must be a multiple of 32 (use the stem kernel)", c_blk);
  const int hin = s2d ? x->h / 2 : x->h, win = s2d ? x->w / 2 : x->w;
  IgemmParams P{};
  P.n_img = x->n;
  P.ho = (hin + 2 * pad - k) / stride + 1;
  P.wo = (win + 2 * pad - k) / stride + 1;
  P.cout = cout;
  P.block_k = (c_blk % 64 == 0) ? 64 : 32;
  P.block_n = pick_block_n(cout);
  P.kc_per_tap = c_blk / P.block_k;
  // 32-channel inputs read through the parity view: the two pixels of a pair are one contiguous 128-byte row, and the
  // filter taps that read them are neighbours in the packed weight row — one K = 64 k-block (SWIZZLE_128B) replaces
  // two K = 32 ones (SWIZZLE_64B rows cost the tensor core twice the cycles per instruction, see halo_mode()).
  //   s2d:      blocks (i, j = 0) and (i, j = 1) of a tap;
  //   stride 2: (kh, kw = 1 | 2) = [even | odd] pixel of column ow; (kh, kw = 0) = the odd pixel of column ow - 1, read
  //             as channels 32..95 of the pair row — TMA zero-fills 64..95, which meets the weights of (kh, 1).
  static const bool no_pair = getenv("UAVDET_IGEMM_NOPAIR") != nullptr;
  const bool pair = !no_pair && parity && c_blk == 32 && x->ld == 32 && (s2d || (k == 3 && pad == 1));
  if (pair) { P.block_k = 64; P.kc_per_tap = 1; }
  int nt = 0;
  for (int kh = 0; kh < k; ++kh)
    for (int kw = 0; kw < k; ++kw) {
      if (pair && s2d) {
        for (int i = 0; i < 2; ++i) {
          UAVDET_CHECK_ARG(nt < kMaxTaps, "conv_fwd: too many taps");
          P.taps[nt++] = ConvTap{0, kw - pad, i, kh - pad, ((kh * k + kw) * 4 + i * 2) * c_blk};
        }
      } else if (pair) {
        if (kw == 2) continue;                         // merged into the kw = 1 block
        const int th = kh - pad;
        const int ph = ((th % 2) + 2) % 2;
        P.taps[nt++] = ConvTap{kw == 0 ? 32 : 0, kw == 0 ? -1 : 0, ph, (th - ph) / 2, (kh * k + kw) * cin};
      } else if (s2d) {
        for (int i = 0; i < 2; ++i)
          for (int j = 0; j < 2; ++j) {
            UAVDET_CHECK_ARG(nt < kMaxTaps, "conv_fwd: too many taps");
            P.taps[nt++] = ConvTap{j * x->ld, kw - pad, i, kh - pad, ((kh * k + kw) * 4 + (i * 2 + j)) * c_blk};
          }
      } else if (stride == 1) {
        UAVDET_CHECK_ARG(nt < kMaxTaps, "conv_fwd: too many taps");
        P.taps[nt++] = ConvTap{0, kw - pad, 0, kh - pad, (kh * k + kw) * cin};
      } else {
        const int th = kh - pad, tw = kw - pad;
        const int ph = ((th % 2) + 2) % 2, pw = ((tw % 2) + 2) % 2;
        UAVDET_CHECK_ARG(nt < kMaxTaps, "conv_fwd: too many taps");
        P.taps[nt++] = ConvTap{pw * x->ld, (tw - pw) / 2, ph, (th - ph) / 2, (kh * k + kw) * cin};
      }
    }
  P.num_taps = nt;
  int rc = fill_epilogue(P, epi, y, cout);
  if (rc) return rc;
  if (P.epi != UAVDET_EPI_HEAD) {
    UAVDET_CHECK_ARG(y->n == x->n && y->h == P.ho && y->w == P.wo && y->c == cout,
                     "conv_fwd: output view (%d,%d,%d,%d) != expected (%d,%d,%d,%d)", y->n, y->h, y->w, y->c,
                     x->n, P.ho, P.wo, cout);
    P.out = (__nv_bfloat16*)y->ptr;
    P.out_sw = y->ld; P.out_sh = (long long)y->w * y->ld; P.out_sn = (long long)y->h * y->w * y->ld;
    if (P.res) { P.res_sw = epi->res_ld; P.res_sh = (long long)y->w * epi->res_ld; P.res_sn = (long long)y->h * y->w * epi->res_ld; }
  }
  choose_tile(P.ho, P.wo, true, &P.tile_w, &P.tile_h, &P.epi_mode);
  const int w_rows = (P.epi == UAVDET_EPI_HEAD) ? 16 : cout;
  return launch_igemm(x, parity, w_packed, w_rows, k * k * cin, w_batch, P, (cudaStream_t)stream);
}

extern "C" int uavdet_conv_dgrad(const uavdet_act* dy, const void* w_packed_t, int w_batch, int cin, int k,
                                 int stride, int pad, const uavdet_act* dx, const uavdet_epilogue* epi,
                                 void* stream) {
  UAVDET_CHECK_ARG(dy && dy->ptr && w_packed_t && dx && dx->ptr, "conv_dgrad: null input");
  UAVDET_CHECK_ARG(k >= 1 && k <= 5 && (stride == 1 || stride == 2), "conv_dgrad: k=%d stride=%d unsupported", k, stride);
  UAVDET_CHECK_ARG(dy->ld % 8 == 0 && ((uintptr_t)dy->ptr & 15) == 0, "conv_dgrad: dy must be 16-byte aligned");
  UAVDET_CHECK_ARG(w_batch == 1 || w_batch == dy->n, "conv_dgrad: w_batch must be 1 or n");
  const int cout = dy->c;
  UAVDET_CHECK_ARG(cout % 32 == 0 && cin % 32 == 0, "conv_dgrad: channels must be multiples of 32");
  UAVDET_CHECK_ARG(dx->n == dy->n && dx->c == cin, "conv_dgrad: dx view mismatch");
  UAVDET_CHECK_ARG((dx->h + 2 * pad - k) / stride + 1 == dy->h && (dx->w + 2 * pad - k) / stride + 1 == dy->w,
                   "conv_dgrad: spatial sizes inconsistent");
  if (stride == 2) UAVDET_CHECK_ARG(dx->h % 2 == 0 && dx->w % 2 == 0, "conv_dgrad: stride 2 needs even H,W");
  cudaStream_t st = (cudaStream_t)stream;
  const int planes = stride;  // parity planes per axis
  for (int ph = 0; ph < planes; ++ph)
    for (int pw = 0; pw < planes; ++pw) {
      IgemmParams P{};
      P.n_img = dy->n;
      P.ho = dx->h / stride;
      P.wo = dx->w / stride;
      P.cout = cin;  // GEMM N = input channels of the forward conv
      P.block_k = (cout % 64 == 0) ? 64 : 32;
      P.block_n = pick_block_n(cin);
      P.kc_per_tap = cout / P.block_k;
      int nt = 0;
      // dx[2i+ph] gathers dy[oh] for every kh with (ph + pad - kh) % stride == 0, oh = i + (ph+pad-kh)/stride
      for (int kh = 0; kh < k; ++kh) {
        const int th = ph + pad - kh;
        if (((th % stride) + stride) % stride != 0) continue;
        for (int kw = 0; kw < k; ++kw) {
          const int tw = pw + pad - kw;
          if (((tw % stride) + stride) % stride != 0) continue;
          UAVDET_CHECK_ARG(nt < kMaxTaps, "conv_dgrad: too many taps");
          P.taps[nt++] = ConvTap{0, tw / stride, 0, th / stride, (kh * k + kw) * cout};
        }
      }
      P.num_taps = nt;
      uavdet_act dxv = *dx;
      dxv.n = dx->n; dxv.h = P.ho; dxv.w = P.wo;
      int rc = fill_epilogue(P, epi, &dxv, cin);
      if (rc) return rc;
      UAVDET_CHECK_ARG(P.epi == UAVDET_EPI_AFFINE, "conv_dgrad: only the AFFINE epilogue is supported");
      const long long ld = dx->ld;
      P.out = (__nv_bfloat16*)dx->ptr + ((long long)ph * dx->w + pw) * ld;
      P.out_sw = stride * ld; P.out_sh = (long long)stride * dx->w * ld; P.out_sn = (long long)dx->h * dx->w * ld;
      if (P.res) {
        const long long rl = epi->res_ld;
        P.res = P.res + ((long long)ph * dx->w + pw) * rl;
        P.res_sw = stride * rl; P.res_sh = (long long)stride * dx->w * rl; P.res_sn = (long long)dx->h * dx->w * rl;
      }
      choose_tile(P.ho, P.wo, stride == 1, &P.tile_w, &P.tile_h, &P.epi_mode);
      if (nt == 0) {
        // no filter tap reaches this parity plane (1x1 stride-2): the gradient is the residual or zero
        rc = fill_plane(P, st);
      } else {
        rc = launch_igemm(dy, 0, w_packed_t, cin, k * k * cout, w_batch, P, st);
      }
      if (rc) return rc;
    }
  return UAVDET_OK;
}

// Weights of the plane-fused stride-2 data gradient: from the transposed pack wt [cin][3][3][cout] to
// wf [(ph, pw, ci)][(sh, sw, co)] = wt[ci][ph + 1 - 2 sh][pw + 1 - 2 sw][co], zero where the filter index leaves 0..2.
__global__ void __launch_bounds__(256)
pack_dgrad_s2_fused_kernel(const __nv_bfloat16* __restrict__ wt, int cin, int cout, __nv_bfloat16* __restrict__ wf) {
  const int c8 = cout >> 3;
  const long long total = 16ll * cin * c8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % c8) * 8;
    long long r = i / c8;
    const int s = (int)(r % 4); r /= 4;          // (sh, sw)
    const int ci = (int)(r % cin);
    const int q = (int)(r / cin);                // (ph, pw)
    const int kh = (q >> 1) + 1 - 2 * (s >> 1), kw = (q & 1) + 1 - 2 * (s & 1);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (kh >= 0 && kh <= 2 && kw >= 0 && kw <= 2)
      v = *reinterpret_cast<const uint4*>(wt + ((long long)(ci * 3 + kh) * 3 + kw) * cout + co);
    *reinterpret_cast<uint4*>(wf + ((long long)(q * cin + ci) * 4 + s) * cout + co) = v;
  }
}

extern "C" int uavdet_pack_dgrad_s2_fused(const void* w_packed_t, int cin, int cout, void* w_fused, void* stream) {
  UAVDET_CHECK_ARG(w_packed_t && w_fused && cin % 32 == 0 && cout % 32 == 0, "pack_dgrad_s2_fused: bad arguments");
  UAVDET_CHECK_ARG((((uintptr_t)w_packed_t | (uintptr_t)w_fused) & 15) == 0, "pack_dgrad_s2_fused: 16-byte alignment");
  const long long total = 16ll * cin * (cout / 8);
  int blocks = (int)((total + 255) / 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  pack_dgrad_s2_fused_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)w_packed_t, cin, cout,
                                                                        (__nv_bfloat16*)w_fused);
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}

extern "C" int uavdet_conv_dgrad_s2_fused(const uavdet_act* dy, const void* w_fused, int cin, const uavdet_act* dx,
                                          const uavdet_epilogue* epi, void* stream) {
  UAVDET_CHECK_ARG(dy && dy->ptr && w_fused && dx && dx->ptr, "conv_dgrad_s2_fused: null input");
  UAVDET_CHECK_ARG(dy->ld % 8 == 0 && ((uintptr_t)dy->ptr & 15) == 0, "conv_dgrad_s2_fused: dy must be 16-byte aligned");
  const int cout = dy->c;
  UAVDET_CHECK_ARG(cout % 32 == 0 && cin % 32 == 0 && 4 * cin <= 256, "conv_dgrad_s2_fused: needs cin in {32, 64}, cout %% 32 == 0");
  UAVDET_CHECK_ARG(dx->n == dy->n && dx->c == cin && dx->ld == cin && dx->h == 2 * dy->h && dx->w == 2 * dy->w,
                   "conv_dgrad_s2_fused: dx must be the dense (n, 2*ho, 2*wo, cin) tensor of a 3x3 stride-2 pad-1 conv");
  UAVDET_CHECK_ARG(!epi || !epi->res || epi->res_ld == cin, "conv_dgrad_s2_fused: the residual must be dense");
  IgemmParams P{};
  P.n_img = dy->n;
  P.ho = dy->h;
  P.wo = dy->w;
  P.cout = 4 * cin;
  P.block_k = (cout % 64 == 0) ? 64 : 32;
  P.block_n = pick_block_n(4 * cin);
  P.kc_per_tap = cout / P.block_k;
  int nt = 0;
  for (int sh = 0; sh < 2; ++sh)
    for (int sw = 0; sw < 2; ++sw) P.taps[nt++] = ConvTap{0, sw, 0, sh, (sh * 2 + sw) * cout};
  P.num_taps = nt;
  uavdet_act dxv = *dx;
  dxv.h = P.ho; dxv.w = P.wo;
  int rc = fill_epilogue(P, epi, &dxv, 4 * cin);
  if (rc) return rc;
  UAVDET_CHECK_ARG(P.epi == UAVDET_EPI_AFFINE && !P.scale && !P.shift, "conv_dgrad_s2_fused: plain AFFINE epilogue (residual only)");
  const long long ld = dx->ld;
  P.out = (__nv_bfloat16*)dx->ptr;
  P.out_sw = 2 * ld; P.out_sh = 2ll * dx->w * ld; P.out_sn = (long long)dx->h * dx->w * ld;
  P.out_cspan = 2 * cin; P.out_sp = (long long)dx->w * ld;
  if (P.res) {
    P.res_sw = 2 * ld; P.res_sh = 2ll * dx->w * ld; P.res_sn = (long long)dx->h * dx->w * ld;
    P.res_sp = (long long)dx->w * ld;
  }
  choose_tile(P.ho, P.wo, false, &P.tile_w, &P.tile_h, &P.epi_mode);
  return launch_igemm(dy, 0, w_fused, 4 * cin, 4 * cout, 1, P, (cudaStream_t)stream);
}

// 3x3 stride-1 pad-1 convolution with few output channels, two adjacent output pixels per GEMM row: N = 2 * cout columns
// [column parity][cout] over K = 3 x 4 input-column shifts x cin (the tensor core needs as long for an N = 64 instruction
// as for N = 128, so a 64-channel layer runs at half rate the plain way; here it issues 12 taps per pixel PAIR instead of
// 2 x 9).  The input is read as pixel pairs ([2 pixels][ld] rows of a 5-D map, like the stride-2 parity view but in the
// column direction only): shift dx in {-1, 0, 1, 2} of pair j is pixel 2j + dx = pair j + floor(dx / 2), channel
// offset (dx & 1) * ld; TMA zero-fills the pairs outside the image (the padding).  The output map puts the column
// parity into its plane dimension (stride ld), so channel slices of a concat buffer work.
extern "C" int uavdet_conv3x3_pair_fwd(const uavdet_act* x, const void* w_pair, int cout, const uavdet_act* y,
                                       const uavdet_epilogue* epi, void* stream) {
  UAVDET_CHECK_ARG(x && x->ptr && w_pair && y && y->ptr, "conv3x3_pair_fwd: null input");
  UAVDET_CHECK_ARG(x->ld % 8 == 0 && ((uintptr_t)x->ptr & 15) == 0, "conv3x3_pair_fwd: input must be 16-byte aligned");
  const bool dense_out = y->ld == cout;       // the two pixels of an output pair are one contiguous row of 2 * cout channels
  UAVDET_CHECK_ARG((x->c % 64 == 0 || (x->c == 32 && x->ld == 32)) && (cout % 64 == 0 || (cout == 32 && dense_out)) &&
                       2 * cout <= 256 && x->w % 2 == 0,
                   "conv3x3_pair_fwd: needs cin %% 64 == 0 (or a dense 32-channel input), cout in {64, 128} (or 32 into a "
                   "dense tensor), even width (cin=%d ld=%d cout=%d w=%d)", x->c, x->ld, cout, x->w);
  UAVDET_CHECK_ARG(y->n == x->n && y->h == x->h && y->w == x->w && y->c == cout, "conv3x3_pair_fwd: output view mismatch");
  const int C = x->c;
  IgemmParams P{};
  P.n_img = x->n;
  P.ho = x->h;
  P.wo = x->w / 2;
  P.cout = 2 * cout;
  P.block_k = 64;
  P.block_n = pick_block_n(2 * cout);
  int nt = 0;
  if (C == 32) {
    // a dense 32-channel input: a pixel pair is ONE 128-byte row, so a k-block is a whole pair (K = 64, SWIZZLE_128B; 64-byte
    // rows cost the tensor core twice the cycles): pairs j - 1, j, j + 1 = column shifts -2 .. 3, the outer two with zero weights
    P.kc_per_tap = 1;
    for (int dy = -1; dy <= 1; ++dy)
      for (int b = -1; b <= 1; ++b) P.taps[nt++] = ConvTap{0, b, 0, dy, ((dy + 1) * 3 + (b + 1)) * 64};
  } else {
    P.kc_per_tap = C / 64;
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 2; ++dx)
        P.taps[nt++] = ConvTap{(dx & 1) ? x->ld : 0, dx < 0 ? -1 : dx / 2, 0, dy, ((dy + 1) * 4 + (dx + 1)) * C};
  }
  P.num_taps = nt;
  uavdet_act yv = *y;
  yv.w = P.wo;
  int rc = fill_epilogue(P, epi, &yv, 2 * cout);
  if (rc) return rc;
  UAVDET_CHECK_ARG((P.epi == UAVDET_EPI_AFFINE || (P.epi == UAVDET_EPI_STATS && !P.shift)) && !P.res && P.shift_sn == 0 &&
                       !P.sample_affine,
                   "conv3x3_pair_fwd: AFFINE epilogue with shared [2*cout] scale / shift vectors, or STATS; no residual");
  const long long ld = y->ld;
  P.out = (__nv_bfloat16*)y->ptr;
  P.out_sw = 2 * ld; P.out_sh = (long long)y->w * ld; P.out_sn = (long long)y->h * y->w * ld;
  if (!dense_out) { P.out_cspan = cout; P.out_sp = ld; }      // a channel slice: the column parity is a plane of the output map
  P.stat_mod = cout;                                          // STATS: both pixels of a pair add to the same channel
  choose_tile(P.ho, P.wo, dense_out, &P.tile_w, &P.tile_h, &P.epi_mode);
  uavdet_act xp = *x;                 // the pair view: rows of [2 pixels][ld], the second pixel's channels at offset ld
  xp.w = x->w / 2;
  xp.c = x->ld + C;
  xp.ld = 2 * x->ld;
  return launch_igemm(&xp, 0, w_pair, 2 * cout, C == 32 ? 9 * 64 : 12 * C, 1, P, (cudaStream_t)stream);
}

extern "C" int uavdet_conv_dgrad_s2d(const uavdet_act* dy, const void* w_packed_t, int w_batch, int c, int k, int pad,
                                     const uavdet_act* dx, const uavdet_epilogue* epi, void* stream) {
  UAVDET_CHECK_ARG(dy && dy->ptr && w_packed_t && dx && dx->ptr, "conv_dgrad_s2d: null input");
  UAVDET_CHECK_ARG(k >= 1 && k <= 5, "conv_dgrad_s2d: k=%d unsupported", k);
  UAVDET_CHECK_ARG(dy->ld % 8 == 0 && ((uintptr_t)dy->ptr & 15) == 0, "conv_dgrad_s2d: dy must be 16-byte aligned");
  UAVDET_CHECK_ARG(w_batch == 1 || w_batch == dy->n, "conv_dgrad_s2d: w_batch must be 1 or n");
  const int cout = dy->c;
  UAVDET_CHECK_ARG(cout % 32 == 0 && c % 32 == 0, "conv_dgrad_s2d: channels must be multiples of 32");
  UAVDET_CHECK_ARG(dx->n == dy->n && dx->c == c && dx->h % 2 == 0 && dx->w % 2 == 0, "conv_dgrad_s2d: dx view mismatch");
  UAVDET_CHECK_ARG((dx->h / 2 + 2 * pad - k) + 1 == dy->h && (dx->w / 2 + 2 * pad - k) + 1 == dy->w,
                   "conv_dgrad_s2d: spatial sizes inconsistent");
  cudaStream_t st = (cudaStream_t)stream;
  const long long k_total = (long long)k * k * cout;
  // All four parity planes read the same dy taps and differ only in their weight rows (q*c .. q*c+c of the [4c][K]
  // matrix) and in where they store: ONE implicit GEMM with N = 4c whose output map takes the row parity as its third
  // dimension and the column parity as the upper half of a 2c-channel pixel-pair row.  A quarter of the MMAs (the
  // tensor core needs as long for an N = 32 instruction as for an N = 128 one) and dy is read once instead of four times.
  static const bool no_fuse = getenv("UAVDET_IGEMM_NOFUSE_PLANES") != nullptr;
  if (!no_fuse && dx->ld == c && (!epi || !epi->res || epi->res_ld == c)) {
    IgemmParams P{};
    P.n_img = dy->n;
    P.ho = dx->h / 2;
    P.wo = dx->w / 2;
    P.cout = 4 * c;
    P.block_k = (cout % 64 == 0) ? 64 : 32;
    P.block_n = pick_block_n(4 * c);
    P.kc_per_tap = cout / P.block_k;
    int nt = 0;
    for (int kh = 0; kh < k; ++kh)
      for (int kw = 0; kw < k; ++kw) {
        UAVDET_CHECK_ARG(nt < kMaxTaps, "conv_dgrad_s2d: too many taps");
        P.taps[nt++] = ConvTap{0, pad - kw, 0, pad - kh, (kh * k + kw) * cout};
      }
    P.num_taps = nt;
    uavdet_act dxv = *dx;
    dxv.h = P.ho; dxv.w = P.wo;
    int rc = fill_epilogue(P, epi, &dxv, 4 * c);
    if (rc) return rc;
    UAVDET_CHECK_ARG(P.epi == UAVDET_EPI_AFFINE, "conv_dgrad_s2d: only the AFFINE epilogue is supported");
    const long long ld = dx->ld;
    P.out = (__nv_bfloat16*)dx->ptr;
    P.out_sw = 2 * ld; P.out_sh = 2ll * dx->w * ld; P.out_sn = (long long)dx->h * dx->w * ld;
    P.out_cspan = 2 * c; P.out_sp = (long long)dx->w * ld;
    if (P.res) {
      const long long rl = epi->res_ld;
      P.res_sw = 2 * rl; P.res_sh = 2ll * dx->w * rl; P.res_sn = (long long)dx->h * dx->w * rl;
      P.res_sp = (long long)dx->w * rl;
    }
    if (P.shift && (epi->shift_per_sample & 1)) P.shift_sn = 4 * c;
    choose_tile(P.ho, P.wo, false, &P.tile_w, &P.tile_h, &P.epi_mode);
    // a store slab (64 | 32 columns) must not straddle two row-parity planes; a shared (not per-sample) shift is [c]
    const bool slab_ok = (2 * c) % ((P.block_n % 64 == 0) ? 64 : 32) == 0;
    const bool shift_ok = !P.shift || (epi->shift_per_sample & 1);
    if (slab_ok && shift_ok)
      return launch_igemm(dy, 0, w_packed_t, 4 * c, (int)k_total, w_batch, P, st, 4 * c);
  }
  for (int q = 0; q < 4; ++q) {
    const int pi = q >> 1, pj = q & 1;   // row / column parity of this channel block (DySOEM_SimFPN.py:71-73)
    IgemmParams P{};
    P.n_img = dy->n;
    P.ho = dx->h / 2;
    P.wo = dx->w / 2;
    P.cout = c;
    P.block_k = (cout % 64 == 0) ? 64 : 32;
    P.block_n = pick_block_n(c);
    P.kc_per_tap = cout / P.block_k;
    int nt = 0;
    for (int kh = 0; kh < k; ++kh)
      for (int kw = 0; kw < k; ++kw) {
        UAVDET_CHECK_ARG(nt < kMaxTaps, "conv_dgrad_s2d: too many taps");
        P.taps[nt++] = ConvTap{0, pad - kw, 0, pad - kh, (kh * k + kw) * cout};
      }
    P.num_taps = nt;
    uavdet_act dxv = *dx;
    dxv.h = P.ho; dxv.w = P.wo;
    int rc = fill_epilogue(P, epi, &dxv, c);
    if (rc) return rc;
    UAVDET_CHECK_ARG(P.epi == UAVDET_EPI_AFFINE, "conv_dgrad_s2d: only the AFFINE epilogue is supported");
    const long long ld = dx->ld;
    P.out = (__nv_bfloat16*)dx->ptr + ((long long)pi * dx->w + pj) * ld;
    P.out_sw = 2 * ld; P.out_sh = 2ll * dx->w * ld; P.out_sn = (long long)dx->h * dx->w * ld;
    if (P.res) {
      const long long rl = epi->res_ld;
      P.res = P.res + ((long long)pi * dx->w + pj) * rl;
      P.res_sw = 2 * rl; P.res_sh = 2ll * dx->w * rl; P.res_sn = (long long)dx->h * dx->w * rl;
    }
    if (P.shift && (epi->shift_per_sample & 1)) { P.shift = P.shift + q * c; P.shift_sn = 4 * c; }
    choose_tile(P.ho, P.wo, false, &P.tile_w, &P.tile_h, &P.epi_mode);
    const __nv_bfloat16* wq = (const __nv_bfloat16*)w_packed_t + (long long)q * c * k_total;
    rc = launch_igemm(dy, 0, wq, c, (int)k_total, w_batch, P, st, 4 * c);
    if (rc) return rc;
  }
  return UAVDET_OK;
}
