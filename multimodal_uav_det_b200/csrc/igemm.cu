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
//   * warp roles: warp0 = TMA producer, warp1 = MMA issuer (+TMEM alloc), warps 2-5 =
//     epilogue (tcgen05.ld -> BN-stat partial sums / affine+activation+residual -> bf16).
#include "common.cuh"
#include "sm100.cuh"
#include "igemm.h"

namespace uavdet {
using namespace sm100;

constexpr int kIgemmThreads = 192;
constexpr int kAccStride = 256;  // TMEM columns per accumulator buffer

template <int ACT>
__device__ __forceinline__ void affine_act_store(const uint32_t (&r)[32], const IgemmParams& P, int cg,
                                                 __nv_bfloat16* out_px, const __nv_bfloat16* res_px,
                                                 const float* shift) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
  if (P.scale) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      float4 s = __ldg(reinterpret_cast<const float4*>(P.scale + cg + i));
      v[i] *= s.x; v[i + 1] *= s.y; v[i + 2] *= s.z; v[i + 3] *= s.w;
    }
  }
  if (shift) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      float4 s = __ldg(reinterpret_cast<const float4*>(shift + cg + i));
      v[i] += s.x; v[i + 1] += s.y; v[i + 2] += s.z; v[i + 3] += s.w;
    }
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = act_fwd<ACT>(v[i]);
  if (res_px) {
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
      uint4 rr = __ldg(reinterpret_cast<const uint4*>(res_px + cg + i));
      v[i] += bf16_lo(rr.x); v[i + 1] += bf16_hi(rr.x);
      v[i + 2] += bf16_lo(rr.y); v[i + 3] += bf16_hi(rr.y);
      v[i + 4] += bf16_lo(rr.z); v[i + 5] += bf16_hi(rr.z);
      v[i + 6] += bf16_lo(rr.w); v[i + 7] += bf16_hi(rr.w);
    }
  }
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    uint4 o;
    o.x = pack_bf16x2(v[i], v[i + 1]);
    o.y = pack_bf16x2(v[i + 2], v[i + 3]);
    o.z = pack_bf16x2(v[i + 4], v[i + 5]);
    o.w = pack_bf16x2(v[i + 6], v[i + 7]);
    *reinterpret_cast<uint4*>(out_px + cg + i) = o;
  }
}

// Sum v[c] over the 32 lanes of the warp for 32 columns; lane l returns column l's total.
__device__ __forceinline__ float warp_column_sums(float (&v)[32], int lane) {
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

__global__ void __launch_bounds__(kIgemmThreads, 1)
igemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
             const __grid_constant__ IgemmParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int a_bytes = 128 * P.block_k * 2;             // smem reserved for A per stage
  const int b_bytes = P.block_n * P.block_k * 2;
  const int stage_bytes = a_bytes + b_bytes;
  const int a_tx = P.tile_w * P.tile_h * P.block_k * 2;  // bytes the A box actually delivers
  uint8_t* ctrl = smem + (size_t)P.stages * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ctrl);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tfull_bar = empty_bar + kMaxStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  volatile uint32_t* dead = tmem_ptr + 1;
  float* stat = reinterpret_cast<float*>(tmem_ptr + 4);  // [2 acc][2 (sum,sumsq)][256]

  if (threadIdx.x == 0) {
    for (int s = 0; s < P.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tfull_bar[a]), 1);
      mbar_init(smem_u32(&tempty_bar[a]), 4);
    }
    *dead = 0;
    fence_barrier_init();
    prefetch_tensormap(&mapA);
    prefetch_tensormap(&mapB);
  }
  for (int i = threadIdx.x; i < 2 * 2 * 256; i += kIgemmThreads) stat[i] = 0.f;
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_ptr), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int num_kb = P.num_taps * P.kc_per_tap;

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
        const int n_tile = tile % P.n_tiles;
        int m_tile = tile / P.n_tiles;
        const int tw = m_tile % P.tiles_w; m_tile /= P.tiles_w;
        const int th = m_tile % P.tiles_h;
        const int img = m_tile / P.tiles_h;
        const int ow0 = tw * P.tile_w, oh0 = th * P.tile_h, n0 = n_tile * P.block_n;
        for (int t = 0; t < P.num_taps; ++t) {
          const ConvTap tap = P.taps[t];
          for (int kc = 0; kc < P.kc_per_tap; ++kc) {
            mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u, dead, P.watchdog, 0x1u);
            const uint32_t fb = smem_u32(&full_bar[stage]);
            mbar_arrive_expect_tx(fb, (uint32_t)(a_tx + b_bytes));
            uint8_t* sa = smem + (size_t)stage * stage_bytes;
            tma_load_5d(smem_u32(sa), &mapA, fb, tap.c_off + kc * P.block_k, ow0 + tap.dw, tap.p,
                        oh0 + tap.dh, img);
            tma_load_3d(smem_u32(sa + a_bytes), &mapB, fb, tap.w_koff + kc * P.block_k, n0,
                        P.w_batch > 1 ? img : 0);
            if (++stage == P.stages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ==============================
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16(P.block_n, 0, 0);
      const uint32_t layout = (P.block_k == 64) ? 2u : 4u;      // SWIZZLE_128B : SWIZZLE_64B
      const uint32_t sbo = 8u * (uint32_t)P.block_k * 2u;        // 8 rows of one swizzle atom
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
        mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1u, dead, P.watchdog, 0x2u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kAccStride);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(smem_u32(&full_bar[stage]), phase, dead, P.watchdog, 0x4u);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint64_t a_desc = make_smem_desc(sa, 16, sbo, layout);
          const uint64_t b_desc = make_smem_desc(sa + a_bytes, 16, sbo, layout);
          const int ksteps = P.block_k / 16;
          for (int k = 0; k < ksteps; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the swizzle row: +2 in the >>4 field
            tc_mma_bf16(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc,
                        (kb | k) != 0 ? 1u : 0u);
          }
          tc_commit(smem_u32(&empty_bar[stage]));
          if (++stage == P.stages) { stage = 0; phase ^= 1u; }
        }
        tc_commit(smem_u32(&tfull_bar[acc]));
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ============================ epilogue (warps 2..5) ====================
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;
    const int hl = row / P.tile_w, wl = row - hl * P.tile_w;
    const int e_tid = (warp - 2) * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
      const int n_tile = tile % P.n_tiles;
      int m_tile = tile / P.n_tiles;
      const int tw = m_tile % P.tiles_w; m_tile /= P.tiles_w;
      const int th = m_tile % P.tiles_h;
      const int img = m_tile / P.tiles_h;
      const int oh = th * P.tile_h + hl, ow = tw * P.tile_w + wl;
      const int n0 = n_tile * P.block_n;
      const bool valid = (row < P.tile_w * P.tile_h) && (oh < P.ho) && (ow < P.wo);

      mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase, dead, P.watchdog, 0x8u);
      tc_fence_after();
      const uint32_t tbase = tmem_base + (uint32_t)(acc * kAccStride) + ((uint32_t)(q * 32) << 16);

      if (P.epi == UAVDET_EPI_HEAD) {
        uint32_t r[16];
        tmem_ld_32x16(tbase, r);
        tmem_ld_wait();
        if (valid) {
          const int A = P.head_anchors;
          const size_t hw = (size_t)P.ho * P.wo;
          const size_t px = (size_t)oh * P.wo + ow;
          for (int a = 0; a < A; ++a) {
            float o = __uint_as_float(r[a]) + (P.shift ? __ldg(P.shift + a) : 0.f);
            P.head_obj[((size_t)img * A + a) * hw + px] = o;
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
            reinterpret_cast<float4*>(P.head_bbox)[((size_t)img * A + a) * hw + px] = b;
          }
        }
      } else {
        __nv_bfloat16* out_px =
            P.out + (size_t)img * P.out_sn + (size_t)oh * P.out_sh + (size_t)ow * P.out_sw;
        const __nv_bfloat16* res_px =
            P.res ? P.res + (size_t)img * P.res_sn + (size_t)oh * P.res_sh + (size_t)ow * P.res_sw
                  : nullptr;
        const float* shift = P.shift ? P.shift + (size_t)img * P.shift_sn : nullptr;
        for (int c0 = 0; c0 < P.block_n; c0 += 32) {
          const int cg = n0 + c0;
          if (cg >= P.cout) break;  // uniform: last n-tile of a cout that is not a block_n multiple
          uint32_t r[32];
          tmem_ld_32x32(tbase + (uint32_t)c0, r);
          tmem_ld_wait();
          if (P.epi == UAVDET_EPI_STATS) {
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = valid ? __uint_as_float(r[i]) : 0.f;
            if (shift && valid) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                float4 s = __ldg(reinterpret_cast<const float4*>(shift + cg + i));
                v[i] += s.x; v[i + 1] += s.y; v[i + 2] += s.z; v[i + 3] += s.w;
              }
            }
            if (valid) {
#pragma unroll
              for (int i = 0; i < 32; i += 8) {
                uint4 o;
                o.x = pack_bf16x2(v[i], v[i + 1]);
                o.y = pack_bf16x2(v[i + 2], v[i + 3]);
                o.z = pack_bf16x2(v[i + 4], v[i + 5]);
                o.w = pack_bf16x2(v[i + 6], v[i + 7]);
                *reinterpret_cast<uint4*>(out_px + cg + i) = o;
              }
            }
            float sq[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) sq[i] = v[i] * v[i];
            const float s1 = warp_column_sums(v, lane);
            const float s2 = warp_column_sums(sq, lane);
            atomicAdd(&stat[(acc * 2 + 0) * 256 + c0 + lane], s1);
            atomicAdd(&stat[(acc * 2 + 1) * 256 + c0 + lane], s2);
          } else if (valid) {
            switch (P.act) {
              case UAVDET_ACT_LEAKY: affine_act_store<UAVDET_ACT_LEAKY>(r, P, cg, out_px, res_px, shift); break;
              case UAVDET_ACT_SILU: affine_act_store<UAVDET_ACT_SILU>(r, P, cg, out_px, res_px, shift); break;
              case UAVDET_ACT_RELU: affine_act_store<UAVDET_ACT_RELU>(r, P, cg, out_px, res_px, shift); break;
              case UAVDET_ACT_GELU: affine_act_store<UAVDET_ACT_GELU>(r, P, cg, out_px, res_px, shift); break;
              default: affine_act_store<UAVDET_ACT_NONE>(r, P, cg, out_px, res_px, shift); break;
            }
          }
        }
      }
      // TMEM buffer drained: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[acc]));

      if (P.epi == UAVDET_EPI_STATS) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
        for (int c = e_tid; c < P.block_n; c += 128) {
          if (n0 + c < P.cout) {
            atomicAdd(P.sum + n0 + c, stat[(acc * 2 + 0) * 256 + c]);
            atomicAdd(P.sumsq + n0 + c, stat[(acc * 2 + 1) * 256 + c]);
          }
          stat[(acc * 2 + 0) * 256 + c] = 0.f;
          stat[(acc * 2 + 1) * 256 + c] = 0.f;
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
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
int make_act_map(CUtensorMap* m, const uavdet_act* x, int parity, int box_c, int box_w, int box_h) {
  const uint64_t eb = 2;
  uint64_t dims[5], str[4];
  uint32_t box[5] = {(uint32_t)box_c, (uint32_t)box_w, 1u, (uint32_t)box_h, 1u};
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

// Pick the output tile rectangle (tile_w * tile_h <= 128) that wastes the fewest MMA rows;
// ties go to the squarest tile (smallest halo, best L2 reuse across the filter taps).
void choose_tile(int ho, int wo, int* tile_w, int* tile_h) {
  double best_eff = -1.0;
  int best_halo = 1 << 30, bw = 1, bh = 1;
  const int max_w = wo < 128 ? wo : 128;
  for (int tw = 1; tw <= max_w; ++tw) {
    int th = 128 / tw;
    if (th > ho) th = ho;
    const double tiles = (double)ceil_div(ho, th) * ceil_div(wo, tw);
    const double eff = (double)ho * wo / (tiles * 128.0);
    const int halo = (tw + 2) * (th + 2);
    if (eff > best_eff + 1e-9 || (eff > best_eff - 1e-9 && halo < best_halo)) {
      best_eff = eff; best_halo = halo; bw = tw; bh = th;
    }
  }
  *tile_w = bw;
  *tile_h = bh;
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

static int pick_block_n(int cout) {
  if (cout <= 16) return 16;
  for (int bn = 256; bn >= 32; bn -= 32)
    if (cout % bn == 0) return bn;
  return cout < 256 ? ((cout + 31) / 32) * 32 : 256;
}

int launch_igemm(const uavdet_act* a_src, int parity, const void* w_packed, int w_rows, int k_total,
                 int w_batch, IgemmParams& P, cudaStream_t st) {
  CUtensorMap mapA, mapB;
  int rc = make_act_map(&mapA, a_src, parity, P.block_k, P.tile_w, P.tile_h);
  if (rc) return rc;
  {
    uint64_t dims[3] = {(uint64_t)k_total, (uint64_t)w_rows, (uint64_t)w_batch};
    uint64_t str[2] = {(uint64_t)k_total * 2, (uint64_t)k_total * 2 * w_rows};
    uint32_t box[3] = {(uint32_t)P.block_k, (uint32_t)P.block_n, 1u};
    rc = encode_tensor_map(&mapB, const_cast<void*>(w_packed), 3, dims, str, box, P.block_k * 2);
    if (rc) return rc;
  }
  P.tiles_w = ceil_div(P.wo, P.tile_w);
  P.tiles_h = ceil_div(P.ho, P.tile_h);
  P.n_tiles = ceil_div(P.cout, P.block_n);
  P.total_tiles = P.n_img * P.tiles_h * P.tiles_w * P.n_tiles;
  P.w_batch = w_batch;
  const int stage_bytes = 128 * P.block_k * 2 + P.block_n * P.block_k * 2;
  const int ctrl_bytes = 8 * (2 * kMaxStages + 4) + 16 + 2 * 2 * 256 * 4;
  const int max_smem = 227 * 1024;
  int stages = (max_smem - 1024 - ctrl_bytes) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  UAVDET_CHECK_ARG(stages >= 2, "igemm: tile does not fit shared memory");
  P.stages = stages;
  P.watchdog = watchdog_word();
  // Always request (almost) the whole shared memory so exactly one CTA is resident per SM:
  // each CTA allocates all 512 TMEM columns.
  const int smem_bytes = max_smem;
  static bool attr_set = false;
  if (!attr_set) {
    UAVDET_CUDA(cudaFuncSetAttribute(igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    attr_set = true;
  }
  if (P.total_tiles <= 0) return UAVDET_OK;
  int grid = P.total_tiles < kNumSMs ? P.total_tiles : kNumSMs;
  igemm_kernel<<<grid, kIgemmThreads, smem_bytes, st>>>(mapA, mapB, P);
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}

static int fill_epilogue(IgemmParams& P, const uavdet_epilogue* epi, const uavdet_act* y, int cout) {
  P.epi = epi ? epi->epi : UAVDET_EPI_AFFINE;
  P.act = epi ? epi->act : UAVDET_ACT_NONE;
  P.scale = epi ? epi->scale : nullptr;
  P.shift = epi ? epi->shift : nullptr;
  P.shift_sn = (epi && epi->shift_per_sample) ? cout : 0;
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
  }
  return UAVDET_OK;
}

}  // namespace uavdet

using namespace uavdet;

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
  int nt = 0;
  for (int kh = 0; kh < k; ++kh)
    for (int kw = 0; kw < k; ++kw) {
      if (s2d) {
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
  choose_tile(P.ho, P.wo, &P.tile_w, &P.tile_h);
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
      choose_tile(P.ho, P.wo, &P.tile_w, &P.tile_h);
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
