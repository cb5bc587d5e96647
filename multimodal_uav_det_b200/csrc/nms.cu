// K9 — greedy NMS, bit-exact with torchvision.ops.nms (CPU kernel semantics).
// Replaces the call at reference model/_base.py:203.
//
// One thread-block cluster (1-8 CTAs of 1024 threads, chosen from the batch size) per image:
//   1. key build      : key = ~orderable(score)  (NaN first, -0 == +0), value = index
//   2. stable LSD radix sort, 4 x 8-bit passes, ping-pong in the workspace             (CTA 0)
//      (warp match_any ranking keeps equal keys in index order == torch stable sort)
//   3. gather boxes into score order (float4, coalesced afterwards)                      (CTA 0)
//   4. chunked greedy suppression, 64 sorted candidates at a time; chunk c is owned by CTA c mod cluster size
//        a. 64x64 pair mask by warp ballot (each warp: 2 rows x 64 columns)             (owner)
//        b. serial resolve of the chunk against its own mask                            (owner)
//        c. the <=64 kept boxes are stored into every CTA's shared memory (DSMEM), one cluster barrier
//        d. every CTA tests the still-alive candidates of its own later chunks against them; suppressed ones
//           are marked in the CTA's shared-memory bit array
// All IoU arithmetic uses explicit round-to-nearest intrinsics so no FMA contraction can
// change a rounding relative to the CPU reference.
#include "common.cuh"
#include "sm100.cuh"

namespace uavdet {

constexpr int kNmsThreads = 1024;
constexpr int kNmsWarps = kNmsThreads / 32;
constexpr int kChunk = 64;

__device__ __forceinline__ uint32_t score_key(float s) {
  uint32_t b = __float_as_uint(s);
  uint32_t ord;
  if (s != s) {
    ord = 0xffffffffu;  // NaN sorts above everything (torch descending sort puts NaN first)
  } else {
    if (b == 0x80000000u) b = 0u;  // -0.0 == +0.0
    ord = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  }
  return ~ord;  // ascending key == descending score
}

// std::max(a, b) / std::min(a, b) exactly as libstdc++ evaluates them (NaN behaviour incl.)
__device__ __forceinline__ float std_max(float a, float b) { return (a < b) ? b : a; }
__device__ __forceinline__ float std_min(float a, float b) { return (b < a) ? b : a; }

__device__ __forceinline__ float box_area(const float4& b) {
  return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
}

// i = the kept (earlier) box, j = the candidate.  torchvision/csrc/ops/cpu/nms_kernel.cpp
// The reference decides `fl(inter / (ai + aj - inter)) > thr` with an IEEE division.  The division is ~30
// instructions, so the exact quotient is only formed when the outcome is not already certain.  With
// p = fl(thr * u) and u > 0, |fl(x) - x| <= 2^-24 |x| for the product and for the quotient (rounding is monotone),
// so inter > fl(p (1 + 2^-20)) implies fl(inter / u) > thr and inter < fl(p (1 - 2^-20)) implies the opposite
// (inter == 0 falls in the second case).  That needs p to be a normal, positive, finite float -- the exponent
// range check below, which also rejects NaN and u <= 0 -- and a sane threshold: the caller passes thr_fast = thr
// when thr is in [1e-6, 1e6] and NaN otherwise, which sends every pair down the exact path.
// kNaN = false (no NaN coordinate among the candidates of this image, established once per launch) lets the
// std::max / std::min selects collapse to single FMNMX instructions: the two only differ on NaN operands
// (signed zeros change at most the sign of a zero width, never `inter`).
template <bool kNaN>
__device__ __forceinline__ bool iou_exceeds(const float4& bi, float ai, const float4& bj, float aj,
                                            float thr, float thr_fast) {
  float xx1, yy1, xx2, yy2;
  if (kNaN) {
    xx1 = std_max(bi.x, bj.x);
    yy1 = std_max(bi.y, bj.y);
    xx2 = std_min(bi.z, bj.z);
    yy2 = std_min(bi.w, bj.w);
  } else {
    xx1 = fmaxf(bi.x, bj.x);
    yy1 = fmaxf(bi.y, bj.y);
    xx2 = fminf(bi.z, bj.z);
    yy2 = fminf(bi.w, bj.w);
  }
  // std::max(0.f, x) == fmaxf(0.f, x) for every x (NaN -> 0, -0 -> +0 in both)
  const float w = fmaxf(0.f, __fsub_rn(xx2, xx1));
  const float h = fmaxf(0.f, __fsub_rn(yy2, yy1));
  const float inter = __fmul_rn(w, h);
  const float u = __fsub_rn(__fadd_rn(ai, aj), inter);
  const float p = __fmul_rn(thr_fast, u);
  const bool in_range = (__float_as_uint(p) - 0x0D800000u) < 0x64000000u;       // 2^-100 <= p < 2^100
  const bool hi = inter > __fmul_rn(p, 1.00000095367431640625f);                // 1 + 2^-20
  const bool lo = inter < __fmul_rn(p, 0.99999904632568359375f);                // 1 - 2^-20
  if (in_range && (hi || lo)) return hi;
  return __fdiv_rn(inter, u) > thr;
}

constexpr int kSlots = 8;  // ring of kept-box lists; slot = chunk % 8, so a slot always has the same owner CTA
// bytes one owner sends into one CTA's slot: 64 boxes + 64 areas + the count (always the full slot, so that the
// receiving mbarrier's transaction count is a constant)
constexpr uint32_t kSlotTxBytes = kChunk * 16 + kChunk * 4 + 4;

struct NmsSmem {
  uint32_t bin[256];              // digit histogram / running offsets
  uint32_t warp_off[kNmsWarps][256];  // per-warp digit counts, then scatter offsets
  float4 cbox[kChunk];
  float carea[kChunk];
  unsigned long long cmask[kChunk];
  float4 kbox[kSlots][kChunk];    // kept boxes of chunk c live in slot c % 8 of EVERY CTA of the cluster (st.async
  float karea[kSlots][kChunk];    //  by the chunk's owner, completion counted on the receiver's full[] barrier)
  uint32_t nk[kSlots];
  unsigned long long full[kSlots];   // armed with kSlotTxBytes by thread 0, completed by the owner's st.async
  unsigned long long empty[kSlots];  // in the slot's owner CTA: one arrival per warp of the cluster when it is done
  int n_valid;
  int has_nan;
  unsigned long long kept_mask;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory variable in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t dsmem_addr(const void* p, uint32_t rank) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p), r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
  return r;
}
// store into another CTA's shared memory; the bytes are counted on that CTA's mbarrier when they have landed
__device__ __forceinline__ void st_async_f4(uint32_t addr, const float4& v, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(addr),
               "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)),
               "r"(__float_as_uint(v.w)), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void st_async_u32(uint32_t addr, uint32_t v, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(addr), "r"(v),
               "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Bounded wait: a barrier bug traps (the host sees a launch error) instead of hanging the GPU.
__device__ __forceinline__ void nms_mbar_wait(uint32_t bar, uint32_t parity) {
  if (sm100::mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!sm100::mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}

#ifdef UAVDET_NMS_PROFILE
#define NMS_PROF(acc) do { pt = clock64(); acc += pt - pt0; pt0 = pt; } while (0)
#else
#define NMS_PROF(acc) do { } while (0)
#endif

// The fast part of iou_exceeds: bit 0 = the outcome if certain, bit 1 = uncertain (take the exact path).  Kept
// separate so that several tests can be in flight per thread without a branch between them.
template <bool kNaN>
__device__ __forceinline__ uint32_t iou_code(const float4& bi, float ai, const float4& bj, float aj, float thr_fast) {
  float xx1, yy1, xx2, yy2;
  if (kNaN) {
    xx1 = std_max(bi.x, bj.x);
    yy1 = std_max(bi.y, bj.y);
    xx2 = std_min(bi.z, bj.z);
    yy2 = std_min(bi.w, bj.w);
  } else {
    xx1 = fmaxf(bi.x, bj.x);
    yy1 = fmaxf(bi.y, bj.y);
    xx2 = fminf(bi.z, bj.z);
    yy2 = fminf(bi.w, bj.w);
  }
  const float w = fmaxf(0.f, __fsub_rn(xx2, xx1));
  const float h = fmaxf(0.f, __fsub_rn(yy2, yy1));
  const float inter = __fmul_rn(w, h);
  const float u = __fsub_rn(__fadd_rn(ai, aj), inter);
  const float p = __fmul_rn(thr_fast, u);
  const bool in_range = (__float_as_uint(p) - 0x0D800000u) < 0x64000000u;
  const bool hi = inter > __fmul_rn(p, 1.00000095367431640625f);
  const bool lo = inter < __fmul_rn(p, 0.99999904632568359375f);
  return (hi ? 1u : 0u) | ((in_range && (hi || lo)) ? 0u : 2u);
}
// exact reference arithmetic, for the pairs iou_code could not decide
template <bool kNaN>
__device__ __noinline__ bool iou_exact(const float4& bi, float ai, const float4& bj, float aj, float thr) {
  const float xx1 = std_max(bi.x, bj.x);
  const float yy1 = std_max(bi.y, bj.y);
  const float xx2 = std_min(bi.z, bj.z);
  const float yy2 = std_min(bi.w, bj.w);
  const float w = std_max(0.f, __fsub_rn(xx2, xx1));
  const float h = std_max(0.f, __fsub_rn(yy2, yy1));
  const float inter = __fmul_rn(w, h);
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(ai, aj), inter)) > thr;
}

// Step 4 of the kernel below.  Returns the number of kept boxes (identical in every CTA of the cluster).
//
// Candidates are split over the cluster by 64-candidate chunk: chunk c belongs to CTA c % ncta, and inside that
// CTA own chunk number i (= c / ncta) belongs for good to the two warps of thread group i % 16, so a word of the
// removed[] bit array is only ever touched by one warp and the sweeps need no CTA-wide barrier.  The kept boxes
// of chunk c travel through slot c % 8 of a ring present in every CTA: the owner waits for the slot's empty
// barrier (every warp of the cluster has finished with its previous content), st.async's the list into every
// CTA, and each warp waits on its CTA's full barrier before sweeping.  No cluster-wide barrier in the loop.
template <bool kNaN>
__device__ __noinline__ int suppress(NmsSmem& S, uint32_t* removed, const float4* sbox, const uint32_t* order,
                                     int64_t* keep, int nv, float thr, float thr_fast, uint32_t ncta,
                                     uint32_t rank) {
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int grp = tid >> 6;            // thread group: 64 threads = one chunk wide
  const int off = tid & (kChunk - 1);
  constexpr int kGroups = kNmsThreads / kChunk;
#ifdef UAVDET_NMS_PROFILE
  long long pt0 = clock64(), p_o1 = 0, p_o2 = 0, p_o3 = 0, p_o4 = 0, p_bar = 0, p_sweep = 0, pt;
#endif
  const int n_chunks = (nv + kChunk - 1) / kChunk;
  int kept_total = 0;
  // the boxes / original indices of this CTA's next own chunk, fetched one own-chunk ahead (threads 0..63)
  float4 pf_box = make_float4(0.f, 0.f, 0.f, 0.f);
  uint32_t pf_ord = 0;
  if (tid < kChunk && (int)rank * kChunk + tid < nv) {
    pf_box = sbox[rank * kChunk + tid];
    pf_ord = order[rank * kChunk + tid];
  }
  for (int c = 0; c < n_chunks; ++c) {
    const int c0 = c * kChunk;
    const int slot = c & (kSlots - 1);
    const uint32_t use = (uint32_t)c / kSlots;
    const uint32_t owner = (uint32_t)c % ncta;
    if (rank == owner) {
      const int cn = min(kChunk, nv - c0);
      const uint32_t my_ord = pf_ord;
      if (tid < kChunk) {
        S.cbox[tid] = pf_box;
        S.carea[tid] = box_area(pf_box);
        const int nj = c0 + (int)ncta * kChunk + tid;
        if (nj < nv) {
          pf_box = sbox[nj];
          pf_ord = order[nj];
        }
      }
      __syncthreads();  // cbox ready; every warp's sweep of the previous chunks has updated removed[]
      const unsigned long long rem_in =
          (unsigned long long)removed[c0 >> 5] | ((unsigned long long)removed[(c0 >> 5) + 1] << 32);
      const unsigned long long live_mask = (cn == 64) ? ~0ull : ((1ull << cn) - 1ull);
      const unsigned long long alive_in = ~rem_in & live_mask;
      NMS_PROF(p_o1);
      unsigned long long kept = 0ull;
      if (alive_in != 0ull) {  // (CTA-uniform) otherwise the whole chunk is already suppressed
        // a. pair mask over the still-alive candidates: warp w owns rows 2w and 2w+1, lanes cover columns lane
        //    and lane + 32
        const float4 bc0 = S.cbox[lane], bc1 = S.cbox[lane + 32];
        const float ac0 = S.carea[lane], ac1 = S.carea[lane + 32];
        uint32_t code[4];
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int r = warp * 2 + rr;
          const float4 bi = S.cbox[r];
          const float ai = S.carea[r];
          const bool row_alive = (alive_in >> r) & 1ull;
          const bool t0 = row_alive && lane > r && ((alive_in >> lane) & 1ull);
          const bool t1 = row_alive && lane + 32 > r && ((alive_in >> (lane + 32)) & 1ull);
          code[rr * 2 + 0] = t0 ? iou_code<kNaN>(bi, ai, bc0, ac0, thr_fast) : 0u;
          code[rr * 2 + 1] = t1 ? iou_code<kNaN>(bi, ai, bc1, ac1, thr_fast) : 0u;
        }
        if ((code[0] | code[1] | code[2] | code[3]) & 2u) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (code[q] & 2u) {
              const int r = warp * 2 + (q >> 1);
              code[q] = iou_exact<kNaN>(S.cbox[r], S.carea[r], (q & 1) ? bc1 : bc0, (q & 1) ? ac1 : ac0, thr) ? 1u : 0u;
            }
        }
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const uint32_t m0 = __ballot_sync(0xffffffffu, code[rr * 2 + 0] & 1u);
          const uint32_t m1 = __ballot_sync(0xffffffffu, code[rr * 2 + 1] & 1u);
          if (lane == 0) S.cmask[warp * 2 + rr] = (unsigned long long)m0 | ((unsigned long long)m1 << 32);
        }
        __syncthreads();
        NMS_PROF(p_o2);
        // b. serial resolve by warp 0 (warp-uniform): hop from one surviving candidate to the next
        if (warp == 0) {
          unsigned long long alive = alive_in;
          while (alive) {
            const int r = __ffsll((long long)alive) - 1;
            kept |= 1ull << r;
            alive &= ~(S.cmask[r] | (1ull << r));  // rows only hold columns > r
          }
          if (lane == 0) S.kept_mask = kept;
        }
        __syncthreads();
        kept = S.kept_mask;
        NMS_PROF(p_o3);
      }
      const int nk = __popcll(kept);
      // c. the slot must be free in every CTA, then: kept boxes -> every CTA's slot; indices -> keep[]
      nms_mbar_wait(sm100::smem_u32(&S.empty[slot]), (use & 1u) ^ 1u);
      {
        const uint32_t q = (uint32_t)grp;  // destination CTA (16 thread groups >= cluster size)
        if (q < ncta) {
          const uint32_t bar = dsmem_addr(&S.full[slot], q);
          // entry `off` of the slot: the off-th kept box (entries >= nk are never read; any bytes will do)
          int src = off;
          if (off < nk) {
            unsigned long long m = kept;
            for (int t = 0; t < off; ++t) m &= m - 1;  // drop the `off` lowest set bits
            src = __ffsll((long long)m) - 1;
          }
          st_async_f4(dsmem_addr(&S.kbox[slot][off], q), S.cbox[src], bar);
          st_async_u32(dsmem_addr(&S.karea[slot][off], q), __float_as_uint(S.carea[src]), bar);
          if (off == 0) st_async_u32(dsmem_addr(&S.nk[slot], q), (uint32_t)nk, bar);
        }
        if (tid < kChunk && ((kept >> tid) & 1ull))
          keep[kept_total + __popcll(kept & ((1ull << tid) - 1ull))] = (int64_t)my_ord;
      }
      NMS_PROF(p_o4);
    }
    const uint32_t full_bar = sm100::smem_u32(&S.full[slot]);
    nms_mbar_wait(full_bar, use & 1u);
    if (tid == 0) sm100::mbar_arrive_expect_tx(full_bar, kSlotTxBytes);  // arm the slot's next use
    NMS_PROF(p_bar);
    const int nk = (int)S.nk[slot];
    kept_total += nk;
    if (nk != 0) {
      // d. suppress this warp's later candidates: own chunks i = grp (mod 16) with chunk number > c
      const int i_min = (c >= (int)rank) ? (c - (int)rank) / (int)ncta + 1 : 0;
      int i = i_min + ((grp - i_min) & (kGroups - 1));
      const int stride = kGroups * (int)ncta * kChunk;
      int j = ((int)rank + i * (int)ncta) * kChunk + off;
      const int j_end = n_chunks * kChunk;
      float4 nb = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j < nv) nb = sbox[j];
      for (; j < j_end; j += stride) {
        const float4 bj = nb;
        if (j + stride < nv) nb = sbox[j + stride];
        const uint32_t word = removed[j >> 5];
        bool hit = false;
        if (j < nv && !((word >> (j & 31)) & 1u)) {
          const float aj = box_area(bj);
          int k = 0;
          for (; k + 4 <= nk; k += 4) {
            const uint32_t q0 = iou_code<kNaN>(S.kbox[slot][k + 0], S.karea[slot][k + 0], bj, aj, thr_fast);
            const uint32_t q1 = iou_code<kNaN>(S.kbox[slot][k + 1], S.karea[slot][k + 1], bj, aj, thr_fast);
            const uint32_t q2 = iou_code<kNaN>(S.kbox[slot][k + 2], S.karea[slot][k + 2], bj, aj, thr_fast);
            const uint32_t q3 = iou_code<kNaN>(S.kbox[slot][k + 3], S.karea[slot][k + 3], bj, aj, thr_fast);
            uint32_t any = q0 | q1 | q2 | q3;
            if (any & 2u) {
              any = 0u;
              const uint32_t qq[4] = {q0, q1, q2, q3};
#pragma unroll
              for (int t = 0; t < 4; ++t)
                any |= (qq[t] & 2u) ? (iou_exact<kNaN>(S.kbox[slot][k + t], S.karea[slot][k + t], bj, aj, thr) ? 1u : 0u)
                                    : qq[t];
            }
            if (any & 1u) { hit = true; break; }
          }
          if (!hit)
            for (; k < nk; ++k)
              if (iou_exceeds<kNaN>(S.kbox[slot][k], S.karea[slot][k], bj, aj, thr, thr_fast)) { hit = true; break; }
        }
        const uint32_t hits = __ballot_sync(0xffffffffu, hit);
        if (lane == 0 && hits) removed[j >> 5] = word | hits;
      }
    }
    // this warp is done with the slot: tell the slot's owner (every lane has read what it needed)
    __syncwarp();
    if (lane == 0) mbar_arrive_remote(dsmem_addr(&S.empty[slot], (uint32_t)slot % ncta));
    NMS_PROF(p_sweep);
  }
#ifdef UAVDET_NMS_PROFILE
  if (tid == 0 && blockIdx.x < ncta)
    printf("nms rank %u/%u: o1 %lld mask %lld resolve %lld bcast %lld wait %lld sweep %lld cycles, chunks %d\n",
           rank, ncta, p_o1, p_o2, p_o3, p_o4, p_bar, p_sweep, n_chunks);
#endif
  return kept_total;
}

// One image per thread-block CLUSTER (1, 2, 4 or 8 CTAs).  CTA 0 sorts; the suppression sweep -- the O(n * kept)
// part -- is split over the CTAs by 64-candidate chunk (chunk c belongs to CTA c mod cluster size, which keeps
// that chunk's `removed` bits in its own shared memory).  Per chunk: the owner resolves it and stores the kept
// boxes into every CTA's shared memory, one cluster barrier, then every CTA sweeps its own later chunks.
__global__ void __launch_bounds__(kNmsThreads, 1)
nms_image_kernel(const float* __restrict__ boxes, const float* __restrict__ scores, int n,
                 float thr, float score_floor, int64_t* __restrict__ keep,
                 int32_t* __restrict__ keep_count, uint8_t* __restrict__ workspace,
                 size_t ws_per_image) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  NmsSmem& S = *reinterpret_cast<NmsSmem*>(smem_raw);
  uint32_t* removed = reinterpret_cast<uint32_t*>(smem_raw + sizeof(NmsSmem));

  const uint32_t ncta = cluster_nctarank();
  const uint32_t rank = cluster_ctarank();
  const int img = blockIdx.x / ncta;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  boxes += (size_t)img * n * 4;
  scores += (size_t)img * n;
  keep += (size_t)img * n;

  // workspace carve-up (per image): keys A/B, vals A/B, sorted boxes
  uint8_t* ws = workspace + (size_t)img * ws_per_image;
  const size_t n_pad = ((size_t)n + 3) & ~(size_t)3;
  uint32_t* keyA = reinterpret_cast<uint32_t*>(ws);
  uint32_t* keyB = keyA + n_pad;
  uint32_t* valA = keyB + n_pad;
  uint32_t* valB = valA + n_pad;
  float4* sbox = reinterpret_cast<float4*>(valB + n_pad);

#ifdef UAVDET_NMS_PROFILE
  long long pt0 = clock64(), p_sort = 0, p_loop = 0, pt;
#endif
  // ---- 1. keys (CTA 0) and the count of candidates above the floor (every CTA) ------------
  if (tid == 0) {
    S.n_valid = 0;
    for (int q = 0; q < kSlots; ++q) {
      sm100::mbar_init(sm100::smem_u32(&S.full[q]), 1);
      sm100::mbar_init(sm100::smem_u32(&S.empty[q]), ncta * kNmsWarps);
    }
    sm100::fence_barrier_init();
    for (int q = 0; q < kSlots; ++q) sm100::mbar_arrive_expect_tx(sm100::smem_u32(&S.full[q]), kSlotTxBytes);
  }
  __syncthreads();
  int local_valid = 0;
  for (int i = tid; i < n; i += kNmsThreads) {
    float s = scores[i];
    if (rank == 0) {
      keyA[i] = score_key(s);
      valA[i] = (uint32_t)i;
    }
    local_valid += (s != s || s > score_floor) ? 1 : 0;
  }
  local_valid = __reduce_add_sync(0xffffffffu, local_valid);
  if (lane == 0 && local_valid) atomicAdd(&S.n_valid, local_valid);

  // ---- 2. stable LSD radix sort (CTA 0) ---------------------------------------------------
  uint32_t* kin = keyA; uint32_t* kout = keyB;
  uint32_t* vin = valA; uint32_t* vout = valB;
  for (int pass = 0; pass < 4 && rank == 0; ++pass) {
    const int shift = pass * 8;
    if (tid < 256) S.bin[tid] = 0;
    __syncthreads();  // also orders the previous pass' global writes within the CTA
    for (int base = 0; base < n; base += kNmsThreads) {  // warp-aggregated histogram
      const int i = base + tid;
      const uint32_t digit = (i < n) ? ((kin[i] >> shift) & 255u) : 256u;
      const uint32_t peers = __match_any_sync(0xffffffffu, digit);
      if (digit < 256u && (peers & ((1u << lane) - 1u)) == 0u) atomicAdd(&S.bin[digit], __popc(peers));
    }
    __syncthreads();
    // exclusive scan of 256 bins by warp 0 (8 per lane)
    if (warp == 0) {
      uint32_t v[8], sum = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) { v[q] = S.bin[lane * 8 + q]; sum += v[q]; }
      uint32_t incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      uint32_t run = incl - sum;
#pragma unroll
      for (int q = 0; q < 8; ++q) { S.bin[lane * 8 + q] = run; run += v[q]; }
    }
    __syncthreads();
    // stable scatter, one tile of 1024 consecutive items at a time
    for (int base = 0; base < n; base += kNmsThreads) {
      const int i = base + tid;
      const bool valid = i < n;
      uint32_t key = 0, val = 0, digit = 256;  // digit 256 = "no item"
      if (valid) { key = kin[i]; val = vin[i]; digit = (key >> shift) & 255u; }
      // rank among same-digit lanes of this warp (lower lanes first)
      const uint32_t peers = __match_any_sync(0xffffffffu, digit);
      const uint32_t rank_in_warp = __popc(peers & ((1u << lane) - 1u));
#pragma unroll
      for (int q = 0; q < 8; ++q) S.warp_off[warp][lane + 32 * q] = 0;
      __syncwarp();
      if (valid && rank_in_warp == 0) S.warp_off[warp][digit] = __popc(peers);
      __syncthreads();
      // per digit: counts -> exclusive offsets over warps, advance the running bin offset
      if (tid < 256) {
        uint32_t run = S.bin[tid];
#pragma unroll 8
        for (int w = 0; w < kNmsWarps; ++w) {
          uint32_t c = S.warp_off[w][tid];
          S.warp_off[w][tid] = run;
          run += c;
        }
        S.bin[tid] = run;
      }
      __syncthreads();
      if (valid) {
        const uint32_t pos = S.warp_off[warp][digit] + rank_in_warp;
        kout[pos] = key;
        vout[pos] = val;
      }
      __syncthreads();
    }
    uint32_t* t = kin; kin = kout; kout = t;
    t = vin; vin = vout; vout = t;
  }
  __syncthreads();
  const uint32_t* order = valA;  // after 4 passes the result is back in the A buffers

  // ---- 3. gather boxes into score order (CTA 0), clear the removed bits --------------------
  const int nv = S.n_valid;
  const float4* boxes4 = reinterpret_cast<const float4*>(boxes);
  uint32_t* nan_flag = reinterpret_cast<uint32_t*>(sbox + n_pad);
  if (rank == 0) {
    if (tid == 0) S.has_nan = 0;
    __syncthreads();
    bool nan = false;
    for (int i = tid; i < nv; i += kNmsThreads) {
      const float4 b = boxes4[order[i]];
      sbox[i] = b;
      nan |= (b.x != b.x) | (b.y != b.y) | (b.z != b.z) | (b.w != b.w);
    }
    if (__any_sync(0xffffffffu, nan) && lane == 0) S.has_nan = 1;
    __syncthreads();
    if (tid == 0) *nan_flag = (uint32_t)S.has_nan;
  }
  const int n_words = (nv + 31) >> 5;
  for (int i = tid; i < n_words + 2; i += kNmsThreads) removed[i] = 0;
  // makes CTA 0's sorted boxes / order / flag visible to the other CTAs, and guarantees every CTA of the cluster
  // is running before any distributed-shared-memory store below
  if (ncta > 1) cluster_sync_all(); else __syncthreads();
  NMS_PROF(p_sort);

  // ---- 4. chunked greedy suppression ------------------------------------------------------
  const float thr_fast = (thr >= 1e-6f && thr <= 1e6f) ? thr : __int_as_float(0x7fc00000);
  int kept_total;
  if (*nan_flag)
    kept_total = suppress<true>(S, removed, sbox, order, keep, nv, thr, thr_fast, ncta, rank);
  else
    kept_total = suppress<false>(S, removed, sbox, order, keep, nv, thr, thr_fast, ncta, rank);
  if (rank == 0 && tid == 0) keep_count[img] = kept_total;
  // tail of `keep` beyond keep_count is left untouched (caller slices by count)
  // no CTA may leave while another can still arrive on its barriers
  if (ncta > 1) cluster_sync_all();
#ifdef UAVDET_NMS_PROFILE
  NMS_PROF(p_loop);
  if (tid == 0 && img == 0)
    printf("nms rank %u/%u: sort %lld loop %lld cycles, kept %d\n", rank, ncta, p_sort, p_loop, kept_total);
#endif
}

static size_t nms_ws_per_image(int n) {
  size_t n_pad = ((size_t)n + 3) & ~(size_t)3;
  size_t bytes = n_pad * 4 * sizeof(uint32_t) + n_pad * sizeof(float4) + 16;  // + the has-NaN flag
  return (bytes + 255) & ~(size_t)255;
}

}  // namespace uavdet

using namespace uavdet;

extern "C" size_t uavdet_nms_workspace_bytes(int batch, int n) {
  if (batch <= 0 || n <= 0) return 256;
  return nms_ws_per_image(n) * (size_t)batch;
}

extern "C" int uavdet_nms(const float* boxes, const float* scores, int batch, int n, double iou_thr,
                          float score_floor, int64_t* keep, int32_t* keep_count, void* workspace,
                          size_t workspace_bytes, void* stream) {
  UAVDET_CHECK_ARG(batch >= 0 && n >= 0, "nms: negative sizes");
  if (batch == 0) return UAVDET_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    UAVDET_CUDA(cudaMemsetAsync(keep_count, 0, sizeof(int32_t) * batch, st));
    return UAVDET_OK;
  }
  UAVDET_CHECK_ARG(boxes && scores && keep && keep_count && workspace, "nms: null pointer");
  UAVDET_CHECK_ARG(workspace_bytes >= uavdet_nms_workspace_bytes(batch, n),
                   "nms: workspace too small (%zu < %zu)", workspace_bytes,
                   uavdet_nms_workspace_bytes(batch, n));
  UAVDET_CHECK_ARG(((uintptr_t)boxes & 15) == 0 && ((uintptr_t)workspace & 15) == 0,
                   "nms: boxes/workspace must be 16-byte aligned");
  // `ovr > thr` is evaluated in double by the reference (float ovr vs double thr); the
  // largest float <= thr gives the identical predicate in fp32.
  float thr_f = (float)iou_thr;
  if ((double)thr_f > iou_thr) thr_f = nextafterf(thr_f, -INFINITY);
  size_t smem = sizeof(NmsSmem) + (((size_t)n + 31) / 32) * 4 + 32;
  UAVDET_CHECK_ARG(smem <= 227 * 1024, "nms: n=%d too large for the shared bit array", n);
  static bool attr_set = false;
  if (!attr_set) {
    UAVDET_CUDA(cudaFuncSetAttribute(nms_image_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     227 * 1024));
    attr_set = true;
  }
  // CTAs per image: as many as keep the whole batch co-resident (about 16 clusters of 8 fit on 148 SMs)
  int ncta = 1;
  if (n >= 4096) ncta = batch <= 16 ? 8 : batch <= 32 ? 4 : batch <= 64 ? 2 : 1;
  if (const char* e = getenv("UAVDET_NMS_CLUSTER")) {
    int v = atoi(e);
    if (v == 1 || v == 2 || v == 4 || v == 8) ncta = v;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)batch * ncta);
  cfg.blockDim = dim3(kNmsThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = ncta;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  UAVDET_CUDA(cudaLaunchKernelEx(&cfg, nms_image_kernel, boxes, scores, n, thr_f, score_floor, keep, keep_count,
                                 (uint8_t*)workspace, nms_ws_per_image(n)));
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}
