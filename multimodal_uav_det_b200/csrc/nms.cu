// K9 — greedy NMS, bit-exact with torchvision.ops.nms (CPU kernel semantics).
// Replaces the call at reference model/_base.py:203.
//
// One thread-block cluster (1-8 CTAs of 512 threads, chosen from the batch size) per image:
//   1. key build      : key = ~orderable(score)  (NaN first, -0 == +0), value = index          (CTA 0)
//   2. stable LSD radix sort, 4 x 8-bit passes, ping-pong in the workspace                     (CTA 0)
//      (warp match_any ranking keeps equal keys in index order == torch stable sort)
//   3. every warp of the cluster gathers ITS candidates (boxes + sorted position) into a private list:
//      chunk c (64 consecutive sorted positions) belongs to CTA c % ncta, warp (c / ncta) % 16
//   4. greedy suppression as a warp-level data flow, no CTA- or cluster-wide barrier:
//        * the kept boxes of chunk c ("list c") reach every CTA through slot c % 8 of a shared-memory ring
//          (st.async + the receiver's mbarrier; an empty-barrier in the owner CTA recycles the slot);
//        * a warp sweeps its private list against list c and compacts the survivors in place, so its lanes
//          stay densely occupied as candidates die;
//        * the survivors of chunk c + 1 sit at the front of their owner warp's list: right after the first
//          trip of that sweep the owner resolves the chunk inside the warp (hop from survivor to survivor,
//          one ballot per hop), publishes list c + 1, and only then finishes its own sweep -- the serial
//          chain of the greedy algorithm overlaps the bulk sweeps of everybody else.
// All IoU arithmetic uses explicit round-to-nearest intrinsics so no FMA contraction can
// change a rounding relative to the CPU reference.
#include "common.cuh"
#include "sm100.cuh"

namespace uavdet {

#ifndef UAVDET_NMS_THREADS
#define UAVDET_NMS_THREADS 512
#endif
constexpr int kNmsThreads = UAVDET_NMS_THREADS;
constexpr int kNmsWarps = kNmsThreads / 32;
constexpr int kChunk = 64;
#ifndef UAVDET_NMS_SWEEP_WARPS
#define UAVDET_NMS_SWEEP_WARPS 16
#endif
static_assert(UAVDET_NMS_SWEEP_WARPS * 32 <= UAVDET_NMS_THREADS, "more sweeping warps than warps");
// warps per CTA that hold candidate lists and run step 4 (the others only help with the sort): fewer warps per
// scheduler means the warp that is resolving the next chunk -- the serial chain -- waits less for an issue slot
constexpr int kSweepWarps = UAVDET_NMS_SWEEP_WARPS;

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
// The reference decides `fl(inter / fl(fl(ai + aj) - inter)) > thr` with an IEEE division.  The division is ~30
// instructions, so the exact quotient (iou_exact) is only formed when the outcome is not already certain; iou_code
// returns bit 0 = the outcome if certain, bit 1 = uncertain.  Rounding is monotone and every fl() is within
// 2^-24 relative, which gives two sufficient tests, selected per launch from a property of the image's boxes:
//
// kMode 0 ("benign": every candidate box is finite with x2 >= x1, y2 >= y1 and an area that is 0 or within
//   [2^-90, 2^90]).  Then 0 <= inter <= min(ai, aj), s = fl(ai + aj) is exactly the reference's sum, u = s - inter > 0
//   whenever s > 0, and  inter / u > t  <=>  inter > s t / (1 + t).  With the host-made constants
//   T_hi >= th / (1 + th) (1 + 2^-22), th = thr (1 + 2^-21), and T_lo <= tl / (1 + tl) (1 - 2^-22), tl = thr (1 - 2^-21):
//     inter > fl(T_hi s)  =>  inter / u >= th  =>  fl(inter / fl(u)) > thr,
//     inter < fl(T_lo s)  =>  inter / u <= tl  =>  fl(inter / fl(u)) <= thr;
//   anything in between (a sliver 2^-20 wide around the threshold, and s == 0) is uncertain.  14 instructions.
// kMode 1 (finite or not, but no NaN coordinate) and kMode 2 (NaN present): compare inter with p = fl(thr u),
//   u = fl(fl(ai + aj) - inter):  inter > fl(p (1 + 2^-20)) => exceeds, inter < fl(p (1 - 2^-20)) => does not, valid
//   when p is a normal positive float (the exponent-range check, which also rejects NaN and u <= 0) and thr is in
//   [1e-6, 1e6] (else the caller passes NaN as t_a and every pair is uncertain).  kMode 2 additionally keeps
//   std::max / std::min's operand order, which differs from FMNMX only on NaN operands (signed zeros change at most
//   the sign of a zero width, never `inter`).
template <int kMode>
__device__ __forceinline__ uint32_t iou_code(const float4& bi, float ai, const float4& bj, float aj, float t_a, float t_b) {
  float xx1, yy1, xx2, yy2;
  if (kMode == 2) {
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
  if (kMode == 0) {
    const float s = __fadd_rn(ai, aj);
    const bool hi = inter > __fmul_rn(t_a, s);
    const bool lo = inter < __fmul_rn(t_b, s);
    return (hi ? 1u : 0u) | ((hi || lo) ? 0u : 2u);
  } else {
    const float u = __fsub_rn(__fadd_rn(ai, aj), inter);
    const float p = __fmul_rn(t_a, u);
    const bool in_range = (__float_as_uint(p) - 0x0D800000u) < 0x64000000u;  // 2^-100 <= p < 2^100
    const bool hi = inter > __fmul_rn(p, 1.00000095367431640625f);           // 1 + 2^-20
    const bool lo = inter < __fmul_rn(p, 0.99999904632568359375f);           // 1 - 2^-20
    return (hi ? 1u : 0u) | ((in_range && (hi || lo)) ? 0u : 2u);
  }
}
// the reference arithmetic itself, for the pairs iou_code could not decide
__device__ __noinline__ bool iou_exact(float4 bi, float ai, float4 bj, float aj, float thr) {
  const float xx1 = std_max(bi.x, bj.x);
  const float yy1 = std_max(bi.y, bj.y);
  const float xx2 = std_min(bi.z, bj.z);
  const float yy2 = std_min(bi.w, bj.w);
  const float w = std_max(0.f, __fsub_rn(xx2, xx1));
  const float h = std_max(0.f, __fsub_rn(yy2, yy1));
  const float inter = __fmul_rn(w, h);
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(ai, aj), inter)) > thr;
}
__device__ __forceinline__ bool iou_decide(uint32_t code, float4 bi, float ai, float4 bj, float aj, float thr) {
  if (code & 2u) return iou_exact(bi, ai, bj, aj, thr);
  return code & 1u;
}

// list entries beyond n: each of the <= 8 x 32 warps rounds its capacity up to whole 64-entry trips
constexpr size_t kListSlack = 8 * kNmsWarps * kChunk + 1024;
constexpr int kSlots = 8;  // ring of kept-box lists; slot = chunk % 8, so a slot always has the same owner CTA

struct NmsSmem {
  uint32_t bin[256];              // digit histogram / running offsets
  uint32_t warp_off[kNmsWarps][256];  // per-warp digit counts, then scatter offsets
  float4 kbox[kSlots][kChunk];    // kept boxes of chunk c live in slot c % 8 of EVERY CTA of the cluster (st.async
  float karea[kSlots][kChunk];    //  by the chunk's owner warp, completion counted on the receiver's full[] barrier)
  uint32_t nk[kSlots];
  unsigned long long full[kSlots];   // 1 arrival (+ the byte count) per use, both made by the sending warp
  unsigned long long empty[kSlots];  // in the slot's owner CTA: one arrival per warp of the cluster when it is done
  int n_valid;
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
__device__ __forceinline__ void mbar_arrive_expect_tx_remote(uint32_t cluster_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr), "r"(bytes) : "memory");
}
// Bounded wait: a barrier bug traps (the host sees a launch error) instead of hanging the GPU.
__device__ __forceinline__ void nms_mbar_wait(uint32_t bar, uint32_t parity) {
  if (sm100::mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!sm100::mbar_try_wait(bar, parity)) {
    __nanosleep(20);
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}

#ifdef UAVDET_NMS_PROFILE
#define NMS_PROF(acc) do { pt = clock64(); acc += pt - pt0; pt0 = pt; } while (0)
#else
#define NMS_PROF(acc) do { } while (0)
#endif

// entries per warp list: the warp's share of the (at most) ceil(n / 64) chunks, in whole 64-entry trips
__host__ __device__ inline int list_capacity(int n, int ncta) {
  const int n_chunks = (n + kChunk - 1) / kChunk;
  return (((n_chunks + ncta - 1) / ncta + kSweepWarps - 1) / kSweepWarps) * kChunk;
}
constexpr size_t kSmemListOffset = (sizeof(NmsSmem) + 15) & ~(size_t)15;

// Step 4 of the kernel below, run by every warp on its own list `lbox/lpos[0..L)` (sorted-position order).
// Returns the number of kept boxes (every warp of the cluster ends with the same count).
template <int kMode>
__device__ __noinline__ int suppress(NmsSmem& S, float4* lbox, uint32_t* lpos, int L, const uint32_t* order,
                                     int64_t* keep, int n_chunks, float thr, float t_a, float t_b, uint32_t ncta,
                                     uint32_t rank) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
#ifdef UAVDET_NMS_PROFILE
  long long pt0 = clock64(), p_wait = 0, p_sweep = 0, p_res = 0, p_hops = 0, p_empty = 0, p_send = 0, p_t0 = 0, pt; int n_trips = 0, n_iters = 0, n_lists = 0; long long sumL = 0, p_k = 0, p_ld = 0;
#endif
  int kept_total = 0;
  // "list -1" is empty: the owner of chunk 0 resolves it straight away
  for (int c = -1; c < n_chunks; ++c) {
    const int slot = c & (kSlots - 1);
    int nk = 0;
    if (c >= 0) {
      nms_mbar_wait(sm100::smem_u32(&S.full[slot]), ((uint32_t)c / kSlots) & 1u);
      nk = (int)S.nk[slot];
      kept_total += nk;
    }
    NMS_PROF(p_wait);
    // does this warp own chunk c + 1?
    const int cn = c + 1;
    const bool owner = cn < n_chunks && (uint32_t)cn % ncta == rank && (((uint32_t)cn / ncta) & (kSweepWarps - 1)) == (uint32_t)warp;
    if (nk != 0 || owner) {
      int out = 0;
      for (int in = 0; in < L || (owner && in == 0); in += kChunk) {
#ifdef UAVDET_NMS_PROFILE
        ++n_trips; if (in == 0) { sumL += L; }
#endif
        // two entries per lane: A = in + lane, B = in + 32 + lane
        const int eA = in + lane, eB = in + 32 + lane;
        bool liveA = eA < L, liveB = eB < L;
        float4 bA = make_float4(0.f, 0.f, 0.f, 0.f), bB = bA;
        uint32_t pA = 0, pB = 0;
        if (liveA) { bA = lbox[eA]; pA = lpos[eA]; }
        if (liveB) { bB = lbox[eB]; pB = lpos[eB]; }
        const float aA = box_area(bA), aB = box_area(bB);
        const uint32_t validA = __ballot_sync(0xffffffffu, liveA), validB = __ballot_sync(0xffffffffu, liveB);
        NMS_PROF(p_ld);
        // sweep against list c, two kept boxes x two candidates in flight
        int k = 0;
        for (; k + 2 <= nk; k += 2) {
          if (!(liveA | liveB)) break;
#ifdef UAVDET_NMS_PROFILE
          ++n_iters;
#endif
          const float4 k0 = S.kbox[slot][k], k1 = S.kbox[slot][k + 1];
          const float a0 = S.karea[slot][k], a1 = S.karea[slot][k + 1];
          const uint32_t cA0 = iou_code<kMode>(k0, a0, bA, aA, t_a, t_b), cA1 = iou_code<kMode>(k1, a1, bA, aA, t_a, t_b);
          const uint32_t cB0 = iou_code<kMode>(k0, a0, bB, aB, t_a, t_b), cB1 = iou_code<kMode>(k1, a1, bB, aB, t_a, t_b);
          bool hA, hB;
          if ((cA0 | cA1 | cB0 | cB1) & 2u) {
#ifdef UAVDET_NMS_PROFILE
            ++n_lists;
#endif
            hA = iou_decide(cA0, k0, a0, bA, aA, thr) || iou_decide(cA1, k1, a1, bA, aA, thr);
            hB = iou_decide(cB0, k0, a0, bB, aB, thr) || iou_decide(cB1, k1, a1, bB, aB, thr);
          } else {
            hA = (cA0 | cA1) & 1u;
            hB = (cB0 | cB1) & 1u;
          }
          liveA = liveA && !hA;
          liveB = liveB && !hB;
        }
        if (k < nk && (liveA | liveB)) {
          const float4 k0 = S.kbox[slot][k];
          const float a0 = S.karea[slot][k];
          const uint32_t cA0 = iou_code<kMode>(k0, a0, bA, aA, t_a, t_b), cB0 = iou_code<kMode>(k0, a0, bB, aB, t_a, t_b);
          liveA = liveA && !iou_decide(cA0, k0, a0, bA, aA, thr);
          liveB = liveB && !iou_decide(cB0, k0, a0, bB, aB, thr);
        }
        __syncwarp();
        if (owner && in == 0) NMS_PROF(p_t0); else NMS_PROF(p_k);
        if (owner && in == 0) {
          // ---- resolve chunk c + 1: its survivors are exactly the live entries of this trip with that chunk number
          const bool memA = liveA && (int)(pA >> 6) == cn, memB = liveB && (int)(pB >> 6) == cn;
          unsigned long long alive =
              (unsigned long long)__ballot_sync(0xffffffffu, memA) | ((unsigned long long)__ballot_sync(0xffffffffu, memB) << 32);
          unsigned long long kept = 0ull;
          // Greedy in rounds: the (up to) four lowest surviving members are tested against all members at once
          // -- one latency for four candidates -- then accepted in order with bit operations only (a later one
          // of the four may have been suppressed by an earlier one; its test results are then simply unused).
          while (alive) {
            int r[4];
            unsigned long long rest = alive;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              r[t] = rest ? __ffsll((long long)rest) - 1 : -1;
              rest &= rest - 1;
            }
            // absent candidates (fewer than four left) re-use lane 0's entry; their results are never looked at
            float4 br[4];
            float ar[4];
            uint32_t cA[4], cB[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const int rr = max(r[t], 0);
              const bool fromB = rr >= 32;
              br[t].x = __shfl_sync(0xffffffffu, fromB ? bB.x : bA.x, rr & 31);
              br[t].y = __shfl_sync(0xffffffffu, fromB ? bB.y : bA.y, rr & 31);
              br[t].z = __shfl_sync(0xffffffffu, fromB ? bB.z : bA.z, rr & 31);
              br[t].w = __shfl_sync(0xffffffffu, fromB ? bB.w : bA.w, rr & 31);
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              ar[t] = box_area(br[t]);
              cA[t] = iou_code<kMode>(br[t], ar[t], bA, aA, t_a, t_b);
              cB[t] = iou_code<kMode>(br[t], ar[t], bB, aB, t_a, t_b);
            }
            if ((cA[0] | cA[1] | cA[2] | cA[3] | cB[0] | cB[1] | cB[2] | cB[3]) & 2u) {
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                cA[t] = iou_decide(cA[t], br[t], ar[t], bA, aA, thr) ? 1u : 0u;
                cB[t] = iou_decide(cB[t], br[t], ar[t], bB, aB, thr) ? 1u : 0u;
              }
            }
            unsigned long long H[4];
#pragma unroll
            for (int t = 0; t < 4; ++t)
              H[t] = (unsigned long long)__ballot_sync(0xffffffffu, memA && (cA[t] & 1u)) |
                     ((unsigned long long)__ballot_sync(0xffffffffu, memB && (cB[t] & 1u)) << 32);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              if (r[t] >= 0 && ((alive >> r[t]) & 1ull)) {
                kept |= 1ull << r[t];
                // a member never suppresses an earlier one that was kept (the test is symmetric), and `alive`
                // no longer holds the kept ones, so clearing every hit is safe; bit r itself goes too
                alive &= ~(H[t] | (1ull << r[t]));
              }
            }
          }
          NMS_PROF(p_hops);
          const int nk_new = __popcll(kept);
          const bool keptA = (kept >> lane) & 1ull, keptB = (kept >> (lane + 32)) & 1ull;
          const int slotA = __popcll(kept & (unsigned long long)lt_mask);
          const int slotB = __popc((uint32_t)kept) + __popc((uint32_t)(kept >> 32) & lt_mask);
          // ---- publish list c + 1: wait until every warp of the cluster is done with the slot's previous content
          const int nslot = cn & (kSlots - 1);
          nms_mbar_wait(sm100::smem_u32(&S.empty[nslot]), (((uint32_t)cn / kSlots) & 1u) ^ 1u);
          NMS_PROF(p_empty);
          for (uint32_t q = 0; q < ncta; ++q) {
            const uint32_t bar = dsmem_addr(&S.full[nslot], q);
            if (keptA) {
              st_async_f4(dsmem_addr(&S.kbox[nslot][slotA], q), bA, bar);
              st_async_u32(dsmem_addr(&S.karea[nslot][slotA], q), __float_as_uint(aA), bar);
            }
            if (keptB) {
              st_async_f4(dsmem_addr(&S.kbox[nslot][slotB], q), bB, bar);
              st_async_u32(dsmem_addr(&S.karea[nslot][slotB], q), __float_as_uint(aB), bar);
            }
            if (lane == 0) {
              st_async_u32(dsmem_addr(&S.nk[nslot], q), (uint32_t)nk_new, bar);
              mbar_arrive_expect_tx_remote(bar, (uint32_t)nk_new * 20u + 4u);
            }
          }
          NMS_PROF(p_send);
          if (keptA) keep[kept_total + slotA] = (int64_t)order[pA];
          if (keptB) keep[kept_total + slotB] = (int64_t)order[pB];
          // the chunk is finished: its members leave the list
          liveA = liveA && !memA;
          liveB = liveB && !memB;
          NMS_PROF(p_res);
        }
        // ---- compact the survivors in place (A entries precede B entries, as in the list)
        const uint32_t mA = __ballot_sync(0xffffffffu, liveA), mB = __ballot_sync(0xffffffffu, liveB);
        if (out != in || mA != validA || mB != validB) {
          const int oA = out + __popc(mA & lt_mask), oB = out + __popc(mA) + __popc(mB & lt_mask);
          if (liveA) { lbox[oA] = bA; lpos[oA] = pA; }
          if (liveB) { lbox[oB] = bB; lpos[oB] = pB; }
        }
        out += __popc(mA) + __popc(mB);
        __syncwarp();  // the stores above are read back by other lanes of this warp in the next sweep
      }
      L = out;
    }
    // this warp is done with list c: tell the slot's owner CTA
    if (c >= 0) {
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(dsmem_addr(&S.empty[slot], (uint32_t)slot % ncta));
    }
    NMS_PROF(p_sweep);
  }
#ifdef UAVDET_NMS_PROFILE
  const int n_iters_max = __reduce_max_sync(0xffffffffu, n_iters); n_lists = __reduce_add_sync(0xffffffffu, n_lists);
  if (lane == 0 && (warp == 0 || warp == 5) && blockIdx.x < ncta && (rank == 0 || rank == 3))
    printf("nms rank %u/%u warp %d: wait %lld sweep %lld | own: trip0 %lld hops %lld empty %lld send %lld keep %lld cycles, chunks %d; slow-path lane-iterations %d trips %d iters(max lane) %d avgL %lld; load %lld kloop %lld\n", rank, ncta, warp, p_wait,
           p_sweep, p_t0, p_hops, p_empty, p_send, p_res, n_chunks, n_lists, n_trips, n_iters_max, sumL / max(n_lists, 1), p_ld, p_k);
#endif
  return kept_total;
}

// One image per thread-block CLUSTER (1, 2, 4 or 8 CTAs); see the file header for the pipeline.
__global__ void __launch_bounds__(kNmsThreads, 1)
nms_image_kernel(const float* __restrict__ boxes, const float* __restrict__ scores, int n,
                 float thr, float score_floor, int64_t* __restrict__ keep,
                 int32_t* __restrict__ keep_count, uint8_t* __restrict__ workspace,
                 size_t ws_per_image, int lists_in_smem, float t_hi, float t_lo) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  NmsSmem& S = *reinterpret_cast<NmsSmem*>(smem_raw);

  const uint32_t ncta = cluster_nctarank();
  const uint32_t rank = cluster_ctarank();
  const int img = blockIdx.x / ncta;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  boxes += (size_t)img * n * 4;
  scores += (size_t)img * n;
  keep += (size_t)img * n;

  // workspace carve-up (per image): keys A/B, vals A/B, the per-warp candidate lists, the box-property flags
  uint8_t* ws = workspace + (size_t)img * ws_per_image;
  const size_t n_pad = ((size_t)n + 3) & ~(size_t)3;
  uint32_t* keyA = reinterpret_cast<uint32_t*>(ws);
  uint32_t* keyB = keyA + n_pad;
  uint32_t* valA = keyB + n_pad;
  uint32_t* valB = valA + n_pad;
  const size_t n_list = n_pad + kListSlack;
  float4* lists_box = reinterpret_cast<float4*>(valB + n_pad);
  uint32_t* lists_pos = reinterpret_cast<uint32_t*>(lists_box + n_list);
  uint32_t* nan_flag = lists_pos + n_list;

#ifdef UAVDET_NMS_PROFILE
  long long pt0 = clock64(), p_sort = 0, p_loop = 0, pt;
#endif
  // ---- 1. keys (CTA 0) and the count of candidates above the floor (every CTA) ------------
  if (tid == 0) {
    S.n_valid = 0;
    for (int q = 0; q < kSlots; ++q) {
      sm100::mbar_init(sm100::smem_u32(&S.full[q]), 1);
      sm100::mbar_init(sm100::smem_u32(&S.empty[q]), ncta * kSweepWarps);
    }
    sm100::fence_barrier_init();
    if (rank == 0) *nan_flag = 0u;
  }
  __syncthreads();
  // With a score floor only the candidates above it (or NaN: they sort first) are ever looked at again, and a detector's
  // floor leaves a few hundred of 96,000 (RTMUAVDet, batch 128: the full-length sort was 2.1 ms of a 19 ms forward).
  // CTA 0 therefore compacts them first — in index order, so the stable sort below still breaks score ties by index —
  // and sorts n_valid entries instead of n.  Two sweeps over the scores, each warp owning one contiguous segment:
  // count, exclusive prefix over the 16 warps, then ballot-ranked writes.
  const bool compact = score_floor > -INFINITY;
  int n_sort = n;
  if (compact && rank == 0) {
    const int seg = (((n + kNmsWarps - 1) / kNmsWarps) + 31) & ~31;
    const int i0 = warp * seg, i1 = min(n, i0 + seg);
    int cnt = 0;
    for (int base = i0; base < i1; base += 128) {
      float sv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + 32 * u + lane;
        sv[u] = i < i1 ? scores[i] : score_floor;     // the floor itself is neither NaN nor above the floor
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) cnt += (sv[u] != sv[u] || sv[u] > score_floor) ? 1 : 0;
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (lane == 0) S.bin[warp] = (uint32_t)cnt;
    __syncthreads();
    uint32_t off = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kNmsWarps; ++w) {
      const uint32_t c = S.bin[w];
      if (w < warp) off += c;
      total += c;
    }
    if (tid == 0) S.n_valid = (int)total;
    n_sort = (int)total;
    const uint32_t lt = (1u << lane) - 1u;
    for (int base = i0; base < i1; base += 128) {
      float sv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + 32 * u + lane;
        sv[u] = i < i1 ? scores[i] : score_floor;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const bool v = sv[u] != sv[u] || sv[u] > score_floor;
        const uint32_t m = __ballot_sync(0xffffffffu, v);
        if (v) {
          const uint32_t pos = off + __popc(m & lt);
          keyA[pos] = score_key(sv[u]);
          valA[pos] = (uint32_t)(base + 32 * u + lane);
        }
        off += __popc(m);
      }
    }
    __syncthreads();   // the sort reads keys other threads wrote, and reuses S.bin
  } else {
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
  }

  // ---- 2. stable LSD radix sort (CTA 0) ---------------------------------------------------
  uint32_t* kin = keyA; uint32_t* kout = keyB;
  uint32_t* vin = valA; uint32_t* vout = valB;
  for (int pass = 0; pass < 4 && rank == 0; ++pass) {
    const int shift = pass * 8;
    if (tid < 256) S.bin[tid] = 0;
    __syncthreads();  // also orders the previous pass' global writes within the CTA
    for (int base = 0; base < n_sort; base += kNmsThreads) {  // warp-aggregated histogram
      const int i = base + tid;
      const uint32_t digit = (i < n_sort) ? ((kin[i] >> shift) & 255u) : 256u;
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
    for (int base = 0; base < n_sort; base += kNmsThreads) {
      const int i = base + tid;
      const bool valid = i < n_sort;
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

  // ---- 3. every warp gathers its own candidates ------------------------------------------------
  // CTA 0's sort result must be visible to the whole cluster (and every CTA must be running before the first
  // distributed-shared-memory access below)
  if (ncta > 1) cluster_sync_all(); else __syncthreads();
  const int nv = S.n_valid;
  const int n_chunks = (nv + kChunk - 1) / kChunk;
  const int cap = list_capacity(n, (int)ncta);  // entries per warp
  float4* lbox = lists_box + (size_t)(rank * kSweepWarps + warp) * cap;
  uint32_t* lpos = lists_pos + (size_t)(rank * kSweepWarps + warp) * cap;
  if (lists_in_smem) {  // the CTA's lists fit behind NmsSmem (decided by the host from n and the cluster size)
    float4* base = reinterpret_cast<float4*>(smem_raw + kSmemListOffset);
    lbox = base + (size_t)warp * cap;
    lpos = reinterpret_cast<uint32_t*>(base + (size_t)kSweepWarps * cap) + (size_t)warp * cap;
  }
  const float4* boxes4 = reinterpret_cast<const float4*>(boxes);
  int L = 0;
  bool nan = false, odd = false;  // odd: not "benign" in the sense of iou_code's kMode 0
  for (int t = 0; warp < kSweepWarps; ++t) {
    const int c = (int)rank + (t * kSweepWarps + warp) * (int)ncta;
    if (c >= n_chunks) break;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int p = c * kChunk + h * 32 + lane;
      if (p < nv) {
        const float4 b = boxes4[order[p]];
        lbox[t * kChunk + h * 32 + lane] = b;
        lpos[t * kChunk + h * 32 + lane] = (uint32_t)p;
        nan |= (b.x != b.x) | (b.y != b.y) | (b.z != b.z) | (b.w != b.w);
        const float a = box_area(b);
        const bool area_ok = a == 0.f || (a >= 0x1p-90f && a <= 0x1p90f);
        const bool finite = fabsf(b.x) < INFINITY && fabsf(b.y) < INFINITY && fabsf(b.z) < INFINITY && fabsf(b.w) < INFINITY;
        odd |= !(finite && b.z >= b.x && b.w >= b.y && area_ok);
      }
    }
    L += min(kChunk, nv - c * kChunk);
  }
  if (__any_sync(0xffffffffu, nan) && lane == 0) atomicOr(nan_flag, 1u);
  if (__any_sync(0xffffffffu, odd) && lane == 0) atomicOr(nan_flag, 2u);
  if (ncta > 1) cluster_sync_all(); else __syncthreads();
  NMS_PROF(p_sort);

  // ---- 4. greedy suppression --------------------------------------------------------------
  const bool thr_ok = thr >= 1e-6f && thr <= 1e6f;
  const float thr_fast = thr_ok ? thr : __int_as_float(0x7fc00000);
  const uint32_t flags = *reinterpret_cast<volatile uint32_t*>(nan_flag);
  int kept_total = 0;
  if (warp >= kSweepWarps) {
  } else if (flags & 1u) {
    kept_total = suppress<2>(S, lbox, lpos, L, order, keep, n_chunks, thr, thr_fast, 0.f, ncta, rank);
  } else if ((flags & 2u) || !thr_ok) {
    kept_total = suppress<1>(S, lbox, lpos, L, order, keep, n_chunks, thr, thr_fast, 0.f, ncta, rank);
  } else {
    kept_total = suppress<0>(S, lbox, lpos, L, order, keep, n_chunks, thr, t_hi, t_lo, ncta, rank);
  }
  if (rank == 0 && tid == 0) keep_count[img] = kept_total;
  // tail of `keep` beyond keep_count is left untouched (caller slices by count)
#ifdef UAVDET_NMS_PROFILE
  NMS_PROF(p_loop);
  if (tid == 0 && img == 0)
    printf("nms rank %u/%u: sort+gather %lld loop %lld cycles, kept %d\n", rank, ncta, p_sort, p_loop, kept_total);
#endif
  // no CTA may leave while another can still arrive on its barriers
  if (ncta > 1) cluster_sync_all();
}

static size_t nms_ws_per_image(int n) {
  size_t n_pad = ((size_t)n + 3) & ~(size_t)3;
  // keys/values ping-pong, the per-warp lists (box + sorted position; every warp's capacity is rounded up to
  // whole 64-entry trips, hence the slack), the has-NaN flag
  size_t bytes = n_pad * 4 * sizeof(uint32_t) + (n_pad + kListSlack) * (sizeof(float4) + sizeof(uint32_t)) + 16;
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
  size_t smem = kSmemListOffset + 16;
  static PerDeviceOnce attr_once;     // the opt-in is per device
  UAVDET_CUDA(attr_once.run([] {
    return cudaFuncSetAttribute(nms_image_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  }));
  // CTAs per image: as many as keep the whole batch co-resident (about 16 clusters of 8 fit on 148 SMs)
  int ncta = 1;
  if (n >= 4096) ncta = batch <= 16 ? 8 : batch <= 32 ? 4 : batch <= 64 ? 2 : 1;
  if (const char* e = getenv("UAVDET_NMS_CLUSTER")) {
    int v = atoi(e);
    if (v == 1 || v == 2 || v == 4 || v == 8) ncta = v;
  }
  // candidate lists in shared memory when the CTA's share fits, else in the workspace
  const size_t list_bytes = (size_t)kSweepWarps * list_capacity(n, ncta) * (sizeof(float4) + sizeof(uint32_t));
  int lists_in_smem = (kSmemListOffset + list_bytes + 16 <= 227 * 1024) ? 1 : 0;
  if (const char* e = getenv("UAVDET_NMS_SMEM_LISTS")) lists_in_smem = lists_in_smem && atoi(e) != 0;
  if (lists_in_smem) smem += list_bytes;
  // iou_code's kMode 0 constants (see there), rounded away from the threshold
  const double th = (double)thr_f * (1.0 + 0x1p-21), tl = (double)thr_f * (1.0 - 0x1p-21);
  const double t_hi_d = th / (1.0 + th) * (1.0 + 0x1p-22), t_lo_d = tl / (1.0 + tl) * (1.0 - 0x1p-22);
  float t_hi = (float)t_hi_d, t_lo = (float)t_lo_d;
  if ((double)t_hi < t_hi_d) t_hi = nextafterf(t_hi, INFINITY);
  if ((double)t_lo > t_lo_d) t_lo = nextafterf(t_lo, -INFINITY);
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
                                 (uint8_t*)workspace, nms_ws_per_image(n), lists_in_smem, t_hi, t_lo));
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}
