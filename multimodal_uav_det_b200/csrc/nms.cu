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
// instructions and almost every pair does not overlap at all, so the exact quotient is only formed when the
// outcome is not already certain:
//   * inter == 0 (no overlap): the quotient is 0, -0 or NaN -> never > thr for thr >= 0;
//   * otherwise compare inter with p = fl(thr * u): |fl(x) - x| <= 2^-24 |x| for the product and for the
//     quotient, so inter > p (1 + 2^-20) implies fl(inter / u) > thr and inter < p (1 - 2^-20) implies the
//     opposite; only the sliver in between (and non-finite / denormal operands, thr < 0) takes the division.
__device__ __forceinline__ bool iou_exceeds(const float4& bi, float ai, const float4& bj, float aj,
                                            float thr) {
  float xx1 = std_max(bi.x, bj.x);
  float yy1 = std_max(bi.y, bj.y);
  float xx2 = std_min(bi.z, bj.z);
  float yy2 = std_min(bi.w, bj.w);
  float w = std_max(0.f, __fsub_rn(xx2, xx1));
  float h = std_max(0.f, __fsub_rn(yy2, yy1));
  float inter = __fmul_rn(w, h);
  if (inter == 0.f && thr >= 0.f) return false;
  float u = __fsub_rn(__fadd_rn(ai, aj), inter);
  if (thr > 0.f && inter > 1e-30f && inter < 1e30f && u > 1e-30f && u < 1e30f) {
    const float p = __fmul_rn(thr, u);
    if (inter > __fmul_rn(p, 1.00000095367431640625f)) return true;     // 1 + 2^-20
    if (inter < __fmul_rn(p, 0.99999904632568359375f)) return false;    // 1 - 2^-20
  }
  float ovr = __fdiv_rn(inter, u);
  return ovr > thr;
}

struct NmsSmem {
  uint32_t bin[256];              // digit histogram / running offsets
  uint32_t warp_off[kNmsWarps][256];  // per-warp digit counts, then scatter offsets
  float4 cbox[kChunk];
  float carea[kChunk];
  unsigned long long cmask[kChunk];
  float4 kbox[2][kChunk];         // kept boxes of chunk c live in buffer c & 1 (written by the chunk's owner CTA
  float karea[2][kChunk];         //  into every CTA of the cluster through distributed shared memory)
  int nk[2];
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
__device__ __forceinline__ void dsmem_st_f4(uint32_t addr, const float4& v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void dsmem_st_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
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

  // ---- 1. keys (CTA 0) and the count of candidates above the floor (every CTA) ------------
  if (tid == 0) S.n_valid = 0;
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
  if (rank == 0)
    for (int i = tid; i < nv; i += kNmsThreads) sbox[i] = boxes4[order[i]];
  const int n_words = (nv + 31) >> 5;
  for (int i = tid; i < n_words; i += kNmsThreads) removed[i] = 0;
  // makes CTA 0's sorted boxes / order visible to the other CTAs, and guarantees every CTA of the cluster is
  // running before any distributed-shared-memory store below
  if (ncta > 1) cluster_sync_all(); else __syncthreads();

  // ---- 4. chunked greedy suppression ------------------------------------------------------
  const int n_chunks = (nv + kChunk - 1) / kChunk;
  int kept_total = 0;  // identical in every CTA of the cluster
  for (int c = 0; c < n_chunks; ++c) {
    const int c0 = c * kChunk;
    const int buf = c & 1;
    const uint32_t owner = (uint32_t)c % ncta;
    if (rank == owner) {
      __syncthreads();  // this CTA's sweep of the previous chunk has updated removed[]
      const int cn = min(kChunk, nv - c0);
      const unsigned long long rem_in =
          (unsigned long long)removed[c0 >> 5] |
          ((c0 + 32 < nv) ? ((unsigned long long)removed[(c0 >> 5) + 1] << 32) : 0ull);
      const unsigned long long live_mask = (cn == 64) ? ~0ull : ((1ull << cn) - 1ull);
      if ((~rem_in & live_mask) == 0ull) {  // whole chunk already suppressed (CTA-uniform)
        if (ncta == 1) continue;
        if (tid < (int)ncta) dsmem_st_u32(dsmem_addr(&S.nk[buf], tid), 0u);
      } else {
        if (tid < cn) {
          float4 b = sbox[c0 + tid];
          S.cbox[tid] = b;
          S.carea[tid] = box_area(b);
        }
        __syncthreads();
        // a. pair mask: warp w owns rows 2w and 2w+1; lanes cover columns lane and lane+32
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int r = warp * 2 + rr;
          bool hit0 = false, hit1 = false;
          if (r < cn) {
            const float4 bi = S.cbox[r];
            const float ai = S.carea[r];
            if (lane > r && lane < cn) hit0 = iou_exceeds(bi, ai, S.cbox[lane], S.carea[lane], thr);
            if (lane + 32 > r && lane + 32 < cn)
              hit1 = iou_exceeds(bi, ai, S.cbox[lane + 32], S.carea[lane + 32], thr);
          }
          uint32_t m0 = __ballot_sync(0xffffffffu, hit0);
          uint32_t m1 = __ballot_sync(0xffffffffu, hit1);
          if (lane == 0) S.cmask[r] = (unsigned long long)m0 | ((unsigned long long)m1 << 32);
        }
        __syncthreads();
        // b. serial resolve, redundantly by lane 0 of every warp (no broadcast + barrier afterwards)
        unsigned long long kept = 0ull;
        if (lane == 0) {
          unsigned long long rem = rem_in | ~live_mask;
#pragma unroll 8
          for (int r = 0; r < kChunk; ++r) {
            unsigned long long m = S.cmask[r];
            bool alive = !((rem >> r) & 1ull);
            if (alive) { kept |= (1ull << r); rem |= m; }
          }
        }
        kept = __shfl_sync(0xffffffffu, kept, 0);
        const int nk = __popcll(kept);
        // c. kept boxes -> every CTA's buffer; indices -> keep[]
        {
          const int src = tid & (kChunk - 1);
          const uint32_t q = (uint32_t)tid >> 6;  // destination CTA (1024 / 64 = 16 >= cluster size)
          if (q < ncta && src < cn && ((kept >> src) & 1ull)) {
            const int slot = __popcll(kept & ((1ull << src) - 1ull));
            dsmem_st_f4(dsmem_addr(&S.kbox[buf][slot], q), S.cbox[src]);
            dsmem_st_u32(dsmem_addr(&S.karea[buf][slot], q), __float_as_uint(S.carea[src]));
            if (q == 0) keep[kept_total + slot] = (int64_t)order[c0 + src];
          }
          if (tid < (int)ncta) dsmem_st_u32(dsmem_addr(&S.nk[buf], tid), (uint32_t)nk);
        }
      }
    }
    if (ncta > 1) cluster_sync_all(); else __syncthreads();
    const int nk = S.nk[buf];
    kept_total += nk;
    if (nk == 0) continue;
    // d. suppress this CTA's later candidates: own chunks c' > c, c' == rank (mod ncta); a warp covers one
    //    32-bit word of removed[] per trip and is the only writer of that word during this sweep
    const int first = c + 1 + (int)((rank + ncta - (uint32_t)(c + 1) % ncta) % ncta);
    for (int cc = first + (tid >> 6) * (int)ncta; cc < n_chunks; cc += (kNmsThreads / kChunk) * (int)ncta) {
      const int j = cc * kChunk + (tid & (kChunk - 1));
      const uint32_t word = removed[j >> 5];
      bool hit = false;
      if (j < nv && !((word >> (j & 31)) & 1u)) {
        const float4 bj = sbox[j];
        const float aj = box_area(bj);
        for (int k = 0; k < nk; ++k) {
          if (iou_exceeds(S.kbox[buf][k], S.karea[buf][k], bj, aj, thr)) { hit = true; break; }
        }
      }
      const uint32_t hits = __ballot_sync(0xffffffffu, hit);
      if (lane == 0 && hits) removed[j >> 5] = word | hits;
    }
  }
  if (rank == 0 && tid == 0) keep_count[img] = kept_total;
  // tail of `keep` beyond keep_count is left untouched (caller slices by count)
}

static size_t nms_ws_per_image(int n) {
  size_t n_pad = ((size_t)n + 3) & ~(size_t)3;
  size_t bytes = n_pad * 4 * sizeof(uint32_t) + n_pad * sizeof(float4);
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
  size_t smem = sizeof(NmsSmem) + (((size_t)n + 31) / 32) * 4 + 16;
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
