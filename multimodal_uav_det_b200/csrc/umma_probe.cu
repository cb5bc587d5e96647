// Hardware probe (debug entry point, not part of the public header): does a K-major SWIZZLE_128B / SWIZZLE_64B UMMA
// operand descriptor accept (a) a start address that is a whole number of rows — not of 8-row swizzle atoms — into a
// TMA-written tile and (b) a stride between 8-row groups (SBO) that is not a multiple of the atom?  Both hold iff the
// tensor core applies the swizzle XOR to the absolute shared-memory address it computes, the way TMA does when it
// writes the tile.  The "halo tile" form of the 3x3 implicit GEMM (one TMA box per pixel tile, nine shifted
// descriptors into it) relies on it.
//
//   A: [rows_total][BK] bf16 row-major in global memory, loaded by ONE TMA box; B: [N = 64][BK] bf16.
//   D[m][n] = sum_k A[shift + (m / 8) * sbo_rows + (m % 8)][k] * B[n][k],  m < 128.
#include "common.cuh"
#include "sm100.cuh"
#include "igemm.h"

namespace uavdet {
using namespace sm100;

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int rows_total, int bk,
                  int shift, int sbo_rows, float* __restrict__ d_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_full, bar_done;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_bytes = bk * 2;
  const int a_bytes = rows_total * row_bytes;
  uint8_t* sa = smem;
  uint8_t* sb = smem + ((a_bytes + 1023) / 1024) * 1024;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar_full), 1);
    mbar_init(smem_u32(&bar_done), 1);
    fence_barrier_init();
  }
  if (warp == 0) { tmem_alloc(smem_u32(&tmem_slot), 64); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(smem_u32(&bar_full), (uint32_t)(a_bytes + 64 * row_bytes));
    tma_load_3d(smem_u32(sa), &mapA, smem_u32(&bar_full), 0, 0, 0);
    tma_load_3d(smem_u32(sb), &mapB, smem_u32(&bar_full), 0, 0, 0);
    while (!mbar_try_wait(smem_u32(&bar_full), 0)) {}
    tc_fence_after();
    const uint32_t layout = bk == 64 ? 2u : 4u;
    const uint32_t idesc = make_idesc_bf16(64, 0, 0);
    const uint64_t ad = make_smem_desc(smem_u32(sa) + (uint32_t)(shift * row_bytes), 16, (uint32_t)(sbo_rows * row_bytes), layout);
    const uint64_t bd = make_smem_desc(smem_u32(sb), 16, (uint32_t)(8 * row_bytes), layout);
    for (int k = 0; k < bk / 16; ++k) tc_mma_bf16(tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, k > 0 ? 1u : 0u);
    tc_commit(smem_u32(&bar_done));
  }
  __syncwarp();
  while (!mbar_try_wait(smem_u32(&bar_done), 0)) {}
  tc_fence_after();
  uint32_t r[32];
  for (int c0 = 0; c0 < 64; c0 += 32) {
    tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) d_out[(warp * 32 + lane) * 64 + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 64); }
}

}  // namespace uavdet

using namespace uavdet;

// a_bf16: [rows_total][bk], b_bf16: [64][bk] (device, 16-byte aligned), d_out: [128][64] fp32 (device)
extern "C" int uavdet_debug_umma_probe(const void* a_bf16, const void* b_bf16, int rows_total, int bk, int shift, int sbo_rows,
                                       float* d_out, void* stream) {
  UAVDET_CHECK_ARG(a_bf16 && b_bf16 && d_out && (bk == 64 || bk == 32) && rows_total <= 256 && shift >= 0 && sbo_rows >= 1 &&
                       shift + 15 * sbo_rows + 8 <= rows_total,
                   "umma_probe: bad arguments");
  CUtensorMap mapA, mapB;
  uint64_t dims[3] = {(uint64_t)bk, (uint64_t)rows_total, 1};
  uint64_t str[2] = {(uint64_t)bk * 2, (uint64_t)bk * 2 * rows_total};
  uint32_t box[3] = {(uint32_t)bk, (uint32_t)rows_total, 1u};
  int rc = encode_tensor_map(&mapA, const_cast<void*>(a_bf16), 3, dims, str, box, bk * 2);
  if (rc) return rc;
  uint64_t dimsb[3] = {(uint64_t)bk, 64, 1};
  uint64_t strb[2] = {(uint64_t)bk * 2, (uint64_t)bk * 2 * 64};
  uint32_t boxb[3] = {(uint32_t)bk, 64u, 1u};
  rc = encode_tensor_map(&mapB, const_cast<void*>(b_bf16), 3, dimsb, strb, boxb, bk * 2);
  if (rc) return rc;
  const int smem = ((rows_total * bk * 2 + 1023) / 1024) * 1024 + 64 * bk * 2 + 1024;
  UAVDET_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  umma_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(mapA, mapB, rows_total, bk, shift, sbo_rows, d_out);
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}
