// Shared between igemm.cu (fwd / dgrad) and wgrad.cu.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace uavdet {

constexpr int kMaxTaps = 36;   // 3x3 taps x 4 space-to-depth blocks, or 5x5
constexpr int kMaxStages = 16;   // thin layers (12 KB stages) need many in flight to cover the memory latency
constexpr int kMaxKb = 72;       // k-blocks (tap x channel chunk) of a halo-mode tile
constexpr int kMaxHaloBufs = 4;

// One filter tap = one shifted TMA box of the activation tensor map.
struct ConvTap {
  int c_off;   // coordinate offset in dim0 (channel; + parity-pixel * ld in the parity view)
  int dw;      // offset in dim1 (output-column units)
  int p;       // dim2 (row parity) coordinate
  int dh;      // offset in dim3 (output-row units)
  int w_koff;  // K offset of this tap inside the packed weight matrix
};

// n / d by multiply-high: q = umulhi(n, mul) >> shr (d > 1), exact for 0 <= n < 2^31.
struct FastDiv {
  uint32_t d, mul, shr;
};
inline FastDiv make_fast_div(int d) {
  FastDiv f{(uint32_t)d, 0u, 0u};
  if (d > 1) {
    int lg = 0;
    while ((1u << lg) < (uint32_t)d) ++lg;
    const int p = 31 + lg;
    f.mul = (uint32_t)(((1ull << p) + (uint64_t)d - 1) / (uint64_t)d);
    f.shr = (uint32_t)(p - 32);
  }
  return f;
}

struct IgemmParams {
  int n_img, ho, wo;
  int tile_w, tile_h, tiles_w, tiles_h;
  int cout, block_n, n_tiles;
  int num_taps, kc_per_tap, block_k;
  int w_batch, stages, total_tiles;
  int total_pairs;      // CTA-pair kernel: (pairs of consecutive pixel tiles) x channel blocks
  int slab_w;           // columns per TMA-store slab of the bf16 epilogues (64 | 32)
  int epi_mode;         // 0 = CTA-wide slab, 1 = per-warp rectangle (5-D map), 2 = per-warp pixel run (flat 3-D map)
  int epi_bufs;         // staging buffers per epilogue warp (modes 1/2)
  int staging_bytes;    // shared memory reserved for output staging
  int epc_floats;       // > 0: the AFFINE epilogue's scale / shift vectors (padded cout floats each) are staged in shared
                        // memory behind the barriers (one weight-independent copy per CTA) and read from there
  int prod_warps;       // active TMA producer warps (1 | 2 | 4), divides `stages`
  int bres_bytes;       // > 0: the whole weight matrix stays resident in shared memory (bytes); stages hold A only
  // Halo mode (tile 8 wide x 16 tall): ONE TMA box per tile and channel sub-block brings every input pixel any filter
  // tap of the tile touches; tap t reads it through a UMMA descriptor whose start is shifted by whole pixel rows and
  // whose 8-row group stride (SBO) is one halo line — the tensor core applies the swizzle to the absolute address, as
  // TMA did when it wrote the box (probed: umma_probe.cu).  Stages then carry B only.
  int halo;             // 0 | 1
  int halo_bufs;        // A buffers in flight (2..kMaxHaloBufs)
  int halo_buf_bytes;   // bytes per buffer = halo_subs * halo_sub_bytes
  int halo_subs;        // boxes per tile (channel sub-blocks of <= 64 channels)
  int halo_sub_bytes;   // 1024-aligned bytes of one box
  int halo_tx;          // bytes one box delivers
  int halo_row_bytes;   // 64 | 128 (= swizzle span of the box)
  int halo_sbo;         // bytes between the halo rows of consecutive tile rows
  int halo_box_c;       // channels per box
  int halo_w0, halo_p0, halo_h0;   // box origin relative to the tile origin (dims 1, 2, 3)
  uint32_t kb_aoff[kMaxKb];        // per k-block: (byte offset of its first row inside the A buffer) >> 4
  FastDiv fd_n, fd_w, fd_h;   // dividers for the tile decode (n_tiles, tiles_w, tiles_h)
  int epi, act;
  const float* scale;
  const float* shift;
  long long shift_sn;   // per-image stride of `shift` (0 = shared)
  const float* sample_affine;   // GroupNorm-fold epilogue: [n][2] = (rstd, mean * rstd) per image, or NULL
  int sa_staged;                // the pairs are copied to shared memory behind the epilogue constants
  int res_tma;          // 1: the warp-private epilogues TMA-load the residual tile into their staging buffer
  const __nv_bfloat16* res;
  long long res_sn, res_sh, res_sw;
  __nv_bfloat16* out;
  long long out_sn, out_sh, out_sw;
  // Fused parity planes (data gradient through the space-to-depth gather): the N columns are [row parity][2c] and the
  // output / residual maps use their third dimension for the row parity: column n -> (plane n / out_cspan, channel
  // n % out_cspan of the pixel-pair row).  0 = ordinary output.
  int out_cspan;
  long long out_sp, res_sp;     // element stride between the row-parity planes
  int stat_mod;                 // > 0: STATS column n adds to channel n % stat_mod (pixel-pair GEMMs: N = [pixel][channel])
  float* sum;
  float* sumsq;
  float* head_obj;
  float* head_bbox;
  int head_anchors;
  unsigned int* watchdog;
  long long* trace;     // debug: per-tile role timestamps of CTA 0 (NULL = off), 8 slots per tile
  int trace_tiles;
  ConvTap taps[kMaxTaps];
};

int encode_tensor_map(CUtensorMap* m, void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes);
int make_act_map(CUtensorMap* m, const uavdet_act* x, int parity, int box_c, int box_w, int box_h, int box_p = 1);
int make_out_map(CUtensorMap* m, void* ptr, int n, int ho, int wo, int c, long long sn, long long sh, long long sw,
                 int box_c, int box_w, int box_h, int planes = 1, long long sp = 0);
void choose_tile(int ho, int wo, bool dense_rows, int* tile_w, int* tile_h, int* epi_mode);
int fill_plane(const IgemmParams& P, cudaStream_t st);
void get_trace(long long** ptr, int* tiles);

}  // namespace uavdet
