// Shared between igemm.cu (fwd / dgrad) and wgrad.cu.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace uavdet {

constexpr int kMaxTaps = 36;   // 3x3 taps x 4 space-to-depth blocks, or 5x5
constexpr int kMaxStages = 8;

// One filter tap = one shifted TMA box of the activation tensor map.
struct ConvTap {
  int c_off;   // coordinate offset in dim0 (channel; + parity-pixel * ld in the parity view)
  int dw;      // offset in dim1 (output-column units)
  int p;       // dim2 (row parity) coordinate
  int dh;      // offset in dim3 (output-row units)
  int w_koff;  // K offset of this tap inside the packed weight matrix
};

struct IgemmParams {
  int n_img, ho, wo;
  int tile_w, tile_h, tiles_w, tiles_h;
  int cout, block_n, n_tiles;
  int num_taps, kc_per_tap, block_k;
  int w_batch, stages, total_tiles;
  int epi, act;
  const float* scale;
  const float* shift;
  long long shift_sn;   // per-image stride of `shift` (0 = shared)
  const __nv_bfloat16* res;
  long long res_sn, res_sh, res_sw;
  __nv_bfloat16* out;
  long long out_sn, out_sh, out_sw;
  float* sum;
  float* sumsq;
  float* head_obj;
  float* head_bbox;
  int head_anchors;
  unsigned int* watchdog;
  ConvTap taps[kMaxTaps];
};

int encode_tensor_map(CUtensorMap* m, void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes);
int make_act_map(CUtensorMap* m, const uavdet_act* x, int parity, int box_c, int box_w, int box_h);
void choose_tile(int ho, int wo, int* tile_w, int* tile_h);
int fill_plane(const IgemmParams& P, cudaStream_t st);

}  // namespace uavdet
