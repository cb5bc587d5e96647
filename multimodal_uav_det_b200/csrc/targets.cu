// K10 — YOLO target encoder on the GPU.
// Replaces AntiUAVDataset.__generate_yolo_bboxes (reference dataset/AntiUAVDataset.py:141-185) and
// calculate_anchor_iou (dataset/_helper.py:308-330) for a whole batch: one target box per frame (the Anti-UAV
// data has exactly one, AntiUAVDataset.py:52-53) -> dense (B, A, S, S, 5) [obj, cx_off, cy_off, w_cells, h_cells]
// per head.  The reference builds these on the CPU in the data loader, 25,200 x 5 floats per frame of which at
// most 9 x 5 are non-zero, and ships them over PCIe; here the host sends 16 bytes per frame and the dense tensors
// are a memset plus one scatter thread per (frame, head).
// All arithmetic is fp32 with explicit round-to-nearest intrinsics in the reference's operation order, so the
// result is bit-identical to the CPU encoder.
#include "common.cuh"

namespace uavdet {

constexpr int kMaxTargetHeads = 4;
constexpr int kMaxTargetAnchors = 8;

struct TargetHeads {
  float* out[kMaxTargetHeads];
  int grid[kMaxTargetHeads];
  float aw[kMaxTargetHeads][kMaxTargetAnchors];  // anchors / input_size (AntiUAVDataset.py:27)
  float ah[kMaxTargetHeads][kMaxTargetAnchors];
};

__global__ void encode_targets_kernel(const float4* __restrict__ boxes, const uint8_t* __restrict__ valid, int batch,
                                      int heads, int A, float input_size, TargetHeads H,
                                      unsigned int* __restrict__ out_of_grid) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch * heads) return;
  const int b = i / heads, hd = i - b * heads;
  if (valid && !valid[b]) return;  // `if bbox.numel() == 0: return []` (:142-143)
  const float4 box = boxes[b];
  // box_convert xyxy -> cxcywh (torchvision), then `bbox /= input_size` (:148-149)
  const float cx = __fdiv_rn(__fdiv_rn(__fadd_rn(box.x, box.z), 2.f), input_size);
  const float cy = __fdiv_rn(__fdiv_rn(__fadd_rn(box.y, box.w), 2.f), input_size);
  const float w = __fdiv_rn(__fsub_rn(box.z, box.x), input_size);
  const float h = __fdiv_rn(__fsub_rn(box.w, box.y), input_size);
  const int S = H.grid[hd];
  const float fs = (float)S;
  // grid cell and offsets (:156-162)
  const float gcx = __fmul_rn(cx, fs), gcy = __fmul_rn(cy, fs);
  const int gx = (int)gcx, gy = (int)gcy;  // int(): truncation toward zero
  if (gx < 0 || gx >= S || gy < 0 || gy >= S) {  // the reference raises IndexError here
    atomicAdd(out_of_grid, 1u);
    return;
  }
  const float cell[4] = {__fsub_rn(gcx, (float)gx), __fsub_rn(gcy, (float)gy), __fmul_rn(w, fs), __fmul_rn(h, fs)};
  // calculate_anchor_iou (_helper.py:308-330)
  float iou[kMaxTargetAnchors];
  int order[kMaxTargetAnchors];
  const float t_area = __fmul_rn(w, h);
  for (int a = 0; a < A; ++a) {
    const float aw = H.aw[hd][a], ah = H.ah[hd][a];
    const float inter = __fmul_rn(fminf(aw, w), fminf(ah, h));
    const float uni = __fsub_rn(__fadd_rn(__fmul_rn(aw, ah), t_area), inter);
    iou[a] = __fdiv_rn(inter, uni);
    order[a] = a;
  }
  // argsort descending; equal IoUs keep the lower anchor index first
  for (int p = 1; p < A; ++p) {
    const int o = order[p];
    int q = p;
    while (q > 0 && iou[order[q - 1]] < iou[o]) { order[q] = order[q - 1]; --q; }
    order[q] = o;
  }
  float* out = H.out[hd] + (size_t)b * A * S * S * 5;
  auto write = [&](int a, float obj) {
    float* p = out + (((size_t)a * S + gy) * S + gx) * 5;
    p[0] = obj;
    p[1] = cell[0]; p[2] = cell[1]; p[3] = cell[2]; p[4] = cell[3];
  };
  if (iou[order[0]] < 0.5f) {  // only the best anchor (:166-169)
    write(order[0], 1.f);
  } else {  // every anchor gets the box; objectness 1 where its IoU >= 0.5 (:170-179)
    for (int p = 0; p < A; ++p) write(order[p], iou[order[p]] >= 0.5f ? 1.f : 0.f);
  }
}

}  // namespace uavdet

using namespace uavdet;

extern "C" int uavdet_encode_targets(const float* boxes_xyxy, const uint8_t* valid, int batch,
                                     const float* anchors_norm_host, int heads, int num_anchors,
                                     const int* grids_host, float input_size, float* const* targets_host,
                                     unsigned int* out_of_grid, void* stream) {
  UAVDET_CHECK_ARG(batch >= 0 && heads > 0 && heads <= kMaxTargetHeads && num_anchors > 0 &&
                       num_anchors <= kMaxTargetAnchors,
                   "encode_targets: heads <= %d, anchors <= %d", kMaxTargetHeads, kMaxTargetAnchors);
  UAVDET_CHECK_ARG(anchors_norm_host && grids_host && targets_host && out_of_grid, "encode_targets: null pointer");
  if (batch == 0) return UAVDET_OK;
  UAVDET_CHECK_ARG(boxes_xyxy && ((uintptr_t)boxes_xyxy & 15) == 0, "encode_targets: boxes must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  TargetHeads H = {};
  for (int hd = 0; hd < heads; ++hd) {
    UAVDET_CHECK_ARG(targets_host[hd] && grids_host[hd] > 0, "encode_targets: bad head %d", hd);
    H.out[hd] = targets_host[hd];
    H.grid[hd] = grids_host[hd];
    for (int a = 0; a < num_anchors; ++a) {
      H.aw[hd][a] = anchors_norm_host[(hd * num_anchors + a) * 2 + 0];
      H.ah[hd][a] = anchors_norm_host[(hd * num_anchors + a) * 2 + 1];
    }
    const size_t bytes = (size_t)batch * num_anchors * grids_host[hd] * grids_host[hd] * 5 * sizeof(float);
    UAVDET_CUDA(cudaMemsetAsync(targets_host[hd], 0, bytes, st));
  }
  UAVDET_CUDA(cudaMemsetAsync(out_of_grid, 0, sizeof(unsigned int), st));
  const int total = batch * heads;
  encode_targets_kernel<<<(total + 127) / 128, 128, 0, st>>>(reinterpret_cast<const float4*>(boxes_xyxy), valid, batch,
                                                              heads, num_anchors, input_size, H, out_of_grid);
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}
