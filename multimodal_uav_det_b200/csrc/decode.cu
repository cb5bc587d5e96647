// K8 — anchor box decode, fused with the NMS input preparation.
// Replaces YOLOHead.__pred_bbox_decoding + __prepare_nms_preds (reference
// model/_base.py:214-248) and RTMHead.__calculate_bbox_size (model/RTMUAVDet.py:274-291).
// Memory-bound: 20 B in + 20 B out per candidate; one thread per candidate, 128-bit accesses.
#include "common.cuh"

namespace uavdet {

struct Anchors { float w[8]; float h[8]; };

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.f / (1.f + expf(-x)); }

// torchvision box_convert cxcywh->xyxy: x1 = cx - 0.5*w, x2 = cx + 0.5*w (no FMA: the CPU
// oracle evaluates mul then add/sub).
__device__ __forceinline__ float4 cxcywh_to_xyxy(float cx, float cy, float w, float h) {
  float hw = __fmul_rn(0.5f, w), hh = __fmul_rn(0.5f, h);
  return make_float4(__fsub_rn(cx, hw), __fsub_rn(cy, hh), __fadd_rn(cx, hw), __fadd_rn(cy, hh));
}

__global__ void decode_yolo_kernel(const float4* __restrict__ bbox, const float* __restrict__ obj,
                                   int batch, int A, int Sh, int Sw, Anchors anc, int ciou,
                                   float4* __restrict__ boxes, float* __restrict__ scores,
                                   int n_total, int cand_off) {
  const int per_img = A * Sh * Sw;
  const int64_t total = (int64_t)batch * per_img;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / per_img);
    const int r = (int)(i - (int64_t)b * per_img);
    const int a = r / (Sh * Sw);
    const int yx = r - a * (Sh * Sw);
    const int gy = yx / Sw, gx = yx - gy * Sw;
    const float4 t = __ldg(&bbox[i]);
    // _base.py:218-221
    float cx = __fsub_rn(__fmul_rn(sigmoid_acc(t.x), 2.f), 0.5f);
    float cy = __fsub_rn(__fmul_rn(sigmoid_acc(t.y), 2.f), 0.5f);
    float sw = __fmul_rn(sigmoid_acc(t.z), 2.f);
    float sh = __fmul_rn(sigmoid_acc(t.w), 2.f);
    float w = __fmul_rn(sw, sw), h = __fmul_rn(sh, sh);
    if (ciou) {  // _base.py:224-237
      cx = __fadd_rn(cx, (float)gx);
      cy = __fadd_rn(cy, (float)gy);
      w = __fmul_rn(w, anc.w[a]);
      h = __fmul_rn(h, anc.h[a]);
    }
    const int64_t o = (int64_t)b * n_total + cand_off + r;
    boxes[o] = cxcywh_to_xyxy(cx, cy, w, h);
    scores[o] = __ldg(&obj[i]);  // raw logits are the NMS scores (_base.py:203)
  }
}

__global__ void decode_rtm_kernel(const float4* __restrict__ bbox, int batch, int A, int Sh, int Sw,
                                  Anchors anc, float4* __restrict__ out) {
  const int per_img = A * Sh * Sw;
  const int64_t total = (int64_t)batch * per_img;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i % per_img);
    const int a = r / (Sh * Sw);
    const int yx = r - a * (Sh * Sw);
    const int gy = yx / Sw, gx = yx - gy * Sw;
    const float4 t = __ldg(&bbox[i]);
    // RTMUAVDet.py:285-288: (b*2 - 0.5 + grid), (b*2)**2 * anchor
    float px = __fadd_rn(__fsub_rn(__fmul_rn(t.x, 2.f), 0.5f), (float)gx);
    float py = __fadd_rn(__fsub_rn(__fmul_rn(t.y, 2.f), 0.5f), (float)gy);
    float sw = __fmul_rn(t.z, 2.f), sh = __fmul_rn(t.w, 2.f);
    float pw = __fmul_rn(__fmul_rn(sw, sw), anc.w[a]);
    float ph = __fmul_rn(__fmul_rn(sh, sh), anc.h[a]);
    out[i] = make_float4(px, py, pw, ph);
  }
}

__global__ void cxcywh_to_xyxy_kernel(const float4* __restrict__ in, float4* __restrict__ out,
                                      int64_t count) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count;
       i += (int64_t)gridDim.x * blockDim.x) {
    float4 t = __ldg(&in[i]);
    out[i] = cxcywh_to_xyxy(t.x, t.y, t.z, t.w);
  }
}

static int grid_for(int64_t total, int threads) {
  int64_t blocks = ceil_div64(total, threads);
  int64_t cap = (int64_t)kNumSMs * 8;
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace uavdet

using namespace uavdet;

extern "C" int uavdet_decode_yolo(const float* bbox_logits, const float* obj_logits, int batch, int A,
                                  int S_h, int S_w, const float* anchors_scaled_host, int ciou,
                                  float* boxes, float* scores, int n_total, int cand_off,
                                  void* stream) {
  UAVDET_CHECK_ARG(A > 0 && A <= 8, "decode_yolo: A=%d unsupported (1..8)", A);
  UAVDET_CHECK_ARG(batch >= 0 && S_h > 0 && S_w > 0, "decode_yolo: bad sizes");
  UAVDET_CHECK_ARG(cand_off >= 0 && cand_off + A * S_h * S_w <= n_total,
                   "decode_yolo: candidate slice out of range");
  UAVDET_CHECK_ARG((((uintptr_t)bbox_logits | (uintptr_t)boxes) & 15) == 0,
                   "decode_yolo: 16-byte alignment required");
  if (batch == 0) return UAVDET_OK;
  Anchors anc{};
  for (int a = 0; a < A; ++a) { anc.w[a] = anchors_scaled_host[2 * a]; anc.h[a] = anchors_scaled_host[2 * a + 1]; }
  int64_t total = (int64_t)batch * A * S_h * S_w;
  decode_yolo_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      (const float4*)bbox_logits, obj_logits, batch, A, S_h, S_w, anc, ciou, (float4*)boxes, scores,
      n_total, cand_off);
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}

extern "C" int uavdet_decode_rtm(const float* bbox_sig, int batch, int A, int S_h, int S_w,
                                 const float* anchors_host, float* bbox_out, void* stream) {
  UAVDET_CHECK_ARG(A > 0 && A <= 8, "decode_rtm: A=%d unsupported (1..8)", A);
  UAVDET_CHECK_ARG((((uintptr_t)bbox_sig | (uintptr_t)bbox_out) & 15) == 0,
                   "decode_rtm: 16-byte alignment required");
  if (batch == 0) return UAVDET_OK;
  Anchors anc{};
  for (int a = 0; a < A; ++a) { anc.w[a] = anchors_host[2 * a]; anc.h[a] = anchors_host[2 * a + 1]; }
  int64_t total = (int64_t)batch * A * S_h * S_w;
  decode_rtm_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      (const float4*)bbox_sig, batch, A, S_h, S_w, anc, (float4*)bbox_out);
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}

extern "C" int uavdet_cxcywh_to_xyxy(const float* in, float* out, int64_t count, void* stream) {
  UAVDET_CHECK_ARG((((uintptr_t)in | (uintptr_t)out) & 15) == 0, "cxcywh_to_xyxy: alignment");
  if (count <= 0) return UAVDET_OK;
  cxcywh_to_xyxy_kernel<<<grid_for(count, 256), 256, 0, (cudaStream_t)stream>>>(
      (const float4*)in, (float4*)out, count);
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}
