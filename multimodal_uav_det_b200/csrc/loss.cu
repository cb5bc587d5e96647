// YOLO head loss, forward and gradient in three launches per head scale
// (reference model/_base.py:155-192 YOLOHead.compute_metrics loop body, utils/metrics.py:8-84 bbox_loss /
// objectness_loss / no_obj_loss, utils/postprocess.py:51-85 calculate_iou, model/_base.py:214-270 decode and
// target rewrite; torchvision.ops.complete_box_iou_loss for the 'ciou' branch).
//
// The reference walks (sample, head) pairs in Python with boolean-mask indexing (a device sync each); the torch
// restatement in utils/metrics.py is ~100 small launches per head and direction.  Here:
//   1. loss_prepass_kernel  — per sample: number of positive cells and the FIRST positive (a,h,w) cell
//                             (calculate_iou compares every positive with the first target only, :83-85);
//   2. loss_main_kernel     — per cell: decode, IoU vs the first target, CIoU / MSE box term, the two BCE terms,
//                             their analytic gradients w.r.t. the logits, the rewritten target boxes;
//   3. loss_finalize_kernel — per-sample means -> the head's (bbox, objectness) loss sums.
// All arithmetic fp32.  Gradients are "unit" gradients (dL/d(bbox_sum) = dL/d(obj_sum) = 1); the caller scales.
#include "common.cuh"

namespace uavdet {

struct LossParams {
  const float* p_bbox;   // (B,A,H,W,4) logits
  const float* p_obj;    // (B,A,H,W,1) logits
  const float* tgt;      // (B,A,H,W,5) [obj, cx, cy, w, h]
  int B, A, H, W;
  float aw[8], ah[8];    // anchors / head scale
  int ciou;
  float bbox_w, objectness_w, obj_scale_w, no_obj_w;
  float* d_bbox;         // (B,A,H,W,4)
  float* d_obj;          // (B,A,H,W,1)
  float* new_t;          // (B,A,H,W,4) or NULL
  int* first;            // [B]
  float* npos;           // [B]
  float* acc;            // [B][3]
  float* out;            // [2]
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
// BCE-with-logits, the numerically stable form torch uses: max(x,0) - x*t + log1p(exp(-|x|))
__device__ __forceinline__ float bce_logits(float x, float t) { return fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x))); }

// Number of positive cells and index of the first one per sample.  blockIdx.x = sample, blockIdx.y = slice of its cells
// (one block per sample left 64 blocks to scan 6 MB each on DySOEM's 320x320 head); slices combine with one atomicMin /
// atomicAdd each into workspace words initialised by the host (first = 0x7f7f7f7f, npos = 0: cudaMemsetAsync).
__global__ void __launch_bounds__(256) loss_prepass_kernel(LossParams P) {
  __shared__ int s_first;
  __shared__ int s_cnt;
  const int b = blockIdx.x;
  const int cells = P.A * P.H * P.W;
  if (threadIdx.x == 0) { s_first = 0x7f7f7f7f; s_cnt = 0; }
  __syncthreads();
  const float* t = P.tgt + (long long)b * cells * 5;
  int my_first = 0x7f7f7f7f, my_cnt = 0;
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < cells; i += gridDim.y * blockDim.x) {
    if (__ldg(t + (long long)i * 5) == 1.0f) { ++my_cnt; if (i < my_first) my_first = i; }
  }
  if (my_cnt) { atomicAdd(&s_cnt, my_cnt); atomicMin(&s_first, my_first); }
  __syncthreads();
  if (threadIdx.x == 0 && s_cnt) {
    atomicAdd(P.npos + b, (float)s_cnt);          // exact: counts are small integers
    atomicMin(P.first + b, s_first);
  }
}

struct Box4 { float x1, y1, x2, y2; };
__device__ __forceinline__ Box4 to_xyxy(float cx, float cy, float w, float h) {
  return Box4{cx - 0.5f * w, cy - 0.5f * h, cx + 0.5f * w, cy + 0.5f * h};
}

// torchvision complete_box_iou_loss (eps 1e-7) and its gradient w.r.t. the predicted box b1 (alpha detached).
__device__ __forceinline__ float ciou_loss_grad(const Box4& p, const Box4& g, Box4& d) {
  const float eps = 1e-7f;
  const float ix1 = fmaxf(p.x1, g.x1), iy1 = fmaxf(p.y1, g.y1), ix2 = fminf(p.x2, g.x2), iy2 = fminf(p.y2, g.y2);
  const bool overlap = (iy2 > iy1) && (ix2 > ix1);
  const float iw = ix2 - ix1, ih = iy2 - iy1;
  const float inter = overlap ? iw * ih : 0.f;
  const float w = p.x2 - p.x1, h = p.y2 - p.y1, wg = g.x2 - g.x1, hg = g.y2 - g.y1;
  const float uni = w * h + wg * hg - inter + eps;
  const float iou = inter / uni;
  const float ex1 = fminf(p.x1, g.x1), ey1 = fminf(p.y1, g.y1), ex2 = fmaxf(p.x2, g.x2), ey2 = fmaxf(p.y2, g.y2);
  const float ew = ex2 - ex1, eh = ey2 - ey1;
  const float diag = ew * ew + eh * eh + eps;
  const float dcx = 0.5f * (p.x1 + p.x2) - 0.5f * (g.x1 + g.x2), dcy = 0.5f * (p.y1 + p.y2) - 0.5f * (g.y1 + g.y2);
  const float centre = dcx * dcx + dcy * dcy;
  const float c4 = 0.40528473456935109f;   // 4 / pi^2
  const float dth = atanf(wg / hg) - atanf(w / h);
  const float v = c4 * dth * dth;
  const float alpha = v / (1.f - iou + v + eps);
  const float loss = 1.f - iou + centre / diag + alpha * v;
  // ---- gradient ----
  // intersection (torch.max / torch.min pick the larger / smaller operand; ties have measure zero)
  float in_x1 = 0.f, in_y1 = 0.f, in_x2 = 0.f, in_y2 = 0.f;
  if (overlap) {
    in_x1 = p.x1 > g.x1 ? -ih : 0.f;
    in_x2 = p.x2 < g.x2 ? ih : 0.f;
    in_y1 = p.y1 > g.y1 ? -iw : 0.f;
    in_y2 = p.y2 < g.y2 ? iw : 0.f;
  }
  // union' = area_p' - inter'
  const float un_x1 = -h - in_x1, un_x2 = h - in_x2, un_y1 = -w - in_y1, un_y2 = w - in_y2;
  const float inv_u2 = 1.f / (uni * uni);
  const float iou_x1 = (in_x1 * uni - inter * un_x1) * inv_u2, iou_x2 = (in_x2 * uni - inter * un_x2) * inv_u2;
  const float iou_y1 = (in_y1 * uni - inter * un_y1) * inv_u2, iou_y2 = (in_y2 * uni - inter * un_y2) * inv_u2;
  // centre / diag
  const float dg_x1 = p.x1 < g.x1 ? -2.f * ew : 0.f, dg_x2 = p.x2 > g.x2 ? 2.f * ew : 0.f;
  const float dg_y1 = p.y1 < g.y1 ? -2.f * eh : 0.f, dg_y2 = p.y2 > g.y2 ? 2.f * eh : 0.f;
  const float inv_d2 = 1.f / (diag * diag);
  const float cd_x1 = (dcx * diag - centre * dg_x1) * inv_d2, cd_x2 = (dcx * diag - centre * dg_x2) * inv_d2;
  const float cd_y1 = (dcy * diag - centre * dg_y1) * inv_d2, cd_y2 = (dcy * diag - centre * dg_y2) * inv_d2;
  // aspect term: v = c4 (atan(wg/hg) - atan(w/h))^2
  const float wh2 = w * w + h * h;
  const float v_w = -2.f * c4 * dth * (h / wh2), v_h = -2.f * c4 * dth * (-w / wh2);
  d.x1 = -iou_x1 + cd_x1 + alpha * (-v_w);
  d.x2 = -iou_x2 + cd_x2 + alpha * v_w;
  d.y1 = -iou_y1 + cd_y1 + alpha * (-v_h);
  d.y2 = -iou_y2 + cd_y2 + alpha * v_h;
  return loss;
}

__global__ void __launch_bounds__(256) loss_main_kernel(LossParams P) {
  const int cells = P.A * P.H * P.W;
  const long long total = (long long)P.B * cells;
  const int lane = threadIdx.x & 31;
  for (long long i0 = (long long)blockIdx.x * blockDim.x; i0 < total; i0 += (long long)gridDim.x * blockDim.x) {
    const long long i = i0 + threadIdx.x;
    const bool live = i < total;
    int b = -1;
    float s_box = 0.f, s_pos = 0.f, s_neg = 0.f;
    if (live) {
      b = (int)(i / cells);
      const int r = (int)(i - (long long)b * cells);
      const int a = r / (P.H * P.W);
      const int yx = r - a * P.H * P.W;
      const int gy = yx / P.W, gx = yx - gy * P.W;
      const float* t = P.tgt + i * 5;
      const float t_obj = t[0], tx = t[1], ty = t[2], tw = t[3], th = t[4];
      const bool cell = t_obj == 1.0f;
      const float np = P.npos[b];
      const float logit = P.p_obj[i];
      const float aw = P.aw[a], ah = P.ah[a];
      // rewritten target box (__build_target_bbox, _base.py:250-270)
      float nt0, nt1, nt2, nt3;
      if (P.ciou) { nt0 = tx + (float)gx; nt1 = ty + (float)gy; nt2 = tw; nt3 = th; }
      else { nt0 = tx; nt1 = ty; nt2 = sqrtf((1e-16f + tw) / aw) * 0.5f; nt3 = sqrtf((1e-16f + th) / ah) * 0.5f; }
      if (P.new_t) reinterpret_cast<float4*>(P.new_t)[i] = make_float4(nt0, nt1, nt2, nt3);
      float4 db = make_float4(0.f, 0.f, 0.f, 0.f);
      float pos_t = 0.f;
      if (cell || t_obj != 0.f) {
        const float4 lg = reinterpret_cast<const float4*>(P.p_bbox)[i];
        const float s0 = sigmoidf_(lg.x), s1 = sigmoidf_(lg.y), s2 = sigmoidf_(lg.z), s3 = sigmoidf_(lg.w);
        float cx = 2.f * s0 - 0.5f, cy = 2.f * s1 - 0.5f;
        float w = (2.f * s2) * (2.f * s2), h = (2.f * s3) * (2.f * s3);
        if (P.ciou) { cx += (float)gx; cy += (float)gy; w *= aw; h *= ah; }
        // calculate_iou: IoU with the sample's FIRST positive target as stored (before the rewrite)
        {
          const int first = P.first[b] < cells ? P.first[b] : 0;   // torch.argmax of an all-false mask is 0
          const float* t0 = P.tgt + ((long long)b * cells + first) * 5;
          const Box4 tb = to_xyxy(t0[1], t0[2], t0[3], t0[4]);
          const Box4 pb = P.ciou ? to_xyxy(cx, cy, w, h) : to_xyxy(cx, cy, w * aw, h * ah);
          const float lx = fmaxf(pb.x1, tb.x1), ly = fmaxf(pb.y1, tb.y1), rx = fminf(pb.x2, tb.x2), ry = fminf(pb.y2, tb.y2);
          const float iw = fmaxf(rx - lx, 0.f), ih = fmaxf(ry - ly, 0.f);
          const float inter = iw * ih;
          const float ap = (pb.x2 - pb.x1) * (pb.y2 - pb.y1), at = (tb.x2 - tb.x1) * (tb.y2 - tb.y1);
          pos_t = inter / (ap + at - inter) * t_obj;
        }
        if (cell) {
          // derivative of the decode w.r.t. the logits
          const float dcx = 2.f * s0 * (1.f - s0), dcy = 2.f * s1 * (1.f - s1);
          float dw = 8.f * s2 * s2 * (1.f - s2), dh = 8.f * s3 * s3 * (1.f - s3);
          if (P.ciou) { dw *= aw; dh *= ah; }
          float g_cx, g_cy, g_w, g_h;
          if (P.ciou) {
            Box4 gr;
            s_box = ciou_loss_grad(to_xyxy(cx, cy, w, h), to_xyxy(nt0, nt1, nt2, nt3), gr);
            g_cx = gr.x1 + gr.x2; g_cy = gr.y1 + gr.y2;
            g_w = 0.5f * (gr.x2 - gr.x1); g_h = 0.5f * (gr.y2 - gr.y1);
            const float k = P.bbox_w / np;
            db = make_float4(k * g_cx * dcx, k * g_cy * dcy, k * g_w * dw, k * g_h * dh);
          } else {
            const float e0 = cx - nt0, e1 = cy - nt1, e2 = w - nt2, e3 = h - nt3;
            s_box = e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3;
            const float k = P.bbox_w * 2.f / (4.f * np);
            db = make_float4(k * e0 * dcx, k * e1 * dcy, k * e2 * dw, k * e3 * dh);
          }
        }
      }
      reinterpret_cast<float4*>(P.d_bbox)[i] = db;
      const float sg = sigmoidf_(logit);
      if (cell) {
        s_pos = bce_logits(logit, pos_t);
        P.d_obj[i] = P.objectness_w * P.obj_scale_w * (sg - pos_t) / np;
      } else {
        s_neg = bce_logits(logit, t_obj);
        P.d_obj[i] = P.no_obj_w * (sg - t_obj) / ((float)cells - np);
      }
    }
    // per-sample partial sums: whole-warp butterfly when the warp lies inside one sample, else per-lane atomics
    const unsigned peers = __match_any_sync(0xffffffffu, b);
    if (peers == 0xffffffffu) {
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) {
        s_box += __shfl_xor_sync(0xffffffffu, s_box, off);
        s_pos += __shfl_xor_sync(0xffffffffu, s_pos, off);
        s_neg += __shfl_xor_sync(0xffffffffu, s_neg, off);
      }
      if (lane == 0 && b >= 0) {
        if (s_box != 0.f) atomicAdd(P.acc + b * 3 + 0, s_box);
        if (s_pos != 0.f) atomicAdd(P.acc + b * 3 + 1, s_pos);
        atomicAdd(P.acc + b * 3 + 2, s_neg);
      }
    } else if (b >= 0) {
      if (s_box != 0.f) atomicAdd(P.acc + b * 3 + 0, s_box);
      if (s_pos != 0.f) atomicAdd(P.acc + b * 3 + 1, s_pos);
      if (s_neg != 0.f) atomicAdd(P.acc + b * 3 + 2, s_neg);
    }
  }
}

__global__ void loss_finalize_kernel(LossParams P) {
  // one warp: lane-strided over the samples, then a butterfly
  const int cells = P.A * P.H * P.W;
  float bl = 0.f, ol = 0.f;
  for (int b = threadIdx.x; b < P.B; b += 32) {
    const float np = P.npos[b];
    const float box = P.acc[b * 3 + 0], pos = P.acc[b * 3 + 1], neg = P.acc[b * 3 + 2];
    bl += P.ciou ? box / np : box / (4.f * np);                 // mean over an empty set is NaN, like the reference
    ol += P.objectness_w * P.obj_scale_w * (pos / np) + P.no_obj_w * (neg / ((float)cells - np));
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    bl += __shfl_xor_sync(0xffffffffu, bl, off);
    ol += __shfl_xor_sync(0xffffffffu, ol, off);
  }
  if (threadIdx.x == 0) { P.out[0] = P.bbox_w * bl; P.out[1] = ol; }
}

}  // namespace uavdet

using namespace uavdet;

extern "C" int uavdet_yolo_head_loss(const float* p_bbox, const float* p_obj, const float* tgt, int B, int A, int H,
                                     int W, const float* anchors_scaled_host, int ciou, float bbox_w,
                                     float objectness_w, float obj_scale_w, float no_obj_w, float* d_bbox, float* d_obj,
                                     float* new_t, void* workspace, float* out2, void* stream) {
  UAVDET_CHECK_ARG(p_bbox && p_obj && tgt && d_bbox && d_obj && workspace && out2 && anchors_scaled_host,
                   "yolo_head_loss: null pointer");
  UAVDET_CHECK_ARG(A > 0 && A <= 8 && B > 0 && H > 0 && W > 0, "yolo_head_loss: bad shape (A <= 8)");
  UAVDET_CHECK_ARG(((((uintptr_t)p_bbox | (uintptr_t)d_bbox | (uintptr_t)new_t) & 15) == 0), "yolo_head_loss: 16-byte alignment");
  LossParams P{};
  P.p_bbox = p_bbox; P.p_obj = p_obj; P.tgt = tgt;
  P.B = B; P.A = A; P.H = H; P.W = W;
  for (int a = 0; a < A; ++a) { P.aw[a] = anchors_scaled_host[2 * a]; P.ah[a] = anchors_scaled_host[2 * a + 1]; }
  P.ciou = ciou; P.bbox_w = bbox_w; P.objectness_w = objectness_w; P.obj_scale_w = obj_scale_w; P.no_obj_w = no_obj_w;
  P.d_bbox = d_bbox; P.d_obj = d_obj; P.new_t = new_t;
  // workspace: [B] int first | [B] float npos | [B][3] float acc
  P.first = (int*)workspace;
  P.npos = (float*)workspace + B;
  P.acc = (float*)workspace + 2 * B;
  P.out = out2;
  cudaStream_t st = (cudaStream_t)stream;
  UAVDET_CUDA(cudaMemsetAsync(P.first, 0x7f, sizeof(int) * (size_t)B, st));
  UAVDET_CUDA(cudaMemsetAsync(P.npos, 0, sizeof(float) * (size_t)B * 4, st));      // npos + acc
  const int cells = A * H * W;
  int slices = cells / 8192;
  if (slices < 1) slices = 1;
  if (slices > 64) slices = 64;
  loss_prepass_kernel<<<dim3((unsigned)B, (unsigned)slices), 256, 0, st>>>(P);
  UAVDET_LAUNCH_CHECK();
  const long long total = (long long)B * A * H * W;
  long long blocks = (total + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  loss_main_kernel<<<(unsigned)blocks, 256, 0, st>>>(P);
  UAVDET_LAUNCH_CHECK();
  loss_finalize_kernel<<<1, 32, 0, st>>>(P);
  UAVDET_LAUNCH_CHECK();
  return UAVDET_OK;
}

extern "C" size_t uavdet_yolo_head_loss_workspace_bytes(int B) { return sizeof(float) * (size_t)B * 5; }
