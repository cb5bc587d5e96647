/*
 * uavdet_b200 — C-ABI of the B200-native detector hot path (libuavdet_b200.so).
 *
 * This is the drop-in boundary.  The reference
 * (alfialdo/multimodal-uav-det) is pure Python on ATen/cuDNN/torchvision; each entry
 * point below cites the reference call site it replaces (paths relative to the
 * reference root).  The reference-side binding is a ctypes stub — see INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers + sizes, no torch types; all pointers are DEVICE pointers unless
 *     the name ends in `_host`;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), performs
 *     no allocation and no implicit synchronisation unless stated;
 *   - return 0 on success, non-zero error code otherwise; `uavdet_last_error()` gives
 *     the thread-local message;
 *   - activations are NHWC bf16 with an explicit pixel stride (`ld`, in elements) so a
 *     tensor may be a channel slice of a wider buffer (route concat, DyYOLO.py:139-141);
 *   - fp32 "stat" / weight-gradient buffers are accumulated with atomics and must be
 *     zeroed by the caller.
 */
#ifndef UAVDET_B200_H_
#define UAVDET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UAVDET_OK 0
#define UAVDET_ERR_ARG 1
#define UAVDET_ERR_CUDA 2
#define UAVDET_ERR_UNSUPPORTED 3
#define UAVDET_ERR_DEVICE 4 /* kernel reported a pipeline timeout (debug watchdog) */

/* activation codes (BaselineModel.py:16 LeakyReLU(0.1); _base.py:20 SiLU / ReLU) */
#define UAVDET_ACT_NONE 0
#define UAVDET_ACT_LEAKY 1
#define UAVDET_ACT_SILU 2
#define UAVDET_ACT_RELU 3
#define UAVDET_ACT_GELU 4

/* ---- library ---------------------------------------------------------------------- */
const char* uavdet_last_error(void);
int uavdet_version(void);
/* number of kernels launched by this library in this process (bench.py `gpu_launches`). */
uint64_t uavdet_launch_count(void);
/* reads and clears the device-side watchdog word (non-zero = a pipeline wait timed out).
 * Synchronises `stream`. */
int uavdet_check_device(void* stream, int* flag_host);
/* Data-parallel training (new work, SURVEY.md D6 / §8e — the reference has no collective call site): reserve `margin`
 * SMs for the NCCL all-reduce kernels that overlap backward; the persistent tensor-core kernels launched from now on
 * fill (148 - margin) SMs — for the next `launches` of them (the time a bucket's all-reduce is in flight), or until
 * reset when launches < 0.  Returns the previous margin.  Process-wide; margin 0 = use every SM. */
int uavdet_set_sm_margin(int margin, int launches);

/* Measurement aid: a one-thread kernel writes the device's %globaltimer (ns) to *slot_dev in stream order.
 * Captured into a CUDA graph around a kernel it gives that kernel's duration inside the replayed step
 * (bench.py `roofline`).  Not counted by uavdet_launch_count. */
int uavdet_timestamp(unsigned long long* slot_dev, void* stream);

/* NHWC bf16 activation view. */
typedef struct {
  void* ptr;   /* bf16 */
  int n, h, w; /* batch, height, width */
  int c;       /* channels visible through this view */
  int ld;      /* pixel stride in elements (>= c) */
} uavdet_act;

/* ---- K9: NMS  (replaces torchvision.ops.nms at model/_base.py:203) ----------------- */
/* Greedy IoU suppression, bit-exact with torchvision's CPU kernel: stable descending
 * sort (ties: lower index first, NaN scores first), suppress iff
 * inter/(a_i+a_j-inter) > thr with all arithmetic in unfused fp32.
 *   boxes   (batch, n, 4) fp32 xyxy      scores (batch, n) fp32
 *   keep    (batch, n) int64 — kept ORIGINAL indices in score order, first keep_count[b]
 *   keep_count (batch) int32
 *   score_floor: candidates with score <= score_floor are dropped before NMS (extension,
 *   SURVEY §8f-3; pass -INFINITY for the reference semantics: no threshold).            */
size_t uavdet_nms_workspace_bytes(int batch, int n);
int uavdet_nms(const float* boxes, const float* scores, int batch, int n, double iou_thr,
               float score_floor, int64_t* keep, int32_t* keep_count, void* workspace,
               size_t workspace_bytes, void* stream);

/* ---- K8: box decode ---------------------------------------------------------------- */
/* YOLOHead.__pred_bbox_decoding + __prepare_nms_preds (model/_base.py:214-248) for one
 * head scale: logits (batch, A, S_h, S_w, 4) / (batch, A, S_h, S_w, 1) fp32 ->
 * boxes xyxy written at candidate offset `cand_off` of a (batch, n_total, 4) buffer and
 * scores (raw logits, _base.py:203) at (batch, n_total).  anchors_scaled_host: A*(w,h)
 * already divided by the head scale (_base.py:170).  ciou != 0 adds the grid offsets and
 * anchor scaling (_base.py:224-237); ciou == 0 reproduces the 'mse' branch.             */
int uavdet_decode_yolo(const float* bbox_logits, const float* obj_logits, int batch, int A,
                       int S_h, int S_w, const float* anchors_scaled_host, int ciou,
                       float* boxes, float* scores, int n_total, int cand_off, void* stream);
/* RTMHead.__calculate_bbox_size (model/RTMUAVDet.py:274-291): input is already
 * sigmoid-activated (B,A,H,W,4); writes decoded cxcywh in place layout (B,A,H,W,4).     */
int uavdet_decode_rtm(const float* bbox_sig, int batch, int A, int S_h, int S_w,
                      const float* anchors_host, float* bbox_out, void* stream);
/* cxcywh (.., 4) -> xyxy, torchvision box_convert arithmetic (_base.py:246). */
int uavdet_cxcywh_to_xyxy(const float* in, float* out, int64_t count, void* stream);

/* ---- target encoder ----------------------------------------------------- */
/* AntiUAVDataset.__generate_yolo_bboxes (dataset/AntiUAVDataset.py:141-185) with calculate_anchor_iou
 * (dataset/_helper.py:308-330) for a batch: boxes_xyxy (batch,4) fp32 pixels on the device, one target per
 * frame; valid (batch) bytes or NULL (0 = frame without target -> all-zero targets, :142-143);
 * anchors_norm_host (heads, num_anchors, 2) = anchors / input_size evaluated by the caller in fp32
 * (AntiUAVDataset.py:27); grids_host (heads) = S per head; targets_host = `heads` DEVICE pointers to
 * (batch, num_anchors, S, S, 5) fp32 [obj, cx_off, cy_off, w_cells, h_cells], fully written (zero-filled
 * first).  out_of_grid (device u32) counts frames whose centre falls outside the grid -- the reference raises
 * IndexError there; such frames get all-zero targets.  Bit-identical to the CPU encoder.               */
int uavdet_encode_targets(const float* boxes_xyxy, const uint8_t* valid, int batch,
                          const float* anchors_norm_host, int heads, int num_anchors,
                          const int* grids_host, float input_size, float* const* targets_host,
                          unsigned int* out_of_grid, void* stream);

/* ---- detection loss ----------------------------------------------------- */
/* One head scale of YOLOHead.compute_metrics (model/_base.py:155-192 with utils/metrics.py:8-84,
 * utils/postprocess.py:51-85, _base.py:214-270) for the whole batch, forward AND gradient:
 *   out2[0] = bbox_w * sum_b mean_{positives of b} box_loss      (ciou != 0: complete-IoU, else MSE)
 *   out2[1] = sum_b [ objectness_w*obj_scale_w * mean_pos BCE(obj, iou_vs_first_target * t_obj)
 *                     + no_obj_w * mean_neg BCE(obj, t_obj) ]
 *   d_bbox / d_obj = d out2[0] / d p_bbox, d out2[1] / d p_obj   (same shapes as the logits)
 *   new_t (optional) = the target boxes as the reference rewrites them in place (_base.py:257,266-267).
 * p_bbox (B,A,H,W,4), p_obj (B,A,H,W,1), tgt (B,A,H,W,5) fp32; anchors_scaled_host: A*(w,h) / head scale.
 * workspace: uavdet_yolo_head_loss_workspace_bytes(B) bytes of device scratch.                            */
size_t uavdet_yolo_head_loss_workspace_bytes(int B);
int uavdet_yolo_head_loss(const float* p_bbox, const float* p_obj, const float* tgt, int B, int A, int H, int W,
                          const float* anchors_scaled_host, int ciou, float bbox_w, float objectness_w,
                          float obj_scale_w, float no_obj_w, float* d_bbox, float* d_obj, float* new_t,
                          void* workspace, float* out2, void* stream);

/* ---- K1/K2: implicit-GEMM convolution on tcgen05 ------------------------------------ */
/* Epilogue of the implicit GEMM. */
#define UAVDET_EPI_AFFINE 0 /* y = act(acc*scale[c]+shift[c]) (+res) -> bf16 NHWC        */
#define UAVDET_EPI_STATS 1  /* y = acc -> bf16 NHWC; sum[c]+=acc, sumsq[c]+=acc^2 (fp32) */
#define UAVDET_EPI_HEAD 2   /* fp32 (B,A,H,W,1)+(B,A,H,W,4) logits, _base.py:88-120      */

typedef struct {
  int epi;            /* UAVDET_EPI_*                                                    */
  int act;            /* UAVDET_ACT_* (AFFINE only)                                      */
  const float* scale; /* [cout] or NULL (=1)                                             */
  const float* shift; /* [cout] or NULL (=0)  — conv bias or folded BN shift             */
  const void* res;    /* bf16 residual added after activation, NHWC, or NULL             */
  int res_ld;
  float* sum;         /* STATS: [cout] fp32, caller-zeroed                               */
  float* sumsq;
  float* head_obj;    /* HEAD: (B,A,H,W,1) fp32                                          */
  float* head_bbox;   /* HEAD: (B,A,H,W,4) fp32                                          */
  int head_anchors;   /* A; cout must be 5*A packed [A obj | 4A bbox]                    */
  int shift_per_sample; /* != 0: shift is [n][cout] (aggregated expert bias, DySOEM_SimFPN.py:83-91);
                           in STATS mode a given shift is added before the statistics        */
  const float* sample_affine; /* AFFINE + ReLU|GELU only, or NULL.  [n][2] = (rstd_n, mean_n*rstd_n) from
                           uavdet_groupnorm1_fold: y = act(acc*rstd_n + shift[c] - mean_n*rstd_n*scale[c]) — a
                           GroupNorm(1 group) in front of a 1x1 convolution folded into its epilogue: the GEMM
                           reads the un-normalised tensor with weights W' = W*diag(gamma), scale[c] = sum_i W'[c][i],
                           shift[c] = (W beta)[c] + bias[c]  (RTMUAVDet.py:165-177)                              */
} uavdet_epilogue;

/* Forward convolution  y = conv(x, w), square kernel k in {1,3,5}, stride 1|2, zero pad.
 * Replaces nn.Conv2d / F.conv2d at BaselineModel.py:13, _base.py:18,72-74,85,107,
 * DySOEM_SimFPN.py:58,103-111, RTMUAVDet.py:19.
 *   x: NHWC bf16 (cin multiple of 32); y: NHWC bf16 view (n, ho, wo, cout)
 *   w_packed: bf16 [w_batch][cout][k*k*cin] (tap-major, cin fastest) from uavdet_pack_weight;
 *   w_batch = 1 (static) or n (per-sample dynamic kernels, _base.py:65-74).
 *   s2d != 0: x is read through a fused space-to-depth(2) gather (DySOEM_SimFPN.py:71-75):
 *   logical input is (n, h/2, w/2, 4*c) and cin = 4*x.c.                                 */
int uavdet_conv_fwd(const uavdet_act* x, const void* w_packed, int w_batch, int cout, int k,
                    int stride, int pad, int s2d, const uavdet_act* y,
                    const uavdet_epilogue* epi, void* stream);

/* Data gradient dx = conv_transpose(dy, w).  w_packed_t: bf16 [w_batch][cin][k*k*cout]
 * from uavdet_pack_weight(transposed=1).  epi: AFFINE with optional `res` (skip-path
 * gradient accumulated in the epilogue).  (autograd of the sites above)                  */
int uavdet_conv_dgrad(const uavdet_act* dy, const void* w_packed_t, int w_batch, int cin, int k,
                      int stride, int pad, const uavdet_act* dx, const uavdet_epilogue* epi,
                      void* stream);

/* Data gradient of a 3x3 stride-2 pad-1 convolution with cin in {32, 64} as ONE implicit GEMM over the four output
 * parity planes (autograd of the down-sampling layers, BaselineModel.py:63-75 `[C, 3, 2]` entries): N = 4*cin columns
 * [(row parity, column parity, ci)], K = four dy shifts x cout, weights re-laid out (with zero blocks where a parity
 * plane has no filter tap for a shift) by uavdet_pack_dgrad_s2_fused from the transposed pack [cin][3][3][cout].
 * dx: dense (n, 2*ho, 2*wo, cin); epi: NULL or AFFINE with `res` only.  The per-plane route (uavdet_conv_dgrad) issues
 * nine taps x four planes of N = cin instructions; the tensor core needs as long for N = 32 as for N = 128.            */
int uavdet_pack_dgrad_s2_fused(const void* w_packed_t, int cin, int cout, void* w_fused, void* stream);
int uavdet_conv_dgrad_s2_fused(const uavdet_act* dy, const void* w_fused, int cin, const uavdet_act* dx,
                               const uavdet_epilogue* epi, void* stream);

/* 3x3 stride-1 pad-1 convolution with cout in {64, 128}, two adjacent output pixels per GEMM row (nn.Conv2d at
 * RTMUAVDet.py:194, 256 -> 64 at 160x160, and the 32 -> 64 layer of the first residual block, BaselineModel.py:63-75:
 * an N = 64 GEMM runs the tensor core at half rate, 64-byte K rows at a quarter).  w_pair: bf16 [2*cout][3*S*cin], row
 * px*cout + co, column ((ky*S + s)*cin + ci) = w[co][ci][ky][s + first + 1 - px] (zero where that is not in 0..2) for
 * input-column shift s + first relative to the even pixel: S = 4, first = -1 for cin % 64 == 0; S = 6, first = -2 for a
 * dense 32-channel input (k-blocks are whole pixel pairs).  epi: AFFINE with scale / shift of length 2*cout (the layer's
 * vectors twice), or STATS (sum / sumsq of length cout: both pixels of a pair add to the same channel); no residual.
 * x: (n, h, w, cin) with w even; y: (n, h, w, cout) view.                                                             */
int uavdet_conv3x3_pair_fwd(const uavdet_act* x, const void* w_pair, int cout, const uavdet_act* y,
                            const uavdet_epilogue* epi, void* stream);

/* Data gradient through the fused space-to-depth(2) gather of uavdet_conv_fwd(s2d=1)
 * (autograd of DySOEM_SimFPN.py:71-91): dy (n, h/2, w/2, cout) -> dx (n, h, w, c) where the conv's
 * logical input had 4*c channels, block q = 2*(row parity) + (column parity).  w_packed_t: bf16
 * [w_batch][4*c][k*k*cout].  Runs one implicit GEMM per parity plane of dx.  epi: AFFINE; `res` is a
 * full-resolution gradient added in the epilogue; `shift` with shift_per_sample is [n][4*c] (the
 * pooled-attention gradient broadcast over each parity class).                             */
int uavdet_conv_dgrad_s2d(const uavdet_act* dy, const void* w_packed_t, int w_batch, int c, int k, int pad,
                          const uavdet_act* dx, const uavdet_epilogue* epi, void* stream);

/* Weight gradient dw[co][tap][ci] += sum_pixels dy[p][co] * x[p+tap][ci]  (fp32, atomics,
 * caller-zeroed; layout = packed weight layout).  per_sample != 0 keeps one dw per image
 * ([n][cout][k*k*cin]) for the dynamic-kernel contraction.                               */
int uavdet_conv_wgrad(const uavdet_act* x, const uavdet_act* dy, int k, int stride, int pad,
                      int s2d, float* dw_packed, int per_sample, void* stream);

/* fp32 conv weight -> packed bf16.  flags bit 0: 0 -> [O][kh][kw][I], 1 (transposed) -> [I][kh][kw][O];
 * flags bit 1: the fp32 source is stored channels-last ([O][kh][kw][I], torch.channels_last — how the flat
 * trainer keeps conv weights so that the weight gradient needs no re-layout) instead of OIHW.             */
int uavdet_pack_weight(const float* w_oihw, int O, int I, int k, int flags, void* out_bf16,
                       void* stream);
/* The same for a whole model in one launch (weights change every optimiser step: 2 packs per conv).
 * jobs_dev: device array of n_jobs descriptors.                                                           */
typedef struct {
  const float* src; /* fp32 weight                    */
  void* dst;        /* bf16 packed output             */
  int O, I, k;
  int flags;        /* as uavdet_pack_weight          */
  long long chunk0; /* index of this job's first 4096-element chunk: prefix sum of ceil(O*I*k*k / 4096) */
} uavdet_pack_job;
#define UAVDET_PACK_CHUNK 4096
int uavdet_pack_weights_batched(const uavdet_pack_job* jobs_dev, int n_jobs, long long total_chunks,
                                void* stream);
/* packed fp32 grad [O][kh][kw][I] -> OIHW fp32 (accumulate != 0: +=).                      */
int uavdet_unpack_wgrad(const float* dw_packed, int O, int I, int k, float* grad_oihw,
                        int accumulate, void* stream);

/* Stem: direct convolution for cin in {1,3} reading the NCHW fp32 network input
 * (BaselineModel.py:89-97 first layer, DySOEM_SimFPN.py:30, RTMUAVDet.py:31).
 * w: OIHW fp32.  Writes NHWC bf16 raw conv output (+ optional batch stats) or, with
 * scale/shift, the activated output.                                                     */
int uavdet_stem_fwd(const float* x_nchw, int n, int cin, int h, int w, const float* w_oihw,
                    int cout, int k, int stride, int pad, const uavdet_act* y,
                    const uavdet_epilogue* epi, void* stream);
/* Space-to-depth(2) of the 3-channel NCHW fp32 input into NHWC bf16 (n, H/2, W/2, 32): channel (py*2+px)*3 + ci holds
 * x[ci][2*by+py][2*bx+px], channels 12..31 are zero.  RTMUAVDet's 5x5 stride-2 pad-1 stem (RTMUAVDet.py:28-36) is a
 * 3x3 pad-1 stride-1 convolution over this map (filter row kh = 2*tap_y + py - 1), i.e. an ordinary uavdet_conv_fwd.  */
int uavdet_stem_s2d_pack(const float* x_nchw, int n, int h, int w, const uavdet_act* y, void* stream);
/* im2col of the cin<=3 input: (n,cin,h,w) fp32 NCHW -> (n,ho,wo,32) bf16 NHWC whose channel (ci*k+kh)*k+kw holds
 * the tap (zero outside the image, channels >= cin*k*k are zero).  The stem conv then IS a 1x1 convolution over
 * 32 channels with the weight matrix w.flatten(1) zero-padded to 32 columns: uavdet_conv_fwd / uavdet_conv_wgrad
 * (tcgen05) serve BaselineModel.py:89-97's first layer, its DyConv variant and their weight gradients.        */
int uavdet_im2col_stem(const float* x_nchw, int n, int cin, int h, int w, int k, int stride, int pad,
                       const uavdet_act* y, void* stream);
int uavdet_stem_wgrad(const float* x_nchw, int n, int cin, int h, int w, const uavdet_act* dy,
                      int k, int stride, int pad, float* grad_oihw, void* stream);
/* The 3x3 cin <= 3 stem on tcgen05 WITHOUT a patch tensor (BaselineModel.py:89-97 first layer; DyYOLO.py:89-100 with
 * per-sample kernels): the im2col rows are built in shared memory from the NCHW fp32 input (one 64-byte bf16 row per
 * output pixel, K = cin*9 zero-padded to 32) and consumed in place by the tensor core, so the layer moves its
 * algorithmic bytes only (12 B in, 64 B out per pixel; the im2col route writes and re-reads 64 B more).
 *   w_bf16: [w_batch][32][32] bf16 — w.flatten(1) rows zero-padded to 32 columns (uavdet_dyn_aggregate transposed == 2
 *           writes exactly this for the dynamic stem); w_batch = 1 or n;
 *   y:      (n, ho, wo, 32) NHWC bf16; epi: AFFINE (scale/shift/act) or STATS (raw output + batch sums);
 *   wgrad:  dy (n, ho, wo, 32); dw_o32 fp32 [per_sample ? n : 1][32][32] (columns >= cin*9 stay untouched), ACCUMULATED
 *           with atomics (caller-zeroed) — autograd of the same sites.
 * uavdet_stem_mma_supported: 1 when (cin, cout, k) is instantiated (k = 3, cin in 1..3, cout = 32).                 */
int uavdet_stem_mma_supported(int cin, int cout, int k);
int uavdet_stem_mma_fwd(const float* x_nchw, int n, int cin, int h, int w, const void* w_bf16, int w_batch,
                        int k, int stride, int pad, const uavdet_act* y, const uavdet_epilogue* epi,
                        void* stream);
int uavdet_stem_mma_wgrad(const float* x_nchw, int n, int cin, int h, int w, const uavdet_act* dy, int k,
                          int stride, int pad, float* dw_o32, int per_sample, void* stream);

/* ---- K5: batch-norm + activation (two-phase, train mode) ---------------------------- */
/* From batch sums: mean/invstd, running-stat update (momentum, unbiased var —
 * nn.BatchNorm2d at BaselineModel.py:14), and folded scale=gamma*invstd,
 * shift=beta-mean*scale.  count = n*h*w.                                                */
int uavdet_bn_finalize(const float* sum, const float* sumsq, int c, double count, float eps,
                       float momentum, const float* gamma, const float* beta,
                       float* running_mean, float* running_var, float* mean, float* invstd,
                       float* scale, float* shift, void* stream);
/* y = act(raw*scale+shift) (+res).  raw,y,res: NHWC bf16 views of equal n,h,w,c.          */
int uavdet_bn_act_fwd(const uavdet_act* raw, const float* scale, const float* shift, int act,
                      const uavdet_act* res, const uavdet_act* y, void* stream);
/* backward phase 1: dz = dy*act'(raw*scale+shift); sum_dz[c] += dz; sum_dzr[c] += dz*raw (fp32 atomics,
 * caller-zeroed). */
int uavdet_bn_act_bwd_reduce(const uavdet_act* dy, const uavdet_act* raw, const float* scale,
                             const float* shift, int act, float* sum_dz, float* sum_dzr, void* stream);
/* per channel: dgamma = invstd*(sum_dzr - mean*sum_dz), dbeta = sum_dz and the folded phase-2
 * coefficients k1 = -scale*invstd*dgamma/M, k0 = -scale*dbeta/M - k1*mean  (M = count).             */
int uavdet_bn_bwd_finalize(const float* sum_dz, const float* sum_dzr, const float* mean,
                           const float* invstd, const float* scale, int c, double count, float* dgamma,
                           float* dbeta, float* k1, float* k0, void* stream);
/* backward phase 2: d_raw = scale*dz + k1*raw + k0
 * (= gamma*invstd*(dz - mean(dz) - xhat*mean(dz*xhat)), the autograd of nn.BatchNorm2d in training). */
int uavdet_bn_act_bwd_apply(const uavdet_act* dy, const uavdet_act* raw, const float* scale,
                            const float* shift, const float* k1, const float* k0, int act,
                            const uavdet_act* d_raw, void* stream);
/* The same two phases with the per-channel finalize folded into the streaming kernel (one launch less per layer
 * and direction): uavdet_bn_train_fwd = bn_finalize + bn_act_fwd (publishes mean/invstd/scale/shift, updates the
 * running statistics); uavdet_bn_act_bwd_apply_fused = bn_bwd_finalize + bn_act_bwd_apply (publishes dgamma/dbeta;
 * accumulate != 0 adds them into the buffers instead — `param.grad` of bn.weight / bn.bias, as autograd's
 * AccumulateGrad would, so a data-parallel trainer can reduce the bucket as soon as the layer is done). */
int uavdet_bn_train_fwd(const uavdet_act* raw, const float* sum, const float* sumsq, double count, float eps,
                        float momentum, const float* gamma, const float* beta, float* running_mean,
                        float* running_var, float* mean, float* invstd, float* scale, float* shift, int act,
                        const uavdet_act* res, const uavdet_act* y, void* stream);
int uavdet_bn_act_bwd_apply_fused(const uavdet_act* dy, const uavdet_act* raw, const float* scale,
                                  const float* shift, const float* sum_dz, const float* sum_dzr, const float* mean,
                                  const float* invstd, double count, int act, float* dgamma, float* dbeta,
                                  int accumulate, const uavdet_act* d_raw, void* stream);
/* eval-mode / bias-only activation backward: dx = dy*act'(raw*scale+shift)*scale.         */
int uavdet_act_bwd(const uavdet_act* dy, const uavdet_act* raw, const float* scale,
                   const float* shift, int act, const uavdet_act* dx, void* stream);

/* ---- data movement ------------------------------------------------------------------ */
/* nearest x2 upsample of x into y (y.h = 2*x.h) — nn.Upsample(2) + cat (BaselineModel.py:86,
 * 120-122): y is typically a channel slice of the concat buffer.                         */
int uavdet_upsample2x_fwd(const uavdet_act* x, const uavdet_act* y, void* stream);
/* dx = sum over the 2x2 replicas of dy (accumulate: dx += ...).                           */
int uavdet_upsample2x_bwd(const uavdet_act* dy, const uavdet_act* dx, int accumulate, void* stream);
/* y = a_mult*a + nearest_upsample2x(b)  (SimplifiedFPN top-down adds, DySOEM_SimFPN.py:116-118; the
 * 1x1 conv commutes with nearest upsampling, so it runs at low resolution and is upsampled here). */
int uavdet_upsample2x_add(const uavdet_act* b_low, const uavdet_act* a, float a_mult, const uavdet_act* y,
                          void* stream);
/* y = a (+ b) elementwise over views (copy into / accumulate from channel slices).        */
int uavdet_add(const uavdet_act* a, const uavdet_act* b, const uavdet_act* y, void* stream);
/* NHWC bf16 <-> NCHW fp32 (API edge only: parity tests and user-facing feature maps).     */
int uavdet_nhwc_to_nchw_f32(const uavdet_act* x, float* y_nchw, void* stream);
int uavdet_nchw_f32_to_nhwc(const float* x_nchw, const uavdet_act* y, void* stream);

/* ---- K4: dynamic-kernel attention + aggregation -------------------------------------- */
/* Global average pool over h*w -> (n, c) fp32 (nn.AdaptiveAvgPool2d(1), _base.py:42).     */
int uavdet_gap(const uavdet_act* x, int s2d, float* out, void* stream);
int uavdet_gap_nchw(const float* x_nchw, int n, int c, int hw, float* out, void* stream);
/* Two-layer MLP + softmax(./T): pooled (n,c) -> attn (n,K)  (_base.py:41-46,60-62;
 * DySOEM_SimFPN.py:46-52,78-79).  w1 [hid][c], b1 [hid]|NULL, w2 [K][hid], b2 [K].
 * Saves hidden (n,hid) post-ReLU for backward when non-NULL.                             */
int uavdet_attn_mlp_softmax(const float* pooled, int n, int c, const float* w1, const float* b1,
                            int hid, const float* w2, const float* b2, int K, float temperature,
                            float* attn, float* hidden, void* stream);
/* Per-sample kernel aggregation  W_b = sum_k attn[b,k] * bank[k]  (_base.py:65-66) fused
 * with the bf16 pack: bank fp32 [K][O][I][k][k] -> out bf16 [n][O][k*k*I] (or transposed).
 * bias_bank [K][O] -> bias_out [n][O] when non-NULL (DySOEM_SimFPN.py:83-91 by linearity). */
int uavdet_dyn_aggregate(const float* attn, int n, int K, const float* bank, int O, int I, int k,
                         int transposed, void* out_bf16, const float* bias_bank, float* bias_out,
                         void* stream);
/* transposed == 2: the cin <= 3 stem sites (DyYOLO's 3 -> 32 k3, _base.py:36-39) that run as im2col + 1x1 GEMM:
 * out bf16 [n][O][32] = OIHW-flattened aggregated kernels (I*k*k <= 32 columns) zero-padded to the 32 patch channels. */
/* Gradients of the detection-head outputs (the `(B,A,H,W,1)` / `(B,A,H,W,4)` fp32 tensors of _base.py:91-94,112-115,
 * either may be NULL) gathered into the NHWC bf16 operand of the head's backward GEMMs: dyh (n,h,w,32) with channels
 * [A obj | 4A bbox | zero padding]; the bias gradients (sums over n*h*w) are ACCUMULATED into bias_obj_grad[A] /
 * bias_bbox_grad[4A] (NULL: skipped).  anchors <= 3.                                                           */
int uavdet_head_grad_pack(const float* d_obj, const float* d_bbox, int n, int anchors, int h, int w,
                          const uavdet_act* dyh, float* bias_obj_grad, float* bias_bbox_grad, void* stream);
/* Backward of uavdet_attn_mlp_softmax (autograd of _base.py:41-46,60-62 / DySOEM_SimFPN.py:46-52,78-79):
 * g = a*(d_attn - <a,d_attn>)/T; dh = (g @ w2)*[hidden>0]; dW2 += g^T hidden; db2 += sum g; dW1 += dh^T pooled;
 * db1 += sum dh (NULL: no first-layer bias); d_pooled = out_scale * dh @ w1 (NULL: not needed).  The parameter
 * gradients ACCUMULATE (they are `param.grad` buffers).  workspace: n*(K+hid) floats.                          */
int uavdet_attn_mlp_bwd(const float* attn, const float* d_attn, const float* hidden, const float* pooled, int n,
                        int c, const float* w1, int hid, const float* w2, int K, float temperature,
                        float out_scale, float* workspace, float* dw1, float* db1, float* dw2, float* db2,
                        float* d_pooled, void* stream);

/* Backward of the aggregation: from the per-sample kernel gradients dwb [n][O*I*k*k] fp32
 * (packed != 0: the [O][k*k][I] layout uavdet_conv_wgrad(per_sample=1) writes, I may be 4*c for s2d;
 * packed == 0: OIHW, what uavdet_stem_wgrad writes) accumulate
 *   d_bank[kk][e] += sum_b attn[b,kk] * dwb[b][e]      (OIHW fp32, autograd of _base.py:65-66)
 *   d_attn[b,kk]  += sum_e dwb[b][e] * bank[kk][e].                                          */
int uavdet_dyn_bwd_contract(const float* dwb, int n, int K, const float* attn, const float* bank, int O, int I,
                            int k, int packed, float* d_bank, float* d_attn, void* stream);

/* Backward of the per-sample bias bias[b] = attn[b] @ bias_bank of a dynamic conv (DynamicSOEM, DySOEM_SimFPN.py:56-60;
 * autograd of that matmul): with g = scale * pooled_grad (n, O) — the per-sample channel sums of the output gradient —
 *   d_bias_bank[kk][o] = sum_b attn[b][kk] * g[b][o]   (written),   d_attn[b][kk] += sum_o g[b][o] * bias_bank[kk][o]. */
int uavdet_dyn_bias_bwd(const float* pooled_grad, float scale, int n, int K, int O, const float* attn,
                        const float* bias_bank, float* d_bias_bank, float* d_attn, void* stream);

/* ---- K3 / K6 / K7: RTMUAVDet memory-bound ops -------------------------------------------- */
/* Per-sample depthwise dynamic conv + residual (MDyConv.forward, RTMUAVDet.py:80-98):
 * y[b,p,c] = x[b,p,c] + channel_w[b,c] * sum_t kernel_w[b,t] * x[b,p+t,c]; k odd, pad = k/2.      */
int uavdet_dwdynconv_fwd(const uavdet_act* x, const float* channel_w, const float* kernel_w, int k, int pad,
                         const uavdet_act* y, void* stream);
/* The same at the MDyEncoder site (RTMUAVDet.py:163-174) with the encoder's residual folded in:
 * y = dwdynconv(x) + res, and stats[2b] += sum(y), stats[2b+1] += sum(y^2) over the bf16-rounded y of image b
 * (caller-zeroed, shared by the three MDyConv branches) — the statistics of GroupNorm(cat + residual).          */
int uavdet_dwdynconv_res_stats_fwd(const uavdet_act* x, const float* channel_w, const float* kernel_w, int k, int pad,
                                   const uavdet_act* res, float* stats, const uavdet_act* y, void* stream);
/* out[r][o] = act(in[r][:] . w[o][:] + bias[o])  — the 1x1 convs on pooled vectors (RTMUAVDet.py:54-62). */
int uavdet_linear(const float* in, int rows, int c, const float* w, const float* bias, int out_dim, int act,
                  float* out, void* stream);
/* GroupNorm(num_groups=1) of (a [+ b]) with per-channel affine (RTMUAVDet.py:147,153,165,174).
 * stats_ws: 2*n floats of scratch.                                                              */
int uavdet_groupnorm1(const uavdet_act* a, const uavdet_act* b, const float* gamma, const float* beta,
                      float eps, float* stats_ws, const uavdet_act* y, void* stream);
/* GroupNorm(num_groups=1) folded into the 1x1 convolution that consumes it (RTMUAVDet.py:165-169 -> MDyConv.base_conv,
 * :174-177 -> channel_mlp[0]): uavdet_groupnorm1_stats gives stats[n][2] = per-sample sum / sum of squares of (a [+ b]);
 * uavdet_groupnorm1_fold turns stats (from it or from uavdet_dwdynconv_res_stats_fwd) into sample_affine[n][2] =
 * (rstd_n, mean_n*rstd_n), count = h*w*c, for uavdet_epilogue::sample_affine.                                      */
int uavdet_groupnorm1_stats(const uavdet_act* a, const uavdet_act* b, float* stats, void* stream);
int uavdet_groupnorm1_fold(const float* stats, int n, double count, float eps, float* sample_affine, void* stream);
/* nn.Upsample(scale_factor=2, mode='bilinear'), align_corners=False (RTMUAVDet.py:193).           */
int uavdet_bilinear2x_fwd(const uavdet_act* x, const uavdet_act* y, void* stream);
/* RTMHead post: sigmoid on obj / bbox logits (RTMUAVDet.py:234,253) + decode (:285-289).           */
int uavdet_rtm_head_post(const float* bbox_logits, const float* obj_logits, int batch, int A, int S_h, int S_w,
                         const float* anchors_host, float* bbox_out, float* obj_out, void* stream);

/* ---- optimiser ----------------------------------------------------------------------- */
/* torch.optim.SGD(momentum) step over a flat fp32 parameter arena (_base.py:292-293):
 * buf = momentum*buf + grad*grad_scale; p -= lr*buf   (first_step: buf = grad).            */
int uavdet_sgd_momentum(float* param, const float* grad, float* momentum_buf, int64_t count,
                        float lr, float momentum, float grad_scale, int first_step, void* stream);
/* The same step with the hyper-parameters read from device memory at execution time: hyper_dev = {lr, momentum,
 * grad_scale} (3 floats).  A step captured into a CUDA graph then follows a learning-rate scheduler (the optional
 * CyclicLR of _base.py:299-309) without re-capture: the host rewrites hyper_dev before the replay.            */
int uavdet_sgd_momentum_dev(float* param, const float* grad, float* momentum_buf, int64_t count,
                            const float* hyper_dev, int first_step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UAVDET_B200_H_ */
