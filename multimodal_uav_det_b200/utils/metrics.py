"""Detection losses of the hot path (reference utils/metrics.py:8-84), written as batched,
sync-free tensor math so the B x 3 Python loop of `YOLOHead.compute_metrics`
(model/_base.py:163-192) collapses into a handful of launches.
mAP (`calculate_ap`, metrics.py:88-135) is torchmetrics' CPU-side evaluation and out of scope; the thin
wrapper below only exists so that `YOLOHead.compute_metrics(return_ap=True)` can hand it the NMS survivors when
torchmetrics is installed."""
import math

import torch
import torch.nn.functional as F


def cxcywh_to_xyxy(t: torch.Tensor) -> torch.Tensor:
    cx, cy, w, h = t.unbind(-1)
    return torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], dim=-1)


def complete_box_iou_loss(b1: torch.Tensor, b2: torch.Tensor, eps: float = 1e-7) -> torch.Tensor:
    """Element-wise Complete-IoU loss on xyxy boxes (same arithmetic as
    torchvision.ops.complete_box_iou_loss(reduction='none'), which metrics.py:31-35 calls)."""
    x1, y1, x2, y2 = b1.unbind(-1)
    x1g, y1g, x2g, y2g = b2.unbind(-1)
    ix1, iy1 = torch.max(x1, x1g), torch.max(y1, y1g)
    ix2, iy2 = torch.min(x2, x2g), torch.min(y2, y2g)
    overlap = (iy2 > iy1) & (ix2 > ix1)
    inter = torch.where(overlap, (ix2 - ix1) * (iy2 - iy1), torch.zeros_like(x1))
    union = (x2 - x1) * (y2 - y1) + (x2g - x1g) * (y2g - y1g) - inter + eps
    iou = inter / union
    ex1, ey1 = torch.min(x1, x1g), torch.min(y1, y1g)
    ex2, ey2 = torch.max(x2, x2g), torch.max(y2, y2g)
    diag = (ex2 - ex1) ** 2 + (ey2 - ey1) ** 2 + eps
    centre = ((x1 + x2) / 2 - (x1g + x2g) / 2) ** 2 + ((y1 + y2) / 2 - (y1g + y2g) / 2) ** 2
    v = (4 / (math.pi ** 2)) * torch.pow(torch.atan((x2g - x1g) / (y2g - y1g)) - torch.atan((x2 - x1) / (y2 - y1)), 2)
    with torch.no_grad():
        alpha = v / (1 - iou + v + eps)
    return 1 - iou + centre / diag + alpha * v


def bbox_loss(preds_decoded, targets, head_anchors=None, bbox_loss_fn="mse"):
    """metrics.py:8-37 on already-selected rows (N,4) cxcywh."""
    if bbox_loss_fn == "mse":
        return F.mse_loss(preds_decoded, targets, reduction="mean")
    return complete_box_iou_loss(cxcywh_to_xyxy(preds_decoded), cxcywh_to_xyxy(targets)).mean()


def objectness_loss(preds_obj, targets, obj_scale_w, reduction="mean"):
    """metrics.py:40-62."""
    return F.binary_cross_entropy_with_logits(preds_obj.squeeze(dim=-1), targets, reduction=reduction) * obj_scale_w


def no_obj_loss(preds_no_obj, targets, reduction="mean"):
    """metrics.py:65-84."""
    return F.binary_cross_entropy_with_logits(preds_no_obj.squeeze(dim=-1), targets, reduction=reduction)


def decode_head(bbox_logits: torch.Tensor, scaled_anchors: torch.Tensor, ciou: bool) -> torch.Tensor:
    """YOLOHead.__pred_bbox_decoding (model/_base.py:214-241), batched: (...,A,H,W,4) -> cxcywh."""
    sig = torch.sigmoid(bbox_logits)
    cxy = sig[..., :2] * 2 - 0.5
    wh = (sig[..., 2:] * 2) ** 2
    if ciou:
        a, h, w = bbox_logits.shape[-4:-1]
        dev = bbox_logits.device
        gx = torch.arange(w, device=dev).view(1, 1, w).expand(a, h, w)
        gy = torch.arange(h, device=dev).view(1, h, 1).expand(a, h, w)
        cxy = cxy + torch.stack([gx, gy], dim=-1)
        wh = wh * scaled_anchors.to(dev).view(a, 1, 1, 2)
    return torch.cat([cxy, wh], dim=-1)


def yolo_head_loss(p_bbox, p_obj, tgt, scaled_anchors, obj_scale_w, weights, bbox_loss_fn):
    """Loss of ONE head for the whole batch, equal to the reference's per-sample loop summed over
    samples (model/_base.py:163-192).  p_bbox (B,A,H,W,4), p_obj (B,A,H,W,1) logits,
    tgt (B,A,H,W,5) [obj,cx,cy,w,h].  Returns (sum_i bbox_w*bbox_loss_i, sum_i obj terms,
    new target boxes (B,A,H,W,4) as the reference would have rewritten them in place)."""
    bbox_w, objectness_w, no_obj_w = weights
    ciou = bbox_loss_fn == "ciou"
    b, a, h, w, _ = p_bbox.shape
    sa = scaled_anchors.to(p_bbox.device)
    t_obj = tgt[..., 0]
    t_box = tgt[..., 1:]
    cell = t_obj == 1.0                                               # (B,A,H,W)
    cellf = cell.to(p_bbox.dtype)
    npos = cellf.sum(dim=(1, 2, 3))                                   # (B,)
    dec = decode_head(p_bbox, sa, ciou)                               # (B,A,H,W,4)

    # calculate_iou (postprocess.py:51-85): every positive vs the FIRST positive target, with the
    # target boxes as they are BEFORE __build_target_bbox rewrites them.
    pb = dec.detach()
    if not ciou:
        pb = torch.cat([pb[..., :2], pb[..., 2:] * sa.view(1, a, 1, 1, 2)], dim=-1)
    first = torch.argmax(cell.reshape(b, -1).to(torch.uint8), dim=1)  # first True in (a h w) order
    t0 = torch.gather(t_box.reshape(b, -1, 4), 1, first.view(b, 1, 1).expand(b, 1, 4))  # (B,1,4)
    p_xyxy = cxcywh_to_xyxy(pb)
    t_xyxy = cxcywh_to_xyxy(t0).view(b, 1, 1, 1, 4)
    lt = torch.max(p_xyxy[..., :2], t_xyxy[..., :2])
    rb = torch.min(p_xyxy[..., 2:], t_xyxy[..., 2:])
    whi = (rb - lt).clamp(min=0)
    inter = whi[..., 0] * whi[..., 1]
    area_p = (p_xyxy[..., 2] - p_xyxy[..., 0]) * (p_xyxy[..., 3] - p_xyxy[..., 1])
    area_t = (t_xyxy[..., 2] - t_xyxy[..., 0]) * (t_xyxy[..., 3] - t_xyxy[..., 1])
    ious = inter / (area_p + area_t - inter)                          # (B,A,H,W)

    # __build_target_bbox (_base.py:250-270)
    if ciou:
        gx = torch.arange(w, device=tgt.device).view(1, 1, 1, w)
        gy = torch.arange(h, device=tgt.device).view(1, 1, h, 1)
        new_t = torch.stack([t_box[..., 0] + gx, t_box[..., 1] + gy, t_box[..., 2], t_box[..., 3]], dim=-1)
    else:
        new_t = torch.cat([t_box[..., :2], torch.sqrt((1e-16 + t_box[..., 2:]) / sa.view(1, a, 1, 1, 2)) / 2], dim=-1)

    # bbox loss: mean over the positives of each sample, then summed over samples
    if ciou:
        # keep non-positive rows finite (their value is masked out; avoids NaN * 0 in backward)
        safe_dec = torch.where(cell.unsqueeze(-1), dec, torch.ones_like(dec))
        safe_t = torch.where(cell.unsqueeze(-1), new_t, torch.ones_like(new_t))
        per = complete_box_iou_loss(cxcywh_to_xyxy(safe_dec), cxcywh_to_xyxy(safe_t))
        bl = (per * cellf).sum(dim=(1, 2, 3)) / npos
    else:
        sq = ((dec - new_t) ** 2).sum(dim=-1)
        bl = (sq * cellf).sum(dim=(1, 2, 3)) / (4 * npos)
    bbox_sum = bbox_w * bl.sum()

    logits = p_obj.squeeze(-1)
    pos_t = (ious * t_obj).detach()
    bce_pos = F.binary_cross_entropy_with_logits(logits, pos_t, reduction="none")
    ol = (bce_pos * cellf).sum(dim=(1, 2, 3)) / npos
    bce_neg = F.binary_cross_entropy_with_logits(logits, t_obj, reduction="none")
    nl = (bce_neg * (1 - cellf)).sum(dim=(1, 2, 3)) / (a * h * w - npos)
    obj_sum = objectness_w * obj_scale_w * ol.sum() + no_obj_w * nl.sum()
    return bbox_sum, obj_sum, new_t


class _FusedHeadLoss(torch.autograd.Function):
    """One head scale of the loss on the CUDA kernel (csrc/loss.cu): value and gradient come out of the same
    launch, so backward is two scalings."""

    @staticmethod
    def forward(ctx, p_bbox, p_obj, tgt, anchors_scaled, obj_scale_w, weights, ciou, want_new_t):
        from .. import ops
        bbox_w, objectness_w, no_obj_w = weights
        out2, d_bbox, d_obj, new_t = ops.yolo_head_loss(p_bbox, p_obj, tgt, anchors_scaled, ciou, bbox_w, objectness_w,
                                                        obj_scale_w, no_obj_w, want_new_t)
        ctx.save_for_backward(d_bbox, d_obj)
        if new_t is None:
            new_t = out2.new_empty(0)
        ctx.mark_non_differentiable(new_t)
        return out2[0], out2[1], new_t

    @staticmethod
    def backward(ctx, g_bbox, g_obj, _g_new_t):
        d_bbox, d_obj = ctx.saved_tensors
        return d_bbox * g_bbox, d_obj * g_obj, None, None, None, None, None, None


def yolo_head_loss_fused(p_bbox, p_obj, tgt, scaled_anchors, obj_scale_w, weights, bbox_loss_fn, want_new_t=True):
    """Same contract as `yolo_head_loss`, evaluated by the fused CUDA kernel (CUDA fp32 tensors)."""
    anc = [float(v) for v in scaled_anchors.flatten().tolist()]
    bl, ol, new_t = _FusedHeadLoss.apply(p_bbox.contiguous(), p_obj.contiguous(), tgt.contiguous(), anc,
                                         float(obj_scale_w), tuple(float(w) for w in weights), bbox_loss_fn == "ciou",
                                         want_new_t)
    return bl, ol, (new_t if want_new_t else None)


def calculate_ap(pred_boxes, pred_obj, target_boxes, max_det=300, iou_th=None):
    """reference utils/metrics.py:88-135: torchmetrics MeanAveragePrecision(box_format='cxcywh') on one image's
    predictions (single class).  Needs torchmetrics (absent from the build image: the caller gates on the import)."""
    from torchmetrics.detection import MeanAveragePrecision
    if iou_th is None:
        iou_th = [0.5 + 0.05 * i for i in range(10)]
    metric = MeanAveragePrecision(box_format="cxcywh", iou_thresholds=iou_th, max_detection_thresholds=[max_det] * 3)
    device = target_boxes.device
    preds = [dict(boxes=pred_boxes, scores=pred_obj, labels=torch.ones(len(pred_boxes), dtype=torch.int64, device=device))]
    target = [dict(boxes=target_boxes, labels=torch.ones(len(target_boxes), dtype=torch.int64, device=device))]
    metric.update(preds, target)
    return metric.compute()
