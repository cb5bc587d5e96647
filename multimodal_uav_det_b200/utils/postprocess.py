"""Post-processing entry points kept from the reference (utils/postprocess.py): `calculate_iou`
(on the training path, :51-85) and `draw_bbox` (:11-48, OpenCV visualisation, host only)."""
import torch


def _cxcywh_to_xyxy(t: torch.Tensor) -> torch.Tensor:
    cx, cy, w, h = t.unbind(-1)
    return torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], dim=-1)


def calculate_iou(preds, targets, head_anchors, mask=None, bbox_loss_fn="mse"):
    """IoU of every selected prediction against the FIRST selected target (the reference returns
    `box_iou(...)[:, 0]`, postprocess.py:83-85 — reproduced, not fixed).

    preds / targets: (A,H,W,4) cxcywh in grid units; head_anchors (A,2); mask (A,H,W) bool."""
    boxes = preds.detach().clone()
    if bbox_loss_fn == "mse":
        boxes[..., 2:] = boxes[..., 2:] * head_anchors.to(boxes.device).view(-1, 1, 1, 2)
    if mask is not None:
        boxes, targets = boxes[mask], targets[mask]
    else:
        boxes, targets = boxes.reshape(-1, 4), targets.reshape(-1, 4)
    p = _cxcywh_to_xyxy(boxes)
    t = _cxcywh_to_xyxy(targets)[0]
    lt = torch.max(p[:, :2], t[:2])
    rb = torch.min(p[:, 2:], t[2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[:, 0] * wh[:, 1]
    area_p = (p[:, 2] - p[:, 0]) * (p[:, 3] - p[:, 1])
    area_t = (t[2] - t[0]) * (t[3] - t[1])
    return inter / (area_p + area_t - inter)


def draw_bbox(image, bbox, color=(0, 255, 0), thickness=2, label=None, format="xyxy"):
    """Draw one box (and optional label) on a BGR image; host-side OpenCV, not on the GPU path."""
    import cv2
    vals = [int(v) for v in bbox]
    if format == "xywh":
        x1, y1 = vals[0], vals[1]
        x2, y2 = x1 + vals[2], y1 + vals[3]
    else:
        x1, y1, x2, y2 = vals
    cv2.rectangle(image, (x1, y1), (x2, y2), color, thickness)
    if label is not None:
        font, fs = cv2.FONT_HERSHEY_SIMPLEX, 0.5
        (tw, th), base = cv2.getTextSize(label, font, fs, 1)
        cv2.rectangle(image, (x1, y1 - th - base - 5), (x1 + tw, y1), color, -1)
        cv2.putText(image, label, (x1, y1 - base - 3), font, fs, (255, 255, 255), 1)
    return image
