"""Inference entry point: forward -> fused box decode -> batched NMS, all on the GPU.

The reference has no inference script; its only decode+NMS sequence is the `return_ap` branch of
`YOLOHead.compute_metrics` (model/_base.py:196-203): per image, decode each head, flatten
`(a h w)`, cxcywh->xyxy, concatenate the heads, `torchvision.ops.nms(boxes, logits, 0.5)`.
`detect` reproduces exactly that (boxes stay in per-head grid units, scores are raw logits, no
score threshold) for the whole batch with 3 decode launches + 1 NMS launch.  `score_floor` is an
extension: NMS on the subset with score > floor."""
from __future__ import annotations

from typing import List, NamedTuple

import torch

from . import ops


class Detections(NamedTuple):
    boxes: torch.Tensor       # (B, N, 4) xyxy, all candidates
    scores: torch.Tensor      # (B, N)
    keep: torch.Tensor        # (B, N) int64, first keep_count[b] entries valid, score-descending
    keep_count: torch.Tensor  # (B,) int32


@torch.no_grad()
def postprocess(outs, anchors, head_scales, bbox_loss_fn: str = "ciou", iou_threshold: float = 0.5,
                score_floor: float = float("-inf")) -> Detections:
    boxes, scores = ops.decode_yolo(outs, anchors, head_scales, bbox_loss_fn == "ciou")
    keep, count = ops.nms_batched(boxes, scores, iou_threshold, score_floor)
    return Detections(boxes, scores, keep, count)


@torch.no_grad()
def detect(model, x: torch.Tensor, iou_threshold: float = 0.5, score_floor: float = float("-inf")) -> Detections:
    """model: BaselineModel | DyYOLO | DySOEM_SimFPN (anything with `.yolo_head`)."""
    head = model.yolo_head
    outs = model(x)
    return postprocess(outs, head.anchors.tolist(), head.head_scales.tolist(), head.bbox_loss_fn, iou_threshold,
                       score_floor)


@torch.no_grad()
def detect_rtm(model, x: torch.Tensor, iou_threshold: float = 0.5, score_floor: float = float("-inf")) -> Detections:
    """RTMUAVDet: the heads already return sigmoid objectness and decoded cxcywh boxes (RTMUAVDet.py:274-310,
    one fused sigmoid+decode kernel per scale); candidates of both scales are concatenated per image, converted to
    xyxy and suppressed by the batched NMS kernel.  The reference never calls NMS on this model:
    semantics = torchvision.ops.nms per image on the candidates with score > score_floor."""
    return postprocess_rtm(model(x), iou_threshold, score_floor)


@torch.no_grad()
def postprocess_rtm(outs, iou_threshold: float = 0.5, score_floor: float = float("-inf")) -> Detections:
    """RTMHead outputs (decoded cxcywh boxes + sigmoid objectness per scale) -> xyxy candidates + batched NMS."""
    b = outs[0].bbox.shape[0]
    boxes = torch.cat([o.bbox.reshape(b, -1, 4) for o in outs], dim=1).contiguous()
    scores = torch.cat([o.obj.reshape(b, -1) for o in outs], dim=1).contiguous()
    boxes = ops.cxcywh_to_xyxy(boxes)
    keep, count = ops.nms_batched(boxes, scores, iou_threshold, score_floor)
    return Detections(boxes, scores, keep, count)


class GraphedDetect:
    """`detect` / `detect_rtm` for a fixed batch shape, captured into one CUDA graph: at batch 1 the eager path is
    launch-bound (~90 launches from Python for 7 ms of forward on a 154 GFLOP model).

        run = GraphedDetect(model, x_example, iou_threshold=0.5)
        det = run(x)            # Detections in static buffers, overwritten by the next call
    """

    def __init__(self, model, x: torch.Tensor, iou_threshold: float = 0.5, score_floor: float = float("-inf"),
                 warmup: int = 2):
        self.x = x.detach().clone().float().contiguous()
        fn = detect if hasattr(model, "yolo_head") else detect_rtm
        if hasattr(model, "prepare_for_capture"):
            model.prepare_for_capture()
        body = lambda: fn(model, self.x, iou_threshold, score_floor)
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream(device=self.x.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                body()
        cur.wait_stream(side)
        torch.cuda.synchronize(self.x.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = body()

    def __call__(self, x: torch.Tensor) -> Detections:
        if x is not self.x:
            self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.out


def kept_lists(det: Detections) -> List[torch.Tensor]:
    counts = det.keep_count.tolist()
    return [det.keep[b, :c] for b, c in enumerate(counts)]
