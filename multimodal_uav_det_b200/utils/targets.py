"""YOLO target encoding on the GPU.

The reference encodes targets in the data set, on the CPU, one frame at a time
(`AntiUAVDataset.__generate_yolo_bboxes`, dataset/AntiUAVDataset.py:141-185, with `calculate_anchor_iou`,
dataset/_helper.py:308-330) and the loader ships dense `(3, S, S, 5)` tensors -- 25,200 x 5 floats per frame, at most
45 of them non-zero.  `YoloTargetEncoder` produces the same tensors, bit for bit, from the `(B, 4)` pixel boxes on
the device, so a training step only has to receive 16 bytes of target per frame."""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from .. import ops


class YoloTargetEncoder:
    """anchors: pixel anchors (heads, A, 2) as in conf/model/*.yaml; head_size: grid size S per head
    (AntiUAVDataset.py:28: input_size // head_scale); input_size: the square input side (params.yaml:8-10)."""

    def __init__(self, anchors, head_size: Sequence[int], input_size: int = 640):
        self.anchors = [[list(map(float, a)) for a in head] for head in anchors]
        self.head_size = [int(s) for s in head_size]
        self.input_size = int(input_size)
        assert len(self.anchors) == len(self.head_size)

    @classmethod
    def for_head_scales(cls, anchors, head_scales: Sequence[int], input_size: int = 640) -> "YoloTargetEncoder":
        return cls(anchors, [input_size // s for s in head_scales], input_size)

    def __call__(self, boxes_xyxy: torch.Tensor, valid: Optional[torch.Tensor] = None, check_grid: bool = True,
                 out: Optional[List[torch.Tensor]] = None) -> List[torch.Tensor]:
        """boxes_xyxy (B,4) pixels on the device, one target per frame -> per head (B,A,S,S,5)."""
        return ops.encode_targets(boxes_xyxy, self.anchors, self.head_size, self.input_size, valid=valid,
                                  check_grid=check_grid, out=out)
