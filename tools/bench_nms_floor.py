"""GPU-box diagnostic: batched NMS with a score floor at the configs[4] shape (batch 128 x 96,000 candidates, a few hundred
above the floor), CUDA-event timed.  With the library built with -DUAVDET_NMS_PROFILE (see tools/prof_nms.py) the kernel
also prints its phase cycle counters for image 0."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_uav_det_b200 import ops
B, N = 128, 96000
g = torch.Generator(device="cuda").manual_seed(0)
c = torch.rand(B, N, 2, device="cuda", generator=g) * 600
wh = torch.rand(B, N, 2, device="cuda", generator=g) * 60 + 4
boxes = torch.cat([c - wh / 2, c + wh / 2], 2).contiguous()
scores = torch.rand(B, N, device="cuda", generator=g)
for floor in (0.998, 0.99, 0.9, float("-inf")):
    reps = 1 if floor == float("-inf") else 5
    for _ in range(2):
        keep, cnt = ops.nms_batched(boxes, scores, 0.5, score_floor=floor)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        keep, cnt = ops.nms_batched(boxes, scores, 0.5, score_floor=floor)
    e1.record(); torch.cuda.synchronize()
    print(f"floor {floor}: {int((scores > floor).sum()) / B:.0f} candidates/frame above, kept {float(cnt.float().mean()):.0f}: "
          f"{e0.elapsed_time(e1) / reps * 1000:.0f} us per batch of {B}", flush=True)
