"""K9 against the library bar: torchvision.ops.nms on CUDA (the reference's call, model/_base.py:203) on the same B200
and the same inputs — the 25,200 un-thresholded candidates of a random-init BaselineModel frame (C1), the 96,000 of an
RTMUAVDet frame (C5, with and without a score floor), and a tie-heavy set (scores quantised to 1/20).
Kept indices must be identical; times are CUDA-event medians.  Prints one JSON line per input."""
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torchvision

import bench
from multimodal_uav_det_b200 import inference, ops
from multimodal_uav_det_b200.model import BaselineModel, RTMUAVDet
from multimodal_uav_det_b200.utils.datatype import Config

dev = torch.device("cuda", 0)


def med_us(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    return statistics.median(ts)


def compare(tag, boxes, scores, floor=float("-inf")):
    """boxes (B,N,4), scores (B,N)."""
    b, n = scores.shape
    ours = med_us(lambda: ops.nms_batched(boxes, scores, 0.5, floor))
    keep, count = ops.nms_batched(boxes, scores, 0.5, floor)

    def tv_all():
        outs = []
        for i in range(b):
            if floor == float("-inf"):
                outs.append(torchvision.ops.nms(boxes[i], scores[i], 0.5))
            else:
                idx = torch.nonzero(scores[i] > floor).squeeze(1)
                outs.append(idx[torchvision.ops.nms(boxes[i][idx], scores[i][idx], 0.5)])
        return outs

    tv = med_us(tv_all, iters=10)
    want = tv_all()
    same = all(torch.equal(keep[i, : int(count[i])], want[i]) for i in range(b))
    line = dict(input=tag, batch=b, candidates=n, score_floor=None if floor == float("-inf") else floor,
                kept_mean=float(count.float().mean()), ours_us=ours, torchvision_cuda_us=tv, speedup=tv / ours,
                kept_indices_identical=bool(same), torchvision=torchvision.__version__)
    print(json.dumps(line), flush=True)


def main():
    torch.manual_seed(0)
    model = BaselineModel(hparams=Config(bench.HPARAMS)).to(dev).eval()
    for b in (1, 32):
        x, _ = bench.synth_batch(b)
        det = inference.detect(model, x.to(dev))
        compare(f"C1 BaselineModel random-init outputs, batch {b}", det.boxes, det.scores)
        if b == 1:
            q = torch.round(det.scores * 20) / 20
            compare("C1 boxes, scores quantised to 1/20 (tie-heavy)", det.boxes, q.contiguous())
    del model
    anchors = torch.tensor([[[29, 23], [48, 30], [67, 38]], [[91, 54], [120, 75], [157, 60]]]).float()
    torch.manual_seed(0)
    rtm = RTMUAVDet([3, 640, 640], anchors, 1e-4).to(dev).eval()
    x, _ = bench.synth_batch(8)
    det = inference.detect_rtm(rtm, x.to(dev))
    compare("C5 RTMUAVDet outputs, batch 8, no floor", det.boxes, det.scores)
    compare("C5 RTMUAVDet outputs, batch 8, floor 0.5", det.boxes, det.scores, 0.5)
    ops.check_device()


if __name__ == "__main__":
    main()
