"""Phase breakdown of the NMS kernel (needs a library built with -DUAVDET_NMS_PROFILE; see below) on the C1
candidate distribution.  Usage on the GPU box:
  nvcc ... -DUAVDET_NMS_PROFILE (tools/prof_nms.py --build does it into a scratch library and restores the product one)"""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from multimodal_uav_det_b200 import inference, ops
from multimodal_uav_det_b200.model import BaselineModel
from multimodal_uav_det_b200.utils.datatype import Config

dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = BaselineModel(hparams=Config(bench.HPARAMS)).to(dev).eval()
x, _ = bench.synth_batch(1)
det = inference.detect(model, x.to(dev))
torch.cuda.synchronize()
for c in sys.argv[1:] or ["1", "8"]:
    os.environ["UAVDET_NMS_CLUSTER"] = c
    print("cluster", c, flush=True)
    for _ in range(2):
        ops.nms_batched(det.boxes, det.scores, 0.5)
        torch.cuda.synchronize()
