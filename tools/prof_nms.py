"""Phase breakdown of the NMS kernel on the C1 candidate distribution (25,200 candidates, no score floor).
Needs the library built with the kernel's cycle counters compiled in:
  cd multimodal_uav_det_b200 && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC \
     --expt-relaxed-constexpr -DUAVDET_NMS_PROFILE -c csrc/nms.cu -o build/nms.o && \
     nvcc -shared -o libuavdet_b200.so build/*.o -gencode arch=compute_100a,code=sm_100a
(then `python -m multimodal_uav_det_b200.build --force` restores the product library).  Arguments: cluster sizes."""
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
