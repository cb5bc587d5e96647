"""One warm step of a benchmark configuration between cudaProfilerStart/Stop, for
    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file X.csv \
        python tools/profile_step_launches.py <baseline|dyyolo|dysoem|rtm-infer> [batch]
The launch list shows every kernel of the step (ours, ATen, NCCL) with its device time; tools/launch_summary.py
aggregates it."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from multimodal_uav_det_b200 import inference, ops
from multimodal_uav_det_b200.parallel import FlatSGDTrainer
from multimodal_uav_det_b200.utils.datatype import BatchData
from multimodal_uav_det_b200.utils.targets import YoloTargetEncoder

name = sys.argv[1]
wl = bench.WORKLOADS[name]
B = int(sys.argv[2]) if len(sys.argv) > 2 else wl["batch"]
dev = torch.device("cuda", 0)
model = bench._model_container(name).to(dev)
x, boxes = bench.synth_batch(B)
x = x.to(dev)
if name == "rtm-infer":
    model.eval()
    step = lambda: inference.detect_rtm(model, x, 0.5, bench.RTM_SCORE_FLOOR)
else:
    model.train()
    model.yolo_head.mutate_targets = False
    hp = wl["hp"]
    trainer = FlatSGDTrainer(model, lr=hp["lr"], momentum=wl["momentum"])
    tg = YoloTargetEncoder(hp["anchors"], wl["grids"], bench.IMG)(boxes.float().to(dev))
    fkw = {"attn_temp": 30.0} if name == "dysoem" else {}

    def step():
        trainer.zero_grad()
        outs = model(x, **fkw)
        loss, _, _, _ = model.yolo_head.compute_metrics(outs, BatchData(image=x, bbox=tg))
        loss.backward()
        trainer.step()
for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
ops.check_device()
print("profiled one step of", name, "batch", B)
