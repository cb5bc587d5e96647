"""GPU-box diagnostic: where the replayed training step spends its time (device-timer stamps captured in the graph)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from multimodal_uav_det_b200 import ops
from multimodal_uav_det_b200.model import BaselineModel
from multimodal_uav_det_b200.parallel import FlatSGDTrainer
from multimodal_uav_det_b200.utils.datatype import BatchData, Config

dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = BaselineModel(hparams=Config(bench.HPARAMS)).to(dev).train()
model.yolo_head.mutate_targets = False
trainer = FlatSGDTrainer(model, lr=1e-4, momentum=0.7)
x, boxes = bench.synth_batch(32)
tg = [t.to(dev) for t in bench.encode_targets_stacked(boxes)]
x = x.to(dev)
stamps = torch.zeros(64, dtype=torch.int64, device=dev)
names = []

def mark(name):
    ops.timestamp(stamps, len(names))
    names.append(name)

orig_bwd = model._backward_program
def bwd(tape, grads):
    mark("loss backward (autograd) done / trunk backward starts")
    orig_bwd(tape, grads)
    mark("trunk backward done")
model._backward_program = bwd
orig_head_bwd = model.yolo_head.backward_nhwc
cnt = [0]
def hb(*a, **k):
    r = orig_head_bwd(*a, **k)
    cnt[0] += 1
    if cnt[0] % 3 == 0:
        mark("head backward (3 scales) done")
    return r
model.yolo_head.backward_nhwc = hb

def body():
    names.clear()
    mark("start")
    trainer.zero_grad()
    mark("zero_grad done")
    outs = model(x)
    mark("forward (trunk + heads) done")
    loss, _, _, _ = model.yolo_head.compute_metrics(outs, BatchData(image=x, bbox=tg))
    mark("loss forward done")
    loss.backward()
    mark("backward returned")
    trainer.step()
    mark("sgd done")
    return loss

for _ in range(3):
    body()
model.prepare_for_capture()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    body()
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
t = stamps.cpu().tolist()
for i in range(1, len(names)):
    print(f"{(t[i] - t[i - 1]) / 1e3:9.1f} us  {names[i]}")
print(f"{(t[len(names) - 1] - t[0]) / 1e3:9.1f} us  total")
