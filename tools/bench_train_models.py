"""GPU-box benchmark of the other training configurations of BASELINE.json at full size (one GPU):
  C3  DyYOLO (dy-yolo.yaml layer config) training step, batch 32, 640x640
  C4  DySOEM_SimFPN training step, batch 64, 640x640 (heads at 320/160/80: targets encoded on those grids)
Prints one JSON line per model (frames/s of the graph-replayed step, loss trajectory)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from oracle import oracle as O
from multimodal_uav_det_b200 import ops
from multimodal_uav_det_b200.model import DyYOLO, DySOEM_SimFPN
from multimodal_uav_det_b200.parallel import FlatSGDTrainer, GraphedTrainStep
from multimodal_uav_det_b200.utils.datatype import Config

dev = torch.device("cuda", 0)
D = bench.DARKNET53
DYYOLO = [["DyConv", 32, 3, 1], ["DyConv", 64, 3, 2]] + D[2:11] + [["DyConv", 512, 1, 1]] + D[12:16] + \
         [["DyConv", 256, 1, 1]] + D[17:21] + [["DyConv", 128, 1, 1]] + D[22:]


def run(name, model, batch, grids, anchors, head_scales, steps=10):
    model = model.to(dev).train()
    model.yolo_head.mutate_targets = False
    trainer = FlatSGDTrainer(model, lr=1e-4, momentum=0.7)
    x, boxes = bench.synth_batch(batch)
    per = [O.encode_targets(boxes[i:i + 1], anchors, head_scales, bench.IMG, grids=grids) for i in range(batch)]
    tg = [torch.stack([p[h] for p in per]).to(dev) for h in range(3)]
    x = x.to(dev)
    step = GraphedTrainStep(model, trainer, x, tg, warmup=2)
    losses = [step().item() for _ in range(3)]
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        step()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / steps
    losses.append(step().item())
    ops.check_device()
    print(json.dumps({"model": name, "batch": batch, "ms_per_step": ms, "frames_per_s": batch / ms * 1e3,
                      "losses": losses, "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}), flush=True)


if __name__ == "__main__":
    torch.manual_seed(0)
    hp = dict(bench.HPARAMS, layer_config=DYYOLO, attn_temperature=30)
    run("DyYOLO", DyYOLO(hparams=Config(hp)), 32, None, bench.ANCHORS, bench.HEAD_SCALES)
    torch.cuda.empty_cache()
    torch.manual_seed(0)
    hp2 = dict(anchors=[bench.ANCHORS[2], bench.ANCHORS[1], bench.ANCHORS[0]], head_scales=[32, 16, 8], lr=1e-4,
               lr_scheduler=False, attention_temperature=30, num_dy_conv=[3, 3, 3], dy_kernel_size=[3, 3, 3],
               bbox_loss_fn="mse", loss_balancing=dict(obj_scales_w=[2.0, 1.0, 0.5], bbox_w=4.0, objectness_w=1.0, no_obj_w=4.0),
               optim=dict(name="SGD", momentum=0.7))
    m = DySOEM_SimFPN(hparams=Config(hp2))
    m._attn_temp = 30.0
    orig_fwd = m.forward
    m.forward = lambda x, attn_temp=30.0: orig_fwd(x, attn_temp)
    run("DySOEM_SimFPN", m, 64, [320, 160, 80], hp2["anchors"], hp2["head_scales"])
