"""Evidence that the bucketed gradient all-reduce overlaps backward (VERDICT r01 'Next' #3): one eager data-parallel
training step of BaselineModel under torch.profiler (CUPTI kernel records: name, stream, start, duration).  For every
NCCL kernel the script lists the conv / BatchNorm kernels of the compute streams that run while it is in flight.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_overlap_trace.py > profiles/r02_dp_overlap_timeline.txt
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from multimodal_uav_det_b200.parallel import FlatSGDTrainer
from multimodal_uav_det_b200.utils.datatype import BatchData
from multimodal_uav_det_b200.utils.targets import YoloTargetEncoder

rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
os.environ.setdefault("NCCL_MAX_CTAS", os.environ.get("UAVDET_DP_SM_MARGIN", "8"))
torch.cuda.set_device(local)
dist.init_process_group("nccl")
dev = torch.device("cuda", local)
model = bench._model_container("baseline").to(dev).train()
model.yolo_head.mutate_targets = False
trainer = FlatSGDTrainer(model, lr=1e-4, momentum=0.7)
x, boxes = bench.synth_batch(32, seed=1234 + rank)
x = x.to(dev)
tg = YoloTargetEncoder(bench.ANCHORS, [20, 40, 80], 640)(boxes.float().to(dev))


def step():
    trainer.zero_grad()
    outs = model(x)
    loss, _, _, _ = model.yolo_head.compute_metrics(outs, BatchData(image=x, bbox=tg))
    loss.backward()
    trainer.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
dist.barrier()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
if rank == 0:
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
    ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in ev), key=lambda t: t[0])
    t0 = ks[0][0]
    nccl = [k for k in ks if k[2].startswith("ncclDevKernel")]
    comp = [k for k in ks if "nccl" not in k[2].lower() and "memcpy" not in k[2].lower() and "memset" not in k[2].lower()]
    last_bwd = max((k[1] for k in comp if "wgrad" in k[2] or "igemm" in k[2]), default=ks[-1][1])
    print(f"# one eager BaselineModel step, batch 32/GPU, world {dist.get_world_size()}, rank 0; times in us from the first kernel")
    print(f"# step kernels: {len(ks)}; NCCL kernels: {len(nccl)}; last conv kernel ends at {last_bwd - t0:.0f} us; "
          f"step ends at {ks[-1][1] - t0:.0f} us")
    tot_nccl = tot_over = 0.0
    for s, e, name in nccl:
        over = {}
        for cs, ce, cn in comp:
            o = min(e, ce) - max(s, cs)
            if o > 0:
                key = cn.split("(")[0].replace("void ", "")[:48]
                over[key] = over.get(key, 0.0) + o
        covered = sum(over.values())
        tot_nccl += e - s
        tot_over += min(covered, e - s)
        top = ", ".join(f"{k} {v:.0f}us" for k, v in sorted(over.items(), key=lambda kv: -kv[1])[:4])
        print(f"{name[:40]:40s} start {s - t0:9.0f} us  dur {e - s:7.0f} us  ends {'before' if e <= last_bwd else 'AFTER'} "
              f"the last conv kernel | concurrent compute: {top or 'none'}")
    print(f"# NCCL time {tot_nccl:.0f} us, of which {tot_over:.0f} us ({100 * tot_over / max(tot_nccl, 1):.0f} %) ran beside compute kernels")
dist.destroy_process_group()
