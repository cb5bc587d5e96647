"""GPU-box benchmark of the inference configurations of BASELINE.json (not the headline bench.py line):
  C1  BaselineModel forward + decode + NMS, batch 1 (25,200 candidates/frame, no score threshold = the reference's
      `return_ap` path, model/_base.py:196-203) — CUDA product vs the oracle's CPU path on this box's cores;
  C5  RTMUAVDet forward + fused sigmoid/decode + batched NMS, batch 128 (96,000 candidates/frame).
Prints one JSON line per configuration."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from multimodal_uav_det_b200 import inference, ops
from multimodal_uav_det_b200.model import BaselineModel, RTMUAVDet
from multimodal_uav_det_b200.utils.datatype import Config

dev = torch.device("cuda", 0)


@torch.no_grad()          # inference timings: no tape, no autograd node for the trunk
def timed(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        out = fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters, out


def c1():
    torch.manual_seed(0)
    model = BaselineModel(hparams=Config(bench.HPARAMS)).to(dev).eval()
    res = {}
    for b in (1, 32):
        x, _ = bench.synth_batch(b)
        x = x.to(dev)
        ms_f, _ = timed(lambda: model(x), 10)
        ms, det = timed(lambda: inference.detect(model, x), 10)
        run = inference.GraphedDetect(model, x)
        ms_g, det_g = timed(lambda: run(x), 10)
        assert torch.equal(det_g.keep_count, det.keep_count)
        nms_ms = {}
        for c in (1, 2, 4, 8, 0):
            if c:
                os.environ["UAVDET_NMS_CLUSTER"] = str(c)
            else:
                os.environ.pop("UAVDET_NMS_CLUSTER", None)
            nms_ms["auto" if c == 0 else f"cluster{c}"], _ = timed(lambda: ops.nms_batched(det.boxes, det.scores, 0.5), 10)
        res[b] = dict(ms_total=ms, ms_forward=ms_f, ms_total_graphed=ms_g, ms_nms=nms_ms, fps=b / ms * 1e3, fps_graphed=b / ms_g * 1e3,
                      kept=det.keep_count.float().mean().item())
    line = {"config": "C1 BaselineModel forward+decode+NMS (25,200 candidates/frame, no threshold)", "gpu": res}
    if "--cpu" in sys.argv:
        from oracle import oracle as O
        torch.set_num_threads(os.cpu_count() or 1)
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        x, _ = bench.synth_batch(1)
        def cpu_once():
            with torch.no_grad():
                outs = O.darknet_forward(x, sd, bench.DARKNET53, train=False)
                boxes, scores = O.decode_yolo(outs, bench.ANCHORS, bench.HEAD_SCALES, True)
                return O.nms(boxes[0].numpy(), scores[0].numpy(), 0.5)
        cpu_once()
        t0 = time.perf_counter(); k = cpu_once(); dt = time.perf_counter() - t0
        line["cpu_oracle"] = dict(ms_total=dt * 1e3, fps=1 / dt, cores=os.cpu_count(), kept=len(k))
    print(json.dumps(line), flush=True)


def c5():
    anchors = torch.tensor([[[29, 23], [48, 30], [67, 38]], [[91, 54], [120, 75], [157, 60]]]).float()
    torch.manual_seed(0)
    model = RTMUAVDet([3, 640, 640], anchors, 1e-4).to(dev).eval()
    res = {}
    for b, floor in ((128, 0.55), (128, 0.5), (8, float("-inf"))):
        x, _ = bench.synth_batch(b)
        x = x.to(dev)
        ms_f, _ = timed(lambda: model(x), 5, warm=2)
        ms, det = timed(lambda: inference.detect_rtm(model, x, 0.5, floor), 5, warm=2)
        above = (det.scores > floor).float().sum(dim=1).mean().item()
        res[f"b{b}_floor{floor}"] = dict(ms_total=ms, ms_forward=ms_f, fps=b / ms * 1e3, candidates_above_floor=above,
                                         kept=det.keep_count.float().mean().item())
    print(json.dumps({"config": "C5 RTMUAVDet forward + sigmoid/decode + batched NMS (96,000 candidates/frame)", "gpu": res}),
          flush=True)


if __name__ == "__main__":
    c1()
    c5()
    ops.check_device()
