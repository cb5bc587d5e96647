"""Launches every memory-bound kernel of the library once or twice at the shapes of BASELINE.json's configs (C2:
BaselineModel batch 32; C5: RTMUAVDet batch 128), so that `ncu --set full -k regex:<name>` can capture each of them:

    ncu --set full --clock-control none --import-source on -k regex:'dwdynconv|gn_|bilinear2x|decode_yolo|rtm_head_post|gap_kernel|encode_targets|sgd_momentum|bn_act|bn_bwd|upsample2x|cxcywh' \
        -o gpurun_out/membound python tools/run_membound_kernels.py

Without ncu it prints CUDA-event timings and the achieved GB/s of the ALGORITHMIC bytes of each launch (one JSON line
per kernel; peak = MEASURED_PEAKS.json hbm_gbs)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multimodal_uav_det_b200 import ops
from multimodal_uav_det_b200.utils.targets import YoloTargetEncoder

dev = torch.device("cuda", 0)
PEAK = 6543.7
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(name, fn, bytes_, iters=5):
    """Each timed launch follows an L2 flush (a 256 MB fill), timed separately with events."""
    if os.environ.get("UAVDET_MEMBOUND_ONCE"):       # under ncu: exactly one (cold-L2) launch per kernel
        FLUSH.zero_()
        fn()
        torch.cuda.synchronize()
        return
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        FLUSH.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e-3)
    t = sorted(ts)[len(ts) // 2]
    print(json.dumps(dict(kernel=name, algorithmic_MB=bytes_ / 1e6, us=t * 1e6, GBps=bytes_ / t / 1e9,
                          frac_of_hbm_peak=bytes_ / t / 1e9 / PEAK)), flush=True)


def act(n, h, w, c):
    return torch.randn(n, h, w, c, device=dev).to(torch.bfloat16)


def main():
    # ---- C5: RTMUAVDet batch 128 ----
    n = 128
    for c, hw, k in ((128, 160, 3), (64, 160, 5), (128, 80, 3), (128, 80, 5), (128, 80, 1)):
        x = act(n, hw, hw, c)
        cw = torch.rand(n, c, device=dev)
        kw = torch.rand(n, k * k, device=dev)
        out = torch.empty_like(x)
        run(f"dwdynconv_kernel C{c} k{k} @{hw} b{n}", lambda: ops.dwdynconv_fwd(x, cw, kw, k, k // 2, out=out), 2 * x.numel() * 2)
        del x, out
    for c, hw in ((192, 160), (384, 80)):
        a = act(n, hw, hw, c)
        b = act(n, hw, hw, c)
        g, be = torch.ones(c, device=dev), torch.zeros(c, device=dev)
        out = torch.empty_like(a)
        run(f"gn_stats+gn_apply (GroupNorm G=1) C{c} @{hw} b{n}", lambda: ops.groupnorm1(a, g, be, 1e-5, out=out), 3 * a.numel() * 2)
        run(f"gn_stats+gn_apply with residual operand C{c} @{hw} b{n}", lambda: ops.groupnorm1(a, g, be, 1e-5, b=b, out=out),
            5 * a.numel() * 2)
        del a, b, out
    x = act(n, 80, 80, 256)
    out = ops.empty_act(n, 160, 160, 256, dev)
    run("bilinear2x_kernel C256 80->160 b128", lambda: ops.bilinear2x_fwd(x, out=out), (x.numel() + out.numel()) * 2)
    del x, out
    anchors = torch.tensor([[29, 23], [48, 30], [67, 38]]).float()
    bl = torch.randn(n, 3, 160, 160, 4, device=dev)
    ol = torch.randn(n, 3, 160, 160, 1, device=dev)
    run("rtm_head_post_kernel (sigmoid + decode) 160^2 b128", lambda: ops.rtm_head_post(bl, ol, anchors), 2 * (bl.numel() + ol.numel()) * 4)
    boxes = torch.rand(n, 96000, 4, device=dev)
    run("cxcywh_to_xyxy_kernel 96,000 x b128", lambda: ops.cxcywh_to_xyxy(boxes), 2 * boxes.numel() * 4)
    del bl, ol, boxes
    x = act(n, 160, 160, 128)
    run("gap_kernel C128 @160 b128", lambda: ops.gap(x), x.numel() * 2)
    del x
    # ---- C2: BaselineModel batch 32 ----
    n = 32
    from collections import namedtuple
    DR = namedtuple("DR", "bbox obj")
    outs = [DR(torch.randn(n, 3, s, s, 4, device=dev), torch.randn(n, 3, s, s, 1, device=dev)) for s in (20, 40, 80)]
    anchors3 = [[[199, 73], [315, 92], [268, 182]], [[91, 54], [120, 75], [157, 60]], [[29, 23], [48, 30], [67, 38]]]
    run("decode_yolo_kernel 25,200 cand x b32 (3 launches)", lambda: ops.decode_yolo(outs, anchors3, [32, 16, 8], True),
        n * 25200 * 40)
    enc = YoloTargetEncoder.for_head_scales(anchors3, [32, 16, 8], 640)
    bx = torch.tensor([[300., 300., 340., 330.]], device=dev).repeat(n, 1)
    tg = enc(bx)
    run("encode_targets_kernel b32 (zero fill of dense targets)", lambda: enc(bx, check_grid=False, out=tg),
        sum(t.numel() for t in tg) * 4)
    p = torch.randn(61_518_349, device=dev)
    g = torch.randn_like(p)
    m = torch.zeros_like(p)
    run("sgd_momentum_kernel 61.5 M params", lambda: ops.sgd_momentum(p, g, m, 1e-4, 0.7), p.numel() * 20)
    del p, g, m
    for c, hw in ((64, 320), (256, 80), (1024, 20)):
        raw = act(n, hw, hw, c)
        res = act(n, hw, hw, c)
        dy = act(n, hw, hw, c)
        sc, sh = torch.rand(c, device=dev) + 0.5, torch.randn(c, device=dev)
        mean, inv = torch.randn(c, device=dev) * 0.1, torch.rand(c, device=dev) + 0.5
        out = torch.empty_like(raw)
        run(f"bn_act_fwd_kernel (+residual) C{c} @{hw} b32", lambda: ops.bn_act_fwd(raw, sc, sh, "leaky", res=res, out=out),
            3 * raw.numel() * 2)
        run(f"bn_bwd_reduce + bn_bwd_apply_fused C{c} @{hw} b32",
            lambda: ops.bn_act_bwd(dy, raw, sc, sh, mean, inv, sc, "leaky"), 5 * raw.numel() * 2)
        del raw, res, dy, out
    xin = torch.rand(64, 3, 640, 640, device=dev)
    w13 = torch.randn(32, 3, 1, 1, device=dev)
    s1, s2 = torch.zeros(32, device=dev), torch.zeros(32, device=dev)
    from multimodal_uav_det_b200._lib import EPI_STATS
    run("stem1x1_fwd_kernel (3->32, batch statistics) @640 b64", lambda: ops.stem_fwd(xin, w13, 1, 1, 0, epi=EPI_STATS, sum_=s1, sumsq=s2),
        xin.numel() * 4 + 64 * 640 * 640 * 64)
    dy = act(64, 640, 640, 32)
    run("stem1x1_wgrad_kernel (3->32) @640 b64", lambda: ops.stem_wgrad(xin, dy, 1, 1, 0), xin.numel() * 4 + dy.numel() * 2)
    del dy
    run("gap_nchw_kernel 3x640x640 b64", lambda: ops.gap_nchw(xin), xin.numel() * 4)
    x128 = torch.rand(128, 3, 640, 640, device=dev)
    run("stem_s2d_pack_kernel 3x640x640 -> 320x320x32 b128", lambda: ops.stem_s2d_pack(x128), x128.numel() * 4 + 128 * 320 * 320 * 64)
    del x128, xin
    dob = torch.randn(64, 3, 320, 320, 1, device=dev)
    dbb = torch.randn(64, 3, 320, 320, 4, device=dev)
    gbo, gbb = torch.zeros(3, device=dev), torch.zeros(12, device=dev)
    run("head_grad_pack_kernel 320x320 head b64", lambda: ops.head_grad_pack(dob, dbb, 64, 3, 320, 320, gbo, gbb),
        (dob.numel() + dbb.numel()) * 4 + 64 * 320 * 320 * 64)
    del dob, dbb
    x = act(n, 40, 40, 256)
    out = ops.empty_act(n, 80, 80, 256, dev)
    run("upsample2x_fwd_kernel C256 40->80 b32", lambda: ops.upsample2x_fwd(x, out=out), (x.numel() + out.numel()) * 2)
    ops.check_device()


if __name__ == "__main__":
    main()
