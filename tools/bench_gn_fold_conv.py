"""The 192 -> 192 (384 -> 384) channel-MLP convolution + GELU of MDyEncoder at batch 128, alone: plain bias epilogue
(instance 4) against the GroupNorm-fold epilogue (instance 6) on the same tensor, interleaved, CUDA events."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_uav_det_b200 import ops

dev = torch.device("cuda")
for (hw, c) in ((160, 192), (80, 384)):
    n = 128
    xs = [torch.randn(n, hw, hw, c, device=dev).bfloat16() for _ in range(2)]      # 2 x 1.26 GB: larger than L2
    w = (torch.randn(c, c, device=dev) / c ** 0.5).bfloat16()
    wg, b = w.float().sum(1).contiguous(), torch.randn(c, device=dev)
    sa = torch.stack([torch.rand(n, device=dev) + 0.5, torch.randn(n, device=dev)], 1).contiguous()
    out = torch.empty_like(xs[0])
    def plain(i): ops.conv_fwd(xs[i & 1], w, c, 1, 1, 0, act="gelu", shift=b, out=out)
    def fold(i): ops.conv_fwd(xs[i & 1], w, c, 1, 1, 0, act="gelu", scale=wg, shift=b, sample_affine=sa, out=out)
    def gelu_scale(i): ops.conv_fwd(xs[i & 1], w, c, 1, 1, 0, act="gelu", scale=wg, shift=b, out=out)
    def silu_scale(i): ops.conv_fwd(xs[i & 1], w, c, 1, 1, 0, act="silu", scale=wg, shift=b, out=out)
    def none_scale(i): ops.conv_fwd(xs[i & 1], w, c, 1, 1, 0, act="none", scale=wg, shift=b, out=out)
    def relu(i): ops.conv_fwd(xs[i & 1], w, c, 1, 1, 0, act="relu", shift=b, out=out)
    def relu_fold(i): ops.conv_fwd(xs[i & 1], w, c, 1, 1, 0, act="relu", scale=wg, shift=b, sample_affine=sa, out=out)
    res = {}
    for rep in range(2):
        for name, fn in (("gelu", plain), ("gelu_fold", fold), ("gelu_scale", gelu_scale), ("silu_scale", silu_scale), ("none_scale", none_scale), ("relu", relu), ("relu_fold", relu_fold)):
            for i in range(3): fn(i)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for i in range(10): fn(i)
            e1.record(); torch.cuda.synchronize()
            res.setdefault(name, []).append(round(e0.elapsed_time(e1) * 100, 1))
    print(json.dumps({"hw": hw, "c": c, "us_per_launch": res}))
