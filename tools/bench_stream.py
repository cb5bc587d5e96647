"""GPU-box micro-benchmark of the HBM-bound kernels (achieved GB/s of algorithmic bytes), graph-captured launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_uav_det_b200 import ops
from multimodal_uav_det_b200._lib import EPI_STATS


def timeit(fn, reps=10, inner=4):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(inner): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / (reps * inner)


for (n, hw, c) in [(32, 80, 256), (32, 160, 128), (32, 320, 64), (32, 40, 512), (32, 20, 1024)]:
    raw = torch.randn(n, hw, hw, c, device="cuda").bfloat16()
    dy = torch.randn(n, hw, hw, c, device="cuda").bfloat16()
    res = torch.randn(n, hw, hw, c, device="cuda").bfloat16()
    out = torch.empty_like(raw)
    scale = torch.rand(c, device="cuda") + 0.5; shift = torch.randn(c, device="cuda")
    mean = torch.randn(c, device="cuda"); invstd = torch.rand(c, device="cuda") + 0.5
    e = raw.numel() * 2
    t = timeit(lambda: ops.bn_act_fwd(raw, scale, shift, "leaky", out=out))
    print(f"bn_act_fwd        {c:5d}@{hw:3d}: {t*1e6:7.1f} us  {2*e/t/1e9:7.0f} GB/s")
    t = timeit(lambda: ops.bn_act_fwd(raw, scale, shift, "leaky", res=res, out=out))
    print(f"bn_act_fwd+res    {c:5d}@{hw:3d}: {t*1e6:7.1f} us  {3*e/t/1e9:7.0f} GB/s")
    buf = torch.zeros(6, c, device="cuda")
    t = timeit(lambda: ops.bn_act_bwd(dy, raw, scale, shift, mean, invstd, None, "leaky", buf=buf))
    print(f"bn_bwd (red+app)  {c:5d}@{hw:3d}: {t*1e6:7.1f} us  {5*e/t/1e9:7.0f} GB/s")
