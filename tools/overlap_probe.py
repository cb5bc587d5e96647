"""GPU-box probe: do the weight-gradient kernel and the BatchNorm-backward streaming kernels overlap on two streams?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_uav_det_b200 import ops

n, hw = 32, 80
x = torch.randn(n, hw, hw, 128, device="cuda").bfloat16()
dy = torch.randn(n, hw, hw, 256, device="cuda").bfloat16()
out = torch.zeros(256, 9 * 128, device="cuda")
raw = torch.randn(n, hw, hw, 256, device="cuda").bfloat16()
d2 = torch.randn(n, hw, hw, 256, device="cuda").bfloat16()
c = 256
scale = torch.rand(c, device="cuda") + 0.5; shift = torch.randn(c, device="cuda")
mean = torch.randn(c, device="cuda"); invstd = torch.rand(c, device="cuda") + 0.5
side = torch.cuda.Stream(priority=-1)

def wg(): ops.conv_wgrad(x, dy, 3, 1, 1, out=out)
def bn(): ops.bn_act_bwd(d2, raw, scale, shift, mean, invstd, None, "leaky")

def timeit(fn, reps=20, inner=8):
    """Capture `inner` back-to-back invocations into a CUDA graph (no host launch gaps) and time replays."""
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(inner): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * inner)

def both():
    main = torch.cuda.current_stream()
    side.wait_stream(main)
    with torch.cuda.stream(side):
        wg()
    bn(); bn()
    main.wait_stream(side)

t_w, t_b = timeit(wg), timeit(lambda: (bn(), bn()))
t_both = timeit(both)
print(f"wgrad alone {t_w:.1f} us, 2x bn_bwd alone {t_b:.1f} us, concurrent {t_both:.1f} us (serial sum {t_w + t_b:.1f})")
