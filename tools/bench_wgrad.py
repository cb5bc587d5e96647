"""GPU-box micro-benchmark: wgrad / igemm kernel time per layer shape (back-to-back launches, CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_uav_det_b200 import ops
from multimodal_uav_det_b200._lib import EPI_STATS

SHAPES = [  # n, cin, cout, k, stride, hw_in
    (32, 128, 256, 3, 1, 80), (32, 256, 512, 3, 1, 40), (32, 512, 1024, 3, 1, 20), (32, 64, 128, 3, 1, 160),
    (32, 32, 64, 3, 1, 320), (32, 32, 64, 3, 2, 640), (32, 256, 128, 1, 1, 80), (32, 64, 32, 1, 1, 320),
    (32, 128, 256, 3, 2, 160), (32, 1024, 512, 1, 1, 20)]
which = sys.argv[1] if len(sys.argv) > 1 else "wgrad"
if len(sys.argv) > 2:
    SHAPES = [SHAPES[int(i)] for i in sys.argv[2].split(",")]
reps = int(os.environ.get("REPS", "6"))
for (n, cin, cout, k, s, hw) in SHAPES:
    x = torch.randn(n, hw, hw, cin, device="cuda").bfloat16()
    ho = (hw + 2 * (k // 2) - k) // s + 1
    dy = torch.randn(n, ho, ho, cout, device="cuda").bfloat16()
    flops = 2.0 * n * ho * ho * cout * cin * k * k
    if which == "wgrad":
        out = torch.zeros(cout, k * k * cin, device="cuda")
        fn = lambda: ops.conv_wgrad(x, dy, k, s, k // 2, out=out)
    elif which == "fwd":
        w = ops.pack_weight(torch.randn(cout, cin, k, k, device="cuda") * 0.05)
        s1 = torch.zeros(cout, device="cuda"); s2 = torch.zeros(cout, device="cuda")
        y = torch.empty(n, ho, ho, cout, device="cuda", dtype=torch.bfloat16)
        fn = lambda: ops.conv_fwd(x, w, cout, k, s, k // 2, out=y, epi=EPI_STATS, sum_=s1, sumsq=s2)
    else:
        wt = ops.pack_weight(torch.randn(cout, cin, k, k, device="cuda") * 0.05, transposed=True)
        dx = torch.empty(n, hw, hw, cin, device="cuda", dtype=torch.bfloat16)
        res = torch.randn(n, hw, hw, cin, device="cuda").bfloat16() if which == "dgradres" else None
        fn = lambda: ops.conv_dgrad(dy, wt, cin, k, s, k // 2, (hw, hw), out=dx, res=res)
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    torch.cuda._sleep(int(2e7))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print(f"{which} {cin:5d}->{cout:5d} k{k} s{s} @{hw:4d}: {us:8.1f} us  {flops / us / 1e6:7.0f} TFLOP/s")
ops.check_device()
