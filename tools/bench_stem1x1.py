"""GPU-box diagnostic: the streaming 1x1 stems (DySOEM_SimFPN InputStemLayer / AdaptiveStemLayer) at the configs[3] shape
(batch 64, 640 x 640), CUDA-event timed over rotating inputs larger than L2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_uav_det_b200 import ops
from multimodal_uav_det_b200._lib import EPI_STATS
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
for cin in (3, 1):
    xs = [torch.rand(n, cin, 640, 640, device="cuda") for _ in range(2)]
    w = torch.randn(32, cin, 1, 1, device="cuda") * 0.1
    s1 = torch.zeros(32, device="cuda"); s2 = torch.zeros(32, device="cuda")
    def timeit(fn, reps=6):
        for i in range(2): fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps): fn(i)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1000
    mb = (n * cin * 640 * 640 * 4 + n * 640 * 640 * 64) / 1e6
    t = timeit(lambda i: ops.stem_fwd(xs[i % 2], w, 1, 1, 0, epi=EPI_STATS, sum_=s1, sumsq=s2))
    print(f"cin={cin} stem1x1 fwd stats: {t:.0f} us  {mb / t:.2f} TB/s")
    t = timeit(lambda i: ops.stem_fwd(xs[i % 2], w, 1, 1, 0, act="leaky"))
    print(f"cin={cin} stem1x1 fwd affine: {t:.0f} us  {mb / t:.2f} TB/s")
