"""GPU-box diagnostic: the tensor-core stem kernels (stem_mma.cu) alone at the configs[1] shape, CUDA-event timed with
rotating inputs.  UAVDET_STEM_DBG (bit 1: no global loads, 2: no MMA, 4: no store) switches parts of the forward off."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_uav_det_b200 import ops
from multimodal_uav_det_b200._lib import EPI_STATS
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
xs = [torch.rand(n, 3, 640, 640, device="cuda") for _ in range(3)]
w = (torch.randn(32, 32, device="cuda") * 0.1).to(torch.bfloat16)
outs = [torch.empty(n, 640, 640, 32, dtype=torch.bfloat16, device="cuda") for _ in range(3)]
s1 = torch.zeros(32, device="cuda"); s2 = torch.zeros(32, device="cuda")
dw = torch.zeros(32, 32, device="cuda")
def timeit(fn, reps=6):
    for i in range(2): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1000
mb = (n * 3 * 640 * 640 * 4 + n * 640 * 640 * 64) / 1e6
t = timeit(lambda i: ops.stem_mma_fwd(xs[i % 3], w, 3, 1, 1, epi=EPI_STATS, sum_=s1, sumsq=s2, out=outs[i % 3]))
print(f"dbg={os.environ.get('UAVDET_STEM_DBG', '0')} stem_mma_fwd stats: {t:.0f} us  {mb / t / 1e6 * 1e6:.2f} TB/s")
t = timeit(lambda i: ops.stem_mma_fwd(xs[i % 3], w, 3, 1, 1, act="leaky", out=outs[i % 3]))
print(f"dbg={os.environ.get('UAVDET_STEM_DBG', '0')} stem_mma_fwd affine: {t:.0f} us  {mb / t:.2f} TB/s")
t = timeit(lambda i: ops.stem_mma_wgrad(xs[i % 3], outs[i % 3], 3, 1, 1, out=dw))
print(f"dbg={os.environ.get('UAVDET_STEM_DBG', '0')} stem_mma_wgrad: {t:.0f} us  {mb / t:.2f} TB/s")
