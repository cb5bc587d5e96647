"""GPU-box diagnostic: per-role timeline (clock64 stamps of CTA 0) of igemm_kernel on the thin multi-tap layers, to be run
with UAVDET_IGEMM_HALO=0 and =1.  Columns are cycles relative to the MMA thread's first stamp."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_uav_det_b200 import _lib, ops
from multimodal_uav_det_b200._lib import EPI_STATS
lib = _lib.load()
lib.uavdet_debug_set_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
lib.uavdet_debug_set_trace.restype = None
NT = 24
trace = torch.zeros(NT * 16, dtype=torch.int64, device="cuda")
print("UAVDET_IGEMM_HALO =", os.environ.get("UAVDET_IGEMM_HALO", "(default 1)"))
for (n, cin, cout, k, s, hw, stats) in [(32, 32, 64, 3, 1, 320, True), (32, 64, 32, 3, 1, 320, False), (16, 64, 64, 3, 1, 320, True),
                                        (32, 32, 64, 3, 2, 640, True)]:
    x = torch.randn(n, hw, hw, cin, device="cuda").bfloat16()
    w = ops.pack_weight(torch.randn(cout, cin, k, k, device="cuda") * 0.05)
    s1 = torch.zeros(cout, device="cuda"); s2 = torch.zeros(cout, device="cuda")
    kw = dict(epi=EPI_STATS, sum_=s1, sumsq=s2) if stats else dict()
    for _ in range(2):
        ops.conv_fwd(x, w, cout, k, s, k // 2, **kw)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.conv_fwd(x, w, cout, k, s, k // 2, **kw); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1000)
    trace.zero_()
    lib.uavdet_debug_set_trace(trace.data_ptr(), NT)
    ops.conv_fwd(x, w, cout, k, s, k // 2, **kw)
    torch.cuda.synchronize()
    lib.uavdet_debug_set_trace(None, 0)
    t = trace.cpu().view(NT, 16)
    t0 = int(t[0, 2])
    print(f"=== {cin}->{cout} k{k} s{s} @{hw} n={n} stats={stats}: kernel {min(ts):.0f} us (min of 5, hot L2 where it fits)")
    print(" tile | mma_top after_tempty after_afull mma_commit | epi_before_wait epi_after_wait epi_end")
    for i in range(NT):
        if int(t[i, 2]) == 0: break
        r = [int(v) - t0 if int(v) else -1 for v in t[i]]
        print(f" {i:4d} | {r[2]:8d} {r[3]:8d} {r[13]:8d} {r[4]:8d} | {r[5]:8d} {r[6]:8d} {r[7]:8d}")
