"""GPU-box diagnostic: per-role pipeline timeline (clock64) of CTA 0 of igemm_kernel for a few layer shapes."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_uav_det_b200 import _lib, ops
from multimodal_uav_det_b200._lib import EPI_STATS
lib = _lib.load()
lib.uavdet_debug_set_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
lib.uavdet_debug_set_trace.restype = None
NT = 6
trace = torch.zeros(NT * 16, dtype=torch.int64, device="cuda")
NT = 10
trace = torch.zeros(NT * 16, dtype=torch.int64, device="cuda")
CASES = [(8, 32, 32, 1, 1, 640, True), (32, 64, 32, 1, 1, 320, True), (32, 32, 64, 3, 1, 320, True)]
if len(sys.argv) > 1 and sys.argv[1] == "rtm":      # the channel MLPs of RTMUAVDet's MDyEncoder (1x1, GELU / none)
    CASES = [(32, 192, 192, 1, 1, 160, "gelu"), (32, 192, 192, 1, 1, 160, "none"), (32, 192, 128, 1, 1, 160, "none"),
             (32, 384, 384, 1, 1, 80, "gelu"), (32, 256, 256, 1, 1, 160, "gelu"), (32, 128, 128, 1, 1, 160, "gelu")]
if len(sys.argv) > 1 and sys.argv[1] == "rtm2":     # MDyConv's 32 -> 128 1x1 (ReLU) and its neighbours
    CASES = [(32, 32, 128, 1, 1, 160, "relu"), (32, 32, 128, 1, 1, 160, "none"), (32, 32, 128, 1, 1, 160, "silu"),
             (32, 64, 32, 1, 1, 160, "silu"), (32, 128, 32, 1, 1, 160, "silu"), (32, 64, 128, 1, 1, 160, "relu")]
FOLD = len(sys.argv) > 1 and sys.argv[1] == "rtmfold"   # channel-MLP GELU conv: plain / GroupNorm-fold epilogue (sample_affine)
if FOLD:
    CASES = [(32, 192, 192, 1, 1, 160, "gelu"), (32, 192, 192, 1, 1, 160, "gelu+fold"), (32, 192, 192, 1, 1, 160, "relu"),
             (32, 192, 192, 1, 1, 160, "relu+fold")]
for (n, cin, cout, k, s, hw, stats) in CASES:
    x = torch.randn(n, hw, hw, cin, device="cuda").bfloat16()
    w = ops.pack_weight(torch.randn(cout, cin, k, k, device="cuda") * 0.05)
    s1 = torch.zeros(cout, device="cuda"); s2 = torch.zeros(cout, device="cuda")
    kw = dict(epi=EPI_STATS, sum_=s1, sumsq=s2) if stats is True else dict(act=stats if isinstance(stats, str) else "leaky")
    if stats is not True:      # eval-mode BatchNorm folded into the epilogue, as the inference models launch it
        kw.update(scale=torch.rand(cout, device="cuda") + 0.5, shift=torch.randn(cout, device="cuda") * 0.1)
    if isinstance(stats, str) and stats.endswith("+fold"):
        kw["act"] = stats[:-5]
        kw["sample_affine"] = torch.stack([torch.rand(n, device="cuda") + 0.5, torch.randn(n, device="cuda")], 1).contiguous()
    for _ in range(2):
        ops.conv_fwd(x, w, cout, k, s, k // 2, **kw)
    torch.cuda.synchronize()
    trace.zero_()
    lib.uavdet_debug_set_trace(trace.data_ptr(), NT)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.conv_fwd(x, w, cout, k, s, k // 2, **kw); e1.record()
    torch.cuda.synchronize()
    lib.uavdet_debug_set_trace(None, 0)
    t = trace.cpu().view(NT, 16)
    t0 = int(t[0, 2])
    print(f"=== {cin}->{cout} k{k} s{s} @{hw} stats={stats}: kernel {e0.elapsed_time(e1)*1000:.0f} us; cycles rel. to first TMA issue")
    print(" tile | prod_start prod_end | mma_wait_tempty mma_first_full mma_commit | epi_wait epi_start epi_end")
    for i in range(NT):
        if int(t[i, 2]) == 0: break
        r = [int(v) - t0 for v in t[i]]
        e = r[6]
        e = r[6]
        fine = " ".join(f"{(r[j] - e) if int(t[i, j]) else -1:6d}" for j in range(8, 13))
        print(f" {i:4d} | {r[2]:9d} {r[4]:9d} | {r[5]:9d} {r[6]:9d} {r[7]:9d} (unit {r[7]-r[6]}) | rel epi_start: unit_begin after_waitrd after_sts after_fence after_store: {fine}")
