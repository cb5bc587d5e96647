"""GPU-box diagnostic: epilogue timeline (clock64) of CTA 0 of igemm_kernel for a 1x1 data gradient with / without a
residual operand (the residual-block skip gradient)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_uav_det_b200 import _lib, ops
lib = _lib.load()
lib.uavdet_debug_set_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
lib.uavdet_debug_set_trace.restype = None
NT = 10
trace = torch.zeros(NT * 16, dtype=torch.int64, device="cuda")
for (n, cin, cout, hw, with_res) in [(32, 256, 128, 80, False), (32, 256, 128, 80, True), (32, 512, 256, 40, True)]:
    dy = torch.randn(n, hw, hw, cout, device="cuda").bfloat16()
    wt = ops.pack_weight(torch.randn(cout, cin, 1, 1, device="cuda") * 0.05, transposed=True)
    dx = torch.empty(n, hw, hw, cin, device="cuda", dtype=torch.bfloat16)
    res = torch.randn(n, hw, hw, cin, device="cuda").bfloat16() if with_res else None
    fn = lambda: ops.conv_dgrad(dy, wt, cin, 1, 1, 0, (hw, hw), out=dx, res=res)
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    trace.zero_()
    lib.uavdet_debug_set_trace(trace.data_ptr(), NT)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record()
    torch.cuda.synchronize()
    lib.uavdet_debug_set_trace(None, 0)
    t = trace.cpu().view(NT, 16)
    t0 = int(t[0, 2])
    print(f"=== dgrad {cout}->{cin} 1x1 @{hw} res={with_res}: kernel {e0.elapsed_time(e1)*1000:.0f} us; cycles rel. to first TMA issue")
    print(" tile | prod_start prod_end | epi_wait_begin epi_acc_ready epi_tile_end | last unit, rel. acc_ready: begin after_acquire+res after_staging after_fence after_store")
    for i in range(NT):
        if int(t[i, 2]) == 0: break
        r = [int(v) - t0 for v in t[i]]
        e = r[6]
        fine = " ".join(f"{(r[j] - e) if int(t[i, j]) else -1:6d}" for j in range(8, 13))
        print(f" {i:4d} | {r[2]:9d} {r[4]:9d} | {r[5]:9d} {r[6]:9d} {r[7]:9d} (epilogue {r[7]-r[6]}) | {fine}")
