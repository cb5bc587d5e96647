"""Per-layer table: cuDNN (stock PyTorch F.conv2d, bf16, channels_last, cudnn.benchmark) against this repo's
implicit-GEMM kernels, for the distinct convolution shapes of BaselineModel at batch 32 (SURVEY.md §8a shape table).
forward / data gradient / weight gradient are timed separately with CUDA events; the operands rotate through enough
buffers that no launch finds its input in L2 (> 126 MB between reuses).

    python tools/bench_conv_layers.py [--batch 32] > profiles/r02_torch_cudnn_per_layer.json
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from multimodal_uav_det_b200 import ops

# (cin, cout, k, stride, input side) — distinct Baseline shapes (cin >= 32; the cin=3 stem is timed by bench.py)
SHAPES = [
    (32, 64, 3, 2, 640), (64, 32, 1, 1, 320), (32, 64, 3, 1, 320), (64, 128, 3, 2, 320), (128, 64, 1, 1, 160),
    (64, 128, 3, 1, 160), (128, 256, 3, 2, 160), (256, 128, 1, 1, 80), (128, 256, 3, 1, 80), (256, 512, 3, 2, 80),
    (512, 256, 1, 1, 40), (256, 512, 3, 1, 40), (512, 1024, 3, 2, 40), (1024, 512, 1, 1, 20), (512, 1024, 3, 1, 20),
    (512, 256, 1, 1, 20), (768, 256, 1, 1, 40), (256, 128, 1, 1, 40), (384, 128, 1, 1, 80),
]


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3      # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--ours-only", action="store_true", help="skip the cuDNN side (A/B runs of this repo's kernels)")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True
    n = args.batch
    rows = []
    for cin, cout, k, s, side in SHAPES:
        pad = 1 if k == 3 else 0
        ho = (side + 2 * pad - k) // s + 1
        in_bytes = n * side * side * cin * 2
        out_bytes = n * ho * ho * cout * 2
        nbuf = max(2, min(8, int(200e6 // max(in_bytes, 1)) + 2))
        g = torch.Generator(device=dev).manual_seed(0)
        xs = [torch.randn(n, side, side, cin, device=dev, generator=g).to(torch.bfloat16) for _ in range(nbuf)]
        dys = [torch.randn(n, ho, ho, cout, device=dev, generator=g).to(torch.bfloat16) for _ in range(nbuf)]
        w = torch.randn(cout, cin, k, k, device=dev, generator=g) * 0.05
        flops = 2.0 * n * ho * ho * cout * cin * k * k
        # ---- cuDNN ----
        w_cl = w.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        xs_t = [x.permute(0, 3, 1, 2) for x in xs]          # NCHW view of NHWC storage = channels_last
        dys_t = [d.permute(0, 3, 1, 2) for d in dys]
        it = [0]

        def t_fwd():
            it[0] += 1
            return F.conv2d(xs_t[it[0] % nbuf], w_cl, None, s, pad)

        def t_dgrad():
            it[0] += 1
            return torch.ops.aten.convolution_backward(dys_t[it[0] % nbuf], xs_t[it[0] % nbuf], w_cl, None, (s, s), (pad, pad),
                                                       (1, 1), False, (0, 0), 1, (True, False, False))[0]

        def t_wgrad():
            it[0] += 1
            return torch.ops.aten.convolution_backward(dys_t[it[0] % nbuf], xs_t[it[0] % nbuf], w_cl, None, (s, s), (pad, pad),
                                                       (1, 1), False, (0, 0), 1, (False, True, False))[1]

        cud = dict(fwd=float("nan"), dgrad=float("nan"), wgrad=float("nan")) if args.ours_only else \
            dict(fwd=timed(t_fwd), dgrad=timed(t_dgrad), wgrad=timed(t_wgrad))
        # ---- ours ----
        wp = ops.pack_weight(w)
        wt = ops.pack_weight(w, transposed=True)
        outs = [ops.empty_act(n, ho, ho, cout, dev) for _ in range(nbuf)]
        dxs = [ops.empty_act(n, side, side, cin, dev) for _ in range(nbuf)]
        dw = torch.zeros((cout, k * k * cin), dtype=torch.float32, device=dev)

        def o_fwd():
            it[0] += 1
            return ops.conv_fwd(xs[it[0] % nbuf], wp, cout, k, s, pad, out=outs[it[0] % nbuf])

        def o_dgrad():
            it[0] += 1
            return ops.conv_dgrad(dys[it[0] % nbuf], wt, cin, k, s, pad, (side, side), out=dxs[it[0] % nbuf])

        def o_wgrad():
            it[0] += 1
            return ops.conv_wgrad(xs[it[0] % nbuf], dys[it[0] % nbuf], k, s, pad, out=dw)

        ours = dict(fwd=timed(o_fwd), dgrad=timed(o_dgrad), wgrad=timed(o_wgrad))
        # numerics of the two forward results against each other (same bf16 inputs)
        ref = F.conv2d(xs_t[0], w_cl, None, s, pad).permute(0, 2, 3, 1).float()
        got = ops.conv_fwd(xs[0], wp, cout, k, s, pad).float()
        rel = ((got - ref).norm() / ref.norm()).item()
        hbm_us = (in_bytes + out_bytes) / 6543.7e9 * 1e6
        tc_us = flops / 1399.9e12 * 1e6
        rows.append(dict(shape=f"{cin}->{cout} k{k}/s{s} @{side}", gflop=flops / 1e9, roofline_us=max(hbm_us, tc_us),
                         cudnn_us=cud, ours_us=ours, fwd_rel_l2_vs_cudnn=rel,
                         speedup={kk: cud[kk] / ours[kk] for kk in cud}))
        print(json.dumps(rows[-1]), file=sys.stderr, flush=True)
        del xs, dys, xs_t, dys_t, outs, dxs
        torch.cuda.empty_cache()
    ops.check_device()
    tot = {kk: (sum(r["cudnn_us"][kk] for r in rows), sum(r["ours_us"][kk] for r in rows)) for kk in ("fwd", "dgrad", "wgrad")}
    print(json.dumps(dict(batch=n, torch=torch.__version__, cudnn=torch.backends.cudnn.version(), rows=rows,
                          totals_us={kk: dict(cudnn=v[0], ours=v[1]) for kk, v in tot.items()}), indent=1))


if __name__ == "__main__":
    main()
