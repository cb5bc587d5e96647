"""Diagnostic (GPU box): layer-by-layer train-mode forward error of a trunk vs the oracle evaluated
with the same bf16 storage points, to localise train-path discrepancies."""
import copy, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import test_gpu_models as T
from oracle import oracle as O

def rel(a, b): return ((a - b).norm() / b.norm()).item()

cfg = T.SHALLOW
for train in (True,):
    model, hp = T.make("BaselineModel", cfg)
    model.route_repeats = 2
    model.train(train)
    x = T.synth_input(16, 128)
    sd = copy.deepcopy(model.state_dict())
    taps = {}
    with torch.no_grad(), O.bf16_pipeline(True):
        O.darknet_forward(x, sd, cfg, train=train, taps=taps, route_repeats=2)
    model = model.to("cuda")
    model._debug_taps = {}
    with torch.no_grad():
        model(x.cuda())
    print(f"--- train={train} (faithful oracle)")
    for k in sorted(model._debug_taps, key=lambda s: int(s.split('_')[1])):
        if k in taps and taps[k].shape == model._debug_taps[k].shape:
            a, b = model._debug_taps[k], taps[k]
            print(f"{k:>10s} {tuple(b.shape)}  rel_l2={rel(a, b):.5f}  mean got/ref={a.mean():.5f}/{b.mean():.5f} "
                  f"std got/ref={a.std():.5f}/{b.std():.5f}")
    # single-layer check: conv(3x3 s2)+BN(train)+leaky on identical bf16 input
    import torch.nn.functional as F
    from multimodal_uav_det_b200 import ops
    from multimodal_uav_det_b200._lib import EPI_STATS
    g = torch.Generator().manual_seed(5)
    for (cin, cout, k, s, hw) in [(64, 128, 3, 2, 64), (64, 128, 3, 1, 64), (64, 32, 1, 1, 64), (128, 256, 3, 2, 32)]:
        xin = F.leaky_relu(torch.randn(16, cin, hw, hw, generator=g), 0.1).bfloat16().float()
        w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).bfloat16().float()
        raw = F.conv2d(xin, w, None, s, k // 2)
        sums = torch.zeros(2, cout, device="cuda")
        y = ops.conv_fwd(xin.permute(0, 2, 3, 1).contiguous().bfloat16().cuda(), ops.pack_weight(w.cuda()), cout, k, s, k // 2,
                         epi=EPI_STATS, sum_=sums[0], sumsq=sums[1])
        torch.cuda.synchronize()
        s1, s2 = raw.sum(dim=(0, 2, 3)), (raw * raw).sum(dim=(0, 2, 3))
        got_raw = y.float().cpu().permute(0, 3, 1, 2)
        print(f"layer {cin}->{cout} k{k} s{s}: raw rel={rel(got_raw, raw):.5f} sum rel={rel(sums[0].cpu(), s1):.6f} "
              f"sumsq rel={rel(sums[1].cpu(), s2):.6f}  max|dsum|={(sums[0].cpu()-s1).abs().max():.4f} of |sum|~{s1.abs().mean():.2f}")
