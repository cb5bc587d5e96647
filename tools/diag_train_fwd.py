"""Diagnostic (GPU box): layer-by-layer train-mode forward error of the MINI trunk vs the oracle,
in fp32-oracle and bf16-rounding-faithful-oracle variants, to tell bf16 noise from bugs."""
import copy, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import test_gpu_models as T
from oracle import oracle as O

def rel(a, b): return ((a - b).norm() / b.norm()).item()

for train in (False, True):
    model, hp = T.make("BaselineModel", T.MINI)
    model.train(train)
    x = T.synth_input(8, 128)
    sd = copy.deepcopy(model.state_dict())
    taps = {}
    with torch.no_grad():
        O.darknet_forward(x, sd, T.MINI, train=train, taps=taps)
    model = model.to("cuda")
    model._debug_taps = {}
    with torch.no_grad():
        model(x.cuda())
    # oracle taps are keyed by reference layer index too
    print(f"--- train={train}")
    for k in sorted(model._debug_taps, key=lambda s: int(s.split('_')[1])):
        if k in taps and taps[k].shape == model._debug_taps[k].shape:
            print(f"{k:>10s} {tuple(taps[k].shape)}  rel_l2={rel(model._debug_taps[k], taps[k]):.4f}")
