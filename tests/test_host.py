"""CPU tests of the host side: the C-ABI library loads and exports every symbol the header declares,
the product refuses to run without CUDA (no fallback), drop-in API surface, the batched loss against
the reference's golden numbers, and the data-parallel trainer over gloo (world_size 2)."""
import copy
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def load(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


def golden_logits(seed, batch, grids):
    g = torch.Generator().manual_seed(seed)
    return [(torch.randn(batch, 3, s, s, 4, generator=g), torch.randn(batch, 3, s, s, 1, generator=g)) for s in grids]


# ---- C-ABI ------------------------------------------------------------------------------------------
def test_library_exports_every_symbol_declared_in_the_header(lib):
    from multimodal_uav_det_b200 import _lib
    header = open(os.path.join(ROOT, "include", "uavdet_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(uavdet_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 30
    nm = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (uavdet_[a-z0-9_]+)", nm))
    assert declared <= exported, f"declared but not exported: {sorted(declared - exported)}"
    assert declared == set(_lib.SIGNATURES), f"ctypes table out of sync: {sorted(declared ^ set(_lib.SIGNATURES))}"
    assert lib.uavdet_version() >= 100
    assert lib.uavdet_last_error() is not None


def test_public_header_is_plain_c():
    """include/uavdet_b200.h is the drop-in boundary: it must compile on its own as C11 and as C++ (no torch, no
    CUDA types in the signatures)."""
    hdr = os.path.join(ROOT, "include", "uavdet_b200.h")
    for cmd in (["gcc", "-x", "c", "-std=c11", "-fsyntax-only", "-Wall", "-Werror", hdr],
                ["g++", "-x", "c++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", hdr]):
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    code = re.sub(r"/\*.*?\*/", "", open(hdr).read(), flags=re.S)      # declarations only, comments stripped
    assert "torch" not in code.lower() and "cudaStream_t" not in code and "#include <cuda" not in code


def test_library_is_cuda_only_sm100a(lib):
    from multimodal_uav_det_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


def test_no_cpu_fallback():
    from multimodal_uav_det_b200 import ops
    from multimodal_uav_det_b200._lib import UavdetError
    from multimodal_uav_det_b200.model import BaselineModel
    from multimodal_uav_det_b200.utils.datatype import Config
    gold = load("model_forwards.pt")["baseline"]
    model = BaselineModel(hparams=Config(gold["hp"]))
    with pytest.raises(RuntimeError):
        model(torch.rand(1, 3, 64, 64))
    with pytest.raises(UavdetError):
        ops.nms(torch.zeros(4, 4), torch.zeros(4), 0.5)
    with pytest.raises(UavdetError):
        ops.encode_targets(torch.tensor([[10.0, 10.0, 50.0, 40.0]]), gold["hp"]["anchors"], [2, 4, 8], 64)
    # the product package must never import the oracle
    for dirpath, _, files in os.walk(os.path.join(ROOT, "multimodal_uav_det_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f"{f} imports the oracle"


# ---- drop-in surface ------------------------------------------------------------------------------------
def test_model_api_surface_matches_reference_contract():
    from multimodal_uav_det_b200 import model as M
    from multimodal_uav_det_b200.model import _base
    from multimodal_uav_det_b200.utils.datatype import BatchData, Config, DetectionResults
    assert DetectionResults._fields == ("bbox", "obj") and BatchData._fields == ("image", "bbox")
    gold = load("model_forwards.pt")
    for name, cls in (("baseline", M.BaselineModel), ("dy-yolo", M.DyYOLO)):
        hp = gold[name]["hp"]
        m = cls(hparams=Config(hp))
        for attr in ("forward", "training_step", "validation_step", "configure_optimizers", "log"):
            assert callable(getattr(m, attr))
        opt = m.configure_optimizers()
        assert isinstance(opt, torch.optim.SGD) and opt.defaults["momentum"] == hp["optim"]["momentum"]
        assert isinstance(m.yolo_head, _base.YOLOHead) and not any(k.startswith("yolo_head.anchors") for k in m.state_dict())
    bad = dict(gold["baseline"]["hp"], optim=dict(name="RMSprop", momentum=0.0))
    with pytest.raises(ValueError):
        M.BaselineModel(hparams=Config(bad)).configure_optimizers()
    with pytest.raises(ValueError):
        M.BaselineModel(hparams=Config(gold["dy-yolo"]["hp"]))     # DyConv layers need DyYOLO
    cfg = Config({"a": 1, "b": {"c": [1, 2], "d": {"e": "x"}}})
    assert cfg.a == 1 and cfg.b.c == [1, 2] and cfg.b.d.e == "x"


def test_calculate_iou_matches_golden_including_first_target_quirk():
    from multimodal_uav_det_b200.utils.postprocess import calculate_iou
    g = load("blocks.pt")["calculate_iou"]
    for mode in ("mse", "ciou"):
        got = calculate_iou(g["preds"], g["targets"], g["anchors"], g["mask"], mode)
        torch.testing.assert_close(got, g[mode], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("name", ["baseline", "dy-yolo"])
def test_batched_loss_matches_reference_golden(name):
    """YOLOHead.compute_metrics (batched, sync-free) == the reference's per-sample loop: value,
    gradients w.r.t. every logit, and the in-place rewrite of batch.bbox."""
    from multimodal_uav_det_b200.model._base import YOLOHead
    from multimodal_uav_det_b200.utils.datatype import BatchData, Config, DetectionResults
    gold = load("head_loss_decode.pt")[name]
    hp = gold["hp"]
    head = YOLOHead([8, 8, 8], hp["anchors"], hp["head_scales"], Config(hp["loss_balancing"]), hp["bbox_loss_fn"])
    grids = [gold["size"] // s for s in hp["head_scales"]]
    logits = golden_logits(gold["logits_seed"], 3, grids)
    outs = [DetectionResults(bbox=b.clone().requires_grad_(True), obj=o.clone().requires_grad_(True)) for b, o in logits]
    batch = BatchData(image=torch.zeros(3, 3, 8, 8), bbox=copy.deepcopy(gold["targets"]))
    loss, ap, bl, ol = head.compute_metrics(outs, batch)
    loss.backward()
    assert ap is None
    torch.testing.assert_close(loss.detach(), gold["loss"], rtol=2e-6, atol=1e-6)
    torch.testing.assert_close(bl.detach(), gold["bbox_loss"], rtol=2e-6, atol=1e-6)
    torch.testing.assert_close(ol.detach(), gold["obj_loss"], rtol=2e-6, atol=1e-6)
    for o, (gb, go) in zip(outs, gold["grads"]):
        torch.testing.assert_close(o.bbox.grad, gb, rtol=1e-4, atol=1e-7)
        torch.testing.assert_close(o.obj.grad, go, rtol=1e-4, atol=1e-7)
    for per, gper in zip(batch.bbox, gold["mutated_targets"]):
        for t, gt in zip(per, gper):
            torch.testing.assert_close(t, gt, rtol=1e-6, atol=1e-7)
    # pre-stacked targets (the bench / fast path) give the same loss
    stacked = [torch.stack([gold["targets"][i][h] for i in range(3)]) for h in range(3)]
    head.mutate_targets = False
    outs2 = [DetectionResults(bbox=b, obj=o) for b, o in logits]
    loss2, _, _, _ = head.compute_metrics(outs2, BatchData(image=torch.zeros(3, 3, 8, 8), bbox=stacked))
    torch.testing.assert_close(loss2, gold["loss"], rtol=2e-6, atol=1e-6)


# ---- data-parallel trainer over gloo -----------------------------------------------------------------------
_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from multimodal_uav_det_b200.parallel import FlatSGDTrainer
rank = int(os.environ["RANK"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=2)
torch.manual_seed(0)
model = torch.nn.Sequential(torch.nn.Linear(37, 19), torch.nn.BatchNorm1d(19), torch.nn.Linear(19, 5))
ref = [p.detach().clone() for p in model.parameters()]
trainer = FlatSGDTrainer(model, lr=0.1, momentum=0.7, bucket_mb=0.001)
assert len(trainer.buckets) > 1
bufs = [torch.zeros_like(p) for p in ref]
for step in range(3):
    trainer.zero_grad()
    g = torch.Generator().manual_seed(100 * step + rank)
    grads = []
    for p in model.parameters():
        gr = torch.randn(p.shape, generator=g)
        p.grad.add_(gr)                 # accumulate into the arena view, as the executor does
        trainer._on_grad_ready(p)      # grad_ready_hook
        grads.append(gr)
    trainer.step()
    # expected: SGD(momentum) on the mean gradient of both ranks
    for i in range(len(ref)):
        g0 = torch.randn(ref[i].shape, generator=torch.Generator().manual_seed(100 * step + 0))
        g1 = torch.randn(ref[i].shape, generator=torch.Generator().manual_seed(100 * step + 1))
    gens = [torch.Generator().manual_seed(100 * step + r) for r in range(2)]
    for i in range(len(ref)):
        mean_g = sum(torch.randn(ref[i].shape, generator=gens[r]) for r in range(2)) / 2
        bufs[i] = mean_g if step == 0 else 0.7 * bufs[i] + mean_g
        ref[i] = ref[i] - 0.1 * bufs[i]
for p, r in zip(model.parameters(), ref):
    assert torch.allclose(p.detach(), r, rtol=1e-5, atol=1e-6), (p.detach() - r).abs().max()
# parameters are views of the flat arenas and identical on both ranks
flat = torch.cat([b.param for b in trainer.buckets])
other = [torch.zeros_like(flat) for _ in range(2)]
dist.all_gather(other, flat)
assert torch.equal(other[0], other[1])
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_flat_sgd_trainer_gloo_world_size_2(tmp_path):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(os.environ, RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{o}"
        assert f"rank {r} ok" in o


def test_flat_sgd_trainer_single_process_matches_torch_sgd():
    from multimodal_uav_det_b200.parallel import FlatSGDTrainer
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(11, 7), torch.nn.Linear(7, 3))
    twin = copy.deepcopy(model)
    opt = torch.optim.SGD(twin.parameters(), lr=0.05, momentum=0.78)
    trainer = FlatSGDTrainer(model, lr=0.05, momentum=0.78)
    for step in range(4):
        trainer.zero_grad()
        opt.zero_grad()
        g = torch.Generator().manual_seed(step)
        for p, q in zip(model.parameters(), twin.parameters()):
            gr = torch.randn(p.shape, generator=g)
            p.grad.add_(gr)
            q.grad = gr.clone()
        trainer.step()
        opt.step()
    for p, q in zip(model.parameters(), twin.parameters()):
        torch.testing.assert_close(p.detach(), q.detach(), rtol=1e-6, atol=1e-7)
    # state_dict still works on the re-homed parameters
    assert set(model.state_dict()) == set(twin.state_dict())


# ---- gradient accumulation (reference params.yaml:23 accumulate_grad_batches: 2) -----------------------------------
_WORKER_ACCUM = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from multimodal_uav_det_b200.parallel import FlatSGDTrainer
rank = int(os.environ["RANK"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=2)
torch.manual_seed(0)
model = torch.nn.Sequential(torch.nn.Linear(23, 17), torch.nn.Linear(17, 5))
ref = [p.detach().clone() for p in model.parameters()]
K = 2
trainer = FlatSGDTrainer(model, lr=0.1, momentum=0.7, bucket_mb=0.0005, accumulate_grad_batches=K)
assert len(trainer.buckets) > 1
bufs = [torch.zeros_like(p) for p in ref]
seed = lambda step, micro, r: 1000 * step + 10 * micro + r
for step in range(3):
    trainer.zero_grad()
    for micro in range(K):
        g = torch.Generator().manual_seed(seed(step, micro, rank))
        for p in reversed(list(model.parameters())):
            pass
        for p in model.parameters():
            p.grad.add_(torch.randn(p.shape, generator=g))
            trainer._on_grad_ready(p)
        launched = [b.work is not None for b in trainer.buckets]
        # reduce only on the boundary micro-batch
        assert all(launched) if micro == K - 1 else not any(launched), (micro, launched)
    trainer.step()
    gens = {{(m, r): torch.Generator().manual_seed(seed(step, m, r)) for m in range(K) for r in range(2)}}
    for i in range(len(ref)):
        mean_g = sum(torch.randn(ref[i].shape, generator=gens[(m, r)]) for m in range(K) for r in range(2)) / (2 * K)
        bufs[i] = 0.7 * bufs[i] + mean_g
        ref[i] = ref[i] - 0.1 * bufs[i]
for p, r in zip(model.parameters(), ref):
    assert torch.allclose(p.detach(), r, rtol=1e-5, atol=1e-6), (p.detach() - r).abs().max()
# a third backward before step() is an error, not a silent race with the in-flight reduce
trainer.zero_grad()
for micro in range(K):
    for p in model.parameters():
        trainer._on_grad_ready(p)
try:
    trainer._on_grad_ready(next(iter(model.parameters())))
    raise SystemExit("missing guard")
except RuntimeError as e:
    assert "accumulate_grad_batches" in str(e)
trainer.step()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_flat_sgd_trainer_gradient_accumulation_gloo_world_size_2(tmp_path):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker_accum.py"
    script.write_text(_WORKER_ACCUM.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(os.environ, RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{o}"
        assert f"rank {r} ok" in o


def test_flat_sgd_trainer_accumulation_single_process_matches_torch_sgd_on_mean_gradient():
    from multimodal_uav_det_b200.parallel import FlatSGDTrainer
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(11, 7), torch.nn.Linear(7, 3))
    twin = copy.deepcopy(model)
    opt = torch.optim.SGD(twin.parameters(), lr=0.05, momentum=0.78)
    trainer = FlatSGDTrainer(model, lr=0.05, momentum=0.78, accumulate_grad_batches=2)
    for step in range(3):
        trainer.zero_grad()
        opt.zero_grad()
        for micro in range(2):
            g = torch.Generator().manual_seed(10 * step + micro)
            for p, q in zip(model.parameters(), twin.parameters()):
                gr = torch.randn(p.shape, generator=g)
                p.grad.add_(gr)
                trainer._on_grad_ready(p)
                q.grad = gr / 2 if q.grad is None else q.grad + gr / 2      # Lightning: loss / accumulate_grad_batches
        trainer.step()
        opt.step()
    for p, q in zip(model.parameters(), twin.parameters()):
        torch.testing.assert_close(p.detach(), q.detach(), rtol=1e-6, atol=1e-7)


def test_cyclic_lr_matches_torch_scheduler_and_trainer_lr_is_settable():
    """FlatSGDTrainer.cyclic_lr == torch.optim.lr_scheduler.CyclicLR as configured by the reference
    (_base.py:299-309: base lr/10, max lr, step_size_up 4000, triangular2, cycle_momentum False)."""
    from multimodal_uav_det_b200.parallel import FlatSGDTrainer
    lr = 1e-4
    w = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([w], lr=lr, momentum=0.7)
    sched = torch.optim.lr_scheduler.CyclicLR(opt, base_lr=lr / 10, max_lr=lr, step_size_up=40, mode="triangular2",
                                              cycle_momentum=False)
    for step in range(400):
        want = opt.param_groups[0]["lr"]
        got = FlatSGDTrainer.cyclic_lr(step, lr / 10, lr, 40, "triangular2")
        assert abs(got - want) <= 1e-12 + 1e-9 * want, (step, got, want)
        opt.step()
        sched.step()
    model = torch.nn.Linear(3, 2)
    tr = FlatSGDTrainer(model, lr=0.5, momentum=0.0)
    before = model.weight.detach().clone()
    tr.zero_grad()
    model.weight.grad.add_(1.0)
    tr.lr = 0.25                       # a scheduler changes the rate between steps
    tr.step()
    torch.testing.assert_close(model.weight.detach(), before - 0.25)


def test_reference_named_state_dict_loads_strict_into_a_trainer_wrapped_model():
    """a checkpoint `state_dict` (reference key scheme, pinned by tests/test_oracle.py against the
    reference's key list) loads with strict=True AFTER FlatSGDTrainer re-homed the parameters into flat arenas with
    channels-last conv weights: values land in the arenas, the views stay attached, state_dict() round-trips."""
    from multimodal_uav_det_b200.model import BaselineModel
    from multimodal_uav_det_b200.parallel import FlatSGDTrainer
    from multimodal_uav_det_b200.utils.datatype import Config
    cfg = [[32, 3, 1], [64, 3, 2], ["B", 1], [128, 3, 2], ["S"]]
    hp = dict(anchors=[[[29, 23], [48, 30], [67, 38]]], head_scales=[4], lr=1e-4, lr_scheduler=False, bbox_loss_fn="ciou",
              loss_balancing=dict(obj_scales_w=[1.0], bbox_w=4.0, objectness_w=1.0, no_obj_w=4.0),
              optim=dict(name="SGD", momentum=0.7), layer_config=cfg)
    torch.manual_seed(3)
    ckpt_model = BaselineModel(hparams=Config(hp))
    for m in ckpt_model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_()
            m.num_batches_tracked.fill_(7)
    ckpt = {k: v.clone() for k, v in ckpt_model.state_dict().items()}
    torch.manual_seed(4)
    model = BaselineModel(hparams=Config(hp))
    trainer = FlatSGDTrainer(model, lr=1e-4, momentum=0.7)
    w = model.layers[1].conv.weight
    assert not w.is_contiguous() and w.permute(0, 2, 3, 1).is_contiguous()        # channels-last storage in the arena
    missing, unexpected = model.load_state_dict(ckpt, strict=True)
    assert not missing and not unexpected
    for k, v in model.state_dict().items():
        assert torch.equal(v, ckpt[k]), k
    lo = {id(b): (b.param.data_ptr(), b.param.data_ptr() + 4 * b.numel) for b in trainer.buckets}
    for p in model.parameters():
        b = trainer._bucket_of[id(p)]
        assert lo[id(b)][0] <= p.data_ptr() < lo[id(b)][1]                          # still a view of its arena
    assert not w.is_contiguous() and w.permute(0, 2, 3, 1).is_contiguous()
    # the arena itself carries the loaded values (what the fused optimiser and the weight pack read)
    b = trainer._bucket_of[id(w)]
    off = (w.data_ptr() - b.param.data_ptr()) // 4
    torch.testing.assert_close(b.param[off:off + w.numel()].view(w.shape[0], 3, 3, w.shape[1]).permute(0, 3, 1, 2),
                               ckpt["layers.1.conv.weight"])


# ---- host-side algebra of the pixel-pair GEMM and the GroupNorm fold (pure torch, no GPU) -----------------------------
@pytest.mark.parametrize("cin,cout,flip", [(32, 64, False), (64, 64, False), (128, 64, False), (64, 32, True)])
def test_pair_weight_matrix_reproduces_the_3x3_convolution(cin, cout, flip):
    """ops.pack_weight_pair builds the [2*cout][3*S*cin] matrix uavdet_conv3x3_pair_fwd multiplies with the shifted pixel
    columns (include/uavdet_b200.h): emulate that GEMM on the CPU and compare with F.conv2d (flip=True: with the data
    gradient of the transposed filter, RTMUAVDet.py:194 / BaselineModel.py:63-75)."""
    import torch.nn.functional as F
    from multimodal_uav_det_b200 import ops
    g = torch.Generator().manual_seed(5 + cin + cout)
    n, h, w = 2, 5, 8
    if not flip:
        wt = torch.randn(cout, cin, 3, 3, generator=g).bfloat16().float()
        x = torch.randn(n, cin, h, w, generator=g)
        want = F.conv2d(x, wt, None, 1, 1)
        wp = ops.pack_weight_pair(wt).float()
        k_in, n_out = cin, cout
    else:
        # forward layer cout_f -> cin_f = (cin -> cout here read as dy channels -> dx channels)
        wf = torch.randn(cin, cout, 3, 3, generator=g).bfloat16().float()          # forward weight [dy channels][dx channels]
        x = torch.randn(n, cin, h, w, generator=g)                                  # dy
        xin = torch.zeros(n, cout, h, w, requires_grad=True)
        F.conv2d(xin, wf, None, 1, 1).backward(x)
        want = xin.grad
        w_t = wf.permute(1, 2, 3, 0).reshape(cout, 9 * cin).bfloat16()              # the transposed pack [dx ch][ky][kx][dy ch]
        wp = ops.pack_weight_pair(w_t, flip=True).float()
        k_in, n_out = cin, cout
    shifts, first = ops.pair_weight_shifts(k_in)
    assert wp.shape == (2 * n_out, 3 * shifts * k_in)
    xp = F.pad(x, (8, 8, 1, 1))                                                      # zero padding = TMA's out-of-image fill
    cols = []
    for ky in range(3):
        for s in range(shifts):
            dx = first + s
            cols.append(xp[:, :, ky:ky + h, 8 + dx:8 + dx + w:2])                    # pixel 2j + dx of row r + ky - 1
    a = torch.stack(cols, 1).reshape(n, 3 * shifts * k_in, h, w // 2)               # K index = (ky*S + s)*cin + ci
    out = torch.einsum("ok,nkhw->nohw", wp, a).reshape(n, 2, n_out, h, w // 2)       # N index = px*cout + co
    got = torch.stack([out[:, 0], out[:, 1]], -1).reshape(n, n_out, h, w)            # interleave the two pixels of a pair
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-4)
    buf = ops.pack_weight_pair(wt if not flip else w_t, flip=flip)
    assert torch.equal(ops.pack_weight_pair(wt if not flip else w_t, out=buf.clone(), flip=flip), buf)


def test_mdy_encoder_groupnorm_fold_constants_reproduce_the_oracle():
    """MDyEncoder._folded (W' = diag(a) W diag(gamma), row sums, b) with the per-image (rstd, mean*rstd) epilogue
    y = act(acc*rstd + b - mean*rstd*rowsum(W')) is algebraically GroupNorm -> 1x1 conv (-> BatchNorm) -> activation
    (RTMUAVDet.py:163-177): evaluate the folded form in fp32 torch and compare with the oracle's operator form."""
    import torch.nn.functional as F
    from oracle import oracle as O
    from multimodal_uav_det_b200.engine import Executor
    from multimodal_uav_det_b200.model.RTMUAVDet import MDyEncoder
    torch.manual_seed(11)
    enc = MDyEncoder(96, 64).eval()
    g = torch.Generator().manual_seed(12)
    with torch.no_grad():
        for m in enc.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.1, generator=g); m.running_var.uniform_(0.7, 1.3, generator=g)
                m.weight.uniform_(0.7, 1.3, generator=g); m.bias.normal_(0, 0.1, generator=g)
            if isinstance(m, torch.nn.GroupNorm):
                m.weight.uniform_(0.5, 1.5, generator=g); m.bias.normal_(0, 0.2, generator=g)
    sd = {"e." + k: v for k, v in enc.state_dict().items()}
    x = torch.randn(2, 96, 6, 7, generator=g) * torch.tensor([0.5, 2.0]).view(2, 1, 1, 1) + torch.tensor([1.0, -0.5]).view(2, 1, 1, 1)
    k = enc._folded(Executor())
    t = enc.third

    def fold(v, wp, wg, b, eps, act):
        mu = v.mean(dim=(1, 2, 3), keepdim=True)
        rstd = (v.var(dim=(1, 2, 3), unbiased=False, keepdim=True) + eps).rsqrt()
        acc = F.conv2d(v, wp.float()[:, :, None, None])
        return act(acc * rstd + b.view(1, -1, 1, 1) - mu * rstd * wg.view(1, -1, 1, 1))

    with torch.no_grad():
        base = fold(x, k["wp_in"], k["wg_in"], k["b_in"], enc.group_norm_in.eps, F.relu)
        y = F.group_norm(x, 1, sd["e.group_norm_in.weight"], sd["e.group_norm_in.bias"], 1e-5)
        want_base = torch.cat([O.rtm_conv_module(y, sd, f"e.mdy_conv_{n}.base_conv", 1, 0, "relu", False, 1e-5, 0.1)
                               for n in ("1x1", "3x3", "5x5")], 1)
        assert base.shape == (2, 3 * t, 6, 7)
        torch.testing.assert_close(base, want_base, rtol=2e-2, atol=2e-2)            # W' is rounded to bf16
        v = torch.randn(2, 96, 6, 7, generator=g) + 0.3
        z = fold(v, k["wp_out"], k["wg_out"], k["b_out"], enc.group_norm_out.eps, F.gelu)
        want_z = F.gelu(F.conv2d(F.group_norm(v, 1, sd["e.group_norm_out.weight"], sd["e.group_norm_out.bias"], 1e-5),
                                 sd["e.channel_mlp.0.weight"], sd["e.channel_mlp.0.bias"]))
        torch.testing.assert_close(z, want_z, rtol=2e-2, atol=2e-2)
    assert enc._folded(Executor()) is k                                              # cached until a parameter changes
    with torch.no_grad():
        enc.group_norm_out.weight.mul_(1.5)
    assert enc._folded(Executor()) is not k
