"""GPU parity of whole models against the CPU oracle (oracle/oracle.py restates the reference's
forward in fp32 from a reference-named state_dict; it is itself pinned to the reference in
tests/test_oracle.py).  Tolerances follow SURVEY.md §8a: eval-mode rel-L2 <= 2x the reference's
own fp32->bf16 drift (Baseline 0.24 %, DyYOLO 2.4 %)."""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"

ANCHORS = [[[199, 73], [315, 92], [268, 182]], [[91, 54], [120, 75], [157, 60]], [[29, 23], [48, 30], [67, 38]]]
BASE_HP = dict(anchors=ANCHORS, head_scales=[32, 16, 8], lr=1e-4, lr_scheduler=False,
               loss_balancing=dict(obj_scales_w=[0.5, 1.0, 2.0], bbox_w=4.0, objectness_w=1.0, no_obj_w=4.0),
               bbox_loss_fn="ciou", optim=dict(name="SGD", momentum=0.7))
DARKNET53 = [[32, 3, 1], [64, 3, 2], ["B", 1], [128, 3, 2], ["B", 2], [256, 3, 2], ["B", 8], [512, 3, 2], ["B", 8],
             [1024, 3, 2], ["B", 4], [512, 1, 1], [1024, 3, 1], ["S"], [256, 1, 1], ["U"], [256, 1, 1], [512, 3, 1],
             ["S"], [128, 1, 1], ["U"], [128, 1, 1], [256, 3, 1], ["S"]]
DYYOLO = [["DyConv", 32, 3, 1], ["DyConv", 64, 3, 2]] + DARKNET53[2:11] + [["DyConv", 512, 1, 1]] + DARKNET53[12:16] + \
         [["DyConv", 256, 1, 1]] + DARKNET53[17:21] + [["DyConv", 128, 1, 1]] + DARKNET53[22:]
# every op type of the DSL on a 7-conv-deep trunk: shallow enough that bf16 noise is not amplified
# by dozens of tiny-batch BatchNorms, so backward can be compared tightly
MINI = [[32, 3, 1], [64, 3, 2], ["B", 8], [128, 3, 2], ["B", 8], [256, 3, 2], ["B", 1], [128, 1, 1], [256, 3, 1], ["S"],
        [64, 1, 1], ["U"], [64, 1, 1], [128, 3, 1], ["S"], [32, 1, 1], ["U"], [32, 1, 1], [64, 3, 1], ["S"]]


def rel_l2(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def make(cls_name, layer_config, seed=0, **over):
    from multimodal_uav_det_b200.model import BaselineModel, DyYOLO
    from multimodal_uav_det_b200.utils.datatype import Config
    hp = dict(BASE_HP, layer_config=layer_config, **over)
    torch.manual_seed(seed)
    model = {"BaselineModel": BaselineModel, "DyYOLO": DyYOLO}[cls_name](hparams=Config(hp))
    return model, hp


def synth_input(b, size, seed=1234):
    """even indices 'RGB' (3 independent channels), odd 'IR' (one channel replicated)."""
    x = torch.rand(b, 3, size, size, generator=torch.Generator().manual_seed(seed))
    x[1::2] = x[1::2, :1].expand(-1, 3, -1, -1)
    return x


def randomize_bn(model, seed=7):
    """Non-trivial running stats / affine so eval-mode folding is actually exercised."""
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) * 0.5 + 0.75)
            m.weight.data.copy_(torch.rand(m.num_features, generator=g) * 0.5 + 0.75)
            m.bias.data.copy_(torch.randn(m.num_features, generator=g) * 0.1)


@pytest.mark.parametrize("size,batch", [(128, 2), (640, 1)])
def test_baseline_eval_forward_matches_oracle(lib, size, batch):
    from oracle import oracle as O
    model, hp = make("BaselineModel", DARKNET53)
    randomize_bn(model)
    model.eval()
    sd = copy.deepcopy(model.state_dict())
    x = synth_input(batch, size)
    with torch.no_grad():
        want = O.darknet_forward(x, sd, DARKNET53)
        got = model.to(DEV)(x.to(DEV))
    from multimodal_uav_det_b200 import ops
    ops.check_device()
    for s, (g, (wb, wo)) in enumerate(zip(got, want)):
        assert g.bbox.shape == wb.shape and g.obj.shape == wo.shape and g.bbox.dtype == torch.float32
        rb, ro = rel_l2(g.bbox.cpu(), wb), rel_l2(g.obj.cpu(), wo)
        print(f"baseline eval size={size} scale {s}: rel_l2 bbox={rb:.4f} obj={ro:.4f}")
        assert rb < 0.01 and ro < 0.01   # 2 x 0.24 % reference self-drift would be 0.0048; bound 1 %


def test_dyyolo_eval_forward_matches_oracle(lib):
    from oracle import oracle as O
    model, hp = make("DyYOLO", DYYOLO, bbox_loss_fn="mse", attn_temperature=30.0)
    randomize_bn(model)
    model.eval()
    sd = copy.deepcopy(model.state_dict())
    x = synth_input(2, 128)
    with torch.no_grad():
        want = O.darknet_forward(x, sd, DYYOLO, 30.0)
        got = model.to(DEV)(x.to(DEV))
    for s, (g, (wb, wo)) in enumerate(zip(got, want)):
        rb, ro = rel_l2(g.bbox.cpu(), wb), rel_l2(g.obj.cpu(), wo)
        print(f"dyyolo eval scale {s}: rel_l2 bbox={rb:.4f} obj={ro:.4f}")
        assert rb < 0.05 and ro < 0.05   # 2 x 2.4 % reference self-drift (unscaled randn expert banks)


def _targets(hp, b, size, seed=1, grids=None):
    from oracle import oracle as O
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(b):
        cx, cy = (torch.rand(2, generator=g) * 0.6 + 0.2) * size
        w = (torch.rand(1, generator=g) * 60 + 20) * size / 640
        h = (torch.rand(1, generator=g) * 40 + 15) * size / 640
        box = torch.tensor([[cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2]])
        anchors = (torch.tensor(hp["anchors"]).float() * size / 640).tolist()
        out.append(O.encode_targets(box, anchors, hp["head_scales"], input_size=size, grids=grids))
    return out


# ~20 BN layers on the longest path and every op of the DSL (routes are triggered by B-blocks of
# `route_repeats` repeats: 8 in the reference, 2 here so the trunk stays shallow enough for bf16
# noise not to be amplified into chaos by dozens of tiny-batch BatchNorms — SURVEY §8a tolerances iii)
SHALLOW = [[32, 3, 1], [64, 3, 2], ["B", 2], [128, 3, 2], ["B", 2], [256, 3, 2], [128, 1, 1], [256, 3, 1], ["S"],
           [64, 1, 1], ["U"], [64, 1, 1], [128, 3, 1], ["S"], [32, 1, 1], ["U"], [32, 1, 1], [64, 3, 1], ["S"]]


TINY = [[32, 3, 1], [64, 3, 2], ["S"]]          # 6 conv+BN layers, one scale


def cosine(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return (a @ b / (a.norm() * b.norm() + 1e-30)).item()


def _train_step_pair(cfg, loss_fn, size, b, grids, route_repeats, faithful=False, train=True, hp_over=None):
    """One forward+loss+backward on the CUDA product and on the CPU oracle (autograd, fp32 or with the
    product's bf16 storage points emulated).  Returns per-output forward errors and per-parameter
    (rel_l2, cosine, name) gradient comparisons."""
    from oracle import oracle as O
    from multimodal_uav_det_b200.utils.datatype import BatchData
    from multimodal_uav_det_b200 import ops
    model, hp = make("BaselineModel", cfg, bbox_loss_fn=loss_fn, **(hp_over or {}))
    model.route_repeats = route_repeats
    anchors = (torch.tensor(hp["anchors"]).float() * size / 640).tolist()
    model.yolo_head.anchors = torch.tensor(anchors).float()
    if not train:
        randomize_bn(model)
    model.train(train)
    x = synth_input(b, size)
    tg = _targets(hp, b, size, grids=grids)
    sd = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "running" not in k)
          for k, v in model.state_dict(keep_vars=False).items()}
    with O.bf16_pipeline(faithful):     # faithful: the oracle rounds where the CUDA path stores bf16
        outs_ref = O.darknet_forward(x, sd, cfg, train=train, route_repeats=route_repeats)
        loss_ref, _, _ = O.yolo_loss(outs_ref, tg, anchors, hp["head_scales"], hp["loss_balancing"], loss_fn)
        loss_ref.backward()
    model = model.to(DEV)
    outs = model(x.to(DEV))
    batch = BatchData(image=x.to(DEV), bbox=[[t.to(DEV) for t in per] for per in copy.deepcopy(tg)])
    loss, _, _, _ = model.yolo_head.compute_metrics(outs, batch)
    loss.backward()
    ops.check_device()
    fwd = [max(rel_l2(g.bbox.detach().cpu(), wb.detach()), rel_l2(g.obj.detach().cpu(), wo.detach()))
           for g, (wb, wo) in zip(outs, outs_ref)]
    grads = []
    for name, p in model.named_parameters():
        if p.grad is None:
            assert not train and ".bn." in name, f"no grad for {name}"   # frozen BN affine in eval mode
            continue
        assert sd[name].grad is not None, name
        grads.append((rel_l2(p.grad.cpu(), sd[name].grad), cosine(p.grad.cpu(), sd[name].grad), name))
    grads.sort(reverse=True)
    return model, loss.item(), loss_ref.item(), fwd, grads


def _summ(tag, loss, loss_ref, fwd, grads):
    med = grads[len(grads) // 2][0]
    print(f"{tag}: loss ref={loss_ref:.5f} got={loss:.5f} fwd rel_l2={['%.4f' % f for f in fwd]} "
          f"grad rel_l2 median={med:.4f} worst={grads[0][0]:.4f} ({grads[0][2]}) min cos={min(g[1] for g in grads):.4f}")
    return med


@pytest.mark.parametrize("loss_fn", ["ciou", "mse"])
def test_tiny_darknet_train_step_matches_oracle_autograd(lib, loss_fn):
    """Train-mode (batch-stat BN) forward + loss + backward of a 6-layer trunk + head against
    autograd through the oracle.  bf16 storage of activation gradients costs ~13 % rel-L2 on the
    parameter gradients even in the oracle itself (BN backward removes the large common-mode part of
    dz, leaving the rounding noise: oracle-with-bf16-points vs oracle-fp32 = 13 %, cos 0.978, measured
    on CPU), so the tight check is against the oracle evaluated with the SAME bf16 storage points; the
    fp32 oracle bounds the direction (cosine) and the loss."""
    over = dict(anchors=[ANCHORS[2]], head_scales=[8],
                loss_balancing=dict(obj_scales_w=[1.0], bbox_w=4.0, objectness_w=1.0, no_obj_w=4.0))
    model, loss, loss_ref, fwd, grads = _train_step_pair(TINY, loss_fn, 64, 16, [32], 8, faithful=True, hp_over=over)
    med = _summ(f"tiny[{loss_fn}] vs bf16-point oracle", loss, loss_ref, fwd, grads)
    assert abs(loss - loss_ref) <= 2e-3 * abs(loss_ref)
    assert max(fwd) < 0.01
    # residual disagreement = sub-ulp accumulation-order differences flipping bf16 roundings, then
    # amplified by the BN-backward projection (measured 5-6 % median on B200)
    # (atomics make it run-to-run variable: 5-10 % median observed)
    assert med < 0.20 and grads[0][0] < 0.30 and min(g[1] for g in grads) > 0.985
    assert int(model.layers[0].bn.num_batches_tracked) == 1 and model.layers[0].bn.running_mean.abs().sum() > 0
    _, loss2, loss_fp32, fwd2, grads2 = _train_step_pair(TINY, loss_fn, 64, 16, [32], 8, hp_over=over)
    _summ(f"tiny[{loss_fn}] vs fp32 oracle", loss2, loss_fp32, fwd2, grads2)
    assert abs(loss2 - loss_fp32) <= 5e-3 * abs(loss_fp32) and max(fwd2) < 0.02
    assert min(g[1] for g in grads2) > 0.95


@pytest.mark.parametrize("loss_fn", ["ciou", "mse"])
def test_shallow_darknet_frozen_bn_backward_matches_oracle(lib, loss_fn):
    """Every op of the layer DSL (stem, stride-2, residual, scale branches, upsample + route concat,
    fused head) with BN frozen (eval statistics) and grad enabled: non-chaotic, so the gradient
    plumbing — skip/route/branch accumulation in the dgrad epilogues, wgrad, head backward — is
    checked tightly against fp32 autograd."""
    model, loss, loss_ref, fwd, grads = _train_step_pair(SHALLOW, loss_fn, 128, 8, [16, 32, 64], 2, train=False)
    med = _summ(f"frozen[{loss_fn}]", loss, loss_ref, fwd, grads)
    assert abs(loss - loss_ref) <= 5e-3 * abs(loss_ref)
    assert max(fwd) < 0.02
    assert med < 0.03 and grads[0][0] < 0.15 and min(g[1] for g in grads) > 0.985


def test_shallow_darknet_train_step_statistical_agreement(lib):
    """Same trunk in train mode (~24 batch-stat BN layers deep): the chaotic regime.  The loss must
    agree and every gradient must point the same way as the oracle's (evaluated with the same bf16
    storage points); element-wise agreement is not expected here."""
    model, loss, loss_ref, fwd, grads = _train_step_pair(SHALLOW, "ciou", 128, 16, [16, 32, 64], 2, faithful=True)
    _summ("train-shallow", loss, loss_ref, fwd, grads)
    assert abs(loss - loss_ref) <= 5e-3 * abs(loss_ref)
    assert max(fwd) < 0.10
    cos = sorted(g[1] for g in grads)
    assert cos[len(cos) // 2] > 0.9 and cos[0] > 0.6


def test_deep_mini_darknet_train_step_sanity(lib):
    """The reference's own route rule (8-repeat blocks) on a ~45-BN-deep trunk: loss value and
    finite gradients only (the reference's own fp32->bf16 train-mode drift is 31 %)."""
    model, loss, loss_ref, fwd, grads = _train_step_pair(MINI, "ciou", 128, 8, [16, 32, 64], 8)
    _summ("train-deep", loss, loss_ref, fwd, grads)
    assert abs(loss - loss_ref) <= 0.02 * abs(loss_ref)
    assert all(torch.isfinite(p.grad).all() for p in model.parameters())


DYSOEM_HP = dict(anchors=[ANCHORS[2], ANCHORS[1], ANCHORS[0]], head_scales=[32, 16, 8], lr=1e-4, lr_scheduler=False,
                 attention_temperature=30, num_dy_conv=[3, 3, 3], dy_kernel_size=[3, 3, 3], bbox_loss_fn="mse",
                 loss_balancing=dict(obj_scales_w=[2.0, 1.0, 0.5], bbox_w=4.0, objectness_w=1.0, no_obj_w=4.0),
                 optim=dict(name="SGD", momentum=0.7))


@pytest.mark.parametrize("train", [False, True])
def test_dysoem_simfpn_forward_matches_oracle(lib, train):
    """DySOEM_SimFPN forward: aggregate-first dynamic kernels (1 GEMM instead of the reference's 3),
    fused space-to-depth, SimFPN — against the oracle, which executes the reference's formulation."""
    from oracle import oracle as O
    from multimodal_uav_det_b200.model import DySOEM_SimFPN
    from multimodal_uav_det_b200.utils.datatype import Config
    torch.manual_seed(0)
    model = DySOEM_SimFPN(hparams=Config(DYSOEM_HP))
    randomize_bn(model)
    model.train(train)
    sd = copy.deepcopy(model.state_dict())
    x = synth_input(4, 128)
    with torch.no_grad():
        want = O.dysoem_simfpn_forward(x, sd, 30.0, train=train)
        got = model.to(DEV)(x.to(DEV), attn_temp=30.0)
    from multimodal_uav_det_b200 import ops
    ops.check_device()
    for s, (g, (wb, wo)) in enumerate(zip(got, want)):
        assert g.bbox.shape == wb.shape
        rb, ro = rel_l2(g.bbox.cpu(), wb), rel_l2(g.obj.cpu(), wo)
        print(f"dysoem train={train} scale {s}: rel_l2 bbox={rb:.4f} obj={ro:.4f}")
        assert rb < 0.02 and ro < 0.02      # ~2 x 0.77 % reference self-drift


def test_dysoem_golden_forward(lib):
    """Against the committed reference output (tests/golden/model_forwards.pt, eval, 64x64)."""
    import os
    from multimodal_uav_det_b200.model import DySOEM_SimFPN
    from multimodal_uav_det_b200.utils.datatype import Config
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "model_forwards.pt"), weights_only=False)["dy-soem_fpn"]
    torch.manual_seed(gold["seed"])
    model = DySOEM_SimFPN(hparams=Config(gold["hp"])).eval().to(DEV)
    x = torch.rand(2, 3, 64, 64, generator=torch.Generator().manual_seed(gold["x_seed"]))
    x[1] = x[1, :1].expand(3, -1, -1)
    with torch.no_grad():
        got = model(x.to(DEV), 30.0)
    for g, (gb, go) in zip(got, gold["outs"]):
        assert rel_l2(g.bbox.cpu(), gb) < 0.02 and rel_l2(g.obj.cpu(), go) < 0.02


def test_rtmuavdet_forward_matches_oracle(lib):
    """RTMUAVDet eval forward incl. sigmoid heads + in-forward decode, 640x640 (319x319 stem quirk)."""
    from oracle import oracle as O
    from multimodal_uav_det_b200.model import RTMUAVDet
    anchors = torch.tensor([[[29, 23], [48, 30], [67, 38]], [[91, 54], [120, 75], [157, 60]]]).float()
    torch.manual_seed(0)
    model = RTMUAVDet([3, 640, 640], anchors, 1e-4)
    randomize_bn(model)
    model.eval()
    sd = copy.deepcopy(model.state_dict())
    x = synth_input(2, 640)
    with torch.no_grad():
        want = O.rtm_forward(x, sd, anchors)
        got = model.to(DEV)(x.to(DEV))
    from multimodal_uav_det_b200 import ops
    ops.check_device()
    for s, (g, (wb, wo)) in enumerate(zip(got, want)):
        assert g.bbox.shape == wb.shape and g.obj.shape == wo.shape
        rb, ro = rel_l2(g.bbox.cpu(), wb), rel_l2(g.obj.cpu(), wo)
        print(f"rtm scale {s}: rel_l2 bbox={rb:.4f} obj={ro:.4f}")
        assert rb < 0.01 and ro < 0.01      # ~2 x 0.40 % reference self-drift


def test_mdy_encoder_folded_groupnorms_match_oracle(lib):
    """MDyEncoder (RTMUAVDet.py:144-184) with both GroupNorms folded into the 1x1 convolutions behind them and the three
    base convolutions as one GEMM, against the oracle's operator-by-operator fp32 form; samples with different means /
    spreads exercise the per-sample epilogue."""
    from oracle import oracle as O
    from multimodal_uav_det_b200.engine import Executor
    from multimodal_uav_det_b200.model.RTMUAVDet import MDyEncoder
    torch.manual_seed(3)
    enc = MDyEncoder(192, 128)
    randomize_bn(enc)
    with torch.no_grad():
        for gn in (enc.group_norm_in, enc.group_norm_out):
            gn.weight.uniform_(0.5, 1.5)
            gn.bias.normal_(0, 0.2)
    enc.eval()
    sd = copy.deepcopy(enc.state_dict())
    g = torch.Generator().manual_seed(4)
    x = torch.randn(3, 192, 40, 24, generator=g) * torch.tensor([0.5, 1.0, 3.0]).view(3, 1, 1, 1) \
        + torch.tensor([0.0, 1.0, -2.0]).view(3, 1, 1, 1)
    x = x.bfloat16().float()
    with torch.no_grad():
        want = O.mdy_encoder(x, {"e." + k: v for k, v in sd.items()}, "e")
        enc = enc.to(DEV)
        got = enc.forward_nhwc(x.to(DEV).permute(0, 2, 3, 1).contiguous().bfloat16(), Executor())
        unf = enc._forward_nhwc_unfused(x.to(DEV).permute(0, 2, 3, 1).contiguous().bfloat16(), Executor())
    from multimodal_uav_det_b200 import ops
    ops.check_device()
    r_f = rel_l2(got.float().permute(0, 3, 1, 2).cpu(), want)
    r_u = rel_l2(unf.float().permute(0, 3, 1, 2).cpu(), want)
    print(f"mdy_encoder rel_l2: folded {r_f:.4f}, unfused {r_u:.4f}")
    assert r_f < 0.01 and r_f < 1.5 * r_u + 1e-3


def test_detect_decode_nms_bit_exact_on_model_outputs(lib):
    """C1: BaselineModel forward + decode + NMS.  Kept indices must be bit-identical to the oracle's
    NMS on the SAME fp32 boxes/scores (the decode itself is checked to fp32 tolerance)."""
    from oracle import oracle as O
    from multimodal_uav_det_b200 import inference
    model, hp = make("BaselineModel", DARKNET53)
    model.eval().to(DEV)
    x = synth_input(2, 640)
    det = inference.detect(model, x.to(DEV))
    assert det.boxes.shape == (2, 25200, 4)
    with torch.no_grad():
        outs = model(x.to(DEV))
    wb, ws = O.decode_yolo([(o.bbox.cpu(), o.obj.cpu()) for o in outs], hp["anchors"], hp["head_scales"], True)
    torch.testing.assert_close(det.boxes.cpu(), wb, rtol=2e-6, atol=2e-5)
    assert torch.equal(det.scores.cpu(), ws)
    for b, kept in enumerate(inference.kept_lists(det)):
        want = O.nms(det.boxes[b].cpu().numpy(), det.scores[b].cpu().numpy(), 0.5)
        assert np.array_equal(kept.cpu().numpy(), want)
        print(f"image {b}: kept {len(want)} of 25200")


@pytest.mark.parametrize("train_bn", [False, True])
def test_graphed_train_step_equals_eager(lib, train_bn):
    """parallel.GraphedTrainStep (whole step replayed from one CUDA graph) follows the eager trainer: the same
    kernels on the same data.  With frozen BN the comparison is tight; with batch-stat BN the fp32-atomic
    accumulation order is amplified chaotically (two eager runs differ as much), so only the loss trajectory
    and the BN bookkeeping are compared there."""
    from multimodal_uav_det_b200.parallel import FlatSGDTrainer, GraphedTrainStep
    from multimodal_uav_det_b200.utils.datatype import BatchData
    from multimodal_uav_det_b200 import ops
    size, b = 128, 8
    runs = {}
    for mode in ("eager", "graph"):
        model, hp = make("BaselineModel", SHALLOW, lr=1e-3)
        model.route_repeats = 2
        if not train_bn:
            randomize_bn(model)
        model = model.to(DEV).train(train_bn)
        model.yolo_head.mutate_targets = False
        init = {k: v.detach().float().cpu().clone() for k, v in model.named_parameters()}
        trainer = FlatSGDTrainer(model, lr=1e-3, momentum=0.7)
        xs = [synth_input(b, size, seed=100 + i).to(DEV) for i in range(3)]
        per = _targets(hp, b, size, grids=[16, 32, 64])
        tg = [torch.stack([per[i][h] for i in range(b)]).to(DEV) for h in range(3)]
        losses = []
        if mode == "eager":
            for x in xs:
                trainer.zero_grad()
                outs = model(x)
                loss, _, _, _ = model.yolo_head.compute_metrics(outs, BatchData(image=x, bbox=tg))
                loss.backward()
                trainer.step()
                losses.append(loss.item())
        else:
            step = GraphedTrainStep(model, trainer, xs[0], tg, warmup=0)   # capture only: parameters untouched
            for x in xs:
                losses.append(step(x, tg).item())
        ops.check_device()
        delta = torch.cat([(p.detach().float().cpu() - init[k]).flatten() for k, p in model.named_parameters()])
        runs[mode] = (losses, delta, int(model.layers[0].bn.num_batches_tracked))
    le, lg = runs["eager"][0], runs["graph"][0]
    cos = cosine(runs["eager"][1], runs["graph"][1])
    print("train_bn", train_bn, "eager", le, "graph", lg, "update cosine", cos)
    if train_bn:
        assert np.allclose(le, lg, rtol=0.05), (le, lg)
        assert runs["eager"][2] == runs["graph"][2] == 3       # BN counters advance inside the graph too
    else:
        assert np.allclose(le, lg, rtol=2e-3), (le, lg)
        assert cos > 0.995
    # the graph bumps the parameter epoch: an eager eval forward afterwards sees the updated weights
    model.eval()
    with torch.no_grad():
        out = model(xs[0])
    assert torch.isfinite(out[0].bbox).all()


# ------------------------------------------------------------------------------------------------
# training of the dynamic-kernel models (DyYOLO, DySOEM_SimFPN)
# ------------------------------------------------------------------------------------------------
# every DyConv flavour of conf/model/dy-yolo.yaml on a short trunk: cin=3 stem, 3x3 stride 2, 1x1 after a scale
# block and 1x1 on a route concat; routes at 2-repeat blocks
DYMINI = [["DyConv", 32, 3, 1], ["DyConv", 64, 3, 2], ["B", 2], [128, 3, 2], ["DyConv", 64, 1, 1], [128, 3, 1], ["S"],
          [32, 1, 1], ["U"], ["DyConv", 64, 1, 1], [128, 3, 1], ["S"]]


def _grad_report(model, sd, tag):
    grads = []
    bn_params = {f"{mn}.{pn}" for mn, m in model.named_modules() if isinstance(m, torch.nn.BatchNorm2d)
                 for pn, _ in m.named_parameters()}
    for name, p in model.named_parameters():
        ref = sd[name].grad
        if p.grad is None:
            assert ref is None or name in bn_params, name      # frozen BN: affine not trained
            continue
        assert ref is not None, name
        grads.append((rel_l2(p.grad.cpu(), ref), cosine(p.grad.cpu(), ref), name))
    grads.sort(reverse=True)
    med = grads[len(grads) // 2][0]
    print(f"{tag}: {len(grads)} tensors, grad rel_l2 median={med:.4f}; worst:", [(f"{g[0]:.3f}", f"{g[1]:.4f}", g[2]) for g in grads[:4]])
    return med, grads


@pytest.mark.parametrize("train_bn", [False, True])
def test_dyyolo_train_step_matches_oracle_autograd(lib, train_bn):
    """DyYOLO forward + loss + backward (attention MLP, expert-bank aggregation, per-sample kernels) against
    autograd through the oracle's restatement of DyConvModule (_base.py:26-77).  Frozen BN: tight; batch-stat
    BN: direction only (the unscaled randn expert bank makes bf16 drift 10x Baseline's)."""
    from oracle import oracle as O
    from multimodal_uav_det_b200.utils.datatype import BatchData
    from multimodal_uav_det_b200 import ops
    size, b, temp = 64, 8, 30.0
    over = dict(anchors=[ANCHORS[1], ANCHORS[2]], head_scales=[4, 2], attn_temperature=temp,
                loss_balancing=dict(obj_scales_w=[1.0, 2.0], bbox_w=4.0, objectness_w=1.0, no_obj_w=4.0))
    model, hp = make("DyYOLO", DYMINI, **over)
    model.route_repeats = 2
    anchors = (torch.tensor(hp["anchors"]).float() * size / 640).tolist()
    model.yolo_head.anchors = torch.tensor(anchors).float()
    if not train_bn:
        randomize_bn(model)
    model.train(train_bn)
    x = synth_input(b, size)
    tg = _targets(hp, b, size, grids=[16, 32])
    sd = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "running" not in k)
          for k, v in model.state_dict(keep_vars=False).items()}
    with O.bf16_pipeline(True):
        outs_ref = O.darknet_forward(x, sd, DYMINI, attn_temperature=temp, train=train_bn, route_repeats=2)
        loss_ref, _, _ = O.yolo_loss(outs_ref, tg, anchors, hp["head_scales"], hp["loss_balancing"], "ciou")
        loss_ref.backward()
    model = model.to(DEV)
    outs = model(x.to(DEV))
    batch = BatchData(image=x.to(DEV), bbox=[[t.to(DEV) for t in per] for per in copy.deepcopy(tg)])
    loss, _, _, _ = model.yolo_head.compute_metrics(outs, batch)
    loss.backward()
    ops.check_device()
    fwd = [max(rel_l2(g.bbox.detach().cpu(), wb.detach()), rel_l2(g.obj.detach().cpu(), wo.detach()))
           for g, (wb, wo) in zip(outs, outs_ref)]
    print(f"dyyolo train_bn={train_bn}: loss ref={loss_ref.item():.5f} got={loss.item():.5f} fwd={fwd}")
    med, grads = _grad_report(model, sd, f"dyyolo train_bn={train_bn}")
    assert abs(loss.item() - loss_ref.item()) <= 0.02 * abs(loss_ref.item())
    assert all(torch.isfinite(p.grad).all() for p in model.parameters() if p.grad is not None)
    dyn = [g for g in grads if ".weights" in g[2] or ".attention." in g[2]]
    assert len(dyn) >= 12                                   # 4 sites x (bank, attention w1, w2, b2)
    if not train_bn:
        assert max(fwd) < 0.05
        assert med < 0.06 and min(g[1] for g in grads) > 0.97, grads[:5]
    else:
        cos = sorted(g[1] for g in grads)
        assert cos[len(cos) // 2] > 0.9


@pytest.mark.parametrize("train_bn", [False, True])
def test_dysoem_simfpn_train_step_matches_oracle_autograd(lib, train_bn):
    """DySOEM_SimFPN forward + loss + backward (aggregate-first dynamic kernels through the space-to-depth view,
    SimFPN skip algebra, fused head) against autograd through the oracle, which executes the reference's
    K-convs-then-weighted-sum formulation.  The heads come out at size/2, size/4, size/8, so the
    targets are encoded on those grids."""
    from oracle import oracle as O
    from multimodal_uav_det_b200.model import DySOEM_SimFPN
    from multimodal_uav_det_b200.utils.datatype import BatchData, Config
    from multimodal_uav_det_b200 import ops
    size, b, temp = 64, 8, 30.0
    hp = dict(DYSOEM_HP)
    torch.manual_seed(0)
    model = DySOEM_SimFPN(hparams=Config(hp))
    anchors = (torch.tensor(hp["anchors"]).float() * size / 640).tolist()
    model.yolo_head.anchors = torch.tensor(anchors).float()
    if not train_bn:
        randomize_bn(model)
    model.train(train_bn)
    x = synth_input(b, size)
    tg = _targets(hp, b, size, grids=[size // 2, size // 4, size // 8])
    sd = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "running" not in k)
          for k, v in model.state_dict(keep_vars=False).items()}
    with O.bf16_pipeline(True):
        outs_ref = O.dysoem_simfpn_forward(x, sd, temp, train=train_bn)
        loss_ref, _, _ = O.yolo_loss(outs_ref, tg, anchors, hp["head_scales"], hp["loss_balancing"], "mse")
        loss_ref.backward()
    model = model.to(DEV)
    outs = model(x.to(DEV), attn_temp=temp)
    batch = BatchData(image=x.to(DEV), bbox=[[t.to(DEV) for t in per] for per in copy.deepcopy(tg)])
    loss, _, _, _ = model.yolo_head.compute_metrics(outs, batch)
    loss.backward()
    ops.check_device()
    fwd = [max(rel_l2(g.bbox.detach().cpu(), wb.detach()), rel_l2(g.obj.detach().cpu(), wo.detach()))
           for g, (wb, wo) in zip(outs, outs_ref)]
    print(f"dysoem train_bn={train_bn}: loss ref={loss_ref.item():.5f} got={loss.item():.5f} fwd={fwd}")
    med, grads = _grad_report(model, sd, f"dysoem train_bn={train_bn}")
    assert abs(loss.item() - loss_ref.item()) <= 0.01 * abs(loss_ref.item())
    if not train_bn:
        assert max(fwd) < 0.03
        assert med < 0.05 and min(g[1] for g in grads) > 0.97, grads[:5]
    else:
        # the expert biases feed a batch-stat BN: their gradient is the residue of an almost exact cancellation
        # (zero if the attention were constant over the batch), i.e. noise in both implementations
        cos = sorted(g[1] for g in grads if not (".dy_convs." in g[2] and g[2].endswith(".bias")))
        assert cos[len(cos) // 2] > 0.9 and cos[0] > 0.5


def test_dysoem_graphed_train_step(lib):
    """DySOEM_SimFPN through FlatSGDTrainer + GraphedTrainStep: the loss decreases over a few replays."""
    from multimodal_uav_det_b200.model import DySOEM_SimFPN
    from multimodal_uav_det_b200.parallel import FlatSGDTrainer, GraphedTrainStep
    from multimodal_uav_det_b200.utils.datatype import Config
    from multimodal_uav_det_b200 import ops
    size, b = 64, 8
    hp = dict(DYSOEM_HP)
    torch.manual_seed(0)
    model = DySOEM_SimFPN(hparams=Config(hp)).to(DEV).train()
    trainer = FlatSGDTrainer(model, lr=1e-3, momentum=0.7)
    x = synth_input(b, size).to(DEV)
    per = _targets(hp, b, size, grids=[size // 2, size // 4, size // 8])
    tg = [torch.stack([per[i][h] for i in range(b)]).to(DEV) for h in range(3)]
    step = GraphedTrainStep(model, trainer, x, tg, warmup=1)
    losses = [step(x, tg).item() for _ in range(6)]
    ops.check_device()
    print("dysoem graphed losses", losses)
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]


def test_flat_trainer_channels_last_storage_gives_same_gradients(lib):
    """FlatSGDTrainer keeps conv weights / gradients channels-last (wgrad accumulates in place, no unpack):
    same gradients as the plain OIHW path, and state_dict values are unchanged."""
    from multimodal_uav_det_b200.parallel import FlatSGDTrainer
    from multimodal_uav_det_b200.utils.datatype import BatchData
    size, b = 128, 4
    grads = {}
    for mode in ("plain", "flat"):
        model, hp = make("BaselineModel", SHALLOW)
        model.route_repeats = 2
        randomize_bn(model)
        model = model.to(DEV).eval()
        model.yolo_head.mutate_targets = False
        sd_before = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
        if mode == "flat":
            trainer = FlatSGDTrainer(model, lr=1e-3, momentum=0.7)
            trainer.zero_grad()
            w = model.layers[1].conv.weight
            assert not w.is_contiguous() and w.permute(0, 2, 3, 1).is_contiguous()
            for k, v in model.state_dict().items():
                assert torch.equal(v.detach().cpu(), sd_before[k]), k
        x = synth_input(b, size).to(DEV)
        per = _targets(hp, b, size, grids=[16, 32, 64])
        tg = [torch.stack([per[i][h] for i in range(b)]).to(DEV) for h in range(3)]
        outs = model(x)
        loss, _, _, _ = model.yolo_head.compute_metrics(outs, BatchData(image=x, bbox=tg))
        loss.backward()
        grads[mode] = {k: p.grad.detach().float().cpu().clone() for k, p in model.named_parameters() if p.grad is not None}
    assert set(grads["plain"]) <= set(grads["flat"])        # the flat arenas give every parameter a (zero) grad
    worst = max(rel_l2(grads["flat"][k], grads["plain"][k]) for k in grads["plain"])
    print("channels-last vs OIHW gradient rel_l2 (worst)", worst)
    assert worst < 2e-3


def test_graphed_detect_equals_eager_detect(lib):
    """inference.GraphedDetect (forward + decode + NMS replayed from one CUDA graph) returns exactly what the eager
    `detect` returns, on fresh inputs copied into its static buffer."""
    from multimodal_uav_det_b200 import inference
    model, _ = make("BaselineModel", SHALLOW)
    model.route_repeats = 2
    randomize_bn(model)
    model = model.to(DEV).eval()
    xs = [synth_input(2, 128, seed=900 + i).to(DEV) for i in range(3)]
    run = inference.GraphedDetect(model, xs[0])
    for x in xs:
        want = inference.detect(model, x)
        got = run(x)
        assert torch.equal(got.keep_count, want.keep_count)
        for b, c in enumerate(want.keep_count.tolist()):
            assert torch.equal(got.keep[b, :c], want.keep[b, :c])
        assert torch.equal(got.boxes, want.boxes) and torch.equal(got.scores, want.scores)
