#!/usr/bin/env python
"""Generates tests/golden/*.pt from the UNMODIFIED reference (imported via oracle/ref_import.py).

The reference has no golden vectors, known-answer tests or fixtures of its own, so
parity is pinned on outputs of the reference itself executed here; this script is the committed
recipe.  Run in the build container (needs /root/reference):  python tools/make_golden.py
Fixtures are small (seeded inputs + reference outputs); big weights are re-created from seeds and
pinned by a SHA-256 of the state_dict.
"""
import copy
import hashlib
import os
import sys

import torch
import torchvision

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import as R  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def sd_hash(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def gen_nms():
    cases = []
    g = torch.Generator().manual_seed(11)
    for n, quant, thr in [(0, None, 0.5), (1, None, 0.5), (17, None, 0.5), (300, 4, 0.5), (2000, 20, 0.5),
                          (1500, None, 0.3), (1500, None, 0.7)]:
        c = torch.rand(n, 2, generator=g) * 100
        wh = torch.rand(n, 2, generator=g) * 40
        boxes = torch.cat([c - wh / 2, c + wh / 2], 1)
        scores = torch.randn(n, generator=g)
        if quant:
            scores = (scores * quant).round() / quant
        cases.append(dict(boxes=boxes, scores=scores, thr=thr, keep=torchvision.ops.nms(boxes, scores, thr)))
    boxes = torch.tensor([[0, 0, 0, 0], [0, 0, 0, 0], [1, 1, 5, 5], [1, 1, 5, 5], [1, 1, 5, 5.0001], [2, 2, 1, 1],
                          [0, 0, 10, 10], [0, 0, 10, 10]], dtype=torch.float32)
    scores = torch.tensor([0.5, 0.5, float("nan"), 0.0, -0.0, 3.0, float("inf"), float("-inf")])
    cases.append(dict(boxes=boxes, scores=scores, thr=0.5, keep=torchvision.ops.nms(boxes, scores, 0.5)))
    torch.save(dict(torchvision=torchvision.__version__, cases=cases), os.path.join(OUT, "nms_cases.pt"))


def ref_targets(ns, hp, boxes_xyxy, input_size=640):
    ds = object.__new__(ns.dataset.AntiUAVDataset)
    ds.input_size = input_size
    ds.anchors = torch.tensor(hp["anchors"]).float() / input_size
    ds.head_size = torch.tensor([input_size // s for s in hp["head_scales"]])
    return [ds._AntiUAVDataset__generate_yolo_bboxes(b.view(1, 4).clone()) for b in boxes_xyxy]


def golden_logits(seed, batch, grids):
    """Deterministic (CPU generator) head logits shared by this script and tests/test_oracle.py."""
    g = torch.Generator().manual_seed(seed)
    return [(torch.randn(batch, 3, s, s, 4, generator=g), torch.randn(batch, 3, s, s, 1, generator=g)) for s in grids]


def gen_head_loss_decode():
    """YOLOHead decode (+NMS prep), target encoder and compute_metrics of the reference for both
    bbox_loss_fn modes on seeded random logits (input 160 px -> 5/10/20 grids, batch 3)."""
    ns = R.load()
    out = {}
    size = 160
    boxes = torch.tensor([[25.0, 50.0, 40.0, 60.0], [100.3, 22.7, 107.9, 29.2], [75.0, 77.5, 130.0, 117.5]])
    for name in ("baseline", "dy-yolo"):
        cfg, hp = R.hparams(name)
        hp = dict(hp, anchors=(torch.tensor(hp["anchors"]).float() * size / 640).tolist())
        head = ns.base.YOLOHead([8, 8, 8], hp["anchors"], hp["head_scales"], cfg.loss_balancing, hp["bbox_loss_fn"])
        grids = [size // s for s in hp["head_scales"]]
        logits = golden_logits(21, 3, grids)
        tg = ref_targets(ns, hp, boxes, size)
        outs = [ns.datatype.DetectionResults(bbox=b.clone().requires_grad_(True), obj=o.clone().requires_grad_(True))
                for b, o in logits]
        batch = ns.datatype.BatchData(image=torch.zeros(3, 3, 8, 8), bbox=copy.deepcopy(tg))
        loss, _, bl, ol = head.compute_metrics(outs, batch)
        loss.backward()
        dec = []
        for hi, (b, o) in enumerate(logits):
            sa = head.anchors[hi] / head.head_scales[hi]
            d = head._YOLOHead__pred_bbox_decoding(b[0], sa)
            xyxy, sc = head._YOLOHead__prepare_nms_preds(d, o[0])
            dec.append((d, xyxy, sc))
        out[name] = dict(hp=hp, size=size, boxes=boxes, logits_seed=21, targets=tg, mutated_targets=batch.bbox,
                         loss=loss.detach(), bbox_loss=bl.detach(), obj_loss=ol.detach(),
                         grads=[(x.bbox.grad.clone(), x.obj.grad.clone()) for x in outs], decode_img0=dec)
    torch.save(out, os.path.join(OUT, "head_loss_decode.pt"))


def gen_models():
    """End-to-end forwards of the four reference models on small seeded inputs (eval mode) +
    state_dict key lists and hashes pinning the seeded initialisation."""
    ns = R.load()
    out = {}
    x64 = torch.rand(2, 3, 64, 64, generator=torch.Generator().manual_seed(31))
    x64[1] = x64[1, :1].expand(3, -1, -1)
    for name, build in (("baseline", lambda c, d: ns.baseline.BaselineModel(hparams=c)),
                        ("dy-yolo", lambda c, d: ns.dyyolo.DyYOLO(hparams=c)),
                        ("dy-soem_fpn", lambda c, d: R.build_dysoem(c, d))):
        cfg, hp = R.hparams(name)
        torch.manual_seed(0)
        m = build(cfg, hp).eval()
        sd = m.state_dict()
        with torch.no_grad():
            o = m(x64, 30.0) if name == "dy-soem_fpn" else m(x64)
        out[name] = dict(hp=hp, seed=0, keys=list(sd.keys()), shapes=[tuple(v.shape) for v in sd.values()],
                         sha256=sd_hash(sd), x_seed=31, outs=[(t.bbox.clone(), t.obj.clone()) for t in o],
                         n_params=sum(p.numel() for p in m.parameters()))
    torch.manual_seed(0)
    m = R.build_rtm().eval()
    sd = m.state_dict()
    x = torch.rand(1, 3, 128, 128, generator=torch.Generator().manual_seed(32))
    with torch.no_grad():
        o = m(x)
    out["rtm"] = dict(seed=0, anchors=R.rtm_default_anchors(), keys=list(sd.keys()), shapes=[tuple(v.shape) for v in sd.values()],
                      sha256=sd_hash(sd), x_seed=32, outs=[(t.bbox.half(), t.obj.half()) for t in o],
                      n_params=sum(p.numel() for p in m.parameters()))
    torch.save(out, os.path.join(OUT, "model_forwards.pt"))


def gen_blocks():
    """Block-level fixtures with stored (small) weights: DyConvModule, DynamicSOEM, SimplifiedFPN,
    MDyConv, MDyEncoder, ConvModule, CNNBlock/ResidualBlock — forward in eval and train mode."""
    ns = R.load()
    g = torch.Generator().manual_seed(41)
    out = {}

    def pack(mod, x, *args, train=False):
        mod.train(train)
        sd = copy.deepcopy(mod.state_dict())
        with torch.no_grad():
            y = mod(x, *args)
        return dict(sd=sd, x=x, args=args, y=y if isinstance(y, torch.Tensor) else tuple(y), train=train)

    torch.manual_seed(5)
    out["dyconv"] = pack(ns.base.DyConvModule(8, 16, kernel_size=3, stride=2, padding=1), torch.randn(3, 8, 12, 12, generator=g), 30.0)
    out["dyconv_train"] = pack(ns.base.DyConvModule(8, 16, kernel_size=1, stride=1, padding=0), torch.randn(3, 8, 6, 6, generator=g), 30.0, train=True)
    out["dyconv_rgb"] = pack(ns.base.DyConvModule(3, 8, kernel_size=3, stride=1, padding=1), torch.rand(2, 3, 10, 10, generator=g), 30.0)
    out["soem"] = pack(ns.dysoem.DynamicSOEM(in_channels=4), torch.randn(2, 4, 12, 12, generator=g), 30.0)
    out["soem_train"] = pack(ns.dysoem.DynamicSOEM(in_channels=4), torch.randn(2, 4, 8, 8, generator=g), 1.0, train=True)
    fpn = ns.dysoem.SimplifiedFPN([4, 8, 16]).eval()
    feats = [torch.randn(2, 4, 16, 16, generator=g), torch.randn(2, 8, 8, 8, generator=g), torch.randn(2, 16, 4, 4, generator=g)]
    with torch.no_grad():
        out["fpn"] = dict(sd=copy.deepcopy(fpn.state_dict()), x=feats, y=tuple(fpn(feats)))
    out["convmodule"] = pack(ns.base.ConvModule(6, 8, kernel_size=(3, 3), padding=1, activation="relu"), torch.randn(2, 6, 7, 7, generator=g))
    out["cnnblock_train"] = pack(ns.baseline.CNNBlock(6, 8, kernel_size=3, stride=2, padding=1), torch.randn(4, 6, 9, 9, generator=g), train=True)
    out["resblock"] = pack(ns.baseline.ResidualBlock(8, num_repeats=2), torch.randn(2, 8, 6, 6, generator=g))
    out["mdyconv"] = pack(ns.rtm.MDyConv(8, 16, dy_kernel_size=5, dy_padding=2, dy_channel_size=4), torch.randn(2, 8, 9, 9, generator=g))
    out["mdyencoder"] = pack(ns.rtm.MDyEncoder(12, 8), torch.randn(2, 12, 6, 6, generator=g))
    # calculate_iou quirk (first-target column) on a multi-positive case
    preds = torch.rand(3, 4, 4, 4, generator=g) * 3
    tgts = torch.rand(3, 4, 4, 4, generator=g) * 3
    mask = torch.zeros(3, 4, 4, dtype=torch.bool)
    mask[0, 1, 2] = mask[2, 3, 0] = mask[1, 1, 1] = True
    anc = torch.tensor([[1.0, 2.0], [2.0, 1.0], [1.5, 1.5]])
    out["calculate_iou"] = dict(preds=preds, targets=tgts, mask=mask, anchors=anc,
                                mse=ns.postprocess.calculate_iou(preds, tgts, anc, mask, "mse"),
                                ciou=ns.postprocess.calculate_iou(preds, tgts, anc, mask, "ciou"))
    torch.save(out, os.path.join(OUT, "blocks.pt"))


def gen_adaptive_stem():
    """AdaptiveStemLayer (DySOEM_SimFPN.py:14-25; dead code in the reference model) on a 1-channel (IR)
    and a 3-channel (RGB) input, eval and train mode."""
    ns = R.load()
    g = torch.Generator().manual_seed(51)
    torch.manual_seed(9)
    m = ns.dysoem.AdaptiveStemLayer(32)
    for bn in (m.gray_conv.conv[1], m.rgb_conv.conv[1]):      # non-trivial statistics / affine
        bn.running_mean.copy_(torch.randn(32, generator=g) * 0.1)
        bn.running_var.copy_(torch.rand(32, generator=g) * 0.5 + 0.75)
        bn.weight.data.copy_(torch.rand(32, generator=g) * 0.5 + 0.75)
        bn.bias.data.copy_(torch.randn(32, generator=g) * 0.1)
    sd = copy.deepcopy(m.state_dict())
    out = dict(sd=sd, cases=[])
    for cin in (1, 3):
        x = torch.rand(2, cin, 12, 12, generator=g)
        for train in (False, True):
            mm = ns.dysoem.AdaptiveStemLayer(32)
            mm.load_state_dict(sd)
            mm.train(train)
            with torch.no_grad():
                y = mm(x)
            out["cases"].append(dict(x=x, train=train, y=y))
    torch.save(out, os.path.join(OUT, "adaptive_stem.pt"))


if __name__ == "__main__":
    if not R.available():
        raise SystemExit("reference not found; golden fixtures can only be generated in the build container")
    os.makedirs(OUT, exist_ok=True)
    gen_nms()
    gen_head_loss_decode()
    gen_models()
    gen_blocks()
    gen_adaptive_stem()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")
