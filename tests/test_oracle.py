"""CPU tests pinning the oracle (oracle/) — the checker the GPU parity tests rely on — against
(a) the committed golden fixtures generated from the reference by tools/make_golden.py, always, and
(b) the unmodified reference itself when /root/reference is present (build container)."""
import copy
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, os.path.join(ROOT, "tools"))

from oracle import oracle as O  # noqa: E402


def load(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


def golden_logits(seed, batch, grids):
    g = torch.Generator().manual_seed(seed)
    return [(torch.randn(batch, 3, s, s, 4, generator=g), torch.randn(batch, 3, s, s, 1, generator=g)) for s in grids]


# ---- NMS -----------------------------------------------------------------------------------------
def test_nms_oracle_matches_golden_torchvision_outputs():
    gold = load("nms_cases.pt")
    for i, c in enumerate(gold["cases"]):
        got = O.nms(c["boxes"].numpy(), c["scores"].numpy(), c["thr"])
        assert np.array_equal(got, c["keep"].numpy()), f"case {i}"
        if len(c["scores"]) <= 300:
            assert O.nms_py(c["boxes"].numpy(), c["scores"].numpy(), c["thr"]) == c["keep"].tolist(), f"py case {i}"


def test_nms_oracle_matches_installed_torchvision_randomised():
    import torchvision
    g = torch.Generator().manual_seed(3)
    for n in (0, 1, 2, 63, 64, 65, 500, 3000):
        for quant in (None, 10):
            c = torch.rand(n, 2, generator=g) * 50
            wh = torch.rand(n, 2, generator=g) * 30
            boxes = torch.cat([c - wh / 2, c + wh / 2], 1)
            scores = torch.rand(n, generator=g)
            if quant:
                scores = (scores * quant).round() / quant
            for thr in (0.5, 0.45):
                want = torchvision.ops.nms(boxes, scores, thr).numpy()
                assert np.array_equal(O.nms(boxes.numpy(), scores.numpy(), thr), want)


# ---- decode / encoder / loss ----------------------------------------------------------------------
@pytest.mark.parametrize("name", ["baseline", "dy-yolo"])
def test_decode_encoder_loss_match_golden(name):
    gold = load("head_loss_decode.pt")[name]
    hp, size = gold["hp"], gold["size"]
    grids = [size // s for s in hp["head_scales"]]
    logits = golden_logits(gold["logits_seed"], 3, grids)
    ciou = hp["bbox_loss_fn"] == "ciou"
    anc = torch.tensor(hp["anchors"]).float()
    hs = torch.tensor(hp["head_scales"])
    for hi, (b, o) in enumerate(logits):
        d = O.decode_yolo_head(b[0], anc[hi] / hs[hi], ciou)
        gd, gxy, gsc = gold["decode_img0"][hi]
        # 1-ulp: the reference applies sigmoid per strided channel slice, the oracle to the whole tensor
        torch.testing.assert_close(d, gd, rtol=2e-6, atol=1e-6)
        torch.testing.assert_close(O.cxcywh_to_xyxy(d.reshape(-1, 4)), gxy, rtol=2e-6, atol=2e-6)
        assert torch.equal(o[0].reshape(-1), gsc)
    boxes_all, scores_all = O.decode_yolo(logits, hp["anchors"], hp["head_scales"], ciou)
    torch.testing.assert_close(boxes_all[0], torch.cat([t[1] for t in gold["decode_img0"]]), rtol=2e-6, atol=2e-6)
    # target encoder (dataset/AntiUAVDataset.py:141-185)
    tg = [O.encode_targets(gold["boxes"][i:i + 1], hp["anchors"], hp["head_scales"], size) for i in range(3)]
    for a, b_ in zip(tg, gold["targets"]):
        for x, y in zip(a, b_):
            assert torch.equal(x, y)
    # loss (model/_base.py:155-212) incl. gradients
    outs = [(b.clone().requires_grad_(True), o.clone().requires_grad_(True)) for b, o in logits]
    loss, bl, ol = O.yolo_loss(outs, tg, hp["anchors"], hp["head_scales"], hp["loss_balancing"], hp["bbox_loss_fn"])
    loss.backward()
    torch.testing.assert_close(loss.detach(), gold["loss"], rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(bl.detach(), gold["bbox_loss"], rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(ol.detach(), gold["obj_loss"], rtol=1e-6, atol=1e-6)
    for (b, o), (gb, go) in zip(outs, gold["grads"]):
        torch.testing.assert_close(b.grad, gb, rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(o.grad, go, rtol=1e-5, atol=1e-7)


# ---- blocks ----------------------------------------------------------------------------------------
def test_block_restatements_match_golden():
    g = load("blocks.pt")
    c = g["dyconv"]
    y, _ = O.dyconv_module(c["x"], {"m." + k: v for k, v in c["sd"].items()}, "m", 3, 2, 1, 30.0)
    torch.testing.assert_close(y, c["y"], rtol=1e-5, atol=1e-5)
    c = g["dyconv_train"]
    y, _ = O.dyconv_module(c["x"], {"m." + k: v for k, v in c["sd"].items()}, "m", 1, 1, 0, 30.0, train=True)
    torch.testing.assert_close(y, c["y"], rtol=1e-4, atol=1e-4)
    c = g["dyconv_rgb"]
    y, _ = O.dyconv_module(c["x"], {"m." + k: v for k, v in c["sd"].items()}, "m", 3, 1, 1, 30.0)
    torch.testing.assert_close(y, c["y"], rtol=1e-5, atol=1e-5)
    for key, temp, train in (("soem", 30.0, False), ("soem_train", 1.0, True)):
        c = g[key]
        y = O.dynamic_soem(c["x"], {"m." + k: v for k, v in c["sd"].items()}, "m", temp, train)
        torch.testing.assert_close(y, c["y"], rtol=1e-5, atol=1e-5)
    c = g["fpn"]
    ys = O.simplified_fpn(c["x"], {"m." + k: v for k, v in c["sd"].items()}, "m")
    for a, b in zip(ys, c["y"]):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-5)
    c = g["convmodule"]
    torch.testing.assert_close(O.conv_module(c["x"], {"m." + k: v for k, v in c["sd"].items()}, "m", 1, 1, "relu"), c["y"])
    c = g["cnnblock_train"]
    torch.testing.assert_close(O.cnn_block(c["x"], {"m." + k: v for k, v in c["sd"].items()}, "m", 2, 1, train=True),
                               c["y"], rtol=1e-5, atol=1e-5)
    c = g["resblock"]
    torch.testing.assert_close(O.residual_block(c["x"], {"m." + k: v for k, v in c["sd"].items()}, "m", 2, True), c["y"])
    c = g["mdyconv"]
    torch.testing.assert_close(O.mdyconv(c["x"], {"m." + k: v for k, v in c["sd"].items()}, "m", 5, 2), c["y"],
                               rtol=1e-5, atol=1e-5)
    c = g["mdyencoder"]
    torch.testing.assert_close(O.mdy_encoder(c["x"], {"m." + k: v for k, v in c["sd"].items()}, "m"), c["y"],
                               rtol=1e-5, atol=1e-5)


def test_adaptive_stem_matches_golden():
    """AdaptiveStemLayer (DySOEM_SimFPN.py:14-25): 1-channel -> gray_conv, 3-channel -> rgb_conv; fixture generated
    from the reference class by tools/make_golden.py::gen_adaptive_stem."""
    g = load("adaptive_stem.pt")
    seen = set()
    for c in g["cases"]:
        y = O.adaptive_stem(c["x"], g["sd"], "", c["train"])
        torch.testing.assert_close(y, c["y"], rtol=1e-5, atol=1e-6)
        seen.add((c["x"].shape[1], c["train"]))
    assert seen == {(1, False), (1, True), (3, False), (3, True)}


def test_bf16_pipeline_switch_is_off_by_default_and_only_rounds():
    g = load("blocks.pt")["resblock"]
    sd = {"m." + k: v for k, v in g["sd"].items()}
    torch.testing.assert_close(O.residual_block(g["x"], sd, "m", 2, True), g["y"])       # off: exact
    with O.bf16_pipeline():
        y = O.residual_block(g["x"], sd, "m", 2, True)
    assert torch.equal(y, y.bfloat16().float())                                             # on: bf16-valued
    assert ((y - g["y"]).norm() / g["y"].norm()) < 0.02


# ---- whole models -------------------------------------------------------------------------------------
def _x64(seed):
    x = torch.rand(2, 3, 64, 64, generator=torch.Generator().manual_seed(seed))
    x[1] = x[1, :1].expand(3, -1, -1)
    return x


@pytest.mark.parametrize("name", ["baseline", "dy-yolo"])
def test_darknet_forward_matches_golden(name):
    """Seeded construction of OUR parameter containers must reproduce the reference's state_dict
    (keys, shapes, SHA-256 of the values), and the oracle forward on it must equal the reference's."""
    from tools_hash import sd_hash
    from multimodal_uav_det_b200.model import BaselineModel, DyYOLO
    from multimodal_uav_det_b200.utils.datatype import Config
    gold = load("model_forwards.pt")[name]
    torch.manual_seed(gold["seed"])
    model = {"baseline": BaselineModel, "dy-yolo": DyYOLO}[name](hparams=Config(gold["hp"])).eval()
    sd = model.state_dict()
    assert list(sd.keys()) == gold["keys"]
    assert [tuple(v.shape) for v in sd.values()] == gold["shapes"]
    assert sum(p.numel() for p in model.parameters()) == gold["n_params"]
    assert sd_hash(sd) == gold["sha256"], "seeded initialisation differs from the reference"
    with torch.no_grad():
        outs = O.darknet_forward(_x64(gold["x_seed"]), sd, gold["hp"]["layer_config"], gold["hp"].get("attn_temperature"))
    for (bb, ob), (gb, go) in zip(outs, gold["outs"]):
        torch.testing.assert_close(bb, gb, rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(ob, go, rtol=1e-4, atol=1e-5)


# ---- against the live reference (build container only) ---------------------------------------------------
@pytest.mark.needs_reference
def test_oracle_equals_reference_forward_all_models():
    from oracle import ref_import as R
    ns = R.load()
    x = torch.rand(2, 3, 128, 128, generator=torch.Generator().manual_seed(7))
    for name, build in (("baseline", lambda c, d: ns.baseline.BaselineModel(hparams=c)),
                        ("dy-yolo", lambda c, d: ns.dyyolo.DyYOLO(hparams=c))):
        cfg, hp = R.hparams(name)
        torch.manual_seed(1)
        m = build(cfg, hp)
        for train in (False, True):
            m.train(train)
            with torch.no_grad():
                sd = copy.deepcopy(m.state_dict())
                ref = m(x)
                mine = O.darknet_forward(x, sd, hp["layer_config"], hp.get("attn_temperature"), train=train)
            for r, (bb, ob) in zip(ref, mine):
                # DyYOLO: the oracle convolves sample by sample, the reference uses one grouped conv
                torch.testing.assert_close(bb, r.bbox, rtol=1e-4, atol=1e-5)
                torch.testing.assert_close(ob, r.obj, rtol=1e-4, atol=1e-5)
    cfg, hp = R.hparams("dy-soem_fpn")
    torch.manual_seed(1)
    m = R.build_dysoem(cfg, hp).eval()
    with torch.no_grad():
        ref = m(x, 30.0)
        mine = O.dysoem_simfpn_forward(x, m.state_dict(), 30.0)
    for r, (bb, ob) in zip(ref, mine):
        torch.testing.assert_close(bb, r.bbox, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(ob, r.obj, rtol=1e-5, atol=1e-6)
    m = R.build_rtm().eval()
    x2 = torch.rand(1, 3, 256, 256, generator=torch.Generator().manual_seed(8))
    with torch.no_grad():
        ref = m(x2)
        mine = O.rtm_forward(x2, m.state_dict(), torch.tensor(R.rtm_default_anchors()).float())
    for r, (bb, ob) in zip(ref, mine):
        torch.testing.assert_close(bb, r.bbox, rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(ob, r.obj, rtol=1e-5, atol=1e-6)


@pytest.mark.needs_reference
def test_golden_fixtures_are_reproducible_from_the_reference(tmp_path):
    """tools/make_golden.py regenerates byte-identical NMS / loss numbers (fixture provenance)."""
    import make_golden
    make_golden.OUT = str(tmp_path)
    make_golden.gen_nms()
    a = torch.load(os.path.join(str(tmp_path), "nms_cases.pt"), weights_only=False)
    b = load("nms_cases.pt")
    for x, y in zip(a["cases"], b["cases"]):
        assert torch.equal(x["keep"], y["keep"]) and torch.equal(x["boxes"], y["boxes"])
    make_golden.gen_adaptive_stem()
    a = torch.load(os.path.join(str(tmp_path), "adaptive_stem.pt"), weights_only=False)
    b = load("adaptive_stem.pt")
    for x, y in zip(a["cases"], b["cases"]):
        assert torch.equal(x["x"], y["x"]) and torch.equal(x["y"], y["y"])


@pytest.mark.needs_reference
def test_reference_built_state_dict_loads_strict_after_flat_trainer_rehoming():
    """A `state_dict` produced by the UNMODIFIED reference DyYOLO (what train.py's ModelCheckpoint saves) loads with
    strict=True into the product model after FlatSGDTrainer moved its parameters into flat channels-last arenas."""
    from oracle import ref_import as R
    from multimodal_uav_det_b200.model import DyYOLO
    from multimodal_uav_det_b200.parallel import FlatSGDTrainer
    from multimodal_uav_det_b200.utils.datatype import Config
    ns = R.load()
    cfg, hp = R.hparams("dy-yolo")
    torch.manual_seed(5)
    ref_sd = {k: v.clone() for k, v in ns.dyyolo.DyYOLO(hparams=cfg).state_dict().items()}
    torch.manual_seed(6)
    model = DyYOLO(hparams=Config(hp))
    FlatSGDTrainer(model, lr=1e-4, momentum=0.7)
    missing, unexpected = model.load_state_dict(ref_sd, strict=True)
    assert not missing and not unexpected
    got = model.state_dict()
    assert list(got.keys()) == list(ref_sd.keys())
    for k in ref_sd:
        assert torch.equal(got[k], ref_sd[k]), k
