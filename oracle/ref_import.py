"""TEST INFRASTRUCTURE ONLY — loads the *unmodified* reference (`/root/reference`) behind
stub modules so that its Python implementation can be executed as the parity oracle in the
build container.  Never imported by the product package; `/root/reference`
does not exist on the GPU box, so everything here is gated on `available()`.

Used by: tools/make_golden.py (fixture generation) and tests marked `needs_reference`.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("UAVDET_REFERENCE", "/root/reference")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_stubs")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "model", "_base.py"))


_loaded = {}


def load():
    """Import the reference packages `model`, `utils`, `dataset` and return a namespace.

    Work-arounds (all documented in SURVEY.md §0/§8c; none edits reference code):
      * stubs for pytorch_lightning / torchmetrics / albumentations / matplotlib / dotenv
      * `utils.metrics.filter_high_iou_bboxes` injected (RTMUAVDet.py:11 imports a missing name)
    """
    if _loaded:
        return _loaded["ns"]
    if not available():
        raise RuntimeError(f"reference not found at {REF_ROOT}")
    for p in (REF_ROOT, _STUBS):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, REF_ROOT)
    sys.path.insert(0, _STUBS)
    if "dotenv" not in sys.modules:
        try:
            importlib.import_module("dotenv")
        except Exception:
            m = types.ModuleType("dotenv")
            m.load_dotenv = lambda *a, **k: False
            sys.modules["dotenv"] = m
    # our own package also has a `utils`-free layout, so no name clash is possible
    utils_metrics = importlib.import_module("utils.metrics")
    if not hasattr(utils_metrics, "filter_high_iou_bboxes"):
        def _missing(*a, **k):
            raise RuntimeError("filter_high_iou_bboxes does not exist in the reference (D4)")
        utils_metrics.filter_high_iou_bboxes = _missing
    ns = types.SimpleNamespace()
    ns.base = importlib.import_module("model._base")
    ns.baseline = importlib.import_module("model.BaselineModel")
    ns.dyyolo = importlib.import_module("model.DyYOLO")
    ns.dysoem = importlib.import_module("model.DySOEM_SimFPN")
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ns.rtm = importlib.import_module("model.RTMUAVDet")
    ns.datatype = importlib.import_module("utils.datatype")
    ns.metrics = utils_metrics
    ns.postprocess = importlib.import_module("utils.postprocess")
    ns.dataset = importlib.import_module("dataset.AntiUAVDataset")
    _loaded["ns"] = ns
    return ns


def hparams(name: str):
    """`conf/model/<name>.yaml`'s hparams wrapped in the reference's own Config (datatype.py:13)."""
    import yaml
    ns = load()
    with open(os.path.join(REF_ROOT, "conf", "model", f"{name}.yaml")) as f:
        doc = yaml.safe_load(f)
    hp = dict(doc["hparams"])
    if "optim" not in hp and "optim" in doc:          # D5(a): dy-soem_fpn.yaml puts optim outside
        hp["optim"] = doc["optim"]
    return ns.datatype.Config(hp), hp


def build_dysoem(hp_cfg, hp_dict):
    """Construct the reference DySOEM_SimFPN despite D5(b): YOLOHead is called one positional
    short.  We rebind the module-level name to a wrapper that supplies head_scales/bbox_loss_fn."""
    ns = load()
    real = ns.base.YOLOHead

    def head3(x_channels, anchors, loss_balancing):
        return real(x_channels, anchors, hp_dict["head_scales"], loss_balancing, hp_dict["bbox_loss_fn"])

    ns.dysoem.YOLOHead = head3
    try:
        return ns.dysoem.DySOEM_SimFPN(hparams=hp_cfg)
    finally:
        ns.dysoem.YOLOHead = real


def build_rtm(anchors=None):
    import torch
    import warnings
    ns = load()
    if anchors is None:
        anchors = rtm_default_anchors()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return ns.rtm.RTMUAVDet(input_size=[3, 640, 640], anchors=torch.tensor(anchors).float(),
                                learning_rate=1e-4)


def rtm_default_anchors():
    # RTMUAVDet has no shipped config (D4); two scales × three anchors, from the small/medium rows
    # of conf/model/baseline.yaml:3-7.
    return [[[29, 23], [48, 30], [67, 38]], [[91, 54], [120, 75], [157, 60]]]
