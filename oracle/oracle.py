"""ORACLE — TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's hot path (alfialdo/multimodal-uav-det).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may import this
module, and only as the checker or the timed CPU baseline — never the product package.

Two kinds of function live here:
  * integer / bit-exact work (NMS keep indices)          -> plain C (oracle/nms_oracle.c) + a
    pure-Python loop for tiny cases;
  * floating-point work (decode, conv/BN/activation stacks, dynamic convs, necks, heads, loss)
    -> functional fp32 PyTorch-on-CPU restatements that read a reference-named `state_dict`.

Parity pins: the reference has no tests or golden vectors of its own, so each
function here is pinned (a) against the reference itself, imported unmodified through
oracle/ref_import.py when /root/reference is present (tests/test_oracle.py), and (b) against
the fixtures in tests/golden/ generated from the reference by tools/make_golden.py.
Third-party arithmetic (torchvision nms / box_convert / box_iou / complete_box_iou_loss) is
restated from torchvision 0.26.0 (reference pins 0.19.1, requirements.txt:177) and pinned
against the installed torchvision.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_ref", "liboracle.so")


def build(force: bool = False) -> str:
    """Compile oracle/nms_oracle.c -> oracle/_ref/liboracle.so (gcc, -ffp-contract=off)."""
    src = os.path.join(_HERE, "nms_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return _LIB


_clib = None


def _c():
    global _clib
    if _clib is None:
        build()
        lib = ctypes.CDLL(_LIB)
        lib.oracle_nms.restype = ctypes.c_int64
        lib.oracle_nms.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_double,
                                   ctypes.c_void_p]
        _clib = lib
    return _clib


# ------------------------------------------------------------------------------------------------
# NMS  (model/_base.py:203 -> torchvision.ops.nms)
# ------------------------------------------------------------------------------------------------
def nms(boxes, scores, iou_threshold: float) -> np.ndarray:
    """boxes (N,4) xyxy fp32, scores (N,) fp32 -> int64 kept indices in score order."""
    b = np.ascontiguousarray(np.asarray(boxes, dtype=np.float32))
    s = np.ascontiguousarray(np.asarray(scores, dtype=np.float32))
    n = s.shape[0]
    keep = np.empty(max(n, 1), dtype=np.int64)
    k = _c().oracle_nms(b.ctypes.data, s.ctypes.data, n, float(iou_threshold), keep.ctypes.data)
    return keep[:k].copy()


def nms_py(boxes, scores, iou_threshold: float) -> List[int]:
    """Pure-Python/NumPy-scalar greedy loop (small cases only) — same rules as nms_oracle.c."""
    b = np.asarray(boxes, dtype=np.float32)
    s = np.asarray(scores, dtype=np.float32)
    n = len(s)
    keyed = [(-math.inf if math.isnan(float(v)) else -float(v), i) for i, v in enumerate(s)]
    order = [i for _, i in sorted(keyed, key=lambda t: t[0])]  # Python's sort is stable
    f32 = np.float32
    areas = [(f32(b[i, 2]) - f32(b[i, 0])) * (f32(b[i, 3]) - f32(b[i, 1])) for i in range(n)]
    sup = [False] * n
    keep = []
    smax = lambda a, c: c if a < c else a
    smin = lambda a, c: c if c < a else a
    for a_ in range(n):
        i = order[a_]
        if sup[i]:
            continue
        keep.append(i)
        for b_ in range(a_ + 1, n):
            j = order[b_]
            if sup[j]:
                continue
            xx1, yy1 = smax(b[i, 0], b[j, 0]), smax(b[i, 1], b[j, 1])
            xx2, yy2 = smin(b[i, 2], b[j, 2]), smin(b[i, 3], b[j, 3])
            w = smax(f32(0), f32(xx2 - xx1))
            h = smax(f32(0), f32(yy2 - yy1))
            inter = f32(w * h)
            with np.errstate(all="ignore"):
                ovr = f32(inter / f32(f32(areas[i] + areas[j]) - inter))
            if float(ovr) > iou_threshold:
                sup[j] = True
    return keep


# ------------------------------------------------------------------------------------------------
# decode  (model/_base.py:214-248, model/RTMUAVDet.py:274-291)
# ------------------------------------------------------------------------------------------------
def cxcywh_to_xyxy(t: torch.Tensor) -> torch.Tensor:
    """torchvision.ops.box_convert(in_fmt='cxcywh', out_fmt='xyxy')."""
    cx, cy, w, h = t.unbind(-1)
    return torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], dim=-1)


def decode_yolo_head(bbox_logits: torch.Tensor, scaled_anchors: torch.Tensor, ciou: bool) -> torch.Tensor:
    """One image, one head: (A,H,W,4) logits -> decoded cxcywh (_base.py:214-241)."""
    sig = torch.sigmoid(bbox_logits)
    cx = sig[..., 0] * 2 - 0.5
    cy = sig[..., 1] * 2 - 0.5
    w = (sig[..., 2] * 2) ** 2
    h = (sig[..., 3] * 2) ** 2
    if ciou:
        a, hh, ww, _ = bbox_logits.shape
        gx = torch.arange(ww).view(1, 1, ww).expand(a, hh, ww)
        gy = torch.arange(hh).view(1, hh, 1).expand(a, hh, ww)
        cx = cx + gx
        cy = cy + gy
        w = w * scaled_anchors[:, 0].view(a, 1, 1)
        h = h * scaled_anchors[:, 1].view(a, 1, 1)
    return torch.stack([cx, cy, w, h], dim=-1)


def decode_yolo(outs: Sequence[Tuple[torch.Tensor, torch.Tensor]], anchors, head_scales, ciou: bool
                ) -> Tuple[torch.Tensor, torch.Tensor]:
    """outs: per head (bbox (B,A,H,W,4), obj (B,A,H,W,1)).  Returns boxes (B,N,4) xyxy in grid
    units and scores (B,N) raw logits, heads concatenated (_base.py:196-203)."""
    anc = torch.tensor(anchors).float()
    hs = torch.tensor(head_scales)
    b = outs[0][0].shape[0]
    all_b, all_s = [], []
    for i in range(b):
        bb, ss = [], []
        for hi, (bbox, obj) in enumerate(outs):
            dec = decode_yolo_head(bbox[i].float(), anc[hi] / hs[hi], ciou)
            bb.append(cxcywh_to_xyxy(dec.reshape(-1, 4)))
            ss.append(obj[i].float().reshape(-1))
        all_b.append(torch.cat(bb))
        all_s.append(torch.cat(ss))
    return torch.stack(all_b), torch.stack(all_s)


def decode_rtm(bbox_sig: torch.Tensor, anchors_head: torch.Tensor) -> torch.Tensor:
    """(B,A,H,W,4) sigmoid outputs -> decoded (RTMUAVDet.py:285-289); anchors not stride-scaled."""
    b, a, h, w, _ = bbox_sig.shape
    gx = torch.arange(w).view(1, 1, 1, w)
    gy = torch.arange(h).view(1, 1, h, 1)
    aw = anchors_head[:, 0].view(1, a, 1, 1)
    ah = anchors_head[:, 1].view(1, a, 1, 1)
    px = bbox_sig[..., 0] * 2 - 0.5 + gx
    py = bbox_sig[..., 1] * 2 - 0.5 + gy
    pw = (bbox_sig[..., 2] * 2) ** 2 * aw
    ph = (bbox_sig[..., 3] * 2) ** 2 * ah
    return torch.stack([px, py, pw, ph], dim=-1)


# ------------------------------------------------------------------------------------------------
# layer stacks (functional, fp32, reference state_dict names)
# ------------------------------------------------------------------------------------------------
# ---- optional emulation of the product's bf16 storage points ---------------------------------------
# With `bf16_pipeline(True)` the functions below round exactly where the CUDA path stores bf16:
# conv inputs / weights, the pre-BN conv output, the activated output (after the residual add) and,
# in backward, the gradients dx / d_raw.  With it off (default) they are the reference-exact fp32
# restatement.  The switch exists because random-init batch-stat BN stacks amplify rounding noise
# ~1.15x per layer: a tight check of backward needs the same rounding points on both sides.
_Q = {"on": False}


class bf16_pipeline:
    def __init__(self, on=True):
        self.on = on

    def __enter__(self):
        self.prev = _Q["on"]
        _Q["on"] = self.on

    def __exit__(self, *a):
        _Q["on"] = self.prev


class _RoundFwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundBwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(torch.float32)


def qf(x):
    return _RoundFwd.apply(x) if _Q["on"] else x


def qb(x):
    return _RoundBwd.apply(x) if _Q["on"] else x


def _bn(x, sd, p, train, eps=1e-5, momentum=0.1):
    return F.batch_norm(x, None if train else sd[p + ".running_mean"], None if train else sd[p + ".running_var"],
                        sd[p + ".weight"], sd[p + ".bias"], training=train, momentum=momentum, eps=eps)


def cnn_block(x, sd, p, stride=1, pad=0, train=False, res=None):
    """CNNBlock: conv(no bias) -> BN -> LeakyReLU(0.1)   (BaselineModel.py:10-22) (+ res)."""
    w = sd[p + ".conv.weight"]
    stem = w.shape[1] < 32          # network input: no gradient flows back, but it is a bf16 storage point too
    # (the product's stem reads bf16 im2col patches and bf16 weights when cin*k*k <= 32, i.e. for k > 1)
    stem_fp32 = stem and not (w.shape[-1] > 1 and w.shape[1] * w.shape[-1] * w.shape[-1] <= 32)
    raw = F.conv2d(x if stem_fp32 else (qf(x) if stem else qb(x)), w if stem_fp32 else qf(w), None, stride, pad)
    y = F.leaky_relu(_bn(qf(qb(raw)), sd, p + ".bn", train), 0.1)
    if res is not None:
        y = y + res
    return qf(y)


def residual_block(x, sd, p, repeats, use_residual, train=False):
    """ResidualBlock (BaselineModel.py:25-45): x = seq(x) + use_residual * x."""
    for r in range(repeats):
        y = cnn_block(x, sd, f"{p}.layers.{r}.0", 1, 0, train)
        x = cnn_block(y, sd, f"{p}.layers.{r}.1", 1, 1, train, res=float(use_residual) * x)
    return x


def dyconv_module(x, sd, p, k, stride, pad, temperature, train=False):
    """DyConvModule (_base.py:26-77): GAP -> 1x1 -> ReLU -> 1x1(+b) -> softmax(/T) -> expert mix
    -> per-sample conv -> BN -> SiLU."""
    b, c = x.shape[:2]
    pooled = x.mean(dim=(2, 3))
    w1 = sd[p + ".attention.1.weight"].flatten(1)
    w2 = sd[p + ".attention.3.weight"].flatten(1)
    hid = F.relu(pooled @ w1.t())
    logits = hid @ w2.t() + sd[p + ".attention.3.bias"]
    attn = F.softmax(logits / temperature, dim=1)
    bank = sd[p + ".weights"]                                   # (K, O, I, k, k)
    filt = (attn @ bank.flatten(1)).view(b, *bank.shape[1:])    # (B, O, I, k, k)
    outs = [F.conv2d(x[i:i + 1], filt[i], None, stride, pad) for i in range(b)]
    y = torch.cat(outs, 0)
    return F.silu(_bn(y, sd, p + ".bn", train)), attn


def yolo_head(feats, sd, p="yolo_head.detection_head"):
    """YOLOHead.forward (_base.py:144-153): per scale 1x1 convs -> (B,A,H,W,1)/(B,A,H,W,4) logits."""
    outs = []
    for s, f in enumerate(feats):
        f = qb(f)
        o = qb(F.conv2d(f, qf(sd[f"{p}.{s}.obj.conv_obj.weight"]), sd[f"{p}.{s}.obj.conv_obj.bias"]))
        bb = qb(F.conv2d(f, qf(sd[f"{p}.{s}.bbox.conv_bbox.weight"]), sd[f"{p}.{s}.bbox.conv_bbox.bias"]))
        b, _, h, w = o.shape
        a = o.shape[1]
        outs.append((bb.view(b, a, 4, h, w).permute(0, 1, 3, 4, 2).contiguous(),
                     o.view(b, a, 1, h, w).permute(0, 1, 3, 4, 2).contiguous()))
    return outs


def darknet_forward(x, sd, layer_config, attn_temperature=None, train=False, taps=None, route_repeats=8):
    """BaselineModel.forward / DyYOLO.forward (BaselineModel.py:105-124, DyYOLO.py:122-144).
    Returns per-head (bbox, obj).  `taps`, if a dict, receives named intermediate tensors."""
    idx = 0
    feats, routes = [], []
    for entry in layer_config:
        kind = entry[0]
        if kind == "B":
            x = residual_block(x, sd, f"layers.{idx}", entry[1], True, train)
            if entry[1] == route_repeats:   # the reference hard-codes 8 (BaselineModel.py:116)
                x = qb(x)                   # (bf16 pipeline: the summed route gradient is stored as bf16)
                routes.append(x)
            idx += 1
        elif kind == "S":
            x = residual_block(x, sd, f"layers.{idx}", 1, False, train)
            x = cnn_block(x, sd, f"layers.{idx + 1}", 1, 0, train)
            feats.append(cnn_block(x, sd, f"layers.{idx + 2}.conv", 1, 1, train))
            idx += 3
        elif kind == "U":
            x = torch.cat([F.interpolate(x, scale_factor=2, mode="nearest"), routes.pop()], dim=1)
            idx += 1
        elif kind == "DyConv":
            _, _, k, s = entry
            x, _ = dyconv_module(x, sd, f"layers.{idx}", k, s, 1 if k == 3 else 0, attn_temperature, train)
            idx += 1
        else:
            _, k, s = entry
            x = cnn_block(x, sd, f"layers.{idx}", s, 1 if k == 3 else 0, train)
            idx += 1
        if taps is not None:
            taps[f"after_{idx - 1}"] = x
    return yolo_head(feats, sd)


def conv_module(x, sd, p, stride=1, pad=0, act="silu", train=False, eps=1e-5, momentum=0.1):
    """ConvModule: conv -> BN -> SiLU|ReLU (_base.py:14-24; RTMUAVDet.py:15-25 with eps/momentum)."""
    y = F.conv2d(x, sd[p + ".conv.0.weight"], sd.get(p + ".conv.0.bias"), stride, pad)
    y = _bn(y, sd, p + ".conv.1", train, eps, momentum)
    return F.silu(y) if act == "silu" else F.relu(y)


def adaptive_stem(x, sd, p, train=False):
    """AdaptiveStemLayer.forward (DySOEM_SimFPN.py:14-25): a 1-channel input goes through `gray_conv`, anything
    else through `rgb_conv`; each is a 1x1 ConvModule(bias=False, SiLU)."""
    branch = "gray_conv" if x.size(1) == 1 else "rgb_conv"
    return conv_module(x, sd, f"{p}.{branch}" if p else branch, 1, 0, "silu", train)


def space_to_depth2(x):
    """DySOEM_SimFPN.py:71-75: cat of the four stride-2 phase slices, phase n = i*2 + j."""
    return torch.cat([x[..., i::2, j::2] for i in range(2) for j in range(2)], dim=1)


def dynamic_soem(x, sd, p, temperature, train=False):
    """DynamicSOEM.forward (DySOEM_SimFPN.py:66-94) as the reference executes it: K full convs
    each scaled by its attention weight, summed, BN, SiLU."""
    f = space_to_depth2(x)
    pooled = f.mean(dim=(2, 3))
    hid = F.relu(F.linear(pooled, sd[p + ".attention.2.weight"], sd[p + ".attention.2.bias"]))
    attn = F.softmax(F.linear(hid, sd[p + ".attention.4.weight"], sd[p + ".attention.4.bias"]) / temperature, dim=-1)
    k = 0
    acc = None
    while f"{p}.dy_convs.{k}.weight" in sd:
        w = sd[f"{p}.dy_convs.{k}.weight"]
        y = F.conv2d(f, w, sd[f"{p}.dy_convs.{k}.bias"], 1, w.shape[-1] // 2) * attn[:, k].view(-1, 1, 1, 1)
        acc = y if acc is None else acc + y
        k += 1
    return F.silu(_bn(acc, sd, p + ".bn", train))


def simplified_fpn(feats, sd, p="neck", train=False):
    """SimplifiedFPN.forward (DySOEM_SimFPN.py:114-126) — x1 is counted twice on purpose."""
    x0, x1, x2 = feats
    up = lambda t: F.interpolate(t, scale_factor=2, mode="nearest")
    center = x1 + F.conv2d(up(x2), sd[p + ".x2_in_down.weight"], sd[p + ".x2_in_down.bias"]) + x1
    x0 = x0 + F.conv2d(up(center), sd[p + ".center_down.weight"], sd[p + ".center_down.bias"])
    x1 = center + F.conv2d(x0, sd[p + ".x0_out_up.weight"], sd[p + ".x0_out_up.bias"], stride=2)
    x2 = x2 + F.conv2d(x1, sd[p + ".x1_out_up.weight"], sd[p + ".x1_out_up.bias"], stride=2)
    return (conv_module(x0, sd, p + ".x0_conv_out", 1, 1, "silu", train),
            conv_module(x1, sd, p + ".x1_conv_out", 1, 1, "silu", train),
            conv_module(x2, sd, p + ".x2_conv_out", 1, 1, "silu", train))


def dysoem_simfpn_forward(x, sd, attn_temp=1.0, train=False):
    """DySOEM_SimFPN.forward (DySOEM_SimFPN.py:149-170)."""
    x = conv_module(x, sd, "input_stem.conv", 1, 0, "silu", train)
    feats = []
    i = 0
    while f"backbone.{i}.bn.weight" in sd:
        x = dynamic_soem(x, sd, f"backbone.{i}", attn_temp, train)
        feats.append(x)
        i += 1
    return yolo_head(simplified_fpn(feats, sd, "neck", train), sd)


# ------------------------------------------------------------------------------------------------
# loss  (model/_base.py:155-212, utils/metrics.py:8-84, utils/postprocess.py:51-85)
# ------------------------------------------------------------------------------------------------
def _box_iou_first(pred_xyxy, tgt_xyxy):
    """torchvision.ops.box_iou(pred, tgt)[:, 0] — IoU of every row against the FIRST target."""
    t = tgt_xyxy[0]
    area_p = (pred_xyxy[:, 2] - pred_xyxy[:, 0]) * (pred_xyxy[:, 3] - pred_xyxy[:, 1])
    area_t = (t[2] - t[0]) * (t[3] - t[1])
    lt = torch.max(pred_xyxy[:, :2], t[:2])
    rb = torch.min(pred_xyxy[:, 2:], t[2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[:, 0] * wh[:, 1]
    return inter / (area_p + area_t - inter)


def _ciou_loss(b1, b2, eps=1e-7):
    """torchvision.ops.complete_box_iou_loss(reduction='none') on xyxy rows."""
    x1, y1, x2, y2 = b1.unbind(-1)
    x1g, y1g, x2g, y2g = b2.unbind(-1)
    xkis1, ykis1 = torch.max(x1, x1g), torch.max(y1, y1g)
    xkis2, ykis2 = torch.min(x2, x2g), torch.min(y2, y2g)
    inter = torch.zeros_like(x1)
    mask = (ykis2 > ykis1) & (xkis2 > xkis1)
    inter[mask] = (xkis2[mask] - xkis1[mask]) * (ykis2[mask] - ykis1[mask])
    union = (x2 - x1) * (y2 - y1) + (x2g - x1g) * (y2g - y1g) - inter + eps
    iou = inter / union
    xc1, yc1 = torch.min(x1, x1g), torch.min(y1, y1g)
    xc2, yc2 = torch.max(x2, x2g), torch.max(y2, y2g)
    diag = (xc2 - xc1) ** 2 + (yc2 - yc1) ** 2 + eps
    cdist = ((x1 + x2) / 2 - (x1g + x2g) / 2) ** 2 + ((y1 + y2) / 2 - (y1g + y2g) / 2) ** 2
    diou = 1 - iou + cdist / diag
    wp, hp, wg, hg = x2 - x1, y2 - y1, x2g - x1g, y2g - y1g
    v = (4 / (math.pi ** 2)) * torch.pow(torch.atan(wg / hg) - torch.atan(wp / hp), 2)
    with torch.no_grad():
        alpha = v / (1 - iou + v + eps)
    return diou + alpha * v


def yolo_loss(outs, targets, anchors, head_scales, loss_balancing: Dict, bbox_loss_fn: str):
    """YOLOHead.compute_metrics(return_ap=False) (_base.py:155-212).
    outs: per head (bbox (B,A,H,W,4), obj (B,A,H,W,1)); targets: per sample, per head (A,H,W,5)
    [obj, cx, cy, w, h] grid units — NOT mutated here (the reference rewrites batch.bbox in place,
    _base.py:257; callers that need that side effect apply `build_target_bbox` themselves).
    Returns (total, bbox_loss, obj_loss)."""
    anc = torch.tensor(anchors).float()
    hs = torch.tensor(head_scales)
    bsz = outs[0][0].shape[0]
    ciou = bbox_loss_fn == "ciou"
    bbox_l = torch.zeros(())
    obj_l = torch.zeros(())
    for i in range(bsz):
        for hi, (bbox, obj) in enumerate(outs):
            sa = anc[hi] / hs[hi]
            p_bbox, p_obj = bbox[i], obj[i]
            tgt = targets[i][hi]
            cell = tgt[..., 0] == 1.0
            t_bbox, t_obj = tgt[..., 1:].clone(), tgt[..., 0]
            dec = decode_yolo_head(p_bbox, sa, ciou)
            # calculate_iou (postprocess.py:51-85)
            pb = dec.detach().clone()
            if not ciou:
                pb[..., 2:] = pb[..., 2:] * sa.view(-1, 1, 1, 2)
            ious = _box_iou_first(cxcywh_to_xyxy(pb[cell]), cxcywh_to_xyxy(t_bbox[cell]))
            # __build_target_bbox (_base.py:250-270)
            a, hh, ww, _ = t_bbox.shape
            if ciou:
                t_bbox[..., 0] = t_bbox[..., 0] + torch.arange(ww).view(1, 1, ww)
                t_bbox[..., 1] = t_bbox[..., 1] + torch.arange(hh).view(1, hh, 1)
            else:
                t_bbox[..., 2:] = torch.sqrt((1e-16 + t_bbox[..., 2:]) / sa.view(-1, 1, 1, 2)) / 2
            if ciou:
                bl = _ciou_loss(cxcywh_to_xyxy(dec[cell]), cxcywh_to_xyxy(t_bbox[cell])).mean()
            else:
                bl = F.mse_loss(dec[cell], t_bbox[cell])
            bbox_l = bbox_l + loss_balancing["bbox_w"] * bl
            ol = F.binary_cross_entropy_with_logits(p_obj[cell].squeeze(-1), ious * t_obj[cell])
            obj_l = obj_l + loss_balancing["objectness_w"] * ol * loss_balancing["obj_scales_w"][hi]
            nl = F.binary_cross_entropy_with_logits(p_obj[~cell].squeeze(-1), t_obj[~cell])
            obj_l = obj_l + loss_balancing["no_obj_w"] * nl
    bbox_l = bbox_l / bsz
    obj_l = obj_l / bsz
    return bbox_l + obj_l, bbox_l, obj_l


def anchor_iou_order(w, h, head_anchors):
    """dataset/_helper.py:308-330: IoU of a (w,h) box against each anchor (both at the origin),
    returned as (indices sorted by IoU descending, sorted IoUs)."""
    aw, ah = head_anchors[..., 0], head_anchors[..., 1]
    inter = torch.min(aw, w) * torch.min(ah, h)
    iou = inter / (aw * ah + w * h - inter)
    order = torch.argsort(iou, descending=True)
    return order, iou[order]


def encode_targets(box_xyxy_px: torch.Tensor, anchors, head_scales, input_size: int = 640,
                   grids: Sequence[int] = None) -> List[torch.Tensor]:
    """AntiUAVDataset.__generate_yolo_bboxes (dataset/AntiUAVDataset.py:141-185) for ONE image with
    ONE target box (Anti-UAV has exactly one, :52-53).  box (1,4) xyxy pixels -> per head
    (A,S,S,5) [obj, cx_off, cy_off, w_cells, h_cells].  `grids` overrides S (default
    input_size // head_scale, :28)."""
    anc = torch.tensor(anchors).float() / input_size
    sizes = [input_size // s for s in head_scales] if grids is None else list(grids)
    x1, y1, x2, y2 = box_xyxy_px.reshape(4).float().unbind()
    cxcywh = torch.stack([(x1 + x2) / 2, (y1 + y2) / 2, x2 - x1, y2 - y1]) / input_size
    cx, cy, w, h = cxcywh
    out = []
    for hi, size in enumerate(sizes):
        t = torch.zeros(anc.shape[1], size, size, 5)
        gcx, gcy = cx * size, cy * size
        gx, gy = int(gcx), int(gcy)
        cell = torch.stack([gcx - gx, gcy - gy, w * size, h * size])
        order, ious = anchor_iou_order(w, h, anc[hi])
        if ious[0] < 0.5:
            t[order[0], gy, gx, 0] = 1.0
            t[order[0], gy, gx, 1:5] = cell
        else:
            for a, iou in zip(order, ious):
                t[a, gy, gx, 0] = 1.0 if iou >= 0.5 else 0.0
                t[a, gy, gx, 1:5] = cell
        out.append(t)
    return out


# ------------------------------------------------------------------------------------------------
# RTMUAVDet  (model/RTMUAVDet.py — deprecated in the reference but forward-runnable)
# ------------------------------------------------------------------------------------------------
def rtm_conv_module(x, sd, p, stride=1, pad=0, act="silu", train=False, eps=1e-3, momentum=0.03):
    """RTMUAVDet.ConvModule (RTMUAVDet.py:15-25): defaults eps=1e-3, momentum=0.03."""
    return conv_module(x, sd, p, stride, pad, act, train, eps, momentum)


def mdyconv(x, sd, p, k, pad, train=False):
    """MDyConv.forward (RTMUAVDet.py:68-100): 1x1 ConvModule(ReLU, eps 1e-5) -> GAP -> 1x1(+b)+ReLU ->
    channel_fc (C) x kernel_fc (k*k) outer product -> per-sample depthwise conv -> + residual."""
    y = rtm_conv_module(x, sd, p + ".base_conv", 1, 0, "relu", train, 1e-5, 0.1)
    b, c = y.shape[:2]
    pooled = y.mean(dim=(2, 3), keepdim=True)
    a = F.relu(F.conv2d(pooled, sd[p + ".attention.1.weight"], sd[p + ".attention.1.bias"]))
    ch_w = F.conv2d(a, sd[p + ".channel_fc.weight"], sd[p + ".channel_fc.bias"])          # (B,C,1,1)
    k_w = F.conv2d(a, sd[p + ".kernel_fc.weight"], sd[p + ".kernel_fc.bias"]).view(b, 1, k, k)
    filt = (k_w * ch_w).reshape(b * c, 1, k, k)
    out = F.conv2d(y.reshape(1, b * c, *y.shape[2:]), filt, None, 1, pad, groups=b * c).view_as(y)
    return out + y


def mdycsp_module(x, sd, p, train=False):
    """MDyCSPModule.forward (RTMUAVDet.py:122-140)."""
    x = rtm_conv_module(x, sd, p + ".base_conv", 2, 1, "silu", train)
    x1 = rtm_conv_module(x, sd, p + ".conv1", 1, 0, "silu", train)
    x2 = rtm_conv_module(x, sd, p + ".conv2", 1, 0, "silu", train)
    x1 = mdyconv(x1, sd, p + ".mdy_conv", 3, 1, train)
    x1 = rtm_conv_module(x1, sd, p + ".transition1", 1, 0, "silu", train)
    return rtm_conv_module(torch.cat([x1, x2], 1), sd, p + ".transition2", 1, 1, "silu", train)


def mdy_encoder(x, sd, p, train=False):
    """MDyEncoder.forward (RTMUAVDet.py:163-184).  Dropout(0.2) is identity in eval mode."""
    res = x
    c = x.shape[1]
    y = F.group_norm(x, 1, sd[p + ".group_norm_in.weight"], sd[p + ".group_norm_in.bias"], 1e-5)
    y = torch.cat([mdyconv(y, sd, p + ".mdy_conv_1x1", 1, 0, train), mdyconv(y, sd, p + ".mdy_conv_3x3", 3, 1, train),
                   mdyconv(y, sd, p + ".mdy_conv_5x5", 5, 2, train)], 1)
    y = y + res
    y = F.group_norm(y, 1, sd[p + ".group_norm_out.weight"], sd[p + ".group_norm_out.bias"], 1e-5)
    y = F.gelu(F.conv2d(y, sd[p + ".channel_mlp.0.weight"], sd[p + ".channel_mlp.0.bias"]))
    if train:
        raise NotImplementedError("train-mode Dropout(0.2) is stochastic; the oracle covers eval mode")
    return F.conv2d(y, sd[p + ".channel_mlp.3.weight"], sd[p + ".channel_mlp.3.bias"])


def rtm_forward(x, sd, anchors, train=False):
    """RTMUAVDet.forward (RTMUAVDet.py:336-345) incl. RTMHead sigmoid + in-forward decode (:274-310).
    anchors: (2,3,2) tensor.  Returns per head (decoded bbox (B,A,H,W,4), obj (B,A,H,W,1))."""
    x = rtm_conv_module(x, sd, "backbone.MDyCSP_1.0.conv", 2, 1, "silu", train)           # 5x5 s2 p1 stem
    x1 = mdycsp_module(x, sd, "backbone.MDyCSP_1.1", train)
    x2 = mdycsp_module(x1, sd, "backbone.MDyCSP_2", train)
    up = F.interpolate(x2, scale_factor=2, mode="bilinear")
    f = F.conv2d(up, sd["neck.upsample.1.weight"], sd["neck.upsample.1.bias"], 1, 1)
    x1 = mdy_encoder(torch.cat([x1, f], 1), sd, "neck.encoder_x1", train)
    d = F.conv2d(x1, sd["neck.downsample.weight"], sd["neck.downsample.bias"], 2, 1)
    x2 = mdy_encoder(torch.cat([x2, d], 1), sd, "neck.encoder_x2", train)
    outs = []
    for i, f_map in enumerate((x1, x2)):
        p = f"head.detection_head.{i}"
        # attribute names are swapped in the reference (obj head holds `conv_bbox`), RTMUAVDet.py:223,243
        o = torch.sigmoid(F.conv2d(f_map, sd[p + ".obj.conv_bbox.weight"], sd[p + ".obj.conv_bbox.bias"]))
        bb = torch.sigmoid(F.conv2d(f_map, sd[p + ".bbox.conv_obj.weight"], sd[p + ".bbox.conv_obj.bias"]))
        b, a, h, w = o.shape
        o = o.view(b, a, 1, h, w).permute(0, 1, 3, 4, 2).contiguous()
        bb = bb.view(b, a, 4, h, w).permute(0, 1, 3, 4, 2).contiguous()
        outs.append((decode_rtm(bb, anchors[i]), o))
    return outs
