"""Shared building blocks of the detector, B200-native (reference model/_base.py).

Class names, constructor/forward signatures and `state_dict` keys follow the reference so the
models drop into `train.py:22-27`; the arithmetic runs in libuavdet_b200.so.  The torch layers
inside each block are parameter containers created in the reference's order (identical RNG
consumption -> identical seeded init) and are never called.

Tensor convention at module boundaries: standalone blocks (`ConvModule`, `DyConvModule`,
`YOLOHead.forward`) accept/return NCHW fp32 like the reference (API edge: one layout-conversion
kernel each way); inside the model classes everything stays NHWC bf16 (`*_nhwc` methods).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
from torch import nn

from .. import ops
from ..engine import ConvUnit, DynSpec, Executor, param_epoch
from ..utils.datatype import BatchData, DetectionResults
from ..utils.metrics import yolo_head_loss, yolo_head_loss_fused

try:  # Lightning is optional: subclass it when importable (train.py drives it), else a no-op shim
    import pytorch_lightning as pl
    LightningModule = pl.LightningModule
except Exception:  # pragma: no cover - depends on the environment
    class LightningModule(nn.Module):
        def log(self, *args, **kwargs):
            return None

        def log_dict(self, *args, **kwargs):
            return None


_ACT_LAYERS = {"silu": lambda: nn.SiLU(inplace=True), "relu": lambda: nn.ReLU(inplace=True)}


def to_nhwc(x: torch.Tensor, channels: Optional[int] = None) -> torch.Tensor:
    """API edge: NCHW (fp32, or bf16/fp16 under autocast) -> NHWC bf16.  A bf16 tensor is taken as the package's own
    NHWC activation only if its shape says so (`channels` = the consumer's input channels, when known): an NCHW bf16
    tensor handed in under autocast is converted, not misread."""
    if x.dtype == torch.bfloat16 and x.dim() == 4:
        if channels is None or (x.shape[3] == channels and (x.shape[1] != channels or x.stride(3) == 1)):
            return x
    return ops.nchw_f32_to_nhwc(x.float().contiguous())


def to_nchw(x: torch.Tensor) -> torch.Tensor:
    return ops.nhwc_to_nchw_f32(x)


class ConvModule(LightningModule):
    """conv -> BN -> SiLU|ReLU (reference _base.py:14-24).  state_dict: conv.0.weight, conv.1.*"""

    def __init__(self, in_channels, out_channels, kernel_size=(1, 1), stride=(1, 1), padding=0, bias=False,
                 activation="silu", eps=1e-5, momentum=0.1):
        super().__init__()
        self.conv = nn.Sequential(
            nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, bias=bias),
            nn.BatchNorm2d(out_channels, eps=eps, momentum=momentum, affine=True, track_running_stats=True),
            _ACT_LAYERS["silu" if activation == "silu" else "relu"](),
        )
        self.activation = "silu" if activation == "silu" else "relu"
        self._exec = Executor()

    def unit(self, stem: bool = False, s2d: bool = False) -> ConvUnit:
        return ConvUnit(self.conv[0], self.conv[1], self.activation, stem=stem, s2d=s2d)

    def forward(self, x):
        cin = self.conv[0].in_channels
        stem = cin < 32
        y = self._exec.conv_forward(self.unit(stem=stem), x.float().contiguous() if stem else to_nhwc(x, cin),
                                    self.training, None)
        self._exec.end_forward()
        return to_nchw(y)


class DyConvModule(LightningModule):
    """Attention-weighted mixture of `num_dy_conv` expert kernels, applied per sample
    (reference _base.py:26-77).  The per-sample kernel sum_k a_k W_k is aggregated in fp32 and
    fed to the implicit GEMM as a batched B operand (one weight matrix per image)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=0, num_dy_conv=4):
        super().__init__()
        self.num_dy_conv = num_dy_conv
        self.stride = stride
        self.padding = padding
        self.out_channels = out_channels
        self.kernel_size = kernel_size
        self.in_channels = in_channels
        hidden = num_dy_conv if in_channels == 3 else int(in_channels * 0.25) + 1
        self.attention = nn.Sequential(
            nn.AdaptiveAvgPool2d(1),
            nn.Conv2d(in_channels, hidden, kernel_size=1, bias=False),
            nn.ReLU(inplace=True),
            nn.Conv2d(hidden, num_dy_conv, kernel_size=1, bias=True),
        )
        self.weights = nn.Parameter(torch.randn(num_dy_conv, out_channels, in_channels, kernel_size, kernel_size),
                                    requires_grad=True)
        self.bn = nn.BatchNorm2d(num_features=out_channels, affine=True)
        self.silu = nn.SiLU(inplace=True)
        self._exec = Executor()

    # ---- NHWC path used by DyYOLO ----------------------------------------------------------------
    def dyn_spec(self, attn_temp) -> DynSpec:
        return DynSpec(bank=lambda: self.weights.detach(), bias_bank=None,
                       bank_params=lambda: [(self.weights, None, False)],
                       w1=self.attention[1].weight, b1=None, w2=self.attention[3].weight, b2=self.attention[3].bias,
                       temperature=float(attn_temp), bn=self.bn, act="silu", cin=self.in_channels,
                       cout=self.out_channels, k=self.kernel_size, stride=self.stride, pad=self.padding,
                       s2d=False, stem=self.in_channels < 32)

    def forward_nhwc(self, x, attn_temp, train: Optional[bool] = None, tape: Optional[list] = None,
                     ex: Optional[Executor] = None) -> torch.Tensor:
        """x: NHWC bf16, or the NCHW fp32 network input when in_channels == 3."""
        train = self.training if train is None else train
        own = ex is None
        ex = self._exec if own else ex
        y = ex.dyn_forward(self.dyn_spec(attn_temp), x, train, tape)
        if own:
            ex.end_forward()
        return y

    def forward(self, x, attn_temp):
        stem = self.in_channels < 32
        return to_nchw(self.forward_nhwc(x.float().contiguous() if stem else to_nhwc(x, self.in_channels), attn_temp))


class ObjectnessHead(LightningModule):
    """1x1 conv -> (B, A, H, W, 1) logits (reference _base.py:80-99); parameter container."""

    def __init__(self, in_channels, n_anchors):
        super().__init__()
        self.n_anchors = n_anchors
        self.conv_obj = nn.Conv2d(in_channels, n_anchors, kernel_size=(1, 1), stride=(1, 1))
        self.sigmoid = nn.Sigmoid()


class BBoxHead(LightningModule):
    """1x1 conv -> (B, A, H, W, 4) logits (reference _base.py:102-120); parameter container."""

    def __init__(self, in_channels, n_anchors):
        super().__init__()
        self.n_anchors = n_anchors
        self.conv_bbox = nn.Conv2d(in_channels, n_anchors * 4, kernel_size=(1, 1), stride=(1, 1))
        self.sigmoid = nn.Sigmoid()


class YOLOHead(LightningModule):
    """Detection head + loss (reference _base.py:122-270).  The objectness and bbox 1x1 convs of a
    scale are fused into one N=16 implicit GEMM whose epilogue writes both outputs in their final
    permuted layout, so each feature map is read once."""

    def __init__(self, x_channels: List[int], anchors, head_scales, loss_balancing, bbox_loss_fn="mse"):
        super().__init__()
        self.anchors = torch.tensor(anchors).float()
        self.head_scales = torch.tensor(head_scales)
        self.detection_head = nn.ModuleList()
        n_anchors = len(anchors[0])
        self.n_anchors = n_anchors
        self.obj_scales_w = loss_balancing.obj_scales_w
        self.bbox_w = loss_balancing.bbox_w
        self.objectness_w = loss_balancing.objectness_w
        self.no_obj_w = loss_balancing.no_obj_w
        self.bbox_loss_fn = bbox_loss_fn
        self.mutate_targets = True  # the reference rewrites batch.bbox in place (_base.py:257,266)
        self.fused_loss = True      # CUDA tensors: loss + gradient by the fused kernel (False: batched torch math)
        for c in x_channels:
            self.detection_head.append(nn.ModuleDict(dict(obj=ObjectnessHead(c, n_anchors),
                                                          bbox=BBoxHead(c, n_anchors))))
        self._packs = {}
        self._sa_cache = {}

    # ---- fused head conv -------------------------------------------------------------------------
    def fused_weight(self, s: int):
        """[A obj rows | 4A bbox rows] packed to 16 x cin bf16 (+ fp32 bias), cached per version."""
        wo = self.detection_head[s]["obj"].conv_obj
        wb = self.detection_head[s]["bbox"].conv_bbox
        ver = (wo.weight._version, wb.weight._version, wo.bias._version, wb.bias._version, wo.weight.device,
               param_epoch(), wo.weight.data_ptr())
        hit = self._packs.get(s)
        if hit is None or hit[0] != ver:
            w = torch.cat([wo.weight.detach(), wb.weight.detach()], dim=0)
            b = torch.cat([wo.bias.detach(), wb.bias.detach()], dim=0).contiguous()
            hit = (ver, ops.pack_weight(w.contiguous(), rows=16), b)
            self._packs[s] = hit
        return hit[1], hit[2]

    def forward_nhwc(self, f_maps: Sequence[torch.Tensor]) -> List[DetectionResults]:
        outs = []
        for s, f in enumerate(f_maps):
            w16, b15 = self.fused_weight(s)
            obj, bbox = ops.conv_head(f, w16, b15, self.n_anchors)
            outs.append(DetectionResults(obj=obj, bbox=bbox))
        return outs

    def forward(self, f_maps: List[torch.Tensor]):
        return self.forward_nhwc([to_nhwc(f) for f in f_maps])

    def backward_nhwc(self, s: int, feat: torch.Tensor, d_bbox: Optional[torch.Tensor], d_obj: Optional[torch.Tensor],
                      hook=None) -> torch.Tensor:
        """Backward of the fused 1x1 head conv of scale `s`: returns dL/d(feature map) (NHWC bf16) and
        accumulates the weight/bias gradients of conv_obj / conv_bbox."""
        a = self.n_anchors
        n, hh, ww, cin = feat.shape
        dev = feat.device
        conv_o = self.detection_head[s]["obj"].conv_obj
        conv_b = self.detection_head[s]["bbox"].conv_bbox
        for conv in (conv_o, conv_b):
            if conv.bias.grad is None:
                conv.bias.grad = torch.zeros_like(conv.bias)
        # (B,A,H,W,1|4) fp32 -> NHWC bf16 with 32 channels [A obj | 4A bbox | zero pad] + both bias gradients: one pass
        dyh = ops.head_grad_pack(d_obj.contiguous() if d_obj is not None else None,
                                 d_bbox.contiguous() if d_bbox is not None else None, n, a, hh, ww,
                                 conv_o.bias.grad, conv_b.bias.grad)
        dwp = ops.conv_wgrad(feat, dyh, 1, 1, 0)              # packed [32][cin]
        for conv, lo, hi in ((conv_o, 0, a), (conv_b, a, 5 * a)):
            g = dwp[lo:hi].view(hi - lo, cin, 1, 1)
            if conv.weight.grad is None:
                conv.weight.grad = g.clone()
            else:
                conv.weight.grad.add_(g)
        w = torch.zeros((32, cin, 1, 1), dtype=torch.float32, device=dev)
        w[:a] = conv_o.weight.detach()
        w[a:5 * a] = conv_b.weight.detach()
        wt = ops.pack_weight(w, transposed=True)              # W^T packed [cin][32]
        dx = ops.conv_dgrad(dyh, wt, cin, 1, 1, 0, (hh, ww))
        # hooks last: a data-parallel trainer may update a bucket right behind its all-reduce, and the fp32 weights
        # were still being read (packed) above
        if hook is not None:
            for conv in (conv_o, conv_b):
                hook(conv.weight)
                hook(conv.bias)
        return dx

    def _scaled_anchors(self, h: int, device) -> torch.Tensor:
        """anchors[h] / head_scales[h] (reference _base.py:170) as a cached device tensor, so the loss does no
        host->device copy per step (and can be captured into a CUDA graph)."""
        key = (h, str(device))
        hit = self._sa_cache.get(key)
        if hit is None:
            hit = (self.anchors[h] / self.head_scales[h]).to(device)
            self._sa_cache[key] = hit
        return hit

    # ---- loss ------------------------------------------------------------------------------------
    def compute_metrics(self, outs: List[DetectionResults], batch: BatchData, return_ap=False):
        """-> (total_loss, ap|None, bbox_loss, obj_loss), reference _base.py:155-212, evaluated as
        batched tensor math per head (no per-sample loop, no device sync)."""
        bsz = len(batch.image)
        targets = batch.bbox
        per_sample = isinstance(targets, (list, tuple)) and isinstance(targets[0], (list, tuple))
        bbox_losses = 0.0
        obj_losses = 0.0
        weights = (self.bbox_w, self.objectness_w, self.no_obj_w)
        for h, out in enumerate(outs):
            if per_sample:
                tgt = torch.stack([targets[i][h] for i in range(bsz)]).to(out.bbox.device)
            else:
                tgt = targets[h].to(out.bbox.device)
            if out.bbox.is_cuda and self.fused_loss:
                # value + gradient of the head scale in three launches (csrc/loss.cu)
                bl, ol, new_t = yolo_head_loss_fused(out.bbox.float(), out.obj.float(), tgt.float(),
                                                     self.anchors[h] / self.head_scales[h], self.obj_scales_w[h],
                                                     weights, self.bbox_loss_fn,
                                                     want_new_t=self.mutate_targets or return_ap)
            else:
                sa = self._scaled_anchors(h, out.bbox.device)
                bl, ol, new_t = yolo_head_loss(out.bbox.float(), out.obj.float(), tgt, sa, self.obj_scales_w[h],
                                               weights, self.bbox_loss_fn)
            bbox_losses = bbox_losses + bl
            obj_losses = obj_losses + ol
            if self.mutate_targets:
                if per_sample:
                    dst = [targets[i][h][..., 1:] for i in range(bsz)]
                    torch._foreach_copy_(dst, list(new_t.detach().unbind(0)))
                else:
                    targets[h][..., 1:] = new_t.detach()
        bbox_losses = bbox_losses / bsz
        obj_losses = obj_losses / bsz
        ap = self._average_precision(outs, new_t, bsz) if return_ap else None
        return bbox_losses + obj_losses, ap, bbox_losses, obj_losses

    def _average_precision(self, outs: List[DetectionResults], last_head_targets: Optional[torch.Tensor], bsz: int):
        """The `return_ap` branch (reference _base.py:194-204): per image, decode every head, flatten `(a h w)`,
        cxcywh->xyxy, concatenate the heads, `nms(boxes, logits, 0.5)` and hand the kept boxes to `calculate_ap`.
        Decode and NMS run batched on the CUDA kernels (3 decode launches + 1 NMS launch for the whole batch); the kept
        detections stay in `self.last_detections`.  mAP itself is torchmetrics' CPU evaluation (out of scope): it is
        called when torchmetrics is importable — with the reference's arguments, including its quirk of passing the
        LAST head's rebuilt target tensor — otherwise the per-image kept detections are returned in the `ap` slot."""
        from .. import inference
        with torch.no_grad():
            det = inference.postprocess([DetectionResults(bbox=o.bbox.detach(), obj=o.obj.detach()) for o in outs],
                                        self.anchors.tolist(), self.head_scales.tolist(), self.bbox_loss_fn, 0.5)
        self.last_detections = det
        kept = inference.kept_lists(det)
        try:
            from ..utils.metrics import calculate_ap
            import torchmetrics  # noqa: F401
        except Exception:
            return [dict(boxes=det.boxes[i][k], scores=det.scores[i][k], keep=k) for i, k in enumerate(kept)]
        total = torch.zeros((), device=det.boxes.device)
        for i, k in enumerate(kept):
            total = total + calculate_ap(det.boxes[i][k], det.scores[i][k], last_head_targets[i])["map"].to(total.device)
        return total / bsz


class BaseModel(LightningModule):
    """Optimiser / step plumbing shared by the models (reference _base.py:273-326)."""

    def __init__(self, hparams):
        super().__init__()
        self.learning_rate = hparams.lr
        self.optimizer = getattr(hparams, "optim", None)   # dy-soem_fpn.yaml keeps `optim` outside hparams (D5)
        self.head_scales = hparams.head_scales
        self.lr_scheduler = hparams.lr_scheduler
        self.backbone = None
        self.neck = None
        self.head = None

    def forward(self, x):
        return x

    def configure_optimizers(self):
        if self.optimizer is None:
            raise ValueError("Invalid optimizer: hparams.optim is missing")
        if self.optimizer.name == "SGD":
            opt = torch.optim.SGD(self.parameters(), lr=self.learning_rate, momentum=self.optimizer.momentum)
        elif self.optimizer.name == "Adam":
            opt = torch.optim.Adam(self.parameters(), lr=self.learning_rate)
        else:
            raise ValueError(f"Invalid optimizer: {self.optimizer}")
        if self.lr_scheduler:
            sched = torch.optim.lr_scheduler.CyclicLR(opt, base_lr=self.learning_rate / 10, max_lr=self.learning_rate,
                                                      step_size_up=4000, mode="triangular2", cycle_momentum=False)
            return dict(optimizer=opt, lr_scheduler=sched)
        return opt

    def _log_losses(self, prefix, loss, bbox_loss, obj_loss, n, **kw):
        self.log(f"{prefix}_loss", loss, prog_bar=True, batch_size=n, **kw)
        self.log(f"{prefix}_bbox_loss", bbox_loss, prog_bar=True, batch_size=n, **kw)
        self.log(f"{prefix}_obj_loss", obj_loss, prog_bar=True, batch_size=n, **kw)
