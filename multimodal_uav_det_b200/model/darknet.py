"""Darknet-53 style trunk shared by BaselineModel and DyYOLO (reference BaselineModel.py:56-124,
DyYOLO.py:56-144).  The `layer_config` mini-DSL is compiled once into a flat program of conv
units; forward and backward walk that program over NHWC bf16 activations with an explicit tape
(one autograd node for the whole trunk instead of ~400 ATen nodes).

Fusions relative to the reference graph:
  * conv + BN(+batch stats) + LeakyReLU (+ residual add)      -> conv kernel epilogue + 1 pass
  * `layer(x) + use_residual * x` (BaselineModel.py:43)        -> residual read in that pass
  * nn.Upsample + torch.cat (BaselineModel.py:120-122)         -> one write into the concat buffer
  * skip-path gradient accumulation                            -> dgrad epilogue
  * obj + bbox head convs and their permutes (_base.py:88-120) -> one GEMM, final layout
"""
from __future__ import annotations

from typing import List, Optional

import torch
from torch import nn

from .. import ops
from ..engine import ConvUnit, DynRecord, Executor, bump_param_epoch
from ..utils.datatype import BatchData, DetectionResults
from ._base import BaseModel, DyConvModule, LightningModule, YOLOHead, to_nhwc, to_nchw


class CNNBlock(LightningModule):
    """conv(bias = not bn_act) -> BN -> LeakyReLU(0.1) (reference BaselineModel.py:10-22)."""

    def __init__(self, in_channels, out_channels, bn_act=True, **kwargs):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, bias=not bn_act, **kwargs)
        self.bn = nn.BatchNorm2d(out_channels)
        self.leaky = nn.LeakyReLU(0.1)
        self.use_bn_act = bn_act
        self._exec = Executor()

    def unit(self) -> ConvUnit:
        stem = self.conv.in_channels < 32
        if self.use_bn_act:
            return ConvUnit(self.conv, self.bn, "leaky", stem=stem)
        return ConvUnit(self.conv, None, "none", stem=stem)

    def forward(self, x):
        u = self.unit()
        y = self._exec.conv_forward(u, x.float().contiguous() if u.stem else to_nhwc(x, self.conv.in_channels),
                                    self.training, None)
        self._exec.end_forward()
        return to_nchw(y)


class ResidualBlock(nn.Module):
    """`num_repeats` x [1x1 C->C/2, 3x3 C/2->C] with optional skip (reference BaselineModel.py:25-45)."""

    def __init__(self, channels, use_residual=True, num_repeats=1):
        super().__init__()
        self.layers = nn.ModuleList()
        for _ in range(num_repeats):
            self.layers += [nn.Sequential(CNNBlock(channels, channels // 2, kernel_size=1),
                                          CNNBlock(channels // 2, channels, kernel_size=3, padding=1))]
        self.use_residual = use_residual
        self.num_repeats = num_repeats

    def forward(self, x):
        ex = Executor()
        h = to_nhwc(x)
        for pair in self.layers:
            y = ex.conv_forward(pair[0].unit(), h, self.training, None)
            h = ex.conv_forward(pair[1].unit(), y, self.training, None, res=h if self.use_residual else None)
        ex.end_forward()
        return to_nchw(h)


class ScalePrediction(nn.Module):
    """3x3 C->2C block feeding one detection scale (reference BaselineModel.py:47-53)."""

    def __init__(self, in_channels):
        super().__init__()
        self.conv = CNNBlock(in_channels, 2 * in_channels, kernel_size=3, padding=1)

    def forward(self, x):
        return self.conv(x)


class _TrunkFn(torch.autograd.Function):
    """One autograd node for the whole detector trunk.  Parameter gradients are accumulated into
    `.grad` by the executor (like autograd's AccumulateGrad would); the only differentiable inputs
    seen by autograd are the anchor tensor (dummy) and the image."""

    @staticmethod
    def forward(ctx, model, x, anchor):
        tape: list = []
        # grad mode is off inside Function.forward; the caller only routes here when a tape is wanted
        outs = model._forward_program(x, tape)
        ctx.model = model
        ctx.tape = tape
        ctx.n_out = len(outs)
        flat = []
        for o in outs:
            flat += [o.bbox, o.obj]
        return tuple(flat)

    @staticmethod
    def backward(ctx, *grads):
        model = ctx.model
        pairs = [(grads[2 * i], grads[2 * i + 1]) for i in range(ctx.n_out)]
        model._backward_program(ctx.tape, pairs)
        ctx.tape = None
        return None, None, None


class DarknetDetector(BaseModel):
    """Compiles `hparams.layer_config` ([C,k,s] | ["B",n] | ["S"] | ["U"] | ["DyConv",C,k,s]) into
    `self.layers` (same module list / state_dict keys as the reference) and a flat program."""

    supports_dyconv = False
    route_repeats = 8   # a ResidualBlock with this many repeats feeds the next route concat (BaselineModel.py:116)

    def __init__(self, hparams):
        super().__init__(hparams)
        self.layers = nn.ModuleList()
        self.attn_temp = getattr(hparams, "attn_temperature", None)
        x_out_channels = []
        in_channels = 3
        for module in hparams.layer_config:
            kind = module[0]
            if kind == "B":
                self.layers.append(ResidualBlock(in_channels, num_repeats=module[1]))
            elif kind == "S":
                self.layers += [ResidualBlock(in_channels, use_residual=False, num_repeats=1),
                                CNNBlock(in_channels, in_channels // 2, kernel_size=1),
                                ScalePrediction(in_channels // 2)]
                x_out_channels.append(in_channels)
                in_channels = in_channels // 2
            elif kind == "U":
                self.layers.append(nn.Upsample(scale_factor=2))
                in_channels = in_channels * 3
            elif kind == "DyConv":
                if not self.supports_dyconv:
                    raise ValueError("DyConv layers need DyYOLO")
                out_channels, kernel_size, stride = module[1:]
                self.layers.append(DyConvModule(in_channels, out_channels, kernel_size=kernel_size, stride=stride,
                                                padding=1 if kernel_size == 3 else 0))
                in_channels = out_channels
            else:
                out_channels, kernel_size, stride = module
                self.layers.append(CNNBlock(in_channels, out_channels, kernel_size=kernel_size, stride=stride,
                                            padding=1 if kernel_size == 3 else 0))
                in_channels = out_channels
        self.yolo_head = YOLOHead(x_out_channels, hparams.anchors, hparams.head_scales, hparams.loss_balancing,
                                  hparams.bbox_loss_fn)
        self._exec = Executor()
        self.training_graph = True       # record the tape when grad is enabled
        self._anchor = None
        self._debug_taps = None          # tests: dict receiving every intermediate activation
        self._igemm_w = None

    # ---- public API ------------------------------------------------------------------------------
    def forward(self, x) -> List[DetectionResults]:
        """x: (B,3,H,W) float in [0,1] (NCHW, like the reference) -> per scale DetectionResults with
        bbox (B,A,S,S,4) and obj (B,A,S,S,1) fp32 logits."""
        if not x.is_cuda:
            raise RuntimeError("multimodal_uav_det_b200 models run on CUDA only (no CPU fallback)")
        x = x.float().contiguous()
        if torch.is_grad_enabled() and self.training_graph and any(p.requires_grad for p in self.parameters()):
            # a fresh (uninitialised, never read) leaf per call: its AccumulateGrad node then belongs to the
            # stream of THIS forward, which CUDA-graph capture of the step needs
            anchor = torch.empty((), device=x.device).requires_grad_()
            flat = _TrunkFn.apply(self, x, anchor)
            return [DetectionResults(bbox=flat[2 * i], obj=flat[2 * i + 1]) for i in range(len(flat) // 2)]
        return self._forward_program(x, None)

    def training_step(self, batch: BatchData, batch_idx):
        outs = self.forward(batch.image)
        loss, _, bbox_loss, obj_loss = self.yolo_head.compute_metrics(outs, batch)
        self._log_losses("train", loss, bbox_loss, obj_loss, len(batch))
        return loss

    def validation_step(self, batch: BatchData, batch_idx):
        outs = self.forward(batch.image)
        loss, _, bbox_loss, obj_loss = self.yolo_head.compute_metrics(outs, batch, return_ap=False)
        self._log_losses("val", loss, bbox_loss, obj_loss, len(batch), on_epoch=True)
        return loss

    # ---- program ---------------------------------------------------------------------------------
    def _igemm_weights(self):
        """Conv weights the implicit-GEMM kernels read (everything but the cin=3 stem): re-packed to bf16 in one
        launch per step."""
        if self._igemm_w is None:
            self._igemm_w = [m.conv.weight for m in self.modules()
                             if isinstance(m, CNNBlock) and m.conv.in_channels % 32 == 0]
        return self._igemm_w

    def prepare_for_capture(self):
        """Build the host-side tables a CUDA-graph capture must not create (they need host->device copies)."""
        self._exec.begin_step(self._igemm_weights()[0].device)
        self._exec.packs.prepack(self._igemm_weights(), with_transposed=True)
        bump_param_epoch()      # ... but the captured step must re-pack: the weights change at every replay

    def _forward_program(self, x, tape: Optional[list]) -> List[DetectionResults]:
        ex = self._exec
        train = self.training
        ex.begin_step(x.device)
        ex.packs.prepack(self._igemm_weights(), with_transposed=tape is not None)
        feats = []
        routes = []            # (tensor, index of the tape entry that produced it)
        h = x                   # NCHW fp32 until the first (stem) conv
        for li, layer in enumerate(self.layers):
            if isinstance(layer, ScalePrediction):
                if tape is not None:
                    tape.append(("scale_begin",))
                feats.append(ex.conv_forward(layer.conv.unit(), h, train, tape))
                if tape is not None:
                    tape.append(("scale_end", len(feats) - 1))
            elif isinstance(layer, CNNBlock):
                h = ex.conv_forward(layer.unit(), h, train, tape)
            elif isinstance(layer, ResidualBlock):
                for pair in layer.layers:
                    y = ex.conv_forward(pair[0].unit(), h, train, tape)
                    h = ex.conv_forward(pair[1].unit(), y, train, tape, res=h if layer.use_residual else None)
                if layer.num_repeats == self.route_repeats:
                    routes.append(h)
                    if tape is not None:
                        tape.append(("route_push",))
            elif isinstance(layer, nn.Upsample):
                route = routes.pop()
                n, hh, ww, c = h.shape
                cat = ops.empty_act(n, 2 * hh, 2 * ww, c + route.shape[3], h.device)
                ops.upsample2x_fwd(h, out=cat[..., :c])
                ops.add(route, None, out=cat[..., c:])
                h = cat
                if tape is not None:
                    tape.append(("concat", c))
            elif isinstance(layer, DyConvModule):
                h = layer.forward_nhwc(h, self.attn_temp, train, tape, ex)
            else:
                raise TypeError(f"unsupported layer {type(layer)}")
            if self._debug_taps is not None and not isinstance(layer, ScalePrediction) and li > 0:
                self._debug_taps[f"after_{li}"] = to_nchw(h).cpu()
        ex.end_forward()
        return self.yolo_head.forward_nhwc(feats) if tape is None else self._head_forward(feats, tape)

    def _head_forward(self, feats, tape):
        tape.append(("head", feats))
        return self.yolo_head.forward_nhwc(feats)

    def _backward_program(self, tape: list, head_grads) -> None:
        """Reverse walk.  `dy` is the gradient w.r.t. the running activation `h`; branches (scale
        predictions, route concat) hand their contribution to the consumer's dgrad epilogue."""
        ex = self._exec
        head = self.yolo_head
        feats = tape[-1][1]
        d_feats = [self._head_backward(s, feats[s], *head_grads[s]) for s in range(len(feats))]
        dy: Optional[torch.Tensor] = None
        route_grads: List[torch.Tensor] = []
        pending_scale: Optional[int] = None
        i = len(tape) - 2
        while i >= 0:
            rec = tape[i]
            if isinstance(rec, tuple):
                tag = rec[0]
                if tag == "scale_end":
                    pending_scale = rec[1]
                elif tag == "scale_begin":
                    pass
                elif tag == "concat":
                    c = rec[1]
                    route_grads.append(dy[..., c:])
                    dy = ops.upsample2x_bwd(dy[..., :c])
                elif tag == "route_push":
                    # the route tensor also feeds the concat: add that gradient to dy
                    rg = route_grads.pop()
                    dy = ops.add(dy, rg) if dy is not None else ops.add(rg, None)
                i -= 1
                continue
            if isinstance(rec, DynRecord):
                dy = ex.dyn_backward(rec, dy, need_dx=(i > 0))
                i -= 1
                continue
            if pending_scale is not None:
                # scale-prediction conv: its input is the running activation; fuse `+ dy` (gradient
                # from the layers after the branch) into the dgrad epilogue
                dy = ex.conv_backward(rec, d_feats[pending_scale], res=dy)
                pending_scale = None
            else:
                res = None
                if rec.has_res:
                    # y = f2(f1(x)) + x : the skip gradient (= dy) is added by f1's dgrad epilogue
                    d_mid = ex.conv_backward(rec, dy)
                    i -= 1
                    rec1 = tape[i]
                    dy = ex.conv_backward(rec1, d_mid, res=dy)
                    i -= 1
                    continue
                dy = ex.conv_backward(rec, dy, res=res, need_dx=(i > 0))
            i -= 1
        ex.end_backward()

    def _head_backward(self, s: int, feat: torch.Tensor, d_bbox: Optional[torch.Tensor],
                       d_obj: Optional[torch.Tensor]) -> torch.Tensor:
        return self.yolo_head.backward_nhwc(s, feat, d_bbox, d_obj, self._exec.grad_ready_hook)
