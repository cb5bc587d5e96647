"""RTMUAVDet — light two-scale detector with per-sample depthwise dynamic convolutions
(reference model/RTMUAVDet.py:15-418; `@deprecated("INVALID MODEL CONFIGURATION")` there, not
exported, and un-importable as shipped — SURVEY.md D4 — but its forward is well defined and is what
this file reproduces, for inference: BatchNorm uses running statistics, Dropout(0.2) is identity).

Kernel mapping: every dense conv is the tcgen05 implicit GEMM with BN/bias + SiLU/ReLU/GELU fused in
the epilogue; channel concats are never materialised (producers write channel slices); the per-sample
depthwise dynamic conv + residual, GroupNorm(1 group) (+ its residual add), bilinear x2 and the
sigmoid + box decode of the head are single memory-bound kernels (csrc/rtm.cu).
Reference quirks kept: the 5x5 stride-2 pad-1 stem yields 319x319 (:31) — stored as 320x320 with a
zero last row/column, which is exactly the zero padding the next stride-2 conv would read; head
attribute names are swapped (obj head owns `conv_bbox`, :223,243); anchors are not stride-scaled
(:288-289); the second residual of MDyEncoder is commented out (:181-182).
"""
from __future__ import annotations

import os
from typing import List

import torch
from torch import nn

from .. import ops
from ..engine import ConvUnit, Executor
from ..utils.datatype import DetectionResults
from . import _base
from ._base import LightningModule, to_nchw, to_nhwc


_NO_GN_FOLD = bool(os.environ.get("UAVDET_RTM_NO_GN_FOLD"))     # A/B switch: GroupNorm as two streaming passes
_NO_PAIR_CONV = bool(os.environ.get("UAVDET_RTM_NO_PAIR_CONV"))  # A/B switch: the 256 -> 64 neck conv as a plain N = 64 GEMM


class ConvModule(_base.ConvModule):
    """conv -> BN(eps 1e-3, momentum 0.03) -> SiLU|ReLU (reference RTMUAVDet.py:15-25)."""

    def __init__(self, in_channels, out_channels, kernel_size=(1, 1), stride=(1, 1), padding=0, bias=False, eps=1e-3,
                 momentum=0.03, activation="silu"):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, bias, activation, eps, momentum)


class StemLayer(LightningModule):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = ConvModule(in_channels, out_channels, kernel_size=(5, 5), stride=(2, 2), padding=(1, 1), bias=False)


class MDyConv(LightningModule):
    """1x1 ConvModule(ReLU) -> [GAP -> fc+ReLU -> channel_fc x kernel_fc] -> per-sample depthwise
    conv -> + residual (reference :40-100)."""

    def __init__(self, in_channels, attention_out_c, dy_kernel_size=3, dy_padding=1, dy_channel_size=None):
        super().__init__()
        self.dy_channel_size = dy_channel_size if dy_channel_size else in_channels
        self.dy_kernel_size = dy_kernel_size
        self.dy_padding = dy_padding
        self.base_conv = ConvModule(in_channels, self.dy_channel_size, kernel_size=(1, 1), eps=1e-5, momentum=0.1,
                                    activation="relu")
        self.attention = nn.Sequential(nn.AdaptiveAvgPool2d((1, 1)),
                                       nn.Conv2d(self.dy_channel_size, attention_out_c, kernel_size=(1, 1)),
                                       nn.ReLU(inplace=True))
        self.channel_fc = nn.Conv2d(attention_out_c, self.dy_channel_size, kernel_size=(1, 1))
        self.kernel_fc = nn.Conv2d(attention_out_c, int(self.dy_kernel_size ** 2), kernel_size=(1, 1))

    def forward_nhwc(self, x, ex: Executor, out=None):
        return self.dynamic_nhwc(ex.conv_forward(self.base_conv.unit(), x, False, None), out=out)

    def dynamic_nhwc(self, y, res=None, stats=None, out=None):
        """Everything behind base_conv: attention on the pooled map -> per-sample depthwise kernel -> conv + y
        (+ `res` and the per-sample statistics of the result, for the GroupNorm the MDyEncoder applies next)."""
        pooled = ops.gap(y)
        att = self.attention[1]
        a = ops.linear(pooled, att.weight.detach().flatten(1), att.bias.detach(), "relu")
        ch_w = ops.linear(a, self.channel_fc.weight.detach().flatten(1), self.channel_fc.bias.detach())
        k_w = ops.linear(a, self.kernel_fc.weight.detach().flatten(1), self.kernel_fc.bias.detach())
        if res is not None:
            return ops.dwdynconv_res_stats_fwd(y, ch_w, k_w, self.dy_kernel_size, self.dy_padding, res, stats, out=out)
        return ops.dwdynconv_fwd(y, ch_w, k_w, self.dy_kernel_size, self.dy_padding, out=out)

    def forward(self, x):
        return to_nchw(self.forward_nhwc(to_nhwc(x), Executor()))


class MDyCSPModule(LightningModule):
    """Stride-2 3x3 -> two 1x1 branches, one through MDyConv -> concat -> 3x3 (reference :103-140)."""

    def __init__(self, in_channels, out_channels, reduction_ratio=2, dy_channel_size=None):
        super().__init__()
        base_out_c = in_channels * 2
        self.base_conv = ConvModule(in_channels, base_out_c, kernel_size=(3, 3), stride=(2, 2), padding=(1, 1))
        self.conv1 = ConvModule(base_out_c, base_out_c // reduction_ratio, kernel_size=(1, 1))
        self.conv2 = ConvModule(base_out_c, base_out_c // reduction_ratio, kernel_size=(1, 1))
        if dy_channel_size:
            self.mdy_conv = MDyConv(base_out_c // reduction_ratio, 16, dy_kernel_size=3, dy_channel_size=dy_channel_size)
        else:
            self.mdy_conv = MDyConv(base_out_c // reduction_ratio, 16, dy_kernel_size=3)
        transition_c = base_out_c // reduction_ratio
        self.transition1 = ConvModule(128, transition_c, kernel_size=(1, 1))      # hard-coded 128 in the reference (:119)
        self.transition2 = ConvModule(base_out_c, out_channels, kernel_size=(3, 3), padding=(1, 1))
        self.half_c = transition_c
        self.base_out_c = base_out_c

    def forward_nhwc(self, x, ex: Executor, out=None):
        x = ex.conv_forward(self.base_conv.unit(), x, False, None)
        n, h, w, _ = x.shape
        cat = ops.empty_act(n, h, w, self.base_out_c, x.device)
        x1 = ex.conv_forward(self.conv1.unit(), x, False, None)
        ex.conv_forward(self.conv2.unit(), x, False, None, out=cat[..., self.half_c:])
        x1 = self.mdy_conv.forward_nhwc(x1, ex)
        ex.conv_forward(self.transition1.unit(), x1, False, None, out=cat[..., :self.half_c])
        return ex.conv_forward(self.transition2.unit(), cat, False, None, out=out)


class MDyEncoder(LightningModule):
    """GN -> three MDyConvs (1x1 / 3x3 / 5x5 dynamic kernels) -> concat + residual -> GN -> channel MLP
    (reference :144-184)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.group_norm_in = nn.GroupNorm(num_groups=1, num_channels=in_channels, eps=1e-5, affine=True)
        self.mdy_conv_1x1 = MDyConv(in_channels, 16, dy_kernel_size=1, dy_padding=0, dy_channel_size=in_channels // 3)
        self.mdy_conv_3x3 = MDyConv(in_channels, 16, dy_kernel_size=3, dy_padding=1, dy_channel_size=in_channels // 3)
        self.mdy_conv_5x5 = MDyConv(in_channels, 16, dy_kernel_size=5, dy_padding=2, dy_channel_size=in_channels // 3)
        self.group_norm_out = nn.GroupNorm(num_groups=1, num_channels=in_channels, eps=1e-5, affine=True)
        self.channel_mlp = nn.Sequential(nn.Conv2d(in_channels, in_channels, kernel_size=(1, 1)), nn.GELU(), nn.Dropout(0.2),
                                         nn.Conv2d(in_channels, out_channels, kernel_size=(1, 1)))
        self.third = in_channels // 3
        self._fold = None

    def _folded(self, ex: Executor):
        """Constants of the GroupNorm folds, rebuilt when a parameter / running statistic they depend on changes.

        GroupNorm(1 group) is one affine map per sample, z = (v - mean_n) * rstd_n * gamma + beta, and both of its consumers
        are 1x1 convolutions (the three MDyConv.base_conv of :165-169 read GN_in(x); channel_mlp[0] of :174-177 reads
        GN_out(cat + x)), so the map moves into their epilogues: W' = diag(a) * W * diag(gamma) in bf16 (a = the folded
        eval-mode BatchNorm scale behind the convolution, or 1) and y = act(acc * rstd_n + b - mean_n * rstd_n *
        rowsum(W')) with b = a * (W beta) + BatchNorm shift / bias (`sample_affine` epilogue).  The three base
        convolutions also become ONE 192 -> 192 (384 -> 384) GEMM: their input is read once instead of three times, and
        the normalised tensor is never written."""
        from ..engine import param_epoch
        gi, go = self.group_norm_in, self.group_norm_out
        mdys = (self.mdy_conv_1x1, self.mdy_conv_3x3, self.mdy_conv_5x5)
        bns = [m.base_conv.conv[1] for m in mdys]
        deps = [gi.weight, gi.bias, go.weight, go.bias, self.channel_mlp[0].weight, self.channel_mlp[0].bias]
        for m, bn in zip(mdys, bns):
            deps += [m.base_conv.conv[0].weight, bn.weight, bn.bias, bn.running_mean, bn.running_var]
        ver = (param_epoch(),) + tuple((t._version, t.data_ptr()) for t in deps)
        if self._fold is not None and self._fold[0] == ver:
            return self._fold[1]
        with torch.no_grad():
            w_in = torch.cat([m.base_conv.conv[0].weight.detach().flatten(1) for m in mdys])            # (3t, C)
            folds = [ex.bn_fold(bn, m.base_conv.conv[0].bias) for m, bn in zip(mdys, bns)]
            a_in = torch.cat([f[0] for f in folds])
            wp_in = (a_in[:, None] * w_in * gi.weight.detach()[None, :]).to(torch.bfloat16).contiguous()
            b_in = (a_in * (w_in @ gi.bias.detach()) + torch.cat([f[1] for f in folds])).contiguous()
            w_out = self.channel_mlp[0].weight.detach().flatten(1)                                        # (C, C)
            wp_out = (w_out * go.weight.detach()[None, :]).to(torch.bfloat16).contiguous()
            b_out = (w_out @ go.bias.detach() + self.channel_mlp[0].bias.detach()).contiguous()
            # row sums of the ROUNDED weights: the mean term then cancels what the tensor core accumulated
            consts = dict(wp_in=wp_in, wg_in=wp_in.float().sum(1).contiguous(), b_in=b_in,
                          wp_out=wp_out, wg_out=wp_out.float().sum(1).contiguous(), b_out=b_out)
        self._fold = (ver, consts)
        return consts

    def forward_nhwc(self, x, ex: Executor, out=None):
        if _NO_GN_FOLD:
            return self._forward_nhwc_unfused(x, ex, out)
        gi, go = self.group_norm_in, self.group_norm_out
        k = self._folded(ex)
        n, h, w, c = x.shape
        t = self.third
        # GN_in + the three base convolutions: one GEMM on the raw input with a per-sample epilogue
        sa = ops.groupnorm1_fold(ops.groupnorm1_stats(x), h * w * c, gi.eps)
        base = ops.conv_fwd(x, k["wp_in"], 3 * t, 1, 1, 0, act="relu", scale=k["wg_in"], shift=k["b_in"], sample_affine=sa)
        # the dynamic depthwise convolutions write cat + x and its per-sample statistics
        cat = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
        stats = torch.zeros((n, 2), dtype=torch.float32, device=x.device)
        for i, m in enumerate((self.mdy_conv_1x1, self.mdy_conv_3x3, self.mdy_conv_5x5)):
            sl = slice(i * t, (i + 1) * t)
            m.dynamic_nhwc(base[..., sl], res=x[..., sl], stats=stats, out=cat[..., sl])
        # GN_out + channel_mlp[0] + GELU
        sa = ops.groupnorm1_fold(stats, h * w * c, go.eps)
        z = ops.conv_fwd(cat, k["wp_out"], c, 1, 1, 0, act="gelu", scale=k["wg_out"], shift=k["b_out"], sample_affine=sa)
        return ex.conv_forward(ConvUnit(self.channel_mlp[3], None, "none"), z, False, None, out=out)

    def _forward_nhwc_unfused(self, x, ex: Executor, out=None):
        """The operator-by-operator form (UAVDET_RTM_NO_GN_FOLD=1): GroupNorm as a statistics and a normalise pass."""
        gi, go = self.group_norm_in, self.group_norm_out
        y = ops.groupnorm1(x, gi.weight.detach(), gi.bias.detach(), gi.eps)
        cat = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
        t = self.third
        self.mdy_conv_1x1.forward_nhwc(y, ex, out=cat[..., :t])
        self.mdy_conv_3x3.forward_nhwc(y, ex, out=cat[..., t:2 * t])
        self.mdy_conv_5x5.forward_nhwc(y, ex, out=cat[..., 2 * t:])
        z = ops.groupnorm1(cat, go.weight.detach(), go.bias.detach(), go.eps, b=x)      # GN(cat + residual)
        z = ex.conv_forward(ConvUnit(self.channel_mlp[0], None, "gelu"), z, False, None)
        return ex.conv_forward(ConvUnit(self.channel_mlp[3], None, "none"), z, False, None, out=out)

    def forward(self, x):
        return to_nchw(self.forward_nhwc(to_nhwc(x), Executor()))


class MFDFEncoderModule(LightningModule):
    """Two-scale feature fusion neck (reference :186-215)."""

    def __init__(self, x1_c_in, x2_c_in):
        super().__init__()
        self.upsample = nn.Sequential(nn.Upsample(scale_factor=2, mode="bilinear"),
                                      nn.Conv2d(x2_c_in, x2_c_in // 4, kernel_size=(3, 3), padding=(1, 1)))
        self.downsample = nn.Conv2d(x1_c_in, x1_c_in, kernel_size=(3, 3), stride=(2, 2), padding=(1, 1))
        self.encoder_x1 = MDyEncoder((x1_c_in // 2) * 3, x1_c_in)
        self.encoder_x2 = MDyEncoder((x2_c_in // 2) * 3, x2_c_in)


class ObjectnessHead(LightningModule):
    def __init__(self, in_channels, n_anchors):
        super().__init__()
        self.n_anchors = n_anchors
        self.conv_bbox = nn.Conv2d(in_channels, n_anchors, kernel_size=(1, 1), stride=(1, 1))   # sic (:223)
        self.sigmoid = nn.Sigmoid()


class BBoxHead(LightningModule):
    def __init__(self, in_channels, n_anchors):
        super().__init__()
        self.n_anchors = n_anchors
        self.conv_obj = nn.Conv2d(in_channels, n_anchors * 4, kernel_size=(1, 1), stride=(1, 1))  # sic (:243)
        self.sigmoid = nn.Sigmoid()


class RTMHead(LightningModule):
    """Sigmoid heads with the box decode inside forward (reference :258-310)."""

    def __init__(self, x_c_in: list, anchors, det_scales: list):
        super().__init__()
        self.det_scales = det_scales
        self.detection_head = nn.ModuleList()
        self.anchors = anchors
        self.n_anchors = len(anchors[0])
        for c in x_c_in:
            self.detection_head.append(nn.ModuleDict(dict(obj=ObjectnessHead(c, self.n_anchors),
                                                          bbox=BBoxHead(c, self.n_anchors))))

    def forward_nhwc(self, feats) -> List[DetectionResults]:
        outs = []
        for i, f in enumerate(feats):
            wo = self.detection_head[i]["obj"].conv_bbox
            wb = self.detection_head[i]["bbox"].conv_obj
            w16 = ops.pack_weight(torch.cat([wo.weight.detach(), wb.weight.detach()]).contiguous(), rows=16)
            b15 = torch.cat([wo.bias.detach(), wb.bias.detach()]).contiguous()
            obj_l, bbox_l = ops.conv_head(f, w16, b15, self.n_anchors)
            bbox, obj = ops.rtm_head_post(bbox_l, obj_l, torch.as_tensor(self.anchors[i]).float())
            outs.append(DetectionResults(obj=obj, bbox=bbox))
        return outs

    def forward(self, x1, x2):
        return self.forward_nhwc([to_nhwc(x1), to_nhwc(x2)])


class RTMUAVDet(LightningModule):
    """`RTMUAVDet(input_size, anchors, learning_rate, optimizer='Adam', det_scales=[160, 80])`;
    forward(x) -> [DetectionResults(bbox decoded cxcywh (B,3,160,160,4), obj sigmoid), (…80x80…)]."""

    def __init__(self, input_size, anchors, learning_rate, optimizer="Adam", det_scales=[160, 80]):
        super().__init__()
        self.learning_rate = learning_rate
        self.optimizer = optimizer
        self.input_size = input_size
        self.det_scales = det_scales
        self.backbone = nn.ModuleDict(dict(
            MDyCSP_1=nn.Sequential(StemLayer(input_size[0], 32),
                                   MDyCSPModule(in_channels=32, out_channels=128, dy_channel_size=128)),
            MDyCSP_2=MDyCSPModule(in_channels=128, out_channels=256)))
        self.neck = MFDFEncoderModule(x1_c_in=128, x2_c_in=256)
        self.head = RTMHead(x_c_in=[128, 256], anchors=anchors, det_scales=det_scales)
        self._exec = Executor()
        self.stem_on_tensor_cores = True      # False: the direct CUDA-core 5x5 kernel (fp32 input, no bf16 rounding of x)
        self._stem_w3 = None
        self._up_pair = None
        self._stem_w3_pair = None

    def _stem_s2d_weight(self, w: torch.Tensor) -> torch.Tensor:
        from ..engine import param_epoch
        ver = (w._version, param_epoch(), w.data_ptr(), w.device)
        if self._stem_w3 is None or self._stem_w3[0] != ver:
            self._stem_w3 = (ver, ops.pack_weight(ops.s2d_stem_weight(w.detach())))
        return self._stem_w3[1]

    def _stem_pair_weight(self, w3: torch.Tensor) -> torch.Tensor:
        if self._stem_w3_pair is None or self._stem_w3_pair[0] is not w3:
            self._stem_w3_pair = (w3, ops.pack_weight_pair(w3))
        return self._stem_w3_pair[1]

    @torch.no_grad()
    def forward(self, x) -> List[DetectionResults]:
        if not x.is_cuda:
            raise RuntimeError("multimodal_uav_det_b200 models run on CUDA only (no CPU fallback)")
        if self.training:
            raise NotImplementedError("RTMUAVDet is inference-only here (its training path is dead code in the "
                                      "reference, SURVEY D4); call .eval()")
        ex = self._exec
        x = x.float().contiguous()
        n = x.shape[0]
        stem = self.backbone["MDyCSP_1"][0].conv
        bn = stem.conv[1]
        scale = bn.weight.detach() * torch.rsqrt(bn.running_var + bn.eps)
        shift = bn.bias.detach() - bn.running_mean * scale
        if x.shape[1] == 3 and x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0 and self.stem_on_tensor_cores:
            # 5x5 stride-2 stem as space-to-depth + 3x3 implicit GEMM (12 live channels of 32): 3.0 ms -> ~1.4 ms at
            # batch 128.  The GEMM computes a 320th row / column the 5x5 stem does not have; they are zeroed, which is
            # what the zero-padded (even-sized) stem output holds there.
            w3 = self._stem_s2d_weight(stem.conv[0].weight)
            if _NO_PAIR_CONV or (x.shape[3] // 2) % 2:        # the pair GEMM needs an even width of the space-to-depth map
                h = ops.conv_fwd(ops.stem_s2d_pack(x), w3, 32, 3, 1, 1, act="silu", scale=scale, shift=shift)
            else:
                # 32 -> 32 channels: as a pixel-pair GEMM (N = 64, whole pixel pairs as K = 64 k-blocks) the tensor core
                # issues half the instructions at twice the rate of the 64-byte-row form
                h = ops.conv3x3_pair_fwd(ops.stem_s2d_pack(x), self._stem_pair_weight(w3), 32, act="silu", scale=scale, shift=shift)
            ho = (x.shape[2] + 2 - 5) // 2 + 1
            wo = (x.shape[3] + 2 - 5) // 2 + 1
            if ho < h.shape[1]:
                h[:, ho:, :, :] = 0
            if wo < h.shape[2]:
                h[:, :, wo:, :] = 0
        else:
            h = ops.stem_fwd(x, stem.conv[0].weight.detach(), 5, 2, 1, act="silu", scale=scale, shift=shift, pad_to_even=True)
        csp1, csp2, neck = self.backbone["MDyCSP_1"][1], self.backbone["MDyCSP_2"], self.neck
        s1 = h.shape[1] // 2                     # 160 for a 640 input
        cat1 = ops.empty_act(n, s1, s1, 192, x.device)
        x1 = csp1.forward_nhwc(h, ex, out=cat1[..., :128])
        cat2 = ops.empty_act(n, s1 // 2, s1 // 2, 384, x.device)
        x2 = csp2.forward_nhwc(x1, ex, out=cat2[..., :256])
        up = ops.bilinear2x_fwd(x2)
        upc = neck.upsample[1]
        if _NO_PAIR_CONV or upc.out_channels % 64 or upc.out_channels > 128 or up.shape[2] % 2 or upc.in_channels % 64:
            ex.conv_forward(ConvUnit(upc, None, "none"), up, False, None, out=cat1[..., 128:])
        else:
            # 256 -> 64 at 160x160: two output pixels per GEMM row (N = 128; an N = 64 GEMM runs the tensor core at half rate)
            from ..engine import param_epoch
            ver = (upc.weight._version, upc.weight.data_ptr(), param_epoch())
            if self._up_pair is None or self._up_pair[0] != ver:
                self._up_pair = (ver, ops.pack_weight_pair(upc.weight))
            ops.conv3x3_pair_fwd(up, self._up_pair[1], upc.out_channels, shift=None if upc.bias is None else upc.bias.detach(),
                                 out=cat1[..., 128:])
        e1 = neck.encoder_x1.forward_nhwc(cat1, ex)
        ex.conv_forward(ConvUnit(neck.downsample, None, "none"), e1, False, None, out=cat2[..., 256:])
        e2 = neck.encoder_x2.forward_nhwc(cat2, ex)
        return self.head.forward_nhwc([e1, e2])

    def configure_optimizers(self):
        if self.optimizer == "SGD":
            return torch.optim.SGD(self.parameters(), lr=self.learning_rate)
        if self.optimizer == "Adam":
            return torch.optim.Adam(self.parameters(), lr=self.learning_rate)
        raise ValueError(f"Invalid optimizer: {self.optimizer}")
