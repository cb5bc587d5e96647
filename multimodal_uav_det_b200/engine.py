"""Layer executor: runs conv(+BN)(+activation)(+residual) units forward and backward on the
C-ABI kernels and keeps the tape needed for backward.

The torch `nn.Conv2d` / `nn.BatchNorm2d` objects inside the model classes are *parameter
containers only* (same construction order and init as the reference, so seeds and checkpoints
line up); their `forward` is never called.  Activations are NHWC bf16; parameters stay fp32
`nn.Parameter`s and are re-packed to bf16 kernel layout whenever their version changes.

Gradients are accumulated straight into `param.grad` (allocated on demand), the way autograd
would, so torch optimizers / Lightning / the DP bucket reducer all see ordinary `.grad`s.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Tuple

import torch
from torch import nn

from . import ops
from ._lib import EPI_AFFINE, EPI_STATS


# ------------------------------------------------------------------------------------------------
# packed-weight cache
# ------------------------------------------------------------------------------------------------
_PARAM_EPOCH = 0


def bump_param_epoch() -> None:
    """Called by code that rewrites parameters through raw pointers (the fused SGD kernel), which
    torch's version counters cannot see: invalidates every packed-weight cache."""
    global _PARAM_EPOCH
    _PARAM_EPOCH += 1


def param_epoch() -> int:
    return _PARAM_EPOCH


class PackCache:
    """bf16 kernel-layout copies of conv weights, invalidated by the parameter's version counter
    (torch-side updates) or the global parameter epoch (raw-pointer updates)."""

    def __init__(self):
        self._store: Dict[Tuple[int, bool], Tuple[int, torch.Tensor]] = {}

    def get(self, w: torch.Tensor, transposed: bool = False, rows: Optional[int] = None) -> torch.Tensor:
        key = (id(w), transposed)
        ver = (w._version, _PARAM_EPOCH, w.data_ptr())
        hit = self._store.get(key)
        if hit is not None and hit[0] == ver and hit[1].device == w.device:
            return hit[1]
        packed = ops.pack_weight(w.detach(), transposed=transposed, rows=rows)
        self._store[key] = (ver, packed)
        return packed

    def clear(self):
        self._store.clear()


def grad_buffer(p: torch.Tensor) -> Tuple[torch.Tensor, bool]:
    """Returns (tensor to write the gradient into, accumulate flag).  If `p.grad` exists the kernels
    accumulate into it in place; otherwise a fresh zero buffer becomes `p.grad`."""
    if p.grad is None:
        p.grad = torch.zeros_like(p, memory_format=torch.contiguous_format)
        return p.grad, True
    return p.grad, True


# ------------------------------------------------------------------------------------------------
# conv + BN + activation unit
# ------------------------------------------------------------------------------------------------
@dataclass
class ConvUnit:
    """One conv(+BN)(+act) site.  `conv`/`bn` are parameter containers (torch modules)."""
    conv: nn.Conv2d
    bn: Optional[nn.BatchNorm2d]
    act: str                     # 'leaky' | 'silu' | 'relu' | 'gelu' | 'none'
    stem: bool = False           # cin in {1,3}: direct kernel on the NCHW fp32 network input
    s2d: bool = False            # fused space-to-depth(2) gather on the input
    name: str = ""

    @property
    def k(self) -> int:
        return self.conv.kernel_size[0]

    @property
    def stride(self) -> int:
        return self.conv.stride[0]

    @property
    def pad(self) -> int:
        p = self.conv.padding
        return p[0] if isinstance(p, tuple) else int(p)

    @property
    def cout(self) -> int:
        return self.conv.out_channels


@dataclass
class ConvRecord:
    unit: ConvUnit
    x: torch.Tensor                       # input (NHWC bf16, or NCHW fp32 for the stem)
    raw: Optional[torch.Tensor] = None    # pre-BN conv output (train) / pre-activation (bias conv)
    scale: Optional[torch.Tensor] = None
    shift: Optional[torch.Tensor] = None
    mean: Optional[torch.Tensor] = None
    invstd: Optional[torch.Tensor] = None
    has_res: bool = False
    in_hw: Tuple[int, int] = (0, 0)


class Executor:
    """Forward/backward of ConvUnits with a shared pack cache and per-step scratch."""

    def __init__(self):
        self.packs = PackCache()
        self._bn_grads: List[Tuple[torch.Tensor, torch.Tensor]] = []
        self._bn_counters: List[torch.Tensor] = []
        self.grad_ready_hook: Optional[Callable[[torch.Tensor], None]] = None

    # ---- forward ---------------------------------------------------------------------------------
    def conv_forward(self, u: ConvUnit, x: torch.Tensor, train: bool, tape: Optional[list],
                     res: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """y = act(bn(conv(x))) (+ res).  train=True uses batch statistics (two-phase BN) and records
        what backward needs; train=False folds BN into the conv epilogue (one kernel)."""
        w = u.conv.weight
        dev = w.device
        if u.bn is not None and train:
            c = u.cout
            sums = torch.zeros((2, c), dtype=torch.float32, device=dev)
            if u.stem:
                raw = ops.stem_fwd(x, w.detach(), u.k, u.stride, u.pad, epi=EPI_STATS, sum_=sums[0], sumsq=sums[1])
            else:
                raw = ops.conv_fwd(x, self.packs.get(w), c, u.k, u.stride, u.pad, s2d=u.s2d, epi=EPI_STATS,
                                   sum_=sums[0], sumsq=sums[1])
            n, ho, wo, _ = raw.shape
            bn = u.bn
            mom = 0.1 if bn.momentum is None else bn.momentum
            mean, invstd, scale, shift = ops.bn_finalize(
                sums[0], sums[1], n * ho * wo, bn.eps, mom, bn.weight.detach(), bn.bias.detach(),
                bn.running_mean if bn.track_running_stats else None,
                bn.running_var if bn.track_running_stats else None)
            if bn.track_running_stats and bn.num_batches_tracked is not None:
                self._bn_counters.append(bn.num_batches_tracked)
            y = ops.bn_act_fwd(raw, scale, shift, u.act, res=res, out=out)
            if tape is not None:
                tape.append(ConvRecord(u, x, raw, scale, shift, mean, invstd, res is not None, self._in_hw(u, x)))
            return y
        # ---- single-kernel path: eval-mode BN folded / bias-only conv ----
        if u.bn is not None:
            bn = u.bn
            scale = bn.weight.detach() * torch.rsqrt(bn.running_var + bn.eps)
            shift = bn.bias.detach() - bn.running_mean * scale
            if u.conv.bias is not None:
                shift = shift + u.conv.bias.detach() * scale
        else:
            scale = None
            shift = u.conv.bias.detach() if u.conv.bias is not None else None
        need_raw = tape is not None and u.act not in ("none", None)
        if need_raw:
            # training through a non-BN activation: keep the pre-activation for act'(z)
            if u.stem:
                raw = ops.stem_fwd(x, w.detach(), u.k, u.stride, u.pad, scale=scale, shift=shift)
            else:
                raw = ops.conv_fwd(x, self.packs.get(w), u.cout, u.k, u.stride, u.pad, s2d=u.s2d, scale=scale,
                                   shift=shift)
            y = ops.bn_act_fwd(raw, None, None, u.act, res=res, out=out)
            tape.append(ConvRecord(u, x, raw, scale, None, None, None, res is not None, self._in_hw(u, x)))
            return y
        if u.stem:
            y = ops.stem_fwd(x, w.detach(), u.k, u.stride, u.pad, act=u.act, scale=scale, shift=shift)
            if res is not None or out is not None:
                y = ops.add(y, res, out=out)
        else:
            y = ops.conv_fwd(x, self.packs.get(w), u.cout, u.k, u.stride, u.pad, s2d=u.s2d, act=u.act, scale=scale,
                             shift=shift, res=res, out=out)
        if tape is not None:
            tape.append(ConvRecord(u, x, None, scale, None, None, None, res is not None, self._in_hw(u, x)))
        return y

    @staticmethod
    def _in_hw(u: ConvUnit, x: torch.Tensor) -> Tuple[int, int]:
        return (x.shape[2], x.shape[3]) if u.stem else (x.shape[1], x.shape[2])

    def end_forward(self):
        if self._bn_counters:
            torch._foreach_add_(self._bn_counters, 1)
            self._bn_counters = []

    # ---- backward --------------------------------------------------------------------------------
    def conv_backward(self, rec: ConvRecord, dy: torch.Tensor, res: Optional[torch.Tensor] = None,
                      need_dx: bool = True, out: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
        """Given dy = dL/d(output incl. residual), accumulate parameter grads and return
        dx (+ res, an extra gradient flowing into the same tensor, fused into the dgrad epilogue)."""
        u = rec.unit
        w = u.conv.weight
        train_bn = rec.mean is not None
        if train_bn:
            d_raw, dgamma, dbeta = ops.bn_act_bwd(dy, rec.raw, rec.scale, rec.shift, rec.mean, rec.invstd,
                                                  u.bn.weight.detach(), u.act)
            self._bn_grads.append((u.bn.weight, dgamma))
            self._bn_grads.append((u.bn.bias, dbeta))
        else:
            # eval-BN / bias-only conv: dz = dy * act'(raw); raw already holds scale*conv+shift
            d_pre = ops.act_bwd(dy, rec.raw, None, None, u.act) if rec.raw is not None else dy
            if u.conv.bias is not None and u.conv.bias.requires_grad:
                self._bn_grads.append((u.conv.bias, self._channel_sum(d_pre)))
            if rec.scale is not None:
                # frozen-BN scale folded into the conv: d_conv = d_pre * scale
                d_raw = ops.bn_act_fwd(d_pre, rec.scale, None, "none")
            else:
                d_raw = d_pre
        # weight gradient
        if w.requires_grad:
            gbuf, _ = grad_buffer(w)
            if u.stem:
                g = ops.stem_wgrad(rec.x, d_raw, u.k, u.stride, u.pad)
                gbuf.add_(g)
            else:
                dwp = ops.conv_wgrad(rec.x, d_raw, u.k, u.stride, u.pad, s2d=u.s2d)
                ops.unpack_wgrad(dwp, w.shape[0], w.shape[1], u.k, grad=gbuf, accumulate=True)
            if self.grad_ready_hook is not None:
                self.grad_ready_hook(w)
        if not need_dx or u.stem:
            return None
        if u.s2d:
            raise NotImplementedError("dgrad through the fused space-to-depth gather")
        wt = self.packs.get(w, transposed=True)
        return ops.conv_dgrad(d_raw, wt, w.shape[1], u.k, u.stride, u.pad, rec.in_hw, res=res, out=out)

    @staticmethod
    def _channel_sum(t: torch.Tensor) -> torch.Tensor:
        return t.float().sum(dim=(0, 1, 2))

    def end_backward(self):
        """Flush the small per-channel gradients (BN affine, biases) in one multi-tensor pass."""
        if not self._bn_grads:
            return
        acc_p, acc_g = [], []
        for p, g in self._bn_grads:
            if not p.requires_grad:
                continue
            if p.grad is None:
                p.grad = g.clone().reshape(p.shape)
            else:
                acc_p.append(p.grad)
                acc_g.append(g.reshape(p.shape))
        if acc_p:
            torch._foreach_add_(acc_p, acc_g)
        if self.grad_ready_hook is not None:
            for p, _ in self._bn_grads:
                self.grad_ready_hook(p)
        self._bn_grads = []
