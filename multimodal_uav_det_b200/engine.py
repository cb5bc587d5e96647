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

import os
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Tuple

import torch
from torch import nn

from . import ops
from ._lib import EPI_AFFINE, EPI_STATS

_NO_FUSE_S2 = bool(os.environ.get("UAVDET_NO_FUSE_S2"))      # A/B switch: stride-2 data gradients plane by plane
_NO_PAIR = bool(os.environ.get("UAVDET_NO_PAIR_CONV"))        # A/B switch: thin 3x3 layers as plain N = 64 GEMMs
_STEM_IM2COL = bool(os.environ.get("UAVDET_STEM_IM2COL"))    # A/B switch: keep the im2col patch tensor of the 3x3 stems


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
    (torch-side updates) or the global parameter epoch (raw-pointer updates).  `prepack` refreshes a whole
    model's weights (both orientations) with ONE kernel launch — they all change at every optimiser step."""

    def __init__(self):
        self._store: Dict[Tuple[int, bool], Tuple[tuple, torch.Tensor]] = {}
        self._table = None        # (key, device table, n_jobs, total_chunks, [(weight, transposed)])

    @staticmethod
    def _ver(w: torch.Tensor) -> tuple:
        return (w._version, _PARAM_EPOCH, w.data_ptr())

    def get(self, w: torch.Tensor, transposed: bool = False, rows: Optional[int] = None) -> torch.Tensor:
        key = (id(w), transposed)
        ver = self._ver(w)
        hit = self._store.get(key)
        if hit is not None and hit[0] == ver and hit[1].device == w.device:
            return hit[1]
        packed = ops.pack_weight(w.detach(), transposed=transposed, rows=rows)
        self._store[key] = (ver, packed)
        return packed

    def prepack(self, weights, with_transposed: bool) -> None:
        """Make the packed copies of `weights` (4-D conv weights with cin % 32 == 0) current."""
        if not weights:
            return
        orient = (False, True) if with_transposed else (False,)
        stale = False
        for w in weights:
            ver = self._ver(w)
            for t in orient:
                hit = self._store.get((id(w), t))
                if hit is None or hit[0] != ver or hit[1].device != w.device:
                    stale = True
                    break
            if stale:
                break
        if not stale:
            return
        key = (tuple(w.data_ptr() for w in weights), with_transposed, weights[0].device)
        if self._table is None or self._table[0] != key:
            jobs, outs = [], []
            for w in weights:
                o, i, k, _ = w.shape
                for t in orient:
                    out = torch.empty((i if t else o, k * k * (o if t else i)), dtype=torch.bfloat16, device=w.device)
                    jobs.append((w.detach(), out, t))
                    outs.append((w, t, out))
            table, n_jobs, chunks = ops.build_pack_table(jobs)
            self._table = (key, table, n_jobs, chunks, outs)
        _, table, n_jobs, chunks, outs = self._table
        ops.pack_weights_batched(table, n_jobs, chunks)
        for w, t, out in outs:
            self._store[(id(w), t)] = (self._ver(w), out)

    def clear(self):
        self._store.clear()
        self._table = None


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


@dataclass
class DynSpec:
    """One dynamic-kernel convolution site: attention MLP over the pooled input mixes an expert bank into one
    kernel per sample (reference _base.py:26-77 DyConvModule, DySOEM_SimFPN.py:38-94 DynamicSOEM)."""
    bank: Callable[[], torch.Tensor]                       # () -> (K,O,I,k,k) fp32, detached, contiguous
    bias_bank: Optional[Callable[[], torch.Tensor]]        # () -> (K,O) fp32 or None
    bank_params: Callable[[], list]                        # () -> [(param, index into d_bank | None)] see accumulate
    w1: torch.Tensor                                       # first attention layer weight (hid, C[,1,1])
    b1: Optional[torch.Tensor]
    w2: torch.Tensor                                       # (K, hid[,1,1])
    b2: torch.Tensor
    temperature: float
    bn: nn.BatchNorm2d
    act: str
    cin: int                                               # channels of the tensor the site reads (before s2d)
    cout: int
    k: int
    stride: int
    pad: int
    s2d: bool = False
    stem: bool = False


@dataclass
class DynRecord:
    spec: DynSpec
    x: torch.Tensor
    pooled: torch.Tensor
    hidden: torch.Tensor
    attn: torch.Tensor
    bank: torch.Tensor
    bias_bank: Optional[torch.Tensor]
    raw: torch.Tensor
    scale: Optional[torch.Tensor]
    shift: Optional[torch.Tensor]
    mean: Optional[torch.Tensor]
    invstd: Optional[torch.Tensor]
    in_hw: Tuple[int, int] = (0, 0)
    has_res: bool = False


class Executor:
    """Forward/backward of ConvUnits with a shared pack cache and per-step scratch."""

    def __init__(self):
        self.packs = PackCache()
        # weight gradients on a side stream (joined in end_backward); UAVDET_NO_WGRAD_OVERLAP=1 keeps them in line
        self.overlap_wgrad = not os.environ.get("UAVDET_NO_WGRAD_OVERLAP")
        self._side: Optional[torch.cuda.Stream] = None
        self._side_keep: list = []
        self._zero_arena: Optional[torch.Tensor] = None
        self._zero_cursor = 0
        self._zero_need = 0
        self._bn_counters: List[torch.Tensor] = []
        self._const: Dict[Tuple[str, int, str], torch.Tensor] = {}
        self._bn_fold: Dict[int, Tuple[tuple, torch.Tensor, torch.Tensor]] = {}
        self._s2f: Dict[tuple, torch.Tensor] = {}    # plane-fused stride-2 data-gradient weights (rebuilt from the pack every step)
        self._pair: Dict[tuple, tuple] = {}          # pixel-pair weight matrices of thin 3x3 layers: (id(w), transposed) -> (version, buffer)
        self.grad_ready_hook: Optional[Callable[[torch.Tensor], None]] = None

    def producer_streams(self) -> list:
        """Streams that gradient writes of the current backward may be pending on besides the caller's current stream:
        the weight-gradient side stream.  A bucket reducer must wait on all of them before it reads a gradient arena."""
        return [self._side] if (self._side is not None and self._side_keep) else []

    # ---- per-step zeroed fp32 scratch ---------------------------------------------------------------
    def begin_step(self, device) -> None:
        """Zero the scratch arena the step's small accumulators (BN sums, BN-backward sums) are carved from:
        one fill launch instead of one per layer.  Sized from the previous step's demand."""
        need = max(self._zero_need, 1 << 16)
        if self._zero_arena is None or self._zero_arena.device != device or self._zero_arena.numel() < need:
            self._zero_arena = torch.zeros(need + need // 4, dtype=torch.float32, device=device)
        else:
            self._zero_arena.zero_()
        self._zero_cursor = 0
        self._zero_need = 0

    def zeros(self, rows: int, cols: int, device) -> torch.Tensor:
        """(rows, cols) fp32 zeros: a 128-byte-aligned slice of the arena zeroed by begin_step, else a fresh fill."""
        n = rows * cols
        n_al = (n + 31) // 32 * 32
        self._zero_need += n_al
        a = self._zero_arena
        if a is not None and a.device == device and self._zero_cursor + n_al <= a.numel():
            out = a[self._zero_cursor:self._zero_cursor + n].view(rows, cols)
            self._zero_cursor += n_al
            return out
        return torch.zeros((rows, cols), dtype=torch.float32, device=device)

    # ---- forward ---------------------------------------------------------------------------------
    @staticmethod
    def stem_as_gemm(cin: int, cout: int, k: int) -> bool:
        """cin<=3 stems whose taps fit 32 channels run as im2col + 1x1 implicit GEMM on the tensor cores; the
        others (RTMUAVDet's 5x5 stem: 75 taps; DySOEM's 1x1: nothing to gather) keep the direct kernel."""
        return cout % 32 == 0 and k > 1 and cin * k * k <= 32

    @staticmethod
    def stem_in_smem(cin: int, cout: int, k: int) -> bool:
        """3x3 cin<=3 -> 32 stems: the im2col rows are built in shared memory by the stem_mma kernels, nothing is
        materialised (UAVDET_STEM_IM2COL=1 keeps the patch tensor: A/B switch)."""
        return not _STEM_IM2COL and ops.stem_mma_supported(cin, cout, k)

    def _stem_pack(self, w: torch.Tensor) -> torch.Tensor:
        """(O, cin, k, k) fp32 -> bf16 [O][32]: w.flatten(1) zero-padded to the 32 im2col channels."""
        key = (id(w), "stem")
        ver = self.packs._ver(w)
        hit = self.packs._store.get(key)
        if hit is not None and hit[0] == ver:
            return hit[1]
        flat = w.detach().flatten(1)
        packed = torch.nn.functional.pad(flat, (0, 32 - flat.shape[1])).to(torch.bfloat16).contiguous()
        self.packs._store[key] = (ver, packed)
        return packed

    def conv_forward(self, u: ConvUnit, x: torch.Tensor, train: bool, tape: Optional[list],
                     res: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """y = act(bn(conv(x))) (+ res).  train=True uses batch statistics (two-phase BN) and records
        what backward needs; train=False folds BN into the conv epilogue (one kernel)."""
        w = u.conv.weight
        dev = w.device
        in_hw = self._in_hw(u, x)
        direct_stem = u.stem
        k, stride, pad, s2d = u.k, u.stride, u.pad, u.s2d
        stem_mma = False
        if u.stem and self.stem_as_gemm(w.shape[1], u.cout, u.k):
            wp = self._stem_pack(w)
            direct_stem = False
            if self.stem_in_smem(w.shape[1], u.cout, u.k):
                stem_mma = True                                   # patch rows built in shared memory (stem_mma.cu)
            else:
                x = ops.stem_im2col(x, u.k, u.stride, u.pad)      # (n, ho, wo, 32) bf16 patches
                k, stride, pad, s2d = 1, 1, 0, False
        elif not u.stem:
            wp = self.packs.get(w)
        if u.bn is not None and train:
            if u.conv.bias is not None:
                raise NotImplementedError("conv bias in front of a train-mode BatchNorm (the bias would have to enter the "
                                          "batch statistics and receive a gradient); the reference never builds one")
            if u.bn.momentum is None:
                raise NotImplementedError("BatchNorm2d(momentum=None) (cumulative moving average) is not implemented")
            c = u.cout
            sums = self.zeros(2, c, dev)
            if direct_stem:
                raw = ops.stem_fwd(x, w.detach(), k, stride, pad, epi=EPI_STATS, sum_=sums[0], sumsq=sums[1])
            elif stem_mma:
                raw = ops.stem_mma_fwd(x, wp, k, stride, pad, epi=EPI_STATS, sum_=sums[0], sumsq=sums[1])
            elif self._pair_ok(u, x, s2d):
                # thin 3x3 layer (32 -> 64 at 320x320): two output pixels per GEMM row, whole pixel pairs as K = 64 k-blocks
                raw = ops.conv3x3_pair_fwd(x, self._pair_weight(w, wp), c, epi=EPI_STATS, sum_=sums[0], sumsq=sums[1])
            else:
                raw = ops.conv_fwd(x, wp, c, k, stride, pad, s2d=s2d, epi=EPI_STATS, sum_=sums[0], sumsq=sums[1])
            n, ho, wo, _ = raw.shape
            bn = u.bn
            mom = bn.momentum
            y, mean, invstd, scale, shift = ops.bn_train_fwd(
                raw, sums[0], sums[1], n * ho * wo, bn.eps, mom, bn.weight.detach(), bn.bias.detach(),
                bn.running_mean if bn.track_running_stats else None,
                bn.running_var if bn.track_running_stats else None, u.act, res=res, out=out)
            if bn.track_running_stats and bn.num_batches_tracked is not None:
                self._bn_counters.append(bn.num_batches_tracked)
            if tape is not None:
                tape.append(ConvRecord(u, x, raw, scale, shift, mean, invstd, res is not None, in_hw))
            return y
        # ---- single-kernel path: eval-mode BN folded / bias-only conv ----
        if u.bn is not None:
            scale, shift = self.bn_fold(u.bn, u.conv.bias)
        else:
            scale = None
            shift = u.conv.bias.detach() if u.conv.bias is not None else None
        need_raw = tape is not None and u.act not in ("none", None)
        if need_raw:
            # training through a non-BN activation: keep the pre-activation for act'(z)
            if direct_stem:
                raw = ops.stem_fwd(x, w.detach(), k, stride, pad, scale=scale, shift=shift)
            elif stem_mma:
                raw = ops.stem_mma_fwd(x, wp, k, stride, pad, scale=scale, shift=shift)
            else:
                raw = ops.conv_fwd(x, wp, u.cout, k, stride, pad, s2d=s2d, scale=scale, shift=shift)
            y = ops.bn_act_fwd(raw, None, None, u.act, res=res, out=out)
            tape.append(ConvRecord(u, x, raw, scale, None, None, None, res is not None, in_hw))
            return y
        if direct_stem or stem_mma:
            if stem_mma:
                y = ops.stem_mma_fwd(x, wp, k, stride, pad, act=u.act, scale=scale, shift=shift)
            else:
                y = ops.stem_fwd(x, w.detach(), k, stride, pad, act=u.act, scale=scale, shift=shift)
            if res is not None or out is not None:
                y = ops.add(y, res, out=out)
        else:
            y = ops.conv_fwd(x, wp, u.cout, k, stride, pad, s2d=s2d, act=u.act, scale=scale, shift=shift, res=res,
                             out=out)
        if tape is not None:
            tape.append(ConvRecord(u, x, None, scale, None, None, None, res is not None, in_hw))
        return y

    @staticmethod
    def _pair_ok(u: ConvUnit, x: torch.Tensor, s2d: bool) -> bool:
        """3x3 stride-1 pad-1 layers with 64 output channels run as pixel-pair GEMMs (ops.conv3x3_pair_fwd)."""
        if _NO_PAIR or u.stem or s2d or u.k != 3 or u.stride != 1 or u.pad != 1 or u.cout != 64 or x.dim() != 4:
            return False
        cin = x.shape[3]
        # (64-channel inputs stay on the halo-tile mode of the plain kernel: DySOEM_SimFPN's 64 -> 64 layers at 320x320 measured
        # 0.4 ms per step slower as pair GEMMs)
        return x.shape[2] % 2 == 0 and ((cin % 64 == 0 and cin != 64) or (cin == 32 and x.stride(2) == 32))

    def _pair_weight(self, w: torch.Tensor, packed: torch.Tensor, transposed: bool = False) -> torch.Tensor:
        """Pair weight matrix of `w` (of its data gradient: transposed), rebuilt from its current bf16 pack when the
        weight changed (two strided copies into a persistent buffer whose zero blocks never change)."""
        ver = self.packs._ver(w)
        hit = self._pair.get((id(w), transposed))
        if hit is not None and hit[0] == ver and hit[1].device == w.device:
            return hit[1]
        buf = ops.pack_weight_pair(packed, out=hit[1] if hit is not None and hit[1].device == w.device else None,
                                   flip=transposed)
        self._pair[(id(w), transposed)] = (ver, buf)
        return buf

    def bn_fold(self, bn: nn.BatchNorm2d, conv_bias: Optional[torch.Tensor] = None):
        """Eval-mode BatchNorm folded to (scale, shift) = (gamma*rsqrt(var+eps), beta - mean*scale [+ bias*scale]),
        cached until a parameter / running statistic of the layer changes (version counters; the parameter epoch covers
        raw-pointer updates by the fused optimiser)."""
        ver = (bn.weight._version, bn.bias._version, bn.running_mean._version, bn.running_var._version, _PARAM_EPOCH,
               bn.weight.data_ptr(), bn.running_var.data_ptr(), None if conv_bias is None else conv_bias._version)
        hit = self._bn_fold.get(id(bn))
        if hit is not None and hit[0] == ver:
            return hit[1], hit[2]
        scale = bn.weight.detach() * torch.rsqrt(bn.running_var + bn.eps)
        shift = bn.bias.detach() - bn.running_mean * scale
        if conv_bias is not None:
            shift = shift + conv_bias.detach() * scale
        self._bn_fold[id(bn)] = (ver, scale, shift)
        return scale, shift

    def _constant(self, kind: str, c: int, device) -> torch.Tensor:
        key = (kind, c, str(device))
        t = self._const.get(key)
        if t is None:
            t = (torch.ones if kind == "ones" else torch.zeros)(c, dtype=torch.float32, device=device)
            self._const[key] = t
        return t

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
            d_raw = self._bn_backward(u.bn, dy, rec, u.act, u.cout)
        else:
            # eval-BN / bias-only conv: dz = dy * act'(raw); raw already holds scale*conv+shift
            d_pre = ops.act_bwd(dy, rec.raw, None, None, u.act) if rec.raw is not None else dy
            if u.conv.bias is not None and u.conv.bias.requires_grad:
                dbias = self._channel_sum(d_pre)
                self._accumulate(u.conv.bias, dbias * rec.scale if rec.scale is not None else dbias)
            if u.bn is not None:
                self._frozen_bn_affine_grads(u.bn, d_pre, rec.raw, u.conv.bias)
            if rec.scale is not None:
                # frozen-BN scale folded into the conv: d_conv = d_pre * scale
                d_raw = ops.bn_act_fwd(d_pre, rec.scale, None, "none")
            else:
                d_raw = d_pre
        # data gradient first (it is the critical path of backward) ...
        dx = None
        if need_dx and not u.stem:
            if u.s2d:
                raise NotImplementedError("dgrad through the fused space-to-depth gather")
            wt = self.packs.get(w, transposed=True)
            cin = w.shape[1]
            h_in, w_in = rec.in_hw
            fused = (not _NO_FUSE_S2 and u.stride == 2 and u.k == 3 and u.pad == 1 and cin in (32, 64)
                     and h_in % 2 == 0 and w_in % 2 == 0 and (out is None or out.stride(2) == cin)
                     and (res is None or res.stride(2) == cin))
            pair_dgrad = (not _NO_PAIR and not fused and u.k == 3 and u.stride == 1 and u.pad == 1 and cin in (32, 64)
                          and res is None and out is None and w_in % 2 == 0
                          and (u.cout % 64 == 0 or (u.cout == 32 and d_raw.stride(2) == 32))
                          and not (cin == 64 and u.cout == 64))        # 64 -> 64: the halo-tile mode of the plain kernel
            if pair_dgrad:
                # thin data gradient (N = cin <= 64): two pixels per GEMM row, the mirrored transposed filter as a pair matrix
                dx = ops.conv3x3_pair_fwd(d_raw, self._pair_weight(w, wt, transposed=True), cin)
            elif fused:
                # all four output-parity planes as one N = 4*cin GEMM over a re-laid-out (zero-padded) weight matrix
                key = (id(w), "s2f")
                buf = self._s2f.get(key)
                if buf is None or buf.device != wt.device:
                    buf = torch.empty((4 * cin, 4 * w.shape[0]), dtype=torch.bfloat16, device=wt.device)
                    self._s2f[key] = buf
                dx = ops.conv_dgrad_s2_fused(d_raw, ops.pack_dgrad_s2_fused(wt, cin, w.shape[0], out=buf), cin, res=res, out=out)
            else:
                dx = ops.conv_dgrad(d_raw, wt, cin, u.k, u.stride, u.pad, rec.in_hw, res=res, out=out)
        # ... then the weight gradient: nothing downstream in backward depends on it, so it goes to a side stream
        # where its tensor-core work overlaps the HBM-bound BatchNorm backward of the next layer (the kernel
        # leaves shared memory for those blocks to share the SM)
        if w.requires_grad:
            side = self._wgrad_stream(d_raw.device)
            if side is not None:
                side.wait_stream(torch.cuda.current_stream())
                self._side_keep.append((rec.x, d_raw))          # keep the operands alive until the join
                with torch.cuda.stream(side):
                    self._weight_grad(rec, u, w, d_raw)
            else:
                self._weight_grad(rec, u, w, d_raw)
        return dx

    def _bn_backward(self, bn: nn.BatchNorm2d, dy: torch.Tensor, rec, act: str, cout: int) -> torch.Tensor:
        """Train-mode BN(+activation) backward of one layer.  dgamma / dbeta are added straight into `bn.weight.grad` /
        `bn.bias.grad` by the kernel and the gradient-ready hook fires right away, so a bucketed all-reduce can start
        while backward continues (it used to wait for one flush at the end of backward)."""
        gw = gb = None
        if bn.weight.requires_grad and bn.bias.requires_grad:
            gw, _ = grad_buffer(bn.weight)
            gb, _ = grad_buffer(bn.bias)
            if not (gw.is_contiguous() and gb.is_contiguous() and gw.dtype == torch.float32):
                gw = gb = None
        d_raw, dgamma, dbeta = ops.bn_act_bwd(dy, rec.raw, rec.scale, rec.shift, rec.mean, rec.invstd, bn.weight.detach(),
                                              act, buf=self.zeros(6, cout, dy.device), grad_gamma=gw, grad_beta=gb)
        if gw is None:
            self._accumulate(bn.weight, dgamma)
            self._accumulate(bn.bias, dbeta)
        elif self.grad_ready_hook is not None:
            self.grad_ready_hook(bn.weight)
            self.grad_ready_hook(bn.bias)
        return d_raw

    def _frozen_bn_affine_grads(self, bn: nn.BatchNorm2d, d_pre: torch.Tensor, z: Optional[torch.Tensor],
                                conv_bias: Optional[torch.Tensor]) -> None:
        """Eval-mode BatchNorm with trainable affine parameters (frozen-statistics fine-tuning): torch's autograd
        still produces dgamma = sum(dz * xhat) and dbeta = sum(dz).  `z` = gamma*xhat + beta is what the tape keeps, so
        xhat = (z - beta) / gamma (a zero gamma gets a zero gradient: xhat is not recoverable from z there)."""
        if not (bn.weight.requires_grad or bn.bias.requires_grad):
            return
        c = bn.num_features
        if z is None:       # activation-free layer: the forward kept no pre-activation, nothing to correlate with
            raise NotImplementedError("gradient of frozen-BatchNorm affine parameters of an activation-free layer")
        sums = self.zeros(2, c, d_pre.device)
        ops.bn_bwd_reduce(d_pre, z, self._constant("ones", c, d_pre.device), self._constant("zeros", c, d_pre.device),
                          "none", sums[0], sums[1])
        gamma, beta = bn.weight.detach(), bn.bias.detach()
        if bn.bias.requires_grad:
            self._accumulate(bn.bias, sums[0].clone())
        if bn.weight.requires_grad:
            safe = torch.where(gamma == 0, torch.ones_like(gamma), gamma)
            self._accumulate(bn.weight, torch.where(gamma == 0, torch.zeros_like(gamma), (sums[1] - beta * sums[0]) / safe))

    def _wgrad_stream(self, device):
        if not self.overlap_wgrad or device.type != "cuda":
            return None
        if self._side is None or self._side.device != device:
            # high priority: when its CTAs and the (thousands of) BatchNorm blocks of the main stream are both
            # pending, the block scheduler must place the weight-gradient CTAs first or they would only start
            # once the streaming kernel has drained
            self._side = torch.cuda.Stream(device=device, priority=-1)
        return self._side

    def _weight_grad(self, rec: ConvRecord, u: ConvUnit, w: torch.Tensor, d_raw: torch.Tensor) -> None:
        if True:    # (block kept at this indentation: it runs under the side-stream context of the caller)
            gbuf, _ = grad_buffer(w)
            if u.stem and rec.x.dtype == torch.bfloat16:
                # im2col stem: a 1x1 weight gradient over the 32 patch channels; the first cin*k*k columns are dW
                dwp = ops.conv_wgrad(rec.x, d_raw, 1, 1, 0)
                kk = w.shape[1] * u.k * u.k
                gbuf.add_(dwp[:, :kk].reshape(w.shape))
            elif u.stem and self.stem_as_gemm(w.shape[1], u.cout, u.k) and self.stem_in_smem(w.shape[1], u.cout, u.k):
                kk = w.shape[1] * u.k * u.k
                dwp = ops.stem_mma_wgrad(rec.x, d_raw, u.k, u.stride, u.pad, out=self.zeros(32, 32, d_raw.device))
                gbuf.add_(dwp[:, :kk].reshape(w.shape))
            elif u.stem:
                g = ops.stem_wgrad(rec.x, d_raw, u.k, u.stride, u.pad)
                gbuf.add_(g)
            else:
                packed_view = gbuf.permute(0, 2, 3, 1)
                if packed_view.is_contiguous():
                    # the gradient buffer already has the kernel's [O][kh][kw][I] layout (1x1 convs; conv
                    # weights the flat trainer keeps channels-last): accumulate straight into it
                    ops.conv_wgrad(rec.x, d_raw, u.k, u.stride, u.pad, s2d=u.s2d,
                                   out=packed_view.reshape(w.shape[0], -1))
                else:
                    dwp = ops.conv_wgrad(rec.x, d_raw, u.k, u.stride, u.pad, s2d=u.s2d)
                    ops.unpack_wgrad(dwp, w.shape[0], w.shape[1], u.k, grad=gbuf, accumulate=True)
            if self.grad_ready_hook is not None:
                self.grad_ready_hook(w)

    # ---- dynamic-kernel convolution ----------------------------------------------------------------
    def dyn_forward(self, sp: DynSpec, x: torch.Tensor, train: bool, tape: Optional[list]) -> torch.Tensor:
        """y = act(bn(conv(x, sum_k a_k(x) W_k) + sum_k a_k(x) b_k)), one kernel per sample as a batched GEMM
        operand.  x: NHWC bf16 (or the NCHW fp32 network input for a cin=3 site)."""
        pooled = ops.gap_nchw(x) if sp.stem else ops.gap(x, s2d=sp.s2d)
        w1 = sp.w1.detach().flatten(1)
        w2 = sp.w2.detach().flatten(1)
        b1 = sp.b1.detach() if sp.b1 is not None else None
        attn, hidden = ops.attn_mlp_softmax(pooled, w1, b1, w2, sp.b2.detach(), float(sp.temperature), want_hidden=True)
        n = attn.shape[0]
        bank = sp.bank()
        bias_bank = sp.bias_bank() if sp.bias_bank is not None else None
        in_hw = (x.shape[2], x.shape[3]) if sp.stem else (x.shape[1], x.shape[2])
        bn, k, s, p, co = sp.bn, sp.k, sp.stride, sp.pad, sp.cout
        direct_stem = sp.stem
        stem_mma = False
        if sp.stem:
            bias_b = None
            if self.stem_as_gemm(sp.cin, co, k):
                # per-sample [O][32] bf16 kernels (taps zero-padded to 32), mixed from the expert bank by the
                # aggregation kernel; the patch rows are built in shared memory (stem_mma) or, for shapes that kernel
                # does not cover, materialised by im2col and fed to the batched 1x1 implicit GEMM
                w_b = ops.dyn_aggregate_stem(attn, bank)
                direct_stem = False
                if self.stem_in_smem(sp.cin, co, k):
                    stem_mma = True
                else:
                    x = ops.stem_im2col(x, k, s, p)
                    k, s, p = 1, 1, 0
            else:
                # direct CUDA-core stem (no shipped configuration takes this branch): fp32 per-sample kernels
                w_b = torch.mm(attn, bank.flatten(1)).view(n, *bank.shape[1:])
        else:
            w_b, bias_b = ops.dyn_aggregate(attn, bank, bias_bank=bias_bank)
        if train:
            sums = self.zeros(2, co, attn.device)
            if direct_stem:
                raw = ops.stem_fwd(x, w_b, k, s, p, epi=EPI_STATS, sum_=sums[0], sumsq=sums[1], per_sample_w=True)
            elif stem_mma:
                raw = ops.stem_mma_fwd(x, w_b, k, s, p, epi=EPI_STATS, sum_=sums[0], sumsq=sums[1])
            else:
                raw = ops.conv_fwd(x, w_b, co, k, s, p, s2d=sp.s2d, w_batch=n, epi=EPI_STATS, shift=bias_b,
                                   shift_per_sample=bias_b is not None, sum_=sums[0], sumsq=sums[1])
            _, ho, wo, _ = raw.shape
            if bn.momentum is None:
                raise NotImplementedError("BatchNorm2d(momentum=None) (cumulative moving average) is not implemented")
            mom = bn.momentum
            y, mean, invstd, scale, shift = ops.bn_train_fwd(raw, sums[0], sums[1], n * ho * wo, bn.eps, mom,
                                                             bn.weight.detach(), bn.bias.detach(), bn.running_mean,
                                                             bn.running_var, sp.act)
            if bn.num_batches_tracked is not None:
                self._bn_counters.append(bn.num_batches_tracked)
            if tape is not None:
                tape.append(DynRecord(sp, x, pooled, hidden, attn, bank, bias_bank, raw, scale, shift, mean, invstd, in_hw))
            return y
        scale, shift = self.bn_fold(bn)
        per_sample_shift = bias_b is not None
        if per_sample_shift:
            shift = (shift.unsqueeze(0) + bias_b * scale.unsqueeze(0)).contiguous()
        if tape is not None:
            # frozen-BN training: keep z = scale*conv + shift for act'(z)
            if direct_stem:
                z = ops.stem_fwd(x, w_b, k, s, p, scale=scale, shift=shift, per_sample_w=True)
            elif stem_mma:
                z = ops.stem_mma_fwd(x, w_b, k, s, p, scale=scale, shift=shift)
            else:
                z = ops.conv_fwd(x, w_b, co, k, s, p, s2d=sp.s2d, w_batch=n, scale=scale, shift=shift,
                                 shift_per_sample=per_sample_shift)
            y = ops.bn_act_fwd(z, None, None, sp.act)
            tape.append(DynRecord(sp, x, pooled, hidden, attn, bank, bias_bank, z, scale, None, None, None, in_hw))
            return y
        if direct_stem:
            return ops.stem_fwd(x, w_b, k, s, p, act=sp.act, scale=scale, shift=shift, per_sample_w=True)
        if stem_mma:
            return ops.stem_mma_fwd(x, w_b, k, s, p, act=sp.act, scale=scale, shift=shift)
        return ops.conv_fwd(x, w_b, co, k, s, p, s2d=sp.s2d, w_batch=n, act=sp.act, scale=scale, shift=shift,
                            shift_per_sample=per_sample_shift)

    def _accumulate(self, p: torch.Tensor, g: torch.Tensor, fire: bool = True) -> None:
        if not p.requires_grad:
            return
        if p.grad is None:
            p.grad = g.reshape(p.shape).clone()
        else:
            p.grad.add_(g.reshape(p.shape))
        if fire and self.grad_ready_hook is not None:
            self.grad_ready_hook(p)

    def dyn_backward(self, rec: DynRecord, dy: torch.Tensor, res: Optional[torch.Tensor] = None,
                     need_dx: bool = True) -> Optional[torch.Tensor]:
        """Backward of dyn_forward.  The per-sample kernel gradients (n x |W| fp32, <= 67 MB at the largest
        site) are contracted with the attention / the expert bank by one kernel; the attention-MLP chain is
        a handful of (n x C)-sized fp32 products; the pooled-input gradient is broadcast over the pixels by
        the data-gradient epilogue (per-sample shift)."""
        sp = rec.spec
        n = rec.attn.shape[0]
        if rec.mean is not None:
            d_raw = self._bn_backward(sp.bn, dy, rec, sp.act, sp.cout)
        else:
            d_pre = ops.act_bwd(dy, rec.raw, None, None, sp.act)
            self._frozen_bn_affine_grads(sp.bn, d_pre, rec.raw, None)
            d_raw = ops.bn_act_fwd(d_pre, rec.scale, None, "none")
        # per-sample kernel gradient -> expert-bank gradient + attention gradient
        if sp.stem and rec.x.dtype == torch.bfloat16:
            kk = sp.cin * sp.k * sp.k                              # im2col stem: per-sample 1x1 wgrad, first kk columns
            dwb = ops.conv_wgrad(rec.x, d_raw, 1, 1, 0, per_sample=True)[:, :, :kk].contiguous()
        elif sp.stem and self.stem_as_gemm(sp.cin, sp.cout, sp.k) and self.stem_in_smem(sp.cin, sp.cout, sp.k):
            kk = sp.cin * sp.k * sp.k
            dwb = ops.stem_mma_wgrad(rec.x, d_raw, sp.k, sp.stride, sp.pad, per_sample=True)[:, :, :kk].contiguous()
        elif sp.stem:
            dwb = ops.stem_wgrad(rec.x, d_raw, sp.k, sp.stride, sp.pad, per_sample=True)
        else:
            dwb = ops.conv_wgrad(rec.x, d_raw, sp.k, sp.stride, sp.pad, s2d=sp.s2d, per_sample=True)
        d_attn = self.zeros(n, rec.attn.shape[1], dy.device)
        bank_params = sp.bank_params()
        # one parameter holds the whole bank (DyConvModule.weights): its gradient buffer IS d_bank
        whole = (len(bank_params) == 1 and bank_params[0][1] is None and not bank_params[0][2]
                 and bank_params[0][0].requires_grad)
        if whole:
            d_bank, _ = grad_buffer(bank_params[0][0])
            whole = d_bank.is_contiguous() and d_bank.dtype == torch.float32 and d_bank.shape == rec.bank.shape
        if not whole:
            d_bank = torch.zeros_like(rec.bank)
        ops.dyn_bwd_contract(dwb.view(n, -1), rec.attn, rec.bank, d_bank, d_attn, packed=not sp.stem)
        d_bias_bank = None
        if rec.bias_bank is not None:
            # per-sample channel sums of d_raw = pool x pixel count; both contractions in one small kernel
            d_bias_bank = ops.dyn_bias_bwd(ops.gap(d_raw), float(d_raw.shape[1] * d_raw.shape[2]), rec.attn,
                                           rec.bias_bank.contiguous(), d_attn)
        # Gradient-ready hooks are fired at the END of this function: a data-parallel trainer may step a bucket right
        # behind its all-reduce, and the expert bank / attention weights are still read below (fp32 masters, no pack)
        ready = []
        if whole:
            ready.append(bank_params[0][0])
        else:
            for p, idx, is_bias in bank_params:
                src = d_bias_bank if is_bias else d_bank
                self._accumulate(p, src if idx is None else src[idx], fire=False)
                ready.append(p)
        # attention MLP backward (softmax(s/T), Linear/conv1x1, ReLU, Linear/conv1x1): two launches that add the
        # parameter gradients straight into their `.grad` buffers and return the pooled-input gradient, already
        # divided by the pool size
        a = rec.attn
        w1 = sp.w1.detach().flatten(1)
        w2 = sp.w2.detach().flatten(1)
        h, w = rec.in_hw
        want_dx = need_dx and not sp.stem
        mlp_params = [sp.w1, sp.b1, sp.w2, sp.b2]
        bufs = []
        for p in mlp_params:
            if p is None or not p.requires_grad:
                bufs.append(None)
                continue
            gb, _ = grad_buffer(p)
            if not (gb.is_contiguous() and gb.dtype == torch.float32):
                raise RuntimeError("attention-MLP gradient buffers must be contiguous float32")
            bufs.append(gb)
        d_pooled = ops.attn_mlp_bwd(a, d_attn.contiguous(), rec.hidden, rec.pooled, w1, w2, float(sp.temperature),
                                    (4.0 if sp.s2d else 1.0) / (h * w), bufs[0], bufs[1], bufs[2], bufs[3],
                                    want_d_pooled=want_dx)
        ready += [p for p, gb in zip(mlp_params, bufs) if gb is not None]
        dx = None
        if want_dx:
            wt, _ = ops.dyn_aggregate(a, rec.bank, transposed=True)
            if sp.s2d:
                dx = ops.conv_dgrad_s2d(d_raw, wt, sp.cin, sp.k, sp.pad, w_batch=n, res=res, shift=d_pooled)
            else:
                dx = ops.conv_dgrad(d_raw, wt, sp.cin, sp.k, sp.stride, sp.pad, rec.in_hw, w_batch=n, res=res,
                                    shift=d_pooled, shift_per_sample=True)
        if self.grad_ready_hook is not None:
            for p in ready:
                if p.requires_grad:
                    self.grad_ready_hook(p)
        return dx

    @staticmethod
    def _channel_sum(t: torch.Tensor) -> torch.Tensor:
        """Per-channel sum over (n, h, w) of an NHWC bf16 tensor: the streaming pool kernel (one read of t) and a
        reduction of its (n, c) output — a cast to fp32 plus an ATen reduction moved the tensor three times."""
        return ops.gap(t).sum(dim=0) * float(t.shape[1] * t.shape[2])

    def end_backward(self):
        """Join the weight-gradient stream (every per-channel gradient was already written and announced per layer)."""
        if self._side is not None and self._side_keep:
            torch.cuda.current_stream().wait_stream(self._side)
        self._side_keep = []
