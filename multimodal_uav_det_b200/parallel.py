"""Data-parallel training plumbing (new work: the reference trains on one GPU).

One process per GPU.  Parameters and gradients are re-homed into a few large flat fp32 arenas
("buckets", reverse registration order ~ the order backward produces gradients):
  * `param.data` / `param.grad` become views, so the executor's in-place gradient accumulation,
    torch optimizers and checkpoints keep working unchanged;
  * the optimiser step is ONE fused SGD-momentum kernel launch per bucket;
  * with world_size > 1 each bucket is all-reduced (NCCL, sum) on a side stream as soon as the
    last gradient of the bucket has been written, overlapping the rest of backward; the 1/world
    mean is folded into the SGD kernel's `grad_scale`.  BatchNorm statistics stay per replica,
    matching Lightning's default DDP for this model (no SyncBN in the reference).
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from . import ops
from .engine import bump_param_epoch


class _Bucket:
    def __init__(self, params: List[torch.nn.Parameter], device):
        self.params = params
        self.offsets = []
        off = 0
        for p in params:
            self.offsets.append(off)
            off += (p.numel() + 3) // 4 * 4          # keep every view 16-byte aligned
        self.numel = off
        self.param = torch.zeros(off, dtype=torch.float32, device=device)
        self.grad = torch.zeros(off, dtype=torch.float32, device=device)
        self.momentum = torch.zeros(off, dtype=torch.float32, device=device)
        self.pending = 0
        self.work = None
        self.updated = False
        for p, o in zip(params, self.offsets):
            p.data, p.grad = self._views(p, o)

    def _views(self, p, o):
        """Parameter / gradient views into the arenas.  Conv weights of the implicit-GEMM layers are stored
        channels-last ([O][kh][kw][I] — the layout the weight-gradient kernel accumulates in and the bf16
        pack reads), exposed through a permuted (O,I,kh,kw) view: values, shapes and `state_dict` keys are
        unchanged, only the strides differ (torch.channels_last)."""
        n = p.numel()
        flat_p, flat_g = self.param[o:o + n], self.grad[o:o + n]
        if p.dim() == 4 and p.shape[2] * p.shape[3] > 1 and p.shape[1] % 32 == 0:
            oc, ic, kh, kw = p.shape
            view = flat_p.view(oc, kh, kw, ic).permute(0, 3, 1, 2)
            gview = flat_g.view(oc, kh, kw, ic).permute(0, 3, 1, 2)
        else:
            view, gview = flat_p.view(p.shape), flat_g.view(p.shape)
        if p.data.data_ptr() != view.data_ptr():
            view.copy_(p.data)
        return view, gview


class FlatSGDTrainer:
    """Flat-arena SGD(momentum) + bucketed gradient all-reduce for one model replica.

    accumulate_grad_batches = k (reference params.yaml:23 trains with 2): `step()` is called after k backward passes;
    gradients accumulate in the arenas, a bucket is all-reduced once — when its last gradient of the k-th (boundary)
    micro-batch has been written, so the reduce still overlaps that backward — and the optimiser sees the mean over
    the k micro-batches and the world (Lightning divides each micro-batch loss by k; here 1/(k*world) is folded into
    the fused SGD kernel).

    Learning rate and momentum live in a 3-float device tensor read by the SGD kernel at run time, so a step captured
    into a CUDA graph follows `trainer.lr = ...` / `cyclic_lr` (reference _base.py:299-309) without re-capture."""

    def __init__(self, model: torch.nn.Module, lr: float, momentum: float, bucket_mb: float = 32.0,
                 process_group=None, accumulate_grad_batches: int = 1, sm_margin: Optional[int] = None):
        self.model = model
        self.group = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self.accumulate = int(accumulate_grad_batches)
        if self.accumulate < 1:
            raise ValueError("accumulate_grad_batches must be >= 1")
        params = [p for p in model.parameters() if p.requires_grad]
        device = params[0].device
        cap = int(bucket_mb * 1024 * 1024 / 4)
        self.buckets: List[_Bucket] = []
        cur, cur_n = [], 0
        for p in reversed(params):
            if cur and cur_n + p.numel() > cap:
                self.buckets.append(_Bucket(cur, device))
                cur, cur_n = [], 0
            cur.append(p)
            cur_n += p.numel()
        if cur:
            # the LAST bucket (the first layers of the model) completes only when backward ends, so its all-reduce is
            # the one that cannot overlap anything: keep it small (tail_mb) by splitting the earliest layers off
            tail_cap = int(float(os.environ.get("UAVDET_DP_TAIL_MB", "2")) * 1024 * 1024 / 4)
            tail, tail_n = [], 0
            while len(cur) > 1 and tail_n + cur[-1].numel() <= tail_cap:
                tail_n += cur[-1].numel()
                tail.insert(0, cur.pop())
            self.buckets.append(_Bucket(cur, device))
            if tail:
                self.buckets.append(_Bucket(tail, device))
        self._bucket_of: Dict[int, _Bucket] = {id(p): b for b in self.buckets for p in b.params}
        self._device = device
        self._hyper = torch.zeros(3, dtype=torch.float32, device=device)
        self._hyper_host = torch.zeros(3, dtype=torch.float32)
        if device.type == "cuda":
            self._hyper_host = self._hyper_host.pin_memory()
        self._lr, self._momentum = float(lr), float(momentum)
        self._hyper_dirty = True
        self.steps_done = 0
        # the all-reduce runs beside backward on a high-priority stream: its CTAs must get an SM as soon as one frees
        self._comm_stream = torch.cuda.Stream(device=device, priority=-1) if (self.world > 1 and device.type == "cuda") else None
        # SMs left to the collective's CTAs while reduces are in flight (the persistent conv kernels fill the rest)
        if sm_margin is None:
            sm_margin = int(os.environ.get("UAVDET_DP_SM_MARGIN", "8"))
        self.sm_margin = sm_margin if (self.world > 1 and device.type == "cuda") else 0
        # ... for this many conv-kernel launches after a bucket's all-reduce was enqueued (< 0: all of backward)
        self.sm_margin_hold = int(os.environ.get("UAVDET_DP_SM_MARGIN_HOLD", "8"))
        self._margin_on = False
        # fused SGD of a bucket right behind its all-reduce on the communication stream (world > 1 only)
        self.early_update = (self.world > 1 and device.type == "cuda"
                             and os.environ.get("UAVDET_DP_EARLY_UPDATE", "1") != "0")
        bump_param_epoch()
        self._execs = [m._exec for m in model.modules() if hasattr(m, "_exec") and hasattr(m, "_forward_program")]
        for ex in self._execs:
            ex.grad_ready_hook = self._on_grad_ready
        self._arm()

    # ---- hyper-parameters ---------------------------------------------------------------------
    @property
    def lr(self) -> float:
        return self._lr

    @lr.setter
    def lr(self, value: float) -> None:
        self._lr = float(value)
        self._hyper_dirty = True

    @property
    def momentum(self) -> float:
        return self._momentum

    @momentum.setter
    def momentum(self, value: float) -> None:
        self._momentum = float(value)
        self._hyper_dirty = True

    def sync_hyper(self) -> None:
        """Publish lr / momentum / gradient scale to the device tensor the SGD kernel reads (stream-ordered copy from
        pinned memory; a no-op when nothing changed).  Called by `step()` and before every graph replay."""
        if not self._hyper_dirty:
            return
        self._hyper_host[0] = self._lr
        self._hyper_host[1] = self._momentum
        self._hyper_host[2] = 1.0 / (self.world * self.accumulate)
        self._hyper.copy_(self._hyper_host, non_blocking=True)
        self._hyper_dirty = False

    @staticmethod
    def cyclic_lr(step: int, base_lr: float, max_lr: float, step_size_up: int = 4000, mode: str = "triangular2") -> float:
        """torch.optim.lr_scheduler.CyclicLR(base_lr, max_lr, step_size_up, mode, cycle_momentum=False) evaluated at
        `step` scheduler steps (reference _base.py:299-309 uses base = lr/10, max = lr, 4000, 'triangular2')."""
        total = 2.0 * step_size_up
        cycle = math.floor(1 + step / total)
        x = 1.0 + step / total - cycle
        scale = x / 0.5 if x <= 0.5 else (x - 1) / (0.5 - 1)
        height = (max_lr - base_lr) * scale
        if mode == "triangular":
            factor = 1.0
        elif mode == "triangular2":
            factor = 1.0 / (2.0 ** (cycle - 1))
        else:
            raise ValueError("cyclic_lr: mode must be 'triangular' or 'triangular2'")
        return base_lr + height * factor

    # ---- gradient lifecycle ------------------------------------------------------------------
    def _arm(self):
        for b in self.buckets:
            b.pending = len(b.params) * self.accumulate
            b.work = None

    def zero_grad(self):
        for b in self.buckets:
            b.grad.zero_()
            for p, o in zip(b.params, b.offsets):      # re-attach if someone set grads to None
                if p.grad is None or p.grad.data_ptr() != b.grad.data_ptr() + 4 * o:
                    p.grad = b._views(p, o)[1]
        self._arm()

    def _on_grad_ready(self, p):
        b = self._bucket_of.get(id(p))
        if b is None:
            return
        if b.work is not None or b.pending <= 0:
            raise RuntimeError("FlatSGDTrainer: a gradient arrived for a bucket that is already being reduced — more "
                               f"backward passes than accumulate_grad_batches={self.accumulate} before step()")
        b.pending -= 1
        if b.pending == 0 and self.world > 1:
            self._launch_reduce(b)

    def _launch_reduce(self, b: _Bucket):
        if self._comm_stream is not None:
            if self.sm_margin:
                # the persistent conv kernels launched while this bucket is on the wire leave SMs to the collective
                ops.set_sm_margin(self.sm_margin, self.sm_margin_hold)
                self._margin_on = True
            # gradients of the bucket were written on the caller's stream and on the executors' weight-gradient
            # side streams: the collective waits for all of them
            self._comm_stream.wait_stream(torch.cuda.current_stream())
            for ex in self._execs:
                for s in ex.producer_streams():
                    self._comm_stream.wait_stream(s)
            with torch.cuda.stream(self._comm_stream):
                b.work = dist.all_reduce(b.grad, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
                if self.early_update:
                    # the bucket's layers are done with backward (their kernels read bf16 packs, not these fp32 master
                    # copies): update it right behind its all-reduce, beside the rest of backward
                    self.sync_hyper()
                    ops.sgd_momentum_dev(b.param, b.grad, b.momentum, self._hyper)
                    b.updated = True
        else:  # gloo / CPU tests
            b.work = dist.all_reduce(b.grad, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def finish_reduce(self):
        """Reduce any bucket whose hooks did not all fire (unused parameters) and join the side stream."""
        if self.world == 1:
            return
        if self._comm_stream is not None:       # a late bucket must see every write of backward
            for ex in self._execs:
                for s in ex.producer_streams():
                    torch.cuda.current_stream().wait_stream(s)
        for b in self.buckets:
            if b.work is None:
                self._launch_reduce(b)
        for b in self.buckets:
            b.work.wait()
        if self._comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self._comm_stream)
        if self._margin_on:
            ops.set_sm_margin(0)
            self._margin_on = False

    # ---- optimiser ----------------------------------------------------------------------------
    def step(self):
        """All-reduce join + fused SGD(momentum) on every bucket (torch.optim.SGD semantics,
        reference _base.py:292-293; gradient averaged over the data-parallel world and the accumulated
        micro-batches).  The momentum buffers start at zero, so the first step's `buf = grad` of torch is the
        ordinary update."""
        self.finish_reduce()
        scale = 1.0 / (self.world * self.accumulate)
        if self.buckets and self.buckets[0].param.is_cuda:
            self.sync_hyper()
        for b in self.buckets:
            if b.updated:           # already stepped on the communication stream, right behind its all-reduce
                b.updated = False
                continue
            if b.param.is_cuda:
                ops.sgd_momentum_dev(b.param, b.grad, b.momentum, self._hyper)
            else:  # host-side logic tests (gloo): same arithmetic in torch
                g = b.grad * scale
                b.momentum.mul_(self._momentum).add_(g)
                b.param.add_(b.momentum, alpha=-self._lr)
        self.steps_done += 1
        if self.steps_done % 64 == 1 and self.buckets[0].param.is_cuda and not torch.cuda.is_current_stream_capturing():
            ops.poll_watchdog()     # a tripped pipeline wait must not train on silently (costs one 4-byte read)
        bump_param_epoch()
        self._arm()


class GraphedTrainStep:
    """One whole training step (zero-grad, forward, loss, backward, bucketed all-reduce, SGD) captured
    into a CUDA graph and replayed: the step is ~2,000 kernel launches of a fixed shape, so launching it
    from Python costs more than executing it.  Inputs live in static device buffers; `__call__` copies
    the new batch in (device->device, or straight from pinned host memory) and replays.

        step = GraphedTrainStep(model, trainer, x_example, targets_example)
        loss = step(x, targets)          # loss: 0-dim device tensor, overwritten by the next call

    The graph holds the per-step activation memory in its private pool.  Capturing needs a few eager
    warm-up steps first (lazy initialisation, cuTensorMap entry point, allocator warm-up); they are run
    here on a side stream and DO update the parameters (they are ordinary training steps on the example
    batch) unless `warmup=0`.

    With `encoder` (a `utils.targets.YoloTargetEncoder`) the step takes the raw `(B, 4)` pixel boxes instead of the
    dense per-head targets: `targets` is then that box tensor everywhere (constructor, `load`, `prefetch`,
    `__call__`) and the dense targets are produced on the device right before the replay."""

    def __init__(self, model, trainer: FlatSGDTrainer, x: torch.Tensor, targets, warmup: int = 2, encoder=None,
                 forward_kwargs: Optional[dict] = None):
        from .utils.datatype import BatchData
        self.model, self.trainer = model, trainer
        self.x = x.detach().clone().float().contiguous()
        self.encoder = encoder
        if encoder is not None:
            self.boxes = targets.detach().clone().float().contiguous()
            self.targets = encoder(self.boxes)
        else:
            self.boxes = None
            self.targets = [t.detach().clone() for t in targets]
        head = model.yolo_head
        for h in range(len(self.targets)):
            head._scaled_anchors(h, self.x.device)      # host->device constants must exist before capture
        if hasattr(model, "prepare_for_capture"):
            model.prepare_for_capture()
        prev_mut = head.mutate_targets
        head.mutate_targets = False            # targets are inputs of the graph, never rewritten in place

        if trainer.accumulate != 1:
            raise NotImplementedError("GraphedTrainStep captures one micro-batch per optimiser step; use the eager "
                                      "FlatSGDTrainer loop for accumulate_grad_batches > 1")
        trainer.sync_hyper()                   # the H2D copy of the hyper-parameters must not be captured
        fkw = dict(forward_kwargs or {})       # e.g. attn_temp for DySOEM_SimFPN.forward(x, attn_temp)

        def body():
            trainer.zero_grad()
            outs = model(self.x, **fkw)
            loss, _, bbox_loss, obj_loss = head.compute_metrics(outs, BatchData(image=self.x, bbox=self.targets))
            loss.backward()
            trainer.step()
            return loss.detach(), bbox_loss.detach(), obj_loss.detach()

        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream(device=self.x.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                body()
        cur.wait_stream(side)
        torch.cuda.synchronize(self.x.device)
        from . import ops as _ops
        self.graph = torch.cuda.CUDAGraph()
        l0 = _ops.launch_count()
        with torch.cuda.graph(self.graph):
            self.loss, self.bbox_loss, self.obj_loss = body()
        self.captured_launches = _ops.launch_count() - l0     # this library's kernels inside one replay
        trainer.steps_done -= 1                                # the capture itself executed nothing
        self.replays = 0
        head.mutate_targets = prev_mut
        bump_param_epoch()
        self._stage = None

    # ---- input pipeline: overlap the next batch's host->device copy with the current step ------------------
    def prefetch(self, x: torch.Tensor, targets) -> None:
        """Start copying the NEXT batch (pinned host or device tensors) into staging buffers on a side stream;
        the following `run_prefetched()` consumes it.  The copy overlaps whatever the main stream is running."""
        if self._stage is None:
            self._stage = (torch.empty_like(self.x),
                           [torch.empty_like(self.boxes)] if self.encoder is not None else [torch.empty_like(t) for t in self.targets])
            self._copy_stream = torch.cuda.Stream(device=self.x.device)
            self._staged = torch.cuda.Event()
            self._consumed = torch.cuda.Event()
            self._consumed.record()
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._consumed)     # the previous staged batch has been taken over
            self._stage[0].copy_(x, non_blocking=True)
            for dst, src in zip(self._stage[1], [targets] if self.encoder is not None else targets):
                dst.copy_(src, non_blocking=True)
            self._staged.record()

    def run_prefetched(self) -> torch.Tensor:
        """Replay the step on the batch staged by `prefetch` (device-to-device hand-over, then replay)."""
        cur = torch.cuda.current_stream()
        cur.wait_event(self._staged)
        self.x.copy_(self._stage[0], non_blocking=True)
        if self.encoder is not None:
            self.boxes.copy_(self._stage[1][0], non_blocking=True)
            self.encoder(self.boxes, check_grid=False, out=self.targets)
        else:
            for dst, src in zip(self.targets, self._stage[1]):
                dst.copy_(src, non_blocking=True)
        self._consumed.record()
        return self()

    def load(self, x: torch.Tensor, targets) -> None:
        """Stage a batch into the static buffers (asynchronous on the current stream)."""
        if x is not self.x:
            self.x.copy_(x, non_blocking=True)
        if self.encoder is not None:
            if targets is not self.boxes:
                self.boxes.copy_(targets, non_blocking=True)
            self.encoder(self.boxes, check_grid=False, out=self.targets)
            return
        for dst, src in zip(self.targets, targets):
            if src is not dst:
                dst.copy_(src, non_blocking=True)

    def __call__(self, x: Optional[torch.Tensor] = None, targets=None) -> torch.Tensor:
        if x is not None:
            self.load(x, targets if targets is not None else (self.boxes if self.encoder is not None else self.targets))
        self.trainer.sync_hyper()     # lr / momentum set since the last replay reach the captured SGD kernels
        self.graph.replay()
        self.trainer.steps_done += 1
        self.replays += 1
        if self.replays % 64 == 1:
            from . import ops as _ops
            _ops.poll_watchdog()
        bump_param_epoch()      # parameters were rewritten on the device: packed-weight caches are stale
        return self.loss
