"""Tensor-level wrappers over the C-ABI (include/uavdet_b200.h).

PyTorch is used only for device memory and streams: every function takes CUDA tensors, passes
raw pointers + the current stream to libuavdet_b200.so and returns tensors it allocated with
torch.empty.  Activations are NHWC bf16 tensors of shape (N,H,W,C) whose last-dim stride is 1
and whose pixel stride (`stride(2)`) may exceed C (channel-slice views of a concat buffer).
No op has a PyTorch/CPU fallback: a missing library or a non-CUDA tensor raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import ACT, EPI_AFFINE, EPI_HEAD, EPI_STATS, Act, Epilogue, UavdetError, check


import os as _os
_NO_BN_FUSE = bool(_os.environ.get("UAVDET_NO_BN_FUSE_FWD"))      # A/B switches for tuning
_NO_BN_FUSE_BWD = bool(_os.environ.get("UAVDET_NO_BN_FUSE_BWD"))


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _require_cuda(*ts: Optional[torch.Tensor]) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise UavdetError("uavdet ops need CUDA tensors (there is no CPU fallback)")


def _f32(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise UavdetError("expected a contiguous float32 tensor")
    return t


def act_view(t: torch.Tensor) -> Act:
    """Describe an NHWC bf16 tensor (possibly a channel slice) to the C-ABI."""
    if t.dtype != torch.bfloat16 or t.dim() != 4:
        raise UavdetError(f"expected a 4-D NHWC bfloat16 tensor, got {t.dtype} {tuple(t.shape)}")
    n, h, w, c = t.shape
    sn, sh, sw, sc = t.stride()
    if c > 1 and sc != 1:
        raise UavdetError("channel stride must be 1")
    ld = sw if w > 1 else (sh if h > 1 else max(c, sw))
    if w > 1 and h > 1 and sh != w * ld or (n > 1 and sn != h * w * ld):
        raise UavdetError(f"NHWC view must be dense in pixels: strides {t.stride()} shape {tuple(t.shape)}")
    return Act(t.data_ptr(), n, h, w, c, ld)


def empty_act(n: int, h: int, w: int, c: int, device) -> torch.Tensor:
    return torch.empty((n, h, w, c), dtype=torch.bfloat16, device=device)


def launch_count() -> int:
    return int(_lib.load().uavdet_launch_count())


def set_sm_margin(margin: int, launches: int = -1) -> int:
    """Reserve `margin` SMs for collectives that overlap the tensor-core kernels, for the next `launches` persistent
    kernel launches (< 0: until reset); returns the previous margin."""
    return int(_lib.load().uavdet_set_sm_margin(int(margin), int(launches)))


def check_device() -> None:
    """Synchronise and raise if a kernel's pipeline watchdog tripped."""
    flag = C.c_int(0)
    check(_lib.load().uavdet_check_device(_stream(), C.byref(flag)), "device watchdog")


def poll_watchdog() -> None:
    """check_device() for training loops: raises UavdetError if a bounded pipeline wait of a tensor-core kernel expired
    since the last poll (the kernel drains instead of hanging; its results are then garbage)."""
    check_device()


def timestamp(slots: torch.Tensor, index: int) -> None:
    """Write the device global timer (ns) into slots[index] (int64 tensor) in stream order."""
    check(_lib.load().uavdet_timestamp(C.c_void_p(slots.data_ptr() + 8 * index), _stream()), "timestamp")


# --------------------------------------------------------------------------------------------
# NMS / decode
# --------------------------------------------------------------------------------------------
def nms_batched(boxes: torch.Tensor, scores: torch.Tensor, iou_threshold: float,
                score_floor: float = float("-inf")) -> Tuple[torch.Tensor, torch.Tensor]:
    """boxes (B,N,4) fp32 xyxy, scores (B,N) fp32 -> keep (B,N) int64 (first count[b] valid), count (B,) int32."""
    _require_cuda(boxes, scores)
    boxes = _f32(boxes)
    scores = _f32(scores)
    b, n = scores.shape
    lib = _lib.load()
    keep = torch.empty((b, n), dtype=torch.int64, device=boxes.device)
    count = torch.zeros((b,), dtype=torch.int32, device=boxes.device)
    ws_bytes = lib.uavdet_nms_workspace_bytes(b, n)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=boxes.device)
    check(lib.uavdet_nms(_ptr(boxes), _ptr(scores), b, n, float(iou_threshold), float(score_floor), _ptr(keep),
                         _ptr(count), _ptr(ws), ws_bytes, _stream()), "nms")
    return keep, count


def nms(boxes: torch.Tensor, scores: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    """Drop-in for torchvision.ops.nms(boxes (N,4), scores (N,), thr) -> int64 kept indices."""
    keep, count = nms_batched(boxes.reshape(1, -1, 4).contiguous(), scores.reshape(1, -1).contiguous(),
                              iou_threshold)
    return keep[0, : int(count[0].item())]


def decode_yolo(outs: Sequence, anchors, head_scales, ciou: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    """outs: list of (bbox (B,A,S,S,4), obj (B,A,S,S,1)) fp32 logits per head ->
    boxes (B, sum A*S*S, 4) xyxy in grid units, scores (B, sum) raw logits (model/_base.py:196-203)."""
    lib = _lib.load()
    bboxes = [_f32(o.bbox if hasattr(o, "bbox") else o[0]) for o in outs]
    objs = [_f32(o.obj if hasattr(o, "obj") else o[1]) for o in outs]
    _require_cuda(*bboxes, *objs)
    b = bboxes[0].shape[0]
    counts = [t.shape[1] * t.shape[2] * t.shape[3] for t in bboxes]
    n_total = sum(counts)
    boxes = torch.empty((b, n_total, 4), dtype=torch.float32, device=bboxes[0].device)
    scores = torch.empty((b, n_total), dtype=torch.float32, device=bboxes[0].device)
    off = 0
    for hi, (tb, to) in enumerate(zip(bboxes, objs)):
        _, a, sh, sw, _ = tb.shape
        # anchors / head_scale evaluated in fp32 exactly like torch.tensor(anchors).float() / scale
        anc = (torch.tensor(anchors[hi]).float() / torch.tensor(head_scales)[hi]).flatten().tolist()
        arr = (C.c_float * len(anc))(*anc)
        check(lib.uavdet_decode_yolo(_ptr(tb), _ptr(to), b, a, sh, sw, arr, 1 if ciou else 0, _ptr(boxes),
                                     _ptr(scores), n_total, off, _stream()), "decode_yolo")
        off += counts[hi]
    return boxes, scores


def encode_targets(boxes_xyxy: torch.Tensor, anchors, grids: Sequence[int], input_size: int = 640,
                   valid: torch.Tensor = None, check_grid: bool = True, out: list = None) -> list:
    """GPU form of AntiUAVDataset.__generate_yolo_bboxes (dataset/AntiUAVDataset.py:141-185) for a batch with one
    target box per frame: boxes_xyxy (B,4) fp32 pixels on the device -> per head (B,A,S,S,5) fp32
    [obj, cx_off, cy_off, w_cells, h_cells], bit-identical to the CPU encoder.  `anchors` are the pixel anchors
    (heads, A, 2) of the model config; `grids` the S of every head; `valid` (B,) bool marks frames that have a
    target.  check_grid=True reads back the out-of-grid counter (one sync) and raises IndexError like the reference;
    pass False inside CUDA graphs / latency-critical loops.  `out`: preallocated per-head tensors to write into."""
    lib = _lib.load()
    boxes_xyxy = _f32(boxes_xyxy)
    _require_cuda(boxes_xyxy)
    b = boxes_xyxy.shape[0]
    assert boxes_xyxy.shape == (b, 4), f"Expected bbox shape (B,4), got {tuple(boxes_xyxy.shape)}"
    anc = torch.tensor(anchors).float() / input_size   # fp32, exactly as the data set normalises them (:27)
    heads, a = anc.shape[0], anc.shape[1]
    assert len(grids) == heads
    dev = boxes_xyxy.device
    outs = out if out is not None else [torch.empty((b, a, s, s, 5), dtype=torch.float32, device=dev) for s in grids]
    for o, s_ in zip(outs, grids):
        assert o.shape == (b, a, s_, s_, 5) and o.dtype == torch.float32 and o.is_contiguous() and o.device == dev
    counter = torch.empty(1, dtype=torch.int32, device=dev)
    flat = anc.flatten().tolist()
    anc_arr = (C.c_float * len(flat))(*flat)
    grid_arr = (C.c_int * heads)(*[int(s) for s in grids])
    ptrs = (C.c_void_p * heads)(*[o.data_ptr() for o in outs])
    v = None
    if valid is not None:
        v = valid.to(device=dev, dtype=torch.uint8).contiguous()
    check(lib.uavdet_encode_targets(_ptr(boxes_xyxy), _ptr(v) if v is not None else None, b, anc_arr, heads, a, grid_arr,
                                    float(input_size), ptrs, _ptr(counter), _stream()), "encode_targets")
    if check_grid and b:
        bad = int(counter.item())
        if bad:
            raise IndexError(f"{bad} target centre(s) outside the grid (the reference encoder raises here too)")
    return outs


def decode_rtm(bbox_sig: torch.Tensor, anchors_head) -> torch.Tensor:
    lib = _lib.load()
    _require_cuda(bbox_sig)
    bbox_sig = _f32(bbox_sig)
    b, a, sh, sw, _ = bbox_sig.shape
    anc = torch.as_tensor(anchors_head).float().flatten().tolist()
    arr = (C.c_float * len(anc))(*anc)
    out = torch.empty_like(bbox_sig)
    check(lib.uavdet_decode_rtm(_ptr(bbox_sig), b, a, sh, sw, arr, _ptr(out), _stream()), "decode_rtm")
    return out


def cxcywh_to_xyxy(t: torch.Tensor) -> torch.Tensor:
    _require_cuda(t)
    t = _f32(t)
    out = torch.empty_like(t)
    check(_lib.load().uavdet_cxcywh_to_xyxy(_ptr(t), _ptr(out), t.numel() // 4, _stream()), "cxcywh_to_xyxy")
    return out


def yolo_head_loss(p_bbox, p_obj, tgt, anchors_scaled, ciou: bool, bbox_w: float, objectness_w: float,
                   obj_scale_w: float, no_obj_w: float, want_new_t: bool):
    """Fused loss + gradient of one head scale.  Returns (out2 [bbox_sum, obj_sum], d_bbox, d_obj, new_t|None)."""
    _require_cuda(p_bbox, p_obj, tgt)
    p_bbox, p_obj, tgt = _f32(p_bbox), _f32(p_obj), _f32(tgt)
    b, a, h, w, _ = p_bbox.shape
    lib = _lib.load()
    dev = p_bbox.device
    d_bbox = torch.empty_like(p_bbox)
    d_obj = torch.empty_like(p_obj)
    new_t = torch.empty_like(p_bbox) if want_new_t else None
    out2 = torch.empty(2, dtype=torch.float32, device=dev)
    ws = torch.empty(lib.uavdet_yolo_head_loss_workspace_bytes(b), dtype=torch.uint8, device=dev)
    anc = [float(v) for v in anchors_scaled]
    arr = (C.c_float * len(anc))(*anc)
    check(lib.uavdet_yolo_head_loss(_ptr(p_bbox), _ptr(p_obj), _ptr(tgt), b, a, h, w, arr, 1 if ciou else 0,
                                    float(bbox_w), float(objectness_w), float(obj_scale_w), float(no_obj_w),
                                    _ptr(d_bbox), _ptr(d_obj), _ptr(new_t), _ptr(ws), _ptr(out2), _stream()),
          "yolo_head_loss")
    return out2, d_bbox, d_obj, new_t


# --------------------------------------------------------------------------------------------
# weights
# --------------------------------------------------------------------------------------------
def _weight_layout(w: torch.Tensor) -> int:
    """0: OIHW-contiguous fp32, 2: channels-last ([O][kh][kw][I]) fp32 — the `flags` bit 1 of uavdet_pack_weight."""
    if w.dtype != torch.float32 or w.dim() != 4:
        raise UavdetError("expected a 4-D float32 conv weight")
    if w.is_contiguous():
        return 0
    if w.permute(0, 2, 3, 1).is_contiguous():
        return 2
    raise UavdetError("conv weight must be OIHW-contiguous or channels-last")


def pack_weight(w: torch.Tensor, transposed: bool = False, rows: Optional[int] = None) -> torch.Tensor:
    """Conv weight fp32 (OIHW or channels-last storage) -> bf16 [O][kh*kw*I] (or [I][kh*kw*O]).
    rows > O zero-pads (head convs)."""
    _require_cuda(w)
    o, i, kh, kw = w.shape
    assert kh == kw
    flags = (1 if transposed else 0) | _weight_layout(w)
    r = i if transposed else o
    kt = kh * kw * (o if transposed else i)
    rows = rows or r
    out = torch.zeros((rows, kt), dtype=torch.bfloat16, device=w.device) if rows != r else \
        torch.empty((rows, kt), dtype=torch.bfloat16, device=w.device)
    check(_lib.load().uavdet_pack_weight(_ptr(w), o, i, kh, flags, _ptr(out), _stream()), "pack_weight")
    return out


class _PackJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("O", C.c_int), ("I", C.c_int), ("k", C.c_int),
                ("flags", C.c_int), ("chunk0", C.c_longlong)]


PACK_CHUNK = 4096


def build_pack_table(jobs):
    """jobs: [(weight fp32 (O,I,k,k), out bf16, transposed)] -> (device table, n_jobs, total_chunks) for
    pack_weights_batched.  Holds raw pointers: rebuild when a tensor is re-allocated."""
    arr = (_PackJob * len(jobs))()
    chunk = 0
    for j, (w, out, transposed) in enumerate(jobs):
        o, i, k, _ = w.shape
        arr[j] = _PackJob(w.data_ptr(), out.data_ptr(), o, i, k, (1 if transposed else 0) | _weight_layout(w), chunk)
        chunk += (o * i * k * k + PACK_CHUNK - 1) // PACK_CHUNK
    table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(jobs[0][0].device)
    return table, len(jobs), chunk


def pack_weights_batched(table, n_jobs: int, total_chunks: int) -> None:
    check(_lib.load().uavdet_pack_weights_batched(_ptr(table), n_jobs, total_chunks, _stream()), "pack_weights_batched")


def unpack_wgrad(dw_packed: torch.Tensor, o: int, i: int, k: int, grad: Optional[torch.Tensor] = None,
                 accumulate: bool = False) -> torch.Tensor:
    if grad is None:
        grad = torch.empty((o, i, k, k), dtype=torch.float32, device=dw_packed.device)
        accumulate = False
    check(_lib.load().uavdet_unpack_wgrad(_ptr(dw_packed), o, i, k, _ptr(grad), 1 if accumulate else 0, _stream()),
          "unpack_wgrad")
    return grad


# --------------------------------------------------------------------------------------------
# convolution
# --------------------------------------------------------------------------------------------
def _epilogue(epi: int = EPI_AFFINE, act=None, scale=None, shift=None, res: Optional[torch.Tensor] = None,
              sum_=None, sumsq=None, head_obj=None, head_bbox=None, head_anchors: int = 0,
              shift_per_sample: bool = False, sample_affine=None) -> Epilogue:
    e = Epilogue()
    e.epi = epi
    e.act = ACT[act] if not isinstance(act, int) else act
    e.scale = scale.data_ptr() if scale is not None else None
    e.shift = shift.data_ptr() if shift is not None else None
    if res is not None:
        rv = act_view(res)
        e.res, e.res_ld = rv.ptr, rv.ld
    else:
        e.res, e.res_ld = None, 0
    e.sum = sum_.data_ptr() if sum_ is not None else None
    e.sumsq = sumsq.data_ptr() if sumsq is not None else None
    e.head_obj = head_obj.data_ptr() if head_obj is not None else None
    e.head_bbox = head_bbox.data_ptr() if head_bbox is not None else None
    e.head_anchors = head_anchors
    e.shift_per_sample = 1 if shift_per_sample else 0
    e.sample_affine = sample_affine.data_ptr() if sample_affine is not None else None
    return e


def conv_out_hw(h: int, w: int, k: int, stride: int, pad: int) -> Tuple[int, int]:
    return (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1


def conv_fwd(x: torch.Tensor, w_packed: torch.Tensor, cout: int, k: int, stride: int, pad: int, *,
             s2d: bool = False, w_batch: int = 1, out: Optional[torch.Tensor] = None, epi: int = EPI_AFFINE,
             act=None, scale=None, shift=None, res=None, sum_=None, sumsq=None,
             shift_per_sample: bool = False, sample_affine=None) -> torch.Tensor:
    """Implicit-GEMM conv.  x NHWC bf16; w_packed from pack_weight / dyn_aggregate.
    shift_per_sample: `shift` is (n, cout) — one bias row per image.
    sample_affine: (n, 2) fp32 from groupnorm1_fold — the GroupNorm-fold epilogue (see the header)."""
    _require_cuda(x, w_packed)
    n, h, w, _ = x.shape
    hin, win = (h // 2, w // 2) if s2d else (h, w)
    ho, wo = conv_out_hw(hin, win, k, stride, pad)
    if out is None:
        out = empty_act(n, ho, wo, cout, x.device)
    xv, yv = act_view(x), act_view(out)
    if sample_affine is not None and (tuple(sample_affine.shape) != (n, 2) or sample_affine.dtype != torch.float32
                                      or not sample_affine.is_contiguous()):
        raise UavdetError(f"sample_affine must be a contiguous ({n}, 2) fp32 tensor")
    e = _epilogue(epi, act, _f32(scale), _f32(shift), res, _f32(sum_), _f32(sumsq), shift_per_sample=shift_per_sample,
                  sample_affine=sample_affine)
    check(_lib.load().uavdet_conv_fwd(C.byref(xv), _ptr(w_packed), w_batch, cout, k, stride, pad, 1 if s2d else 0,
                                      C.byref(yv), C.byref(e), _stream()), "conv_fwd")
    return out


def pair_weight_shifts(cin: int) -> Tuple[int, int]:
    """(S, first): a row of the pair weight matrix covers the S input-column shifts first .. first + S - 1."""
    return (6, -2) if cin == 32 else (4, -1)


def pack_weight_pair(w: torch.Tensor, out: Optional[torch.Tensor] = None, flip: bool = False) -> torch.Tensor:
    """3x3 conv weight -> bf16 [2*cout][3*S*cin] of conv3x3_pair_fwd: row px*cout + co holds the filter shifted by px input
    columns, zero blocks where the shifted filter has no tap.  `w`: (cout, cin, 3, 3) fp32, or the bf16 pack
    [cout][3*3*cin] of pack_weight.  S = 4 column shifts (-1..2) for cin % 64 == 0, 6 (-2..3: whole pixel pairs) for a
    32-channel input.  `out`: a buffer from an earlier call (its zero blocks are kept, only the filter blocks rewritten).
    flip=True mirrors the filter: with the TRANSPOSED pack [cin][3*3*cout] of pack_weight this gives the weights of the data
    gradient, which is the same convolution of dy with the mirrored, transposed filter."""
    if w.dim() == 4:
        cout, cin = w.shape[0], w.shape[1]
        wn = w.detach().permute(0, 2, 3, 1)                               # (cout, ky, kx, cin)
    else:
        cout, cin = w.shape[0], w.shape[1] // 9
        wn = w.view(cout, 3, 3, cin)
    if flip:
        wn = wn.flip(1, 2)
    shifts, first = pair_weight_shifts(cin)
    if out is None:
        out = torch.zeros((2 * cout, 3 * shifts * cin), dtype=torch.bfloat16, device=w.device)
    wp = out.view(2, cout, 3, shifts, cin)
    a = -1 - first                                                        # index of column shift -1
    wp[0, :, :, a:a + 3].copy_(wn)
    wp[1, :, :, a + 1:a + 4].copy_(wn)
    return out


def conv3x3_pair_fwd(x: torch.Tensor, w_pair: torch.Tensor, cout: int, *, act=None, scale=None, shift=None,
                     out: Optional[torch.Tensor] = None, epi: int = EPI_AFFINE, sum_=None, sumsq=None) -> torch.Tensor:
    """3x3 stride-1 pad-1 conv with cout in {64, 128}: two output pixels per GEMM row (see the header).  scale / shift are
    the layer's [cout] vectors (repeated for the two pixels here); epi=EPI_STATS adds the per-channel sum / sum of squares."""
    _require_cuda(x, w_pair)
    n, h, w, cin = x.shape
    if out is None:
        out = empty_act(n, h, w, cout, x.device)
    xv, yv = act_view(x), act_view(out)
    scale2 = None if scale is None else _f32(scale).repeat(2).contiguous()      # alive until the launch below
    shift2 = None if shift is None else _f32(shift).repeat(2).contiguous()
    e = _epilogue(epi, act, scale2, shift2, None, _f32(sum_), _f32(sumsq))
    check(_lib.load().uavdet_conv3x3_pair_fwd(C.byref(xv), _ptr(w_pair), cout, C.byref(yv), C.byref(e), _stream()),
          "conv3x3_pair_fwd")
    return out


def conv_head(x: torch.Tensor, w_packed16: torch.Tensor, bias15: torch.Tensor, anchors: int
              ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Fused objectness+bbox 1x1 head conv (model/_base.py:80-120): one N=16 GEMM writes both
    (B,A,H,W,1) and (B,A,H,W,4) fp32 logits in their final layout."""
    _require_cuda(x, w_packed16, bias15)
    n, h, w, _ = x.shape
    obj = torch.empty((n, anchors, h, w, 1), dtype=torch.float32, device=x.device)
    bbox = torch.empty((n, anchors, h, w, 4), dtype=torch.float32, device=x.device)
    xv = act_view(x)
    e = _epilogue(EPI_HEAD, None, None, _f32(bias15), None, None, None, obj, bbox, anchors)
    check(_lib.load().uavdet_conv_fwd(C.byref(xv), _ptr(w_packed16), 1, 5 * anchors, 1, 1, 0, 0, None, C.byref(e),
                                      _stream()), "conv_head")
    return obj, bbox


def conv_dgrad(dy: torch.Tensor, w_packed_t: torch.Tensor, cin: int, k: int, stride: int, pad: int,
               in_hw: Tuple[int, int], *, w_batch: int = 1, out: Optional[torch.Tensor] = None,
               res: Optional[torch.Tensor] = None, shift: Optional[torch.Tensor] = None,
               shift_per_sample: bool = False) -> torch.Tensor:
    """dx = conv_transpose(dy, w) (+ res) (+ shift[c] or shift[n][c], e.g. the gradient of a global average
    pool broadcast over the pixels)."""
    _require_cuda(dy, w_packed_t)
    n = dy.shape[0]
    if out is None:
        out = empty_act(n, in_hw[0], in_hw[1], cin, dy.device)
    dv, xv = act_view(dy), act_view(out)
    e = _epilogue(EPI_AFFINE, None, None, _f32(shift), res, shift_per_sample=shift_per_sample)
    check(_lib.load().uavdet_conv_dgrad(C.byref(dv), _ptr(w_packed_t), w_batch, cin, k, stride, pad, C.byref(xv),
                                        C.byref(e), _stream()), "conv_dgrad")
    return out


def pack_dgrad_s2_fused(w_packed_t: torch.Tensor, cin: int, cout: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Transposed pack [cin][3*3*cout] of a 3x3 conv -> [4*cin][4*cout] weights of conv_dgrad_s2_fused."""
    _require_cuda(w_packed_t)
    if tuple(w_packed_t.shape) != (cin, 9 * cout) or w_packed_t.dtype != torch.bfloat16 or not w_packed_t.is_contiguous():
        raise UavdetError("pack_dgrad_s2_fused expects the contiguous bf16 transposed pack (cin, 9*cout)")
    if out is None:
        out = torch.empty((4 * cin, 4 * cout), dtype=torch.bfloat16, device=w_packed_t.device)
    check(_lib.load().uavdet_pack_dgrad_s2_fused(_ptr(w_packed_t), cin, cout, _ptr(out), _stream()), "pack_dgrad_s2_fused")
    return out


def conv_dgrad_s2_fused(dy: torch.Tensor, w_fused: torch.Tensor, cin: int, *, out: Optional[torch.Tensor] = None,
                        res: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Data gradient of a 3x3 stride-2 pad-1 conv (cin 32 | 64), all four parity planes in one GEMM."""
    _require_cuda(dy, w_fused)
    n, ho, wo, _ = dy.shape
    if out is None:
        out = empty_act(n, 2 * ho, 2 * wo, cin, dy.device)
    dv, xv = act_view(dy), act_view(out)
    e = _epilogue(EPI_AFFINE, None, None, None, res)
    check(_lib.load().uavdet_conv_dgrad_s2_fused(C.byref(dv), _ptr(w_fused), cin, C.byref(xv), C.byref(e), _stream()),
          "conv_dgrad_s2_fused")
    return out


def conv_dgrad_s2d(dy: torch.Tensor, w_packed_t: torch.Tensor, c: int, k: int, pad: int, *, w_batch: int = 1,
                   out: Optional[torch.Tensor] = None, res: Optional[torch.Tensor] = None,
                   shift: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Data gradient through conv_fwd(s2d=True): dy (n,h/2,w/2,cout) -> dx (n,h,w,c).  `shift` is (n, 4c)."""
    _require_cuda(dy, w_packed_t)
    n, h2, w2, _ = dy.shape
    hin, win = h2 + k - 1 - 2 * pad, w2 + k - 1 - 2 * pad
    if out is None:
        out = empty_act(n, 2 * hin, 2 * win, c, dy.device)
    dv, xv = act_view(dy), act_view(out)
    e = _epilogue(EPI_AFFINE, None, None, _f32(shift), res, shift_per_sample=shift is not None)
    check(_lib.load().uavdet_conv_dgrad_s2d(C.byref(dv), _ptr(w_packed_t), w_batch, c, k, pad, C.byref(xv), C.byref(e),
                                            _stream()), "conv_dgrad_s2d")
    return out


def conv_wgrad(x: torch.Tensor, dy: torch.Tensor, k: int, stride: int, pad: int, *, s2d: bool = False,
               per_sample: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Returns the packed fp32 gradient [cout][k*k*cin] (or [n][cout][k*k*cin])."""
    _require_cuda(x, dy)
    cin = x.shape[3] * (4 if s2d else 1)
    cout = dy.shape[3]
    shape = (x.shape[0], cout, k * k * cin) if per_sample else (cout, k * k * cin)
    if out is None:
        out = torch.zeros(shape, dtype=torch.float32, device=x.device)
    xv, dv = act_view(x), act_view(dy)
    check(_lib.load().uavdet_conv_wgrad(C.byref(xv), C.byref(dv), k, stride, pad, 1 if s2d else 0, _ptr(out),
                                        1 if per_sample else 0, _stream()), "conv_wgrad")
    return out


def stem_fwd(x_nchw: torch.Tensor, w: torch.Tensor, k: int, stride: int, pad: int, *, epi: int = EPI_AFFINE,
             act=None, scale=None, shift=None, sum_=None, sumsq=None, per_sample_w: bool = False,
             pad_to_even: bool = False) -> torch.Tensor:
    _require_cuda(x_nchw, w)
    x_nchw = _f32(x_nchw)
    w = _f32(w)
    n, cin, h, ww = x_nchw.shape
    cout = w.shape[-4]
    ho, wo = conv_out_hw(h, ww, k, stride, pad)
    if pad_to_even and ho % 2 == 1 and wo % 2 == 1:
        ho, wo = ho + 1, wo + 1      # extra last row/column written as zeros (== the next conv's zero padding)
    out = empty_act(n, ho, wo, cout, x_nchw.device)
    yv = act_view(out)
    e = _epilogue(epi, act, _f32(scale), _f32(shift), None, _f32(sum_), _f32(sumsq),
                  head_anchors=-1 if per_sample_w else 0)
    check(_lib.load().uavdet_stem_fwd(_ptr(x_nchw), n, cin, h, ww, _ptr(w), cout, k, stride, pad, C.byref(yv),
                                      C.byref(e), _stream()), "stem_fwd")
    return out


def stem_im2col(x_nchw: torch.Tensor, k: int, stride: int, pad: int) -> torch.Tensor:
    """(n,cin,h,w) fp32 -> (n,ho,wo,32) bf16 patches, channel (ci*k+kh)*k+kw, zero padded (cin*k*k <= 32)."""
    _require_cuda(x_nchw)
    x_nchw = _f32(x_nchw)
    n, cin, h, w = x_nchw.shape
    ho, wo = conv_out_hw(h, w, k, stride, pad)
    out = empty_act(n, ho, wo, 32, x_nchw.device)
    yv = act_view(out)
    check(_lib.load().uavdet_im2col_stem(_ptr(x_nchw), n, cin, h, w, k, stride, pad, C.byref(yv), _stream()),
          "im2col_stem")
    return out


def stem_mma_supported(cin: int, cout: int, k: int) -> bool:
    return bool(_lib.load().uavdet_stem_mma_supported(int(cin), int(cout), int(k)))


def stem_mma_fwd(x_nchw: torch.Tensor, w_o32: torch.Tensor, k: int, stride: int, pad: int, *, epi: int = EPI_AFFINE,
                 act=None, scale=None, shift=None, sum_=None, sumsq=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """3x3 cin<=3 stem on the tensor cores, patch rows built in shared memory (no im2col tensor).
    w_o32: bf16 (32, 32) or per-sample (n, 32, 32): w.flatten(1) zero-padded to 32 columns."""
    _require_cuda(x_nchw, w_o32)
    x_nchw = _f32(x_nchw)
    if w_o32.dtype != torch.bfloat16 or not w_o32.is_contiguous() or tuple(w_o32.shape[-2:]) != (32, 32):
        raise UavdetError("stem_mma_fwd expects contiguous bf16 weights (.., 32, 32)")
    n, cin, h, w = x_nchw.shape
    w_batch = w_o32.shape[0] if w_o32.dim() == 3 else 1
    ho, wo = conv_out_hw(h, w, k, stride, pad)
    if out is None:
        out = empty_act(n, ho, wo, 32, x_nchw.device)
    yv = act_view(out)
    e = _epilogue(epi, act, _f32(scale), _f32(shift), None, _f32(sum_), _f32(sumsq))
    check(_lib.load().uavdet_stem_mma_fwd(_ptr(x_nchw), n, cin, h, w, _ptr(w_o32), w_batch, k, stride, pad, C.byref(yv),
                                          C.byref(e), _stream()), "stem_mma_fwd")
    return out


def stem_mma_wgrad(x_nchw: torch.Tensor, dy: torch.Tensor, k: int, stride: int, pad: int,
                   per_sample: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 (32, 32) [or (n, 32, 32)] gradient of the zero-padded flattened stem kernel; columns < cin*k*k are dW."""
    _require_cuda(x_nchw, dy)
    x_nchw = _f32(x_nchw)
    n, cin, h, w = x_nchw.shape
    if out is None:
        out = torch.zeros((n, 32, 32) if per_sample else (32, 32), dtype=torch.float32, device=dy.device)
    dv = act_view(dy)
    check(_lib.load().uavdet_stem_mma_wgrad(_ptr(x_nchw), n, cin, h, w, C.byref(dv), k, stride, pad, _ptr(out),
                                            1 if per_sample else 0, _stream()), "stem_mma_wgrad")
    return out


def stem_s2d_pack(x_nchw: torch.Tensor) -> torch.Tensor:
    """(n,3,H,W) fp32 -> (n,H/2,W/2,32) bf16 space-to-depth map, channel (py*2+px)*3+ci, channels 12..31 zero."""
    _require_cuda(x_nchw)
    x_nchw = _f32(x_nchw)
    n, cin, h, w = x_nchw.shape
    if cin != 3:
        raise UavdetError("stem_s2d_pack expects a 3-channel input")
    out = empty_act(n, h // 2, w // 2, 32, x_nchw.device)
    yv = act_view(out)
    check(_lib.load().uavdet_stem_s2d_pack(_ptr(x_nchw), n, h, w, C.byref(yv), _stream()), "stem_s2d_pack")
    return out


def s2d_stem_weight(w: torch.Tensor) -> torch.Tensor:
    """(O,3,5,5) weight of a 5x5 stride-2 pad-1 stem -> the equivalent (O,32,3,3) weight over stem_s2d_pack's map:
    filter row kh = 2*tap_y + py - 1 (tap_y in 0..2, py in 0..1), same for columns; rows / columns outside 0..4 and the
    20 padding channels are zero."""
    o = w.shape[0]
    assert tuple(w.shape[1:]) == (3, 5, 5)
    w3 = torch.zeros((o, 32, 3, 3), dtype=torch.float32, device=w.device)
    for ty in range(3):
        for py in range(2):
            kh = 2 * ty + py - 1
            if not 0 <= kh <= 4:
                continue
            for tx in range(3):
                for px in range(2):
                    kw = 2 * tx + px - 1
                    if 0 <= kw <= 4:
                        c0 = (py * 2 + px) * 3
                        w3[:, c0:c0 + 3, ty, tx] = w[:, :, kh, kw]
    return w3


def stem_wgrad(x_nchw: torch.Tensor, dy: torch.Tensor, k: int, stride: int, pad: int,
               per_sample: bool = False) -> torch.Tensor:
    _require_cuda(x_nchw, dy)
    n, cin, h, w = x_nchw.shape
    shape = (n, dy.shape[3], cin, k, k) if per_sample else (dy.shape[3], cin, k, k)
    grad = torch.zeros(shape, dtype=torch.float32, device=dy.device)
    dv = act_view(dy)
    check(_lib.load().uavdet_stem_wgrad(_ptr(_f32(x_nchw)), -n if per_sample else n, cin, h, w, C.byref(dv), k,
                                        stride, pad, _ptr(grad), _stream()), "stem_wgrad")
    return grad


# --------------------------------------------------------------------------------------------
# batch-norm / activation / data movement
# --------------------------------------------------------------------------------------------
def bn_finalize(sum_, sumsq, count: float, eps: float, momentum: float, gamma, beta, running_mean, running_var):
    c = sum_.numel()
    dev = sum_.device
    mean = torch.empty(c, dtype=torch.float32, device=dev)
    invstd = torch.empty_like(mean)
    scale = torch.empty_like(mean)
    shift = torch.empty_like(mean)
    check(_lib.load().uavdet_bn_finalize(_ptr(sum_), _ptr(sumsq), c, float(count), float(eps), float(momentum),
                                         _ptr(gamma), _ptr(beta), _ptr(running_mean), _ptr(running_var), _ptr(mean),
                                         _ptr(invstd), _ptr(scale), _ptr(shift), _stream()), "bn_finalize")
    return mean, invstd, scale, shift


def bn_train_fwd(raw, sum_, sumsq, count: float, eps: float, momentum: float, gamma, beta, running_mean, running_var,
                 act, res=None, out=None):
    """Train-mode BN + activation (+ residual) in one pass; returns (y, mean, invstd, scale, shift)."""
    if _NO_BN_FUSE:
        mean, invstd, scale, shift = bn_finalize(sum_, sumsq, count, eps, momentum, gamma, beta, running_mean, running_var)
        return bn_act_fwd(raw, scale, shift, act, res=res, out=out), mean, invstd, scale, shift
    c = raw.shape[3]
    stats = torch.empty((4, c), dtype=torch.float32, device=raw.device)      # mean, invstd, scale, shift
    if out is None:
        out = torch.empty(raw.shape, dtype=torch.bfloat16, device=raw.device)
    rv, yv = act_view(raw), act_view(out)
    resv = act_view(res) if res is not None else None
    check(_lib.load().uavdet_bn_train_fwd(C.byref(rv), _ptr(sum_), _ptr(sumsq), float(count), float(eps), float(momentum),
                                          _ptr(gamma), _ptr(beta), _ptr(running_mean), _ptr(running_var), _ptr(stats[0]),
                                          _ptr(stats[1]), _ptr(stats[2]), _ptr(stats[3]),
                                          ACT[act] if not isinstance(act, int) else act,
                                          C.byref(resv) if resv is not None else None, C.byref(yv), _stream()),
          "bn_train_fwd")
    return out, stats[0], stats[1], stats[2], stats[3]


def bn_act_fwd(raw, scale, shift, act, res=None, out=None):
    if out is None:
        out = torch.empty(raw.shape, dtype=torch.bfloat16, device=raw.device)
    rv, yv = act_view(raw), act_view(out)
    resv = act_view(res) if res is not None else None
    check(_lib.load().uavdet_bn_act_fwd(C.byref(rv), _ptr(scale), _ptr(shift), ACT[act] if not isinstance(act, int) else act,
                                        C.byref(resv) if resv is not None else None, C.byref(yv), _stream()),
          "bn_act_fwd")
    return out


def bn_act_bwd(dy, raw, scale, shift, mean, invstd, gamma, act, buf=None, grad_gamma=None, grad_beta=None):
    """Train-mode BN(+act) backward.  Returns (d_raw bf16, dgamma fp32, dbeta fp32).
    `gamma` is unused (scale = gamma*invstd already carries it); kept for call-site symmetry.
    grad_gamma / grad_beta: live fp32 gradient buffers (`bn.weight.grad`, `bn.bias.grad`) the kernel ADDS dgamma / dbeta
    into (returned as such), so no separate accumulation pass is needed."""
    c = raw.shape[3]
    a = ACT[act] if not isinstance(act, int) else act
    if buf is None:                                                      # (6, c) zeros: sum_dz, sum_dzr, dgamma, dbeta, k1, k0
        buf = torch.zeros((6, c), dtype=torch.float32, device=raw.device)
    dv, rv = act_view(dy), act_view(raw)
    lib = _lib.load()
    check(lib.uavdet_bn_act_bwd_reduce(C.byref(dv), C.byref(rv), _ptr(scale), _ptr(shift), a, _ptr(buf[0]),
                                       _ptr(buf[1]), _stream()), "bn_act_bwd_reduce")
    count = raw.shape[0] * raw.shape[1] * raw.shape[2]
    d_raw = torch.empty(raw.shape, dtype=torch.bfloat16, device=raw.device)
    ov = act_view(d_raw)
    if _NO_BN_FUSE_BWD:
        check(lib.uavdet_bn_bwd_finalize(_ptr(buf[0]), _ptr(buf[1]), _ptr(mean), _ptr(invstd), _ptr(scale), c,
                                         float(count), _ptr(buf[2]), _ptr(buf[3]), _ptr(buf[4]), _ptr(buf[5]), _stream()),
              "bn_bwd_finalize")
        check(lib.uavdet_bn_act_bwd_apply(C.byref(dv), C.byref(rv), _ptr(scale), _ptr(shift), _ptr(buf[4]), _ptr(buf[5]),
                                          a, C.byref(ov), _stream()), "bn_act_bwd_apply")
        if grad_gamma is not None and grad_beta is not None:
            grad_gamma.add_(buf[2])
            grad_beta.add_(buf[3])
            return d_raw, grad_gamma, grad_beta
        return d_raw, buf[2], buf[3]
    into = grad_gamma is not None and grad_beta is not None
    dg, db = (_f32(grad_gamma), _f32(grad_beta)) if into else (buf[2], buf[3])
    check(lib.uavdet_bn_act_bwd_apply_fused(C.byref(dv), C.byref(rv), _ptr(scale), _ptr(shift), _ptr(buf[0]), _ptr(buf[1]),
                                            _ptr(mean), _ptr(invstd), float(count), a, _ptr(dg), _ptr(db), 1 if into else 0,
                                            C.byref(ov), _stream()), "bn_act_bwd_apply_fused")
    return d_raw, dg, db


def bn_bwd_reduce(dy, raw, scale, shift, act, sum_dz, sum_dzr) -> None:
    """sum_dz[c] += sum_p dz, sum_dzr[c] += sum_p dz*raw with dz = dy * act'(raw*scale + shift) (phase 1 of the
    BatchNorm backward; also the per-channel sums behind the affine gradients of a frozen BatchNorm)."""
    dv, rv = act_view(dy), act_view(raw)
    check(_lib.load().uavdet_bn_act_bwd_reduce(C.byref(dv), C.byref(rv), _ptr(_f32(scale)), _ptr(_f32(shift)),
                                               ACT[act] if not isinstance(act, int) else act, _ptr(sum_dz), _ptr(sum_dzr),
                                               _stream()), "bn_act_bwd_reduce")


def act_bwd(dy, raw, scale, shift, act):
    dx = torch.empty(raw.shape, dtype=torch.bfloat16, device=raw.device)
    dv, rv, ov = act_view(dy), act_view(raw), act_view(dx)
    check(_lib.load().uavdet_act_bwd(C.byref(dv), C.byref(rv), _ptr(scale), _ptr(shift),
                                     ACT[act] if not isinstance(act, int) else act, C.byref(ov), _stream()), "act_bwd")
    return dx


def upsample2x_fwd(x, out=None):
    n, h, w, c = x.shape
    if out is None:
        out = empty_act(n, 2 * h, 2 * w, c, x.device)
    xv, yv = act_view(x), act_view(out)
    check(_lib.load().uavdet_upsample2x_fwd(C.byref(xv), C.byref(yv), _stream()), "upsample2x_fwd")
    return out


def upsample2x_bwd(dy, out=None, accumulate=False):
    n, h2, w2, c = dy.shape
    if out is None:
        out = empty_act(n, h2 // 2, w2 // 2, c, dy.device)
        accumulate = False
    dv, ov = act_view(dy), act_view(out)
    check(_lib.load().uavdet_upsample2x_bwd(C.byref(dv), C.byref(ov), 1 if accumulate else 0, _stream()),
          "upsample2x_bwd")
    return out


def upsample2x_add(b_low, a, a_mult=1.0, out=None):
    """a_mult * a + nearest_upsample2x(b_low)."""
    if out is None:
        out = torch.empty(a.shape, dtype=torch.bfloat16, device=a.device)
    bv, av, ov = act_view(b_low), act_view(a), act_view(out)
    check(_lib.load().uavdet_upsample2x_add(C.byref(bv), C.byref(av), float(a_mult), C.byref(ov), _stream()),
          "upsample2x_add")
    return out


def add(a, b=None, out=None):
    if out is None:
        out = torch.empty(a.shape, dtype=torch.bfloat16, device=a.device)
    av, ov = act_view(a), act_view(out)
    bv = act_view(b) if b is not None else None
    check(_lib.load().uavdet_add(C.byref(av), C.byref(bv) if bv is not None else None, C.byref(ov), _stream()), "add")
    return out


def nhwc_to_nchw_f32(x):
    n, h, w, c = x.shape
    out = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
    xv = act_view(x)
    check(_lib.load().uavdet_nhwc_to_nchw_f32(C.byref(xv), _ptr(out), _stream()), "nhwc_to_nchw")
    return out


def nchw_f32_to_nhwc(x):
    x = _f32(x)
    n, c, h, w = x.shape
    out = empty_act(n, h, w, c, x.device)
    ov = act_view(out)
    check(_lib.load().uavdet_nchw_f32_to_nhwc(_ptr(x), C.byref(ov), _stream()), "nchw_to_nhwc")
    return out


# --------------------------------------------------------------------------------------------
# dynamic-kernel attention
# --------------------------------------------------------------------------------------------
def gap(x, s2d=False):
    n, h, w, c = x.shape
    out = torch.empty((n, (4 if s2d else 1) * c), dtype=torch.float32, device=x.device)
    xv = act_view(x)
    check(_lib.load().uavdet_gap(C.byref(xv), 1 if s2d else 0, _ptr(out), _stream()), "gap")
    return out


def gap_nchw(x):
    x = _f32(x)
    n, c, h, w = x.shape
    out = torch.empty((n, c), dtype=torch.float32, device=x.device)
    check(_lib.load().uavdet_gap_nchw(_ptr(x), n, c, h * w, _ptr(out), _stream()), "gap_nchw")
    return out


def attn_mlp_softmax(pooled, w1, b1, w2, b2, temperature, want_hidden=False):
    n, c = pooled.shape
    hid, k = w1.shape[0], w2.shape[0]
    attn = torch.empty((n, k), dtype=torch.float32, device=pooled.device)
    hidden = torch.empty((n, hid), dtype=torch.float32, device=pooled.device) if want_hidden else None
    check(_lib.load().uavdet_attn_mlp_softmax(_ptr(_f32(pooled)), n, c, _ptr(_f32(w1)), _ptr(b1), hid, _ptr(_f32(w2)),
                                              _ptr(b2), k, float(temperature), _ptr(attn), _ptr(hidden), _stream()),
          "attn_mlp_softmax")
    return (attn, hidden) if want_hidden else attn


def head_grad_pack(d_obj, d_bbox, n, anchors, h, w, bias_obj_grad=None, bias_bbox_grad=None):
    """(B,A,H,W,1) / (B,A,H,W,4) fp32 head-output gradients -> dyh (n,h,w,32) NHWC bf16 [A obj | 4A bbox | 0]; the bias
    gradients are accumulated into the given fp32 buffers."""
    dev = (d_obj if d_obj is not None else d_bbox).device
    dyh = torch.empty((n, h, w, 32), dtype=torch.bfloat16, device=dev)
    dv = act_view(dyh)
    check(_lib.load().uavdet_head_grad_pack(_ptr(_f32(d_obj)), _ptr(_f32(d_bbox)), n, anchors, h, w, C.byref(dv),
                                            _ptr(bias_obj_grad), _ptr(bias_bbox_grad), _stream()), "head_grad_pack")
    return dyh


def attn_mlp_bwd(attn, d_attn, hidden, pooled, w1, w2, temperature, out_scale, dw1, db1, dw2, db2, want_d_pooled=True):
    """Backward of attn_mlp_softmax.  dw1/db1/dw2/db2: fp32 gradient buffers to ACCUMULATE into (db1 may be None).
    Returns d_pooled (n, c) * out_scale, or None."""
    n, c = pooled.shape
    hid, K = w1.shape[0], w2.shape[0]
    dev = pooled.device
    ws = torch.empty(n * (K + hid), dtype=torch.float32, device=dev)
    d_pooled = torch.empty((n, c), dtype=torch.float32, device=dev) if want_d_pooled else None
    check(_lib.load().uavdet_attn_mlp_bwd(_ptr(_f32(attn)), _ptr(_f32(d_attn)), _ptr(_f32(hidden)), _ptr(_f32(pooled)), n, c,
                                          _ptr(_f32(w1)), hid, _ptr(_f32(w2)), K, float(temperature), float(out_scale),
                                          _ptr(ws), _ptr(dw1), _ptr(db1), _ptr(dw2), _ptr(db2), _ptr(d_pooled), _stream()),
          "attn_mlp_bwd")
    return d_pooled


def dyn_aggregate_stem(attn, bank):
    """attn (n,K), bank (K,O,I,k,k) with I*k*k <= 32 -> bf16 (n, O, 32): per-sample OIHW-flat kernels zero-padded to the
    32 im2col channels (the B operand of the stem-as-GEMM path)."""
    n, kk_ = attn.shape
    K, o, i, k, _ = bank.shape
    assert K == kk_ and i * k * k <= 32
    out = torch.empty((n, o, 32), dtype=torch.bfloat16, device=attn.device)
    check(_lib.load().uavdet_dyn_aggregate(_ptr(_f32(attn)), n, K, _ptr(_f32(bank.contiguous())), o, i, k, 2, _ptr(out),
                                           None, None, _stream()), "dyn_aggregate_stem")
    return out


def dyn_aggregate(attn, bank, transposed=False, bias_bank=None):
    """attn (n,K) fp32, bank (K,O,I,k,k) fp32 -> bf16 (n, O, k*k*I) [or (n, I, k*k*O)], bias (n,O)|None."""
    n, kk_ = attn.shape
    K, o, i, k, _ = bank.shape
    assert K == kk_
    rows, kt = (i, k * k * o) if transposed else (o, k * k * i)
    out = torch.empty((n, rows, kt), dtype=torch.bfloat16, device=attn.device)
    bias_out = torch.empty((n, o), dtype=torch.float32, device=attn.device) if bias_bank is not None else None
    check(_lib.load().uavdet_dyn_aggregate(_ptr(_f32(attn)), n, K, _ptr(_f32(bank.contiguous())), o, i, k,
                                           1 if transposed else 0, _ptr(out), _ptr(bias_bank), _ptr(bias_out),
                                           _stream()), "dyn_aggregate")
    return out, bias_out


def dyn_bwd_contract(dwb, attn, bank, d_bank, d_attn, packed: bool):
    """dwb (n, O*I*k*k) per-sample kernel gradients -> d_bank (K,O,I,k,k) += , d_attn (n,K) += ."""
    n, K = attn.shape
    _, o, i, k, _ = bank.shape
    check(_lib.load().uavdet_dyn_bwd_contract(_ptr(_f32(dwb)), n, K, _ptr(_f32(attn)), _ptr(_f32(bank)), o, i, k,
                                              1 if packed else 0, _ptr(_f32(d_bank)), _ptr(_f32(d_attn)), _stream()),
          "dyn_bwd_contract")


def dyn_bias_bwd(pooled_grad, scale: float, attn, bias_bank, d_attn):
    """Backward of bias[b] = attn[b] @ bias_bank: returns d_bias_bank (K, O) and adds into d_attn (n, K) in place;
    `pooled_grad` (n, O) x `scale` = per-sample channel sums of the output gradient."""
    _require_cuda(pooled_grad, attn, bias_bank, d_attn)
    n, K = attn.shape
    O = bias_bank.shape[1]
    if pooled_grad.shape != (n, O) or bias_bank.shape[0] != K or d_attn.shape != (n, K):
        raise UavdetError("dyn_bias_bwd: shape mismatch")
    for t in (pooled_grad, attn, bias_bank, d_attn):
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise UavdetError("dyn_bias_bwd expects contiguous fp32 tensors")
    out = torch.empty((K, O), dtype=torch.float32, device=attn.device)
    check(_lib.load().uavdet_dyn_bias_bwd(_ptr(pooled_grad), float(scale), n, K, O, _ptr(attn), _ptr(bias_bank), _ptr(out),
                                          _ptr(d_attn), _stream()), "dyn_bias_bwd")
    return out


# --------------------------------------------------------------------------------------------
# RTMUAVDet ops
# --------------------------------------------------------------------------------------------
def dwdynconv_fwd(x, channel_w, kernel_w, k, pad, out=None):
    if out is None:
        out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    xv, yv = act_view(x), act_view(out)
    check(_lib.load().uavdet_dwdynconv_fwd(C.byref(xv), _ptr(_f32(channel_w)), _ptr(_f32(kernel_w)), k, pad,
                                           C.byref(yv), _stream()), "dwdynconv_fwd")
    return out


def dwdynconv_res_stats_fwd(x, channel_w, kernel_w, k, pad, res, stats, out=None):
    """out = dwdynconv(x) + res; stats (n, 2) fp32 += per-sample [sum, sum of squares] of the bf16-rounded result."""
    if out is None:
        out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    if stats.dtype != torch.float32 or tuple(stats.shape) != (x.shape[0], 2) or not stats.is_contiguous():
        raise UavdetError("dwdynconv_res_stats_fwd: stats must be a contiguous (n, 2) fp32 tensor")
    xv, rv, yv = act_view(x), act_view(res), act_view(out)
    check(_lib.load().uavdet_dwdynconv_res_stats_fwd(C.byref(xv), _ptr(_f32(channel_w)), _ptr(_f32(kernel_w)), k, pad,
                                                     C.byref(rv), _ptr(stats), C.byref(yv), _stream()),
          "dwdynconv_res_stats_fwd")
    return out


def linear(inp, w, bias=None, act=None):
    rows, c = inp.shape
    o = w.shape[0]
    out = torch.empty((rows, o), dtype=torch.float32, device=inp.device)
    check(_lib.load().uavdet_linear(_ptr(_f32(inp)), rows, c, _ptr(_f32(w)), _ptr(bias), o,
                                    ACT[act] if not isinstance(act, int) else act, _ptr(out), _stream()), "linear")
    return out


def groupnorm1(a, gamma, beta, eps, b=None, out=None):
    if out is None:
        out = torch.empty(a.shape, dtype=torch.bfloat16, device=a.device)
    ws = torch.empty((2 * a.shape[0],), dtype=torch.float32, device=a.device)
    av, yv = act_view(a), act_view(out)
    bv = act_view(b) if b is not None else None
    check(_lib.load().uavdet_groupnorm1(C.byref(av), C.byref(bv) if bv is not None else None, _ptr(_f32(gamma)),
                                        _ptr(_f32(beta)), float(eps), _ptr(ws), C.byref(yv), _stream()), "groupnorm1")
    return out


def groupnorm1_stats(a, b=None):
    """(n, 2) fp32: per-sample sum and sum of squares of (a [+ b]) over h*w*c."""
    stats = torch.empty((a.shape[0], 2), dtype=torch.float32, device=a.device)
    av = act_view(a)
    bv = act_view(b) if b is not None else None
    check(_lib.load().uavdet_groupnorm1_stats(C.byref(av), C.byref(bv) if bv is not None else None, _ptr(stats), _stream()),
          "groupnorm1_stats")
    return stats


def groupnorm1_fold(stats, count, eps):
    """(n, 2) per-sample [sum, sum of squares] over `count` elements -> (n, 2) [rstd, mean * rstd]: the `sample_affine`
    of the 1x1 conv that absorbs GroupNorm(1 group) (see uavdet_epilogue in the header for the algebra)."""
    n = stats.shape[0]
    out = torch.empty((n, 2), dtype=torch.float32, device=stats.device)
    check(_lib.load().uavdet_groupnorm1_fold(_ptr(_f32(stats)), n, float(count), float(eps), _ptr(out), _stream()),
          "groupnorm1_fold")
    return out


def bilinear2x_fwd(x, out=None):
    n, h, w, c = x.shape
    if out is None:
        out = empty_act(n, 2 * h, 2 * w, c, x.device)
    xv, yv = act_view(x), act_view(out)
    check(_lib.load().uavdet_bilinear2x_fwd(C.byref(xv), C.byref(yv), _stream()), "bilinear2x_fwd")
    return out


def rtm_head_post(bbox_logits, obj_logits, anchors_head):
    b, a, sh, sw, _ = bbox_logits.shape
    anc = torch.as_tensor(anchors_head).float().flatten().tolist()
    arr = (C.c_float * len(anc))(*anc)
    bbox = torch.empty_like(bbox_logits)
    obj = torch.empty_like(obj_logits)
    check(_lib.load().uavdet_rtm_head_post(_ptr(_f32(bbox_logits)), _ptr(_f32(obj_logits)), b, a, sh, sw, arr, _ptr(bbox),
                                           _ptr(obj), _stream()), "rtm_head_post")
    return bbox, obj


def sgd_momentum(param, grad, buf, lr, momentum, grad_scale=1.0, first_step=False):
    check(_lib.load().uavdet_sgd_momentum(_ptr(param), _ptr(grad), _ptr(buf), param.numel(), float(lr),
                                          float(momentum), float(grad_scale), 1 if first_step else 0, _stream()),
          "sgd_momentum")


def sgd_momentum_dev(param, grad, buf, hyper, first_step=False):
    """SGD(momentum) step with {lr, momentum, grad_scale} read from the 3-float device tensor `hyper` at run time."""
    check(_lib.load().uavdet_sgd_momentum_dev(_ptr(param), _ptr(grad), _ptr(buf), param.numel(), _ptr(_f32(hyper)),
                                              1 if first_step else 0, _stream()), "sgd_momentum_dev")
