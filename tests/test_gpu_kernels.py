"""GPU parity tests of the individual kernels, called through the C-ABI (via ops.py) and checked
against the CPU oracle / fp32 torch-on-CPU restatements on identical (bf16-rounded) inputs.

Tolerances: conv outputs are bf16 -> |err| <= 2^-7 * ref_scale style bound, stated
per test; fp32 reductions 1e-3 relative; NMS kept indices bit-exact."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _ops(lib):
    from multimodal_uav_det_b200 import ops
    return ops


def bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


def nhwc(t_nchw):
    return t_nchw.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(DEV)


def to_nchw(t_nhwc):
    return t_nhwc.float().cpu().permute(0, 3, 1, 2).contiguous()


def rel_l2(a, b):
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def assert_close_bf16(got, ref, what, rel=6e-3, frac=2.0 ** -6):
    """bf16 output: relative L2 <= rel and every element within frac*max|ref| (+bf16 ulp of itself)."""
    r = rel_l2(got, ref)
    scale = ref.abs().max().item() + 1e-6
    worst = (got - ref).abs().max().item()
    assert math.isfinite(r) and r <= rel and worst <= frac * scale, \
        f"{what}: rel_l2={r:.3e} (<= {rel}), max_abs={worst:.3e} (<= {frac * scale:.3e})"


# ------------------------------------------------------------------------------------------------
# NMS
# ------------------------------------------------------------------------------------------------
def _random_boxes(n, g, span=100.0, size=40.0, quant=None):
    c = torch.rand(n, 2, generator=g) * span
    wh = torch.rand(n, 2, generator=g) * size
    boxes = torch.cat([c - wh / 2, c + wh / 2], 1)
    scores = torch.randn(n, generator=g)
    if quant:
        scores = (scores * quant).round() / quant
    return boxes, scores


@pytest.mark.parametrize("n,quant,thr", [(1, None, 0.5), (7, None, 0.5), (64, None, 0.5), (65, 4, 0.5),
                                         (1000, None, 0.5), (3000, 20, 0.5), (5000, None, 0.3),
                                         (25200, 20, 0.5), (4097, None, 0.45), (120000, None, 0.5)])
def test_nms_matches_oracle_bit_exact(lib, n, quant, thr):
    from oracle import oracle as O
    ops = _ops(lib)
    g = torch.Generator().manual_seed(n)
    # the largest case is the size class of DySOEM's heads (candidate lists no longer fit shared memory)
    boxes, scores = _random_boxes(n, g, quant=quant, span=100.0 if n < 100000 else 2000.0)
    want = O.nms(boxes.numpy(), scores.numpy(), thr)
    got = ops.nms(boxes.to(DEV), scores.to(DEV), thr).cpu().numpy()
    ops.check_device()
    assert got.dtype == np.int64
    assert np.array_equal(got, want), f"n={n}: kept {len(got)} vs {len(want)}"


@pytest.mark.parametrize("thr", [0.5, 1.0 / 3.0, 0.2, 0.6, 0.25])
@pytest.mark.parametrize("scale", [1.0, 2.0 ** -60, 2.0 ** 50, 2.0 ** -47, 3e18])
def test_nms_threshold_ties(lib, thr, scale):
    """Small-integer boxes: a large share of the pairs has an IoU exactly on (or one rounding away from) the
    threshold, which is where the division-free tests must hand over to the exact quotient.  The scales move the
    areas towards / past the bounds of the fast tests' validity range (subnormal and overflowing areas)."""
    from oracle import oracle as O
    ops = _ops(lib)
    g = torch.Generator().manual_seed(int(thr * 1000) + 7)
    n = 3000
    xy = torch.randint(0, 10, (n, 2), generator=g).float()
    wh = torch.randint(1, 7, (n, 2), generator=g).float()
    boxes = (torch.cat([xy, xy + wh], 1) * scale).float()
    scores = torch.randn(n, generator=g)
    want = O.nms(boxes.numpy(), scores.numpy(), thr)
    got = ops.nms(boxes.to(DEV), scores.to(DEV), thr).cpu().numpy()
    ops.check_device()
    assert np.array_equal(got, want), f"kept {len(got)} vs {len(want)}"


def test_nms_odd_boxes(lib):
    """Boxes that leave the benign class: inverted extents, infinities and NaN coordinates mixed into ordinary ones
    (each variant selects another arithmetic mode of the kernel)."""
    from oracle import oracle as O
    import torchvision
    ops = _ops(lib)
    g = torch.Generator().manual_seed(11)
    base, scores = _random_boxes(4000, g)
    variants = {}
    b = base.clone(); b[::7, [0, 2]] = b[::7, [2, 0]]; variants["inverted"] = b
    b = base.clone(); b[::11, 2] = float("inf"); b[5::13, 0] = float("-inf"); variants["inf"] = b
    b = base.clone(); b[::9, 1] = float("nan"); b[3::17, 2] = float("nan"); variants["nan"] = b
    b = base.clone(); b[::5] *= 1e-30; variants["tiny"] = b
    for name, b in variants.items():
        want = torchvision.ops.nms(b, scores, 0.5).numpy()
        assert np.array_equal(O.nms(b.numpy(), scores.numpy(), 0.5), want), name
        got = ops.nms(b.to(DEV), scores.to(DEV), 0.5).cpu().numpy()
        assert np.array_equal(got, want), f"{name}: kept {len(got)} vs {len(want)}"


def test_nms_edge_cases(lib):
    from oracle import oracle as O
    ops = _ops(lib)
    # empty
    k, c = ops.nms_batched(torch.zeros(2, 0, 4, device=DEV), torch.zeros(2, 0, device=DEV), 0.5)
    assert c.tolist() == [0, 0]
    # degenerate (zero-area -> 0/0 NaN never suppresses), identical boxes, NaN score, +-0 ties, inf
    boxes = torch.tensor([[0, 0, 0, 0], [0, 0, 0, 0], [1, 1, 5, 5], [1, 1, 5, 5], [1, 1, 5, 5.0001],
                          [2, 2, 1, 1], [0, 0, 10, 10], [0, 0, 10, 10]], dtype=torch.float32)
    scores = torch.tensor([0.5, 0.5, float("nan"), 0.0, -0.0, 3.0, float("inf"), float("-inf")])
    want = O.nms(boxes.numpy(), scores.numpy(), 0.5)
    import torchvision
    assert np.array_equal(want, torchvision.ops.nms(boxes, scores, 0.5).numpy())
    got = ops.nms(boxes.to(DEV), scores.to(DEV), 0.5).cpu().numpy()
    assert np.array_equal(got, want)


def test_nms_batched_and_score_floor(lib):
    from oracle import oracle as O
    ops = _ops(lib)
    g = torch.Generator().manual_seed(5)
    bs, n = 5, 2000
    boxes = torch.empty(bs, n, 4)
    scores = torch.empty(bs, n)
    for b in range(bs):
        boxes[b], scores[b] = _random_boxes(n, g, quant=8 if b % 2 else None)
    keep, count = ops.nms_batched(boxes.to(DEV), scores.to(DEV), 0.5)
    for b in range(bs):
        want = O.nms(boxes[b].numpy(), scores[b].numpy(), 0.5)
        assert np.array_equal(keep[b, : int(count[b])].cpu().numpy(), want)
    # extension: score floor == torchvision nms on the filtered subset
    floor = 0.25
    keep, count = ops.nms_batched(boxes.to(DEV), scores.to(DEV), 0.5, score_floor=floor)
    for b in range(bs):
        idx = torch.nonzero(scores[b] > floor).flatten()
        want = idx.numpy()[O.nms(boxes[b][idx].numpy(), scores[b][idx].numpy(), 0.5)]
        assert np.array_equal(keep[b, : int(count[b])].cpu().numpy(), want)


@pytest.mark.parametrize("n", [1, 7, 33, 511, 513, 4100, 96000])
def test_nms_score_floor_compaction_edges(lib, n):
    """The floor path compacts the surviving candidates (index order) before the sort: sizes around the warp-segment
    and tile boundaries, heavy score ties (tie order = index order must survive the compaction), NaN scores (kept, they
    sort first), floors that keep nothing / everything, single image and batch — against the oracle on the filtered
    subset, and (floor below every score) against the full-length path."""
    from oracle import oracle as O
    ops = _ops(lib)
    g = torch.Generator().manual_seed(900 + n)
    bs = 3
    boxes = torch.empty(bs, n, 4)
    scores = torch.empty(bs, n)
    for b in range(bs):
        boxes[b], scores[b] = _random_boxes(n, g, quant=8)
    scores = (scores * 16).round() / 16                     # ~100 distinct values: ties everywhere
    if n >= 33:
        scores[1, torch.randperm(n, generator=g)[: max(1, n // 50)]] = float("nan")
    for floor in (0.5, 0.9375, 2.0, -1.0, 100.0):
        keep, count = ops.nms_batched(boxes.to(DEV), scores.to(DEV), 0.5, score_floor=floor)
        for b in range(bs):
            sel = torch.nonzero(torch.isnan(scores[b]) | (scores[b] > floor)).flatten()
            want = sel.numpy()[O.nms(boxes[b][sel].numpy(), scores[b][sel].numpy(), 0.5)] if sel.numel() else np.zeros(0, np.int64)
            assert np.array_equal(keep[b, : int(count[b])].cpu().numpy(), want), (n, floor, b)
    full, full_count = ops.nms_batched(boxes.to(DEV), scores.to(DEV), 0.5)
    if n >= 33:
        scores[1] = torch.nan_to_num(scores[1], nan=0.25)      # (the full-length path is compared without NaN scores)
        full, full_count = ops.nms_batched(boxes.to(DEV), scores.to(DEV), 0.5)
    keep, count = ops.nms_batched(boxes.to(DEV), scores.to(DEV), 0.5, score_floor=-1e30)
    assert torch.equal(count, full_count)
    for b in range(bs):
        assert torch.equal(keep[b, : int(count[b])], full[b, : int(count[b])])
    ops.check_device()


# ------------------------------------------------------------------------------------------------
# decode
# ------------------------------------------------------------------------------------------------
ANCHORS = [[[199, 73], [315, 92], [268, 182]], [[91, 54], [120, 75], [157, 60]], [[29, 23], [48, 30], [67, 38]]]
SCALES = [32, 16, 8]


@pytest.mark.parametrize("ciou", [True, False])
def test_decode_yolo_matches_oracle(lib, ciou):
    from oracle import oracle as O
    ops = _ops(lib)
    g = torch.Generator().manual_seed(3)
    outs = [(torch.randn(2, 3, s, s, 4, generator=g) * 2, torch.randn(2, 3, s, s, 1, generator=g)) for s in (4, 8, 16)]
    wb, ws = O.decode_yolo(outs, ANCHORS, SCALES, ciou)
    gb, gs = ops.decode_yolo([(b.to(DEV), o.to(DEV)) for b, o in outs], ANCHORS, SCALES, ciou)
    assert torch.equal(gs.cpu(), ws)  # scores are a pure copy
    torch.testing.assert_close(gb.cpu(), wb, rtol=2e-6, atol=2e-6)  # fp32; expf differs by <= 2 ulp


def test_decode_rtm_matches_oracle(lib):
    from oracle import oracle as O
    ops = _ops(lib)
    g = torch.Generator().manual_seed(4)
    t = torch.rand(2, 3, 10, 12, 4, generator=g)
    anc = torch.tensor(ANCHORS[2]).float()
    want = O.decode_rtm(t, anc)
    got = ops.decode_rtm(t.to(DEV), anc)
    torch.testing.assert_close(got.cpu(), want, rtol=1e-6, atol=1e-6)
    assert torch.equal(ops.cxcywh_to_xyxy(want.to(DEV)).cpu(), O.cxcywh_to_xyxy(want))


# ------------------------------------------------------------------------------------------------
# implicit-GEMM convolution (tcgen05)
# ------------------------------------------------------------------------------------------------
def _conv_case(n, cin, cout, k, stride, h, w, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = bf16_round(torch.randn(n, cin, h, w, generator=g))
    wt = bf16_round(torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k))
    return x, wt


FWD_CASES = [
    # n, cin, cout, k, stride, h, w
    (1, 64, 64, 1, 1, 16, 8),       # single tile, BK=64
    (2, 64, 128, 1, 1, 16, 16),
    (2, 64, 64, 3, 1, 16, 16),      # 3x3 halo / zero padding through TMA OOB
    (1, 32, 64, 3, 1, 16, 16),      # BK=32 (SWIZZLE_64B)
    (2, 32, 64, 3, 2, 32, 32),      # stride 2 through the parity view
    (1, 128, 256, 3, 2, 16, 16),
    (2, 256, 128, 1, 1, 20, 20),    # 20x20 map: non power-of-two tile
    (1, 512, 1024, 3, 1, 20, 20),   # 4 n-tiles, K = 4608
    (3, 128, 256, 3, 1, 40, 40),    # many tiles per CTA: pipeline wrap-around + TMEM double buffering
    (2, 768, 256, 1, 1, 40, 40),
    (1, 64, 192, 1, 1, 24, 24),     # block_n = 192
    (1, 64, 32, 1, 1, 64, 64),
    (1, 64, 64, 5, 1, 12, 12),      # 25 taps
]


@pytest.mark.parametrize("case", FWD_CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv_fwd_plain(lib, case):
    ops = _ops(lib)
    n, cin, cout, k, stride, h, w = case
    pad = k // 2
    x, wt = _conv_case(*case)
    ref = F.conv2d(x, wt, None, stride, pad)
    wp = ops.pack_weight(wt.to(DEV))
    y = ops.conv_fwd(nhwc(x), wp, cout, k, stride, pad)
    ops.check_device()
    assert_close_bf16(to_nchw(y), ref, f"conv_fwd{case}")


HALO_CASES = [
    # n, cin, cout, k, stride, h, w — multi-tap layers with <= 64 input channels and an output width that is a multiple
    # of 8 run in halo mode (one TMA box per tile, shifted UMMA descriptors per tap)
    (2, 32, 64, 3, 1, 24, 40),      # rows past the image in the last tile row (24 = 16 + 8), resident weights
    (3, 64, 32, 3, 1, 96, 64),      # more tiles than SMs x buffers: A-buffer ring wraps, SWIZZLE_128B rows
    (2, 64, 128, 3, 1, 48, 48),     # weights streamed through the stage ring beside the halo buffers
    (2, 32, 64, 3, 2, 64, 48),      # stride 2: parity box (both pixels of a pair in one 128-byte row, two row parities)
    (2, 64, 128, 3, 2, 64, 64),     # stride 2, 64 channels: one box per pixel parity
    (1, 32, 32, 5, 1, 32, 32),      # 25 taps
    (1, 64, 64, 5, 2, 64, 64),
]


@pytest.mark.parametrize("case", HALO_CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv_halo_mode_fwd_dgrad(lib, case):
    """Forward (plain, statistics epilogue, affine + activation + residual) and data gradient (+ skip gradient) of the
    thin multi-tap layers against F.conv2d / conv2d_input."""
    ops = _ops(lib)
    from multimodal_uav_det_b200._lib import EPI_STATS
    n, cin, cout, k, stride, h, w = case
    pad = k // 2
    x, wt = _conv_case(*case, seed=77)
    ref = F.conv2d(x, wt, None, stride, pad)
    wp = ops.pack_weight(wt.to(DEV))
    s1 = torch.zeros(cout, device=DEV)
    s2 = torch.zeros(cout, device=DEV)
    y = ops.conv_fwd(nhwc(x), wp, cout, k, stride, pad, epi=EPI_STATS, sum_=s1, sumsq=s2)
    ops.check_device()
    assert_close_bf16(to_nchw(y), ref, f"halo fwd{case}")
    yr = to_nchw(y)
    torch.testing.assert_close(s1.cpu(), yr.sum(dim=(0, 2, 3)), rtol=2e-3, atol=5e-2)
    torch.testing.assert_close(s2.cpu(), (yr * yr).sum(dim=(0, 2, 3)), rtol=2e-3, atol=5e-2)
    g = torch.Generator().manual_seed(78)
    res = bf16_round(torch.randn(ref.shape, generator=g))
    scale = torch.rand(cout, generator=g) + 0.5
    shift = torch.rand(cout, generator=g) - 0.5
    y2 = ops.conv_fwd(nhwc(x), wp, cout, k, stride, pad, act="leaky", scale=scale.to(DEV), shift=shift.to(DEV),
                      res=nhwc(res))
    ref2 = F.leaky_relu(ref * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1), 0.1) + res
    assert_close_bf16(to_nchw(y2), ref2, f"halo fwd affine{case}")
    # data gradient of the same layer: A = dy (cout channels), taps flipped, stride 2 as four parity planes
    if cout <= 64:
        dy = bf16_round(torch.randn(ref.shape, generator=g))
        skip = bf16_round(torch.randn(x.shape, generator=g))
        refd = torch.nn.grad.conv2d_input(x.shape, wt, dy, stride, pad) + skip
        dx = ops.conv_dgrad(nhwc(dy), ops.pack_weight(wt.to(DEV), transposed=True), cin, k, stride, pad, (h, w),
                            res=nhwc(skip))
        ops.check_device()
        assert_close_bf16(to_nchw(dx), refd, f"halo dgrad{case}")


def test_conv_halo_mode_space_to_depth_per_sample(lib):
    """DynamicSOEM at a shape with many tiles: the space-to-depth gather (36 taps over a two-parity box) with one
    kernel per sample, and its data gradient."""
    from oracle import oracle as O
    ops = _ops(lib)
    n, c, h, w, cout = 3, 32, 96, 64, 64
    g = torch.Generator().manual_seed(113)
    x = bf16_round(torch.randn(n, c, h, w, generator=g))
    wts = bf16_round(torch.randn(n, cout, 4 * c, 3, 3, generator=g) / math.sqrt(36 * c))
    xs = O.space_to_depth2(x)
    ref = torch.cat([F.conv2d(xs[i:i + 1], wts[i], None, 1, 1) for i in range(n)])
    wp = torch.stack([ops.pack_weight(wts[i].to(DEV)) for i in range(n)]).contiguous()
    y = ops.conv_fwd(nhwc(x), wp, cout, 3, 1, 1, s2d=True, w_batch=n)
    ops.check_device()
    assert_close_bf16(to_nchw(y), ref, "halo s2d per-sample")


def test_pack_weight_layouts(lib):
    ops = _ops(lib)
    wt = torch.randn(8, 6, 3, 3)
    p = ops.pack_weight(wt.to(DEV)).float().cpu()
    assert torch.equal(p, bf16_round(wt.permute(0, 2, 3, 1).reshape(8, -1)))
    pt = ops.pack_weight(wt.to(DEV), transposed=True).float().cpu()
    assert torch.equal(pt, bf16_round(wt.permute(1, 2, 3, 0).reshape(6, -1)))
    g = torch.randn(8, 3 * 3 * 6)
    back = ops.unpack_wgrad(g.to(DEV), 8, 6, 3).cpu()
    assert torch.equal(back, g.view(8, 3, 3, 6).permute(0, 3, 1, 2))


@pytest.mark.parametrize("act", ["leaky", "silu", "relu", "gelu", "none"])
def test_conv_fwd_affine_epilogue(lib, act):
    ops = _ops(lib)
    n, cin, cout, k, stride, h, w = 2, 64, 128, 3, 1, 16, 16
    x, wt = _conv_case(n, cin, cout, k, stride, h, w, seed=1)
    g = torch.Generator().manual_seed(9)
    scale = torch.rand(cout, generator=g) + 0.5
    shift = torch.randn(cout, generator=g) * 0.1
    res = bf16_round(torch.randn(n, cout, h, w, generator=g))
    z = F.conv2d(x, wt, None, stride, 1) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    a = {"leaky": lambda t: F.leaky_relu(t, 0.1), "silu": F.silu, "relu": F.relu, "gelu": F.gelu, "none": lambda t: t}[act](z)
    ref = a + res
    y = ops.conv_fwd(nhwc(x), ops.pack_weight(wt.to(DEV)), cout, k, stride, 1, act=act, scale=scale.to(DEV),
                     shift=shift.to(DEV), res=nhwc(res))
    ops.check_device()
    assert_close_bf16(to_nchw(y), ref, f"affine[{act}]")
    if act == "gelu":
        # the exact (erf) GELU of nn.GELU() through the A&S 26.2.17 normal CDF: over a wide range of pre-activations
        # the streaming kernel's fp32 math must sit within bf16 rounding (2^-9 relative, 1e-6 absolute) of erf
        zz = torch.linspace(-9, 9, 2 * 8 * 64 * 9).view(2, 8, 9, 64).contiguous()
        got = ops.bn_act_fwd(zz.to(DEV).to(torch.bfloat16), None, None, "gelu").float().cpu()
        want = F.gelu(zz.to(torch.bfloat16).float())
        assert ((got - want).abs() <= want.abs() * 2.0 ** -8 + 1e-6).all()


def _check_bn_sums(s1, s2, raw_nhwc, ref_nchw):
    """The STATS epilogue sums the bf16-rounded values it stores (what the BatchNorm that follows normalises;
    the reference's AMP path also takes its batch statistics from the half-precision conv output): exact
    against the stored tensor up to fp32 summation order, and within bf16 rounding noise of the fp32 conv."""
    r = raw_nhwc.float()
    torch.testing.assert_close(s1, r.sum(dim=(0, 1, 2)), rtol=1e-4, atol=1e-2)
    torch.testing.assert_close(s2, (r * r).sum(dim=(0, 1, 2)), rtol=1e-4, atol=1e-2)
    cnt = ref_nchw.numel() / ref_nchw.shape[1]
    mean_ref = ref_nchw.sum(dim=(0, 2, 3)) / cnt
    var_ref = (ref_nchw * ref_nchw).sum(dim=(0, 2, 3)) / cnt - mean_ref ** 2
    mean = s1.cpu() / cnt
    var = s2.cpu() / cnt - mean ** 2
    torch.testing.assert_close(mean, mean_ref, rtol=0, atol=2e-3 * float(var_ref.sqrt().max()))
    torch.testing.assert_close(var, var_ref, rtol=5e-3, atol=1e-5)


def test_conv_fwd_stats_epilogue_and_channel_slice(lib):
    """STATS epilogue: raw bf16 output + fp32 per-channel sum / sum of squares; output written into
    a channel slice of a wider buffer (route concat, BaselineModel.py:120-122)."""
    ops = _ops(lib)
    n, cin, cout, k, stride, h, w = 2, 128, 256, 3, 1, 20, 20
    x, wt = _conv_case(n, cin, cout, k, stride, h, w, seed=2)
    ref = F.conv2d(x, wt, None, stride, 1)
    buf = torch.full((n, h, w, 384), 7.0, dtype=torch.bfloat16, device=DEV)
    s1 = torch.zeros(cout, device=DEV)
    s2 = torch.zeros(cout, device=DEV)
    from multimodal_uav_det_b200._lib import EPI_STATS
    ops.conv_fwd(nhwc(x), ops.pack_weight(wt.to(DEV)), cout, k, stride, 1, out=buf[..., 128:], epi=EPI_STATS,
                 sum_=s1, sumsq=s2)
    ops.check_device()
    assert_close_bf16(to_nchw(buf[..., 128:]), ref, "stats raw")
    assert torch.all(buf[..., :128].float() == 7.0), "wrote outside the channel slice"
    _check_bn_sums(s1, s2, buf[..., 128:], ref)


@pytest.mark.parametrize("cin,cout,k,hw,n", [(64, 32, 1, 40, 3), (32, 96, 3, 24, 2), (64, 128, 1, 80, 4),
                                             (128, 256, 3, 40, 8), (64, 160, 3, 16, 3), (32, 224, 1, 20, 2),
                                             (64, 64, 3, 10, 5), (32, 32, 3, 4, 2), (64, 512, 1, 20, 9),
                                             (32, 64, 3, 96, 2), (64, 192, 1, 48, 3)])
def test_conv_fwd_stats_shapes(lib, cin, cout, k, hw, n):
    """STATS epilogue over both slab widths (64- and 32-channel TMA-store slabs), ragged tiles and more tiles
    than SMs (persistent loop, both accumulator buffers)."""
    ops = _ops(lib)
    from multimodal_uav_det_b200._lib import EPI_STATS
    x, wt = _conv_case(n, cin, cout, k, 1, hw, hw, seed=40 + cout)
    ref = F.conv2d(x, wt, None, 1, k // 2)
    s1, s2 = torch.zeros(cout, device=DEV), torch.zeros(cout, device=DEV)
    raw = ops.conv_fwd(nhwc(x), ops.pack_weight(wt.to(DEV)), cout, k, 1, k // 2, epi=EPI_STATS, sum_=s1, sumsq=s2)
    ops.check_device()
    assert_close_bf16(to_nchw(raw), ref, "stats raw")
    _check_bn_sums(s1, s2, raw, ref)


@pytest.mark.parametrize("cin,cout,hw,n", [(64, 128, 20, 2), (64, 256, 40, 5), (32, 96, 12, 3), (64, 64, 10, 2),
                                           (32, 32, 80, 3), (64, 64, 33, 2)])
def test_conv_fwd_affine_epilogue_tile_modes(lib, cin, cout, hw, n):
    """AFFINE epilogue (scale, shift, SiLU, residual) over the three output paths: per-warp rectangles, per-warp
    pixel runs with a tail, CTA-wide slabs; ragged right/bottom edges."""
    ops = _ops(lib)
    x, wt = _conv_case(n, cin, cout, 3, 1, hw, hw, seed=70 + hw)
    g = torch.Generator().manual_seed(hw)
    scale = torch.rand(cout, generator=g) + 0.5
    shift = torch.randn(cout, generator=g) * 0.1
    res = bf16_round(torch.randn(n, cout, hw, hw, generator=g))
    z = F.conv2d(x, wt, None, 1, 1) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    ref = F.silu(z) + res
    y = ops.conv_fwd(nhwc(x), ops.pack_weight(wt.to(DEV)), cout, 3, 1, 1, act="silu", scale=scale.to(DEV),
                     shift=shift.to(DEV), res=nhwc(res))
    ops.check_device()
    assert_close_bf16(to_nchw(y), ref, f"affine tile modes {hw}")


def test_conv_fwd_strided_input_view(lib):
    """Input is a channel slice (ld > c), stride-2 parity view must honour the pixel stride."""
    ops = _ops(lib)
    n, cin, cout, h, w = 2, 64, 128, 16, 16
    x, wt = _conv_case(n, cin, cout, 3, 2, h, w, seed=3)
    buf = torch.randn(n, h, w, 192, device=DEV).to(torch.bfloat16)
    buf[..., 64:128] = nhwc(x)
    for stride in (1, 2):
        ref = F.conv2d(x, wt, None, stride, 1)
        y = ops.conv_fwd(buf[..., 64:128], ops.pack_weight(wt.to(DEV)), cout, 3, stride, 1)
        ops.check_device()
        assert_close_bf16(to_nchw(y), ref, f"strided view s{stride}")


def test_conv_head_epilogue(lib):
    """Fused obj+bbox 1x1 head (model/_base.py:80-120) in the final (B,A,H,W,{1,4}) fp32 layout."""
    from oracle import oracle as O
    ops = _ops(lib)
    n, cin, h, w, A = 2, 256, 20, 20, 3
    g = torch.Generator().manual_seed(11)
    x = bf16_round(torch.randn(n, cin, h, w, generator=g))
    sd = {"h.0.obj.conv_obj.weight": bf16_round(torch.randn(A, cin, 1, 1, generator=g) / 16),
          "h.0.obj.conv_obj.bias": torch.randn(A, generator=g),
          "h.0.bbox.conv_bbox.weight": bf16_round(torch.randn(4 * A, cin, 1, 1, generator=g) / 16),
          "h.0.bbox.conv_bbox.bias": torch.randn(4 * A, generator=g)}
    (ref_bbox, ref_obj), = O.yolo_head([x], sd, "h")
    w15 = torch.cat([sd["h.0.obj.conv_obj.weight"], sd["h.0.bbox.conv_bbox.weight"]]).to(DEV)
    b15 = torch.cat([sd["h.0.obj.conv_obj.bias"], sd["h.0.bbox.conv_bbox.bias"]]).to(DEV)
    obj, bbox = ops.conv_head(nhwc(x), ops.pack_weight(w15, rows=16), b15, A)
    ops.check_device()
    torch.testing.assert_close(obj.cpu(), ref_obj, rtol=1e-3, atol=1e-3)   # fp32 accumulate, fp32 out
    torch.testing.assert_close(bbox.cpu(), ref_bbox, rtol=1e-3, atol=1e-3)


def test_conv_fwd_per_sample_weights(lib):
    """Dynamic kernels: one aggregated weight matrix per image (model/_base.py:65-74)."""
    ops = _ops(lib)
    n, cin, cout, k, h, w = 3, 64, 64, 3, 16, 16
    g = torch.Generator().manual_seed(12)
    x = bf16_round(torch.randn(n, cin, h, w, generator=g))
    bank = torch.randn(4, cout, cin, k, k, generator=g)
    attn = torch.softmax(torch.randn(n, 4, generator=g), 1)
    wp, _ = ops.dyn_aggregate(attn.to(DEV), bank.to(DEV))
    filt = (attn @ bank.flatten(1)).view(n, cout, cin, k, k)
    assert rel_l2(wp.float().cpu(), filt.permute(0, 1, 3, 4, 2).reshape(n, cout, -1)) < 4e-3
    ref = torch.cat([F.conv2d(x[i:i + 1], bf16_round(filt[i]), None, 1, 1) for i in range(n)])
    y = ops.conv_fwd(nhwc(x), wp, cout, k, 1, 1, w_batch=n)
    ops.check_device()
    assert_close_bf16(to_nchw(y), ref, "per-sample weights", rel=1e-2)


def test_conv_fwd_space_to_depth(lib):
    """DynamicSOEM gather fused into the loader (DySOEM_SimFPN.py:71-75): conv over s2d(x)."""
    from oracle import oracle as O
    ops = _ops(lib)
    n, c, h, w, cout = 2, 32, 32, 32, 64
    g = torch.Generator().manual_seed(13)
    x = bf16_round(torch.randn(n, c, h, w, generator=g))
    wt = bf16_round(torch.randn(cout, 4 * c, 3, 3, generator=g) / math.sqrt(36 * c))
    ref = F.conv2d(O.space_to_depth2(x), wt, None, 1, 1)
    y = ops.conv_fwd(nhwc(x), ops.pack_weight(wt.to(DEV)), cout, 3, 1, 1, s2d=True)
    ops.check_device()
    assert_close_bf16(to_nchw(y), ref, "s2d conv")


DGRAD_CASES = [(2, 64, 64, 1, 1, 16, 16), (2, 64, 128, 3, 1, 16, 16), (2, 32, 64, 3, 2, 32, 32),
               (1, 128, 256, 3, 2, 20, 20), (2, 64, 128, 1, 2, 16, 16), (1, 256, 512, 3, 1, 20, 20)]


@pytest.mark.parametrize("case", DGRAD_CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv_dgrad(lib, case):
    ops = _ops(lib)
    n, cin, cout, k, stride, h, w = case
    pad = k // 2
    x, wt = _conv_case(*case)
    ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    g = torch.Generator().manual_seed(21)
    dy = bf16_round(torch.randn(n, cout, ho, wo, generator=g))
    skip = bf16_round(torch.randn(n, cin, h, w, generator=g))
    ref = torch.nn.grad.conv2d_input(x.shape, wt, dy, stride, pad) + skip
    dx = ops.conv_dgrad(nhwc(dy), ops.pack_weight(wt.to(DEV), transposed=True), cin, k, stride, pad, (h, w),
                        res=nhwc(skip))
    ops.check_device()
    assert_close_bf16(to_nchw(dx), ref, f"dgrad{case}")


@pytest.mark.parametrize("n,cin,cout,h,w,with_res", [(2, 32, 64, 32, 32, False), (2, 64, 128, 64, 48, True), (3, 32, 64, 80, 96, True),
                                                     (1, 64, 96, 40, 40, False), (2, 32, 32, 16, 24, True)])
def test_conv_dgrad_stride2_plane_fused(lib, n, cin, cout, h, w, with_res):
    """3x3 stride-2 data gradient with the four output-parity planes as one N = 4*cin GEMM over the re-laid-out weight
    matrix (zero blocks where a plane has no tap for a dy shift), against conv2d_input and against the plane-by-plane
    kernel."""
    ops = _ops(lib)
    x, wt = _conv_case(n, cin, cout, 3, 2, h, w, seed=31)
    g = torch.Generator().manual_seed(32)
    dy = bf16_round(torch.randn(n, cout, h // 2, w // 2, generator=g))
    skip = bf16_round(torch.randn(n, cin, h, w, generator=g)) if with_res else None
    ref = torch.nn.grad.conv2d_input(x.shape, wt, dy, 2, 1) + (skip if with_res else 0)
    wtp = ops.pack_weight(wt.to(DEV), transposed=True)
    wf = ops.pack_dgrad_s2_fused(wtp, cin, cout)
    # the re-layout itself: row (ph, pw, ci), column (sh, sw, co)
    wfr = wf.float().cpu().view(2, 2, cin, 2, 2, cout)
    for ph in range(2):
        for pw in range(2):
            for sh in range(2):
                for sw in range(2):
                    kh, kw = ph + 1 - 2 * sh, pw + 1 - 2 * sw
                    want = wt[:, :, kh, kw].t() if 0 <= kh <= 2 and 0 <= kw <= 2 else torch.zeros(cin, cout)
                    assert torch.equal(wfr[ph, pw, :, sh, sw, :], want)
    dx = ops.conv_dgrad_s2_fused(nhwc(dy), wf, cin, res=nhwc(skip) if with_res else None)
    ops.check_device()
    assert_close_bf16(to_nchw(dx), ref, "plane-fused stride-2 dgrad")
    dx_planes = ops.conv_dgrad(nhwc(dy), wtp, cin, 3, 2, 1, (h, w), res=nhwc(skip) if with_res else None)
    assert rel_l2(dx.float().cpu(), dx_planes.float().cpu()) < 3e-3


WGRAD_CASES = [(2, 64, 64, 1, 1, 16, 16), (2, 64, 128, 3, 1, 16, 16), (2, 32, 64, 3, 2, 32, 32),
               (1, 128, 256, 3, 2, 20, 20), (2, 256, 128, 1, 1, 20, 20), (1, 64, 32, 1, 1, 32, 32),
               (4, 512, 1024, 3, 1, 20, 20), (2, 64, 128, 1, 2, 16, 16),
               (3, 32, 64, 3, 1, 40, 40), (2, 64, 64, 3, 1, 8, 8), (2, 32, 32, 3, 1, 4, 4), (2, 128, 128, 3, 1, 24, 24),
               (2, 256, 512, 3, 1, 40, 40), (2, 32, 64, 3, 2, 80, 80), (2, 64, 96, 5, 1, 16, 16), (2, 32, 64, 5, 2, 32, 32),
               (3, 96, 160, 3, 1, 12, 20), (2, 512, 256, 1, 1, 20, 20), (5, 64, 32, 1, 1, 40, 40)]


@pytest.mark.parametrize("case", WGRAD_CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv_wgrad(lib, case):
    ops = _ops(lib)
    n, cin, cout, k, stride, h, w = case
    pad = k // 2
    x, wt = _conv_case(*case)
    ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    g = torch.Generator().manual_seed(22)
    dy = bf16_round(torch.randn(n, cout, ho, wo, generator=g))
    ref = torch.nn.grad.conv2d_weight(x, wt.shape, dy, stride, pad)
    dwp = ops.conv_wgrad(nhwc(x), nhwc(dy), k, stride, pad)
    ops.check_device()
    got = ops.unpack_wgrad(dwp, cout, cin, k).cpu()
    r = rel_l2(got, ref)
    assert r < 2e-3, f"wgrad{case}: rel_l2={r:.3e}"   # fp32 accumulate of bf16 products


def test_conv_wgrad_s2d_and_per_sample(lib):
    from oracle import oracle as O
    ops = _ops(lib)
    n, c, h, w, cout = 2, 32, 32, 32, 64
    g = torch.Generator().manual_seed(23)
    x = bf16_round(torch.randn(n, c, h, w, generator=g))
    dy = bf16_round(torch.randn(n, cout, h // 2, w // 2, generator=g))
    ref = torch.nn.grad.conv2d_weight(O.space_to_depth2(x), (cout, 4 * c, 3, 3), dy, 1, 1)
    got = ops.unpack_wgrad(ops.conv_wgrad(nhwc(x), nhwc(dy), 3, 1, 1, s2d=True), cout, 4 * c, 3).cpu()
    ops.check_device()
    assert rel_l2(got, ref) < 2e-3
    # per-sample gradients (dynamic-kernel contraction input)
    x2 = bf16_round(torch.randn(3, 64, 16, 16, generator=g))
    dy2 = bf16_round(torch.randn(3, 64, 16, 16, generator=g))
    per = ops.conv_wgrad(nhwc(x2), nhwc(dy2), 3, 1, 1, per_sample=True)
    ops.check_device()
    for i in range(3):
        ref_i = torch.nn.grad.conv2d_weight(x2[i:i + 1], (64, 64, 3, 3), dy2[i:i + 1], 1, 1)
        got_i = ops.unpack_wgrad(per[i].contiguous(), 64, 64, 3).cpu()
        assert rel_l2(got_i, ref_i) < 2e-3, f"per-sample {i}"


# ------------------------------------------------------------------------------------------------
# stem + memory-bound kernels
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cin,k,stride,pad", [(3, 3, 1, 1), (3, 1, 1, 0), (1, 1, 1, 0), (3, 5, 2, 1)])
def test_stem_fwd_and_wgrad(lib, cin, k, stride, pad):
    ops = _ops(lib)
    from multimodal_uav_det_b200._lib import EPI_STATS
    g = torch.Generator().manual_seed(31)
    x = torch.rand(2, cin, 40, 36, generator=g)
    wt = torch.randn(32, cin, k, k, generator=g)
    ref = F.conv2d(x, wt, None, stride, pad)
    s1 = torch.zeros(32, device=DEV)
    s2 = torch.zeros(32, device=DEV)
    y = ops.stem_fwd(x.to(DEV), wt.to(DEV), k, stride, pad, epi=EPI_STATS, sum_=s1, sumsq=s2)
    assert_close_bf16(to_nchw(y), ref, "stem raw", rel=4e-3)
    torch.testing.assert_close(s1.cpu(), ref.sum(dim=(0, 2, 3)), rtol=1e-3, atol=1e-2)
    torch.testing.assert_close(s2.cpu(), (ref * ref).sum(dim=(0, 2, 3)), rtol=1e-3, atol=1e-2)
    y2 = ops.stem_fwd(x.to(DEV), wt.to(DEV), k, stride, pad, act="silu", scale=torch.full((32,), 0.5, device=DEV),
                      shift=torch.full((32,), 0.1, device=DEV))
    assert_close_bf16(to_nchw(y2), F.silu(ref * 0.5 + 0.1), "stem affine", rel=4e-3)
    dy = bf16_round(torch.randn(ref.shape, generator=g))
    gw = ops.stem_wgrad(x.to(DEV), nhwc(dy), k, stride, pad).cpu()
    ref_gw = torch.nn.grad.conv2d_weight(x, wt.shape, dy, stride, pad)
    assert rel_l2(gw, ref_gw) < 1e-3


@pytest.mark.parametrize("act", ["leaky", "silu", "relu"])
@pytest.mark.parametrize("c", [32, 64, 192, 1024])
def test_bn_act_train_fwd_bwd(lib, act, c, monkeypatch):
    """Two-phase train-mode BN + activation (+residual) against autograd on the CPU."""
    ops = _ops(lib)
    n, h, w = 2, 10, 12
    g = torch.Generator().manual_seed(41 + c)
    raw = bf16_round(torch.randn(n, c, h, w, generator=g) * 2 + 0.3).requires_grad_(True)
    gamma = (torch.rand(c, generator=g) + 0.5).requires_grad_(True)
    beta = (torch.randn(c, generator=g) * 0.2).requires_grad_(True)
    res = bf16_round(torch.randn(n, c, h, w, generator=g))
    rm, rv = torch.zeros(c), torch.ones(c)
    fn = {"leaky": lambda t: F.leaky_relu(t, 0.1), "silu": F.silu, "relu": F.relu}[act]
    out_ref = fn(F.batch_norm(raw, rm, rv, gamma, beta, True, 0.1, 1e-5)) + res
    dy = bf16_round(torch.randn(n, c, h, w, generator=g))
    out_ref.backward(dy)
    # GPU
    raw_d = nhwc(raw.detach())
    s1 = raw.detach().sum(dim=(0, 2, 3)).to(DEV)
    s2 = (raw.detach() ** 2).sum(dim=(0, 2, 3)).to(DEV)
    rm_d, rv_d = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
    mean, invstd, scale, shift = ops.bn_finalize(s1, s2, n * h * w, 1e-5, 0.1, gamma.detach().to(DEV),
                                                 beta.detach().to(DEV), rm_d, rv_d)
    torch.testing.assert_close(rm_d.cpu(), rm, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(rv_d.cpu(), rv, rtol=1e-4, atol=1e-5)
    y = ops.bn_act_fwd(raw_d, scale, shift, act, res=nhwc(res))
    assert_close_bf16(to_nchw(y), out_ref.detach(), "bn_act_fwd")
    # fused single-pass variant (finalize folded into the streaming kernel): same outputs, same running stats
    # (both settings of the A/B switch: the one-launch kernel and the finalize + apply pair behind the same call)
    for no_fuse in (False, True):
        monkeypatch.setattr(ops, "_NO_BN_FUSE", no_fuse)
        rm_f, rv_f = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
        y_f, mean_f, invstd_f, scale_f, shift_f = ops.bn_train_fwd(raw_d, s1, s2, n * h * w, 1e-5, 0.1, gamma.detach().to(DEV),
                                                                   beta.detach().to(DEV), rm_f, rv_f, act, res=nhwc(res))
        assert_close_bf16(to_nchw(y_f), to_nchw(y), "fused bn fwd vs two-kernel", rel=1e-3)
        for a_, b_ in ((mean_f, mean), (invstd_f, invstd), (scale_f, scale), (shift_f, shift), (rm_f, rm_d), (rv_f, rv_d)):
            torch.testing.assert_close(a_, b_, rtol=1e-5, atol=1e-6)
        y_n, *_ = ops.bn_train_fwd(raw_d, s1, s2, n * h * w, 1e-5, 0.1, gamma.detach().to(DEV), beta.detach().to(DEV),
                                   torch.zeros(c, device=DEV), torch.ones(c, device=DEV), act)
        assert_close_bf16(to_nchw(y_n), (out_ref - res).detach(), "fused bn fwd without residual")
    d_raw, dgamma, dbeta = ops.bn_act_bwd(nhwc(dy), raw_d, scale, shift, mean, invstd, gamma.detach().to(DEV), act)
    assert_close_bf16(to_nchw(d_raw), raw.grad, "bn d_raw", rel=1e-2, frac=2.0 ** -5)
    torch.testing.assert_close(dgamma.cpu(), gamma.grad, rtol=5e-3, atol=5e-2)
    torch.testing.assert_close(dbeta.cpu(), beta.grad, rtol=5e-3, atol=5e-2)


def test_upsample_add_layout_gap(lib):
    ops = _ops(lib)
    g = torch.Generator().manual_seed(51)
    x = bf16_round(torch.randn(2, 64, 6, 5, generator=g))
    buf = torch.zeros(2, 12, 10, 192, dtype=torch.bfloat16, device=DEV)
    ops.upsample2x_fwd(nhwc(x), out=buf[..., :64])
    assert torch.equal(to_nchw(buf[..., :64]), F.interpolate(x, scale_factor=2, mode="nearest"))
    dy = bf16_round(torch.randn(2, 64, 12, 10, generator=g))
    dx = ops.upsample2x_bwd(nhwc(dy))
    ref = dy.view(2, 64, 6, 2, 5, 2).sum(dim=(3, 5))
    assert_close_bf16(to_nchw(dx), ref, "upsample bwd")
    a, b = bf16_round(torch.randn(2, 64, 6, 5, generator=g)), bf16_round(torch.randn(2, 64, 6, 5, generator=g))
    assert_close_bf16(to_nchw(ops.add(nhwc(a), nhwc(b))), a + b, "add")
    assert torch.equal(ops.nhwc_to_nchw_f32(nhwc(a)).cpu(), a)
    assert torch.equal(to_nchw(ops.nchw_f32_to_nhwc(a.to(DEV))), a)
    big = bf16_round(torch.randn(3, 64, 20, 20, generator=g))
    torch.testing.assert_close(ops.gap(nhwc(big)).cpu(), big.mean(dim=(2, 3)), rtol=1e-4, atol=1e-5)
    from oracle import oracle as O
    torch.testing.assert_close(ops.gap(nhwc(big), s2d=True).cpu(), O.space_to_depth2(big).mean(dim=(2, 3)),
                               rtol=1e-4, atol=1e-5)
    img = torch.rand(3, 3, 20, 20, generator=g)
    torch.testing.assert_close(ops.gap_nchw(img.to(DEV)).cpu(), img.mean(dim=(2, 3)), rtol=1e-5, atol=1e-6)


def test_dyn_bias_bwd_matches_matmul_autograd(lib):
    """bias[b] = attn[b] @ bias_bank (DynamicSOEM's per-sample bias, DySOEM_SimFPN.py:56-60): its backward in one kernel
    against autograd of the matmul, including the accumulate-into-d_attn contract and the pooled-sum scale."""
    ops = _ops(lib)
    g = torch.Generator().manual_seed(77)
    for n, K, O in [(5, 3, 64), (64, 4, 256), (2, 8, 1000), (1, 1, 8)]:
        attn = torch.softmax(torch.randn(n, K, generator=g), 1).requires_grad_(True)
        bank = torch.randn(K, O, generator=g).requires_grad_(True)
        gsum = torch.randn(n, O, generator=g)                 # per-sample channel sums of the output gradient
        (attn @ bank).backward(gsum)
        d_attn0 = torch.randn(n, K, generator=g)
        d_attn = d_attn0.clone().to(DEV)
        scale = 400.0
        d_bank = ops.dyn_bias_bwd((gsum / scale).to(DEV), scale, attn.detach().to(DEV), bank.detach().to(DEV), d_attn)
        torch.testing.assert_close(d_bank.cpu(), bank.grad, rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(d_attn.cpu(), d_attn0 + attn.grad, rtol=1e-4, atol=1e-4)
    ops.check_device()


def test_attention_mlp_softmax(lib):
    ops = _ops(lib)
    g = torch.Generator().manual_seed(61)
    pooled = torch.randn(5, 128, generator=g)
    w1, w2, b2 = torch.randn(33, 128, generator=g), torch.randn(4, 33, generator=g), torch.randn(4, generator=g)
    ref = torch.softmax((F.relu(pooled @ w1.t()) @ w2.t() + b2) / 30.0, 1)
    got = ops.attn_mlp_softmax(pooled.to(DEV), w1.to(DEV), None, w2.to(DEV), b2.to(DEV), 30.0)
    torch.testing.assert_close(got.cpu(), ref, rtol=1e-4, atol=1e-6)


def test_sgd_momentum(lib):
    ops = _ops(lib)
    g = torch.Generator().manual_seed(71)
    p = torch.randn(1003, generator=g)
    opt_p = p.clone().requires_grad_(True)
    opt = torch.optim.SGD([opt_p], lr=0.01, momentum=0.7)
    pd, buf = p.clone().to(DEV), torch.zeros(1003, device=DEV)
    for step in range(3):
        gr = torch.randn(1003, generator=g)
        opt_p.grad = gr.clone()
        opt.step()
        ops.sgd_momentum(pd, gr.to(DEV), buf, 0.01, 0.7, first_step=(step == 0))
    torch.testing.assert_close(pd.cpu(), opt_p.detach(), rtol=1e-6, atol=1e-6)


# ------------------------------------------------------------------------------------------------
# DySOEM / RTMUAVDet kernels
# ------------------------------------------------------------------------------------------------
def test_conv_fwd_per_sample_shift_and_stats(lib):
    """Aggregated per-sample expert bias (DySOEM_SimFPN.py:83-91 by linearity) in both epilogues."""
    ops = _ops(lib)
    from multimodal_uav_det_b200._lib import EPI_STATS
    n, cin, cout, h, w = 3, 64, 64, 16, 16
    x, wt = _conv_case(n, cin, cout, 3, 1, h, w, seed=5)
    g = torch.Generator().manual_seed(55)
    bias = torch.randn(n, cout, generator=g)
    ref = F.conv2d(x, wt, None, 1, 1) + bias.view(n, cout, 1, 1)
    wp = ops.pack_weight(wt.to(DEV))
    y = ops.conv_fwd(nhwc(x), wp, cout, 3, 1, 1, act="silu", shift=bias.to(DEV), shift_per_sample=True)
    assert_close_bf16(to_nchw(y), F.silu(ref), "per-sample shift")
    s1, s2 = torch.zeros(cout, device=DEV), torch.zeros(cout, device=DEV)
    raw = ops.conv_fwd(nhwc(x), wp, cout, 3, 1, 1, epi=EPI_STATS, shift=bias.to(DEV), shift_per_sample=True, sum_=s1, sumsq=s2)
    ops.check_device()
    assert_close_bf16(to_nchw(raw), ref, "per-sample shift stats raw")
    _check_bn_sums(s1, s2, raw, ref)


def test_upsample2x_add(lib):
    ops = _ops(lib)
    g = torch.Generator().manual_seed(56)
    a = bf16_round(torch.randn(2, 64, 12, 10, generator=g))
    b = bf16_round(torch.randn(2, 64, 6, 5, generator=g))
    got = ops.upsample2x_add(nhwc(b), nhwc(a), 2.0)
    assert_close_bf16(to_nchw(got), 2 * a + F.interpolate(b, scale_factor=2, mode="nearest"), "upsample2x_add")


@pytest.mark.parametrize("k", [1, 3, 5])
def test_dwdynconv_matches_oracle_mdyconv_core(lib, k):
    """MDyConv's per-sample depthwise conv + residual (RTMUAVDet.py:80-98)."""
    ops = _ops(lib)
    g = torch.Generator().manual_seed(57 + k)
    n, c, h, w = 3, 64, 14, 12
    x = bf16_round(torch.randn(n, c, h, w, generator=g))
    ch_w, k_w = torch.randn(n, c, generator=g), torch.randn(n, k * k, generator=g)
    filt = (k_w.view(n, 1, k, k) * ch_w.view(n, c, 1, 1)).reshape(n * c, 1, k, k)
    ref = F.conv2d(x.reshape(1, n * c, h, w), filt, None, 1, k // 2, groups=n * c).view(n, c, h, w) + x
    buf = torch.zeros(n, h, w, 192, dtype=torch.bfloat16, device=DEV)
    ops.dwdynconv_fwd(nhwc(x), ch_w.to(DEV), k_w.to(DEV), k, k // 2, out=buf[..., 64:128])
    assert_close_bf16(to_nchw(buf[..., 64:128]), ref, f"dwdynconv k={k}")
    assert torch.all(buf[..., :64] == 0) and torch.all(buf[..., 128:] == 0)


def test_linear_groupnorm_bilinear_rtm_head(lib):
    from oracle import oracle as O
    ops = _ops(lib)
    g = torch.Generator().manual_seed(58)
    inp, wgt, b = torch.randn(5, 70, generator=g), torch.randn(16, 70, generator=g), torch.randn(16, generator=g)
    torch.testing.assert_close(ops.linear(inp.to(DEV), wgt.to(DEV), b.to(DEV), "relu").cpu(), F.relu(inp @ wgt.t() + b),
                               rtol=1e-5, atol=1e-5)
    x = bf16_round(torch.randn(3, 96, 10, 12, generator=g) * 2 + 0.5)
    r = bf16_round(torch.randn(3, 96, 10, 12, generator=g))
    gamma, beta = torch.rand(96, generator=g) + 0.5, torch.randn(96, generator=g) * 0.1
    assert_close_bf16(to_nchw(ops.groupnorm1(nhwc(x), gamma.to(DEV), beta.to(DEV), 1e-5)),
                      F.group_norm(x, 1, gamma, beta, 1e-5), "groupnorm")
    assert_close_bf16(to_nchw(ops.groupnorm1(nhwc(x), gamma.to(DEV), beta.to(DEV), 1e-5, b=nhwc(r))),
                      F.group_norm(x + r, 1, gamma, beta, 1e-5), "groupnorm(a+b)")
    assert_close_bf16(to_nchw(ops.bilinear2x_fwd(nhwc(x))), F.interpolate(x, scale_factor=2, mode="bilinear"), "bilinear")
    bl, ol = torch.randn(2, 3, 6, 7, 4, generator=g), torch.randn(2, 3, 6, 7, 1, generator=g)
    anc = torch.tensor([[29.0, 23.0], [48.0, 30.0], [67.0, 38.0]])
    bbox, obj = ops.rtm_head_post(bl.to(DEV), ol.to(DEV), anc)
    torch.testing.assert_close(bbox.cpu(), O.decode_rtm(torch.sigmoid(bl), anc), rtol=2e-6, atol=2e-5)
    torch.testing.assert_close(obj.cpu(), torch.sigmoid(ol), rtol=2e-6, atol=1e-6)


@pytest.mark.parametrize("k", [1, 3, 5])
def test_dwdynconv_res_stats_adds_residual_and_sums(lib, k):
    """The MDyEncoder form of the depthwise dynamic conv (RTMUAVDet.py:163-174): + the encoder's residual, and the per-sample
    sum / sum of squares of the bf16-rounded result (the statistics of the GroupNorm that follows)."""
    ops = _ops(lib)
    g = torch.Generator().manual_seed(157 + k)
    n, c, h, w = 3, 64, 37, 21
    x = bf16_round(torch.randn(n, c, h, w, generator=g))
    r = bf16_round(torch.randn(n, c, h, w, generator=g) * 2 + 0.3)
    ch_w, k_w = torch.randn(n, c, generator=g), torch.randn(n, k * k, generator=g)
    filt = (k_w.view(n, 1, k, k) * ch_w.view(n, c, 1, 1)).reshape(n * c, 1, k, k)
    ref = F.conv2d(x.reshape(1, n * c, h, w), filt, None, 1, k // 2, groups=n * c).view(n, c, h, w) + x + r
    buf = torch.zeros(n, h, w, 192, dtype=torch.bfloat16, device=DEV)
    rbuf = torch.zeros(n, h, w, 128, dtype=torch.bfloat16, device=DEV)
    rbuf[..., 64:] = nhwc(r)
    stats = torch.zeros(n, 2, dtype=torch.float32, device=DEV)
    stats[:, 0] = 5.0                                                  # the kernel accumulates
    ops.dwdynconv_res_stats_fwd(nhwc(x), ch_w.to(DEV), k_w.to(DEV), k, k // 2, rbuf[..., 64:], stats, out=buf[..., 64:128])
    got = to_nchw(buf[..., 64:128])
    assert_close_bf16(got, ref, f"dwdynconv+res k={k}")
    assert torch.all(buf[..., :64] == 0) and torch.all(buf[..., 128:] == 0)
    want = torch.stack([got.double().sum(dim=(1, 2, 3)) + 5.0, (got.double() ** 2).sum(dim=(1, 2, 3))], 1)
    torch.testing.assert_close(stats.cpu().double(), want, rtol=2e-5, atol=1e-2)


def test_groupnorm_folded_into_the_1x1_conv_behind_it(lib):
    """GroupNorm(1 group) -> 1x1 conv (-> eval BatchNorm) -> activation as ONE GEMM on the un-normalised tensor with a
    per-image (rstd, mean * rstd) epilogue (uavdet_groupnorm1_stats / _fold, uavdet_epilogue::sample_affine), against F.group_norm +
    F.conv2d in fp32 (RTMUAVDet.py:165-177)."""
    ops = _ops(lib)
    g = torch.Generator().manual_seed(258)
    n, c, h, w, cout = 3, 96, 18, 22, 64
    x = bf16_round(torch.randn(n, c, h, w, generator=g) * torch.tensor([0.5, 2.0, 1.0]).view(n, 1, 1, 1)
                   + torch.tensor([0.7, -1.5, 0.0]).view(n, 1, 1, 1))
    r = bf16_round(torch.randn(n, c, h, w, generator=g))
    gamma, beta = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.3
    wt, bias = torch.randn(cout, c, 1, 1, generator=g) / c ** 0.5, torch.randn(cout, generator=g) * 0.2
    a = torch.rand(cout, generator=g) + 0.5                          # a folded BatchNorm scale behind the conv
    for res, act, a_vec in ((None, "relu", a), (r, "gelu", None)):
        v = x if res is None else x + r
        z = F.conv2d(F.group_norm(v, 1, gamma, beta, 1e-5), wt, None)
        ref = (F.relu if act == "relu" else F.gelu)((z * a_vec.view(1, -1, 1, 1) if a_vec is not None else z) + bias.view(1, -1, 1, 1))
        rows = a_vec[:, None] if a_vec is not None else 1.0
        wp = (rows * wt.flatten(1) * gamma[None, :]).to(torch.bfloat16).to(DEV)
        wg = wp.float().sum(1)
        wb = wt.flatten(1) @ beta
        b_vec = ((a_vec * wb if a_vec is not None else wb) + bias).to(DEV)
        stats = ops.groupnorm1_stats(nhwc(x), None if res is None else nhwc(r))
        vd = (nhwc(x).float() + (0 if res is None else nhwc(r).float())).to(torch.bfloat16)
        want_stats = torch.stack([v.double().sum(dim=(1, 2, 3)), (v.double() ** 2).sum(dim=(1, 2, 3))], 1)
        torch.testing.assert_close(stats.cpu().double(), want_stats, rtol=2e-5, atol=1e-2)
        sa = ops.groupnorm1_fold(stats, h * w * c, 1e-5)
        mean, var = v.double().mean(dim=(1, 2, 3)), v.double().var(dim=(1, 2, 3), unbiased=False)
        rstd = (var + 1e-5).rsqrt()
        torch.testing.assert_close(sa.cpu().double(), torch.stack([rstd, mean * rstd], 1), rtol=1e-5, atol=1e-5)
        y = ops.conv_fwd(vd, wp, cout, 1, 1, 0, act=act, scale=wg, shift=b_vec, sample_affine=sa)
        assert_close_bf16(to_nchw(y), ref, f"groupnorm fold ({act})", rel=8e-3)


@pytest.mark.parametrize("cin,cout,h,w", [(64, 64, 10, 12), (256, 64, 40, 48), (128, 128, 9, 34)])
def test_conv3x3_pair_matches_conv2d(lib, cin, cout, h, w):
    """3x3 stride-1 pad-1 conv computed two output pixels per GEMM row (N = 2*cout), bias, channel-slice output
    (RTMUAVDet.py:194)."""
    ops = _ops(lib)
    g = torch.Generator().manual_seed(321 + cin + h)
    n = 3
    x = bf16_round(torch.randn(n, cin, h, w, generator=g))
    wt = bf16_round(torch.randn(cout, cin, 3, 3, generator=g) / (3 * cin ** 0.5))
    bias = torch.randn(cout, generator=g) * 0.3
    ref = F.conv2d(x, wt, bias, 1, 1)
    buf = torch.zeros(n, h, w, cout + 64, dtype=torch.bfloat16, device=DEV)
    ops.conv3x3_pair_fwd(nhwc(x), ops.pack_weight_pair(wt.to(DEV)), cout, shift=bias.to(DEV), out=buf[..., 64:])
    assert_close_bf16(to_nchw(buf[..., 64:]), ref, f"conv3x3 pair {cin}->{cout}")
    assert torch.all(buf[..., :64] == 0)
    plain = ops.conv_fwd(nhwc(x), ops.pack_weight(wt.to(DEV)), cout, 3, 1, 1, shift=bias.to(DEV))
    assert_close_bf16(to_nchw(buf[..., 64:]), to_nchw(plain).float(), "pair vs plain igemm", rel=2e-3)


@pytest.mark.parametrize("cin,h,w", [(32, 20, 24), (64, 11, 16)])
def test_conv3x3_pair_stats_epilogue(lib, cin, h, w):
    """The training forward of the thin 32 -> 64 layer (BaselineModel.py:63-75, first residual block) as a pixel-pair GEMM:
    raw output and the per-channel batch statistics (both pixels of a pair add to the same channel)."""
    from multimodal_uav_det_b200._lib import EPI_STATS
    ops = _ops(lib)
    g = torch.Generator().manual_seed(77 + cin)
    n, cout = 3, 64
    x = bf16_round(torch.randn(n, cin, h, w, generator=g))
    wt = bf16_round(torch.randn(cout, cin, 3, 3, generator=g) / (3 * cin ** 0.5))
    ref = F.conv2d(x, wt, None, 1, 1)
    s1 = torch.zeros(cout, device=DEV)
    s2 = torch.zeros(cout, device=DEV)
    packed = ops.pack_weight(wt.to(DEV))
    wp = ops.pack_weight_pair(packed)
    assert torch.equal(wp, ops.pack_weight_pair(wt.to(DEV)))            # from the bf16 pack == from the fp32 weight
    assert torch.equal(ops.pack_weight_pair(packed, out=wp.clone()), wp)
    raw = ops.conv3x3_pair_fwd(nhwc(x), wp, cout, epi=EPI_STATS, sum_=s1, sumsq=s2)
    got = to_nchw(raw)
    assert_close_bf16(got, ref, f"pair stats conv cin={cin}")
    torch.testing.assert_close(s1.cpu(), got.sum(dim=(0, 2, 3)), rtol=1e-4, atol=1e-2)
    torch.testing.assert_close(s2.cpu(), (got * got).sum(dim=(0, 2, 3)), rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("cin,cout,h,w", [(32, 64, 12, 20), (64, 128, 9, 16), (32, 32, 8, 8)])
def test_conv3x3_pair_data_gradient_matches_autograd(lib, cin, cout, h, w):
    """Data gradient of a thin 3x3 stride-1 layer as the pixel-pair convolution of dy with the mirrored, transposed filter
    (autograd of BaselineModel.py:63-75 residual blocks), against torch autograd and the plain igemm data gradient."""
    ops = _ops(lib)
    g = torch.Generator().manual_seed(99 + cin + cout)
    n = 2
    wt = bf16_round(torch.randn(cout, cin, 3, 3, generator=g) / (3 * cin ** 0.5))
    dy = bf16_round(torch.randn(n, cout, h, w, generator=g))
    x = torch.zeros(n, cin, h, w, requires_grad=True)
    F.conv2d(x, wt, None, 1, 1).backward(dy)
    w_t = ops.pack_weight(wt.to(DEV), transposed=True)
    got = ops.conv3x3_pair_fwd(nhwc(dy), ops.pack_weight_pair(w_t, flip=True), cin)
    assert got.shape == (n, h, w, cin)
    assert_close_bf16(to_nchw(got), x.grad, f"pair dgrad {cout}->{cin}")
    plain = ops.conv_dgrad(nhwc(dy), w_t, cin, 3, 1, 1, (h, w))
    assert_close_bf16(to_nchw(got), to_nchw(plain).float(), "pair vs plain dgrad", rel=2e-3)


def test_stem_zero_padded_odd_output(lib):
    """RTM stem 5x5 s2 p1: 39 -> 18.. odd outputs are stored with a zero last row/column."""
    ops = _ops(lib)
    g = torch.Generator().manual_seed(59)
    x = torch.rand(2, 3, 42, 42, generator=g)          # (42+2-5)//2+1 = 20 -> even, no padding
    wt = torch.randn(32, 3, 5, 5, generator=g) * 0.2
    y = ops.stem_fwd(x.to(DEV), wt.to(DEV), 5, 2, 1, act="silu", pad_to_even=True)
    assert y.shape == (2, 20, 20, 32)
    x = torch.rand(2, 3, 40, 40, generator=g)          # (40+2-5)//2+1 = 19 -> padded to 20
    ref = F.silu(F.conv2d(x, wt, None, 2, 1))
    y = ops.stem_fwd(x.to(DEV), wt.to(DEV), 5, 2, 1, act="silu", pad_to_even=True)
    assert y.shape == (2, 20, 20, 32)
    got = to_nchw(y)
    assert_close_bf16(got[:, :, :19, :19], ref, "stem padded interior", rel=4e-3)
    assert torch.all(got[:, :, 19, :] == 0) and torch.all(got[:, :, :, 19] == 0)


# ------------------------------------------------------------------------------------------------
# backward of the dynamic-kernel convolutions
# ------------------------------------------------------------------------------------------------
def _s2d(x):
    return torch.cat([x[..., i::2, j::2] for i in range(2) for j in range(2)], dim=1)


@pytest.mark.parametrize("per_sample", [False, True])
@pytest.mark.parametrize("c,cout,hw", [(32, 64, 32), (64, 128, 16), (32, 32, 24), (32, 64, 96), (128, 256, 16), (64, 64, 40)])
def test_conv_dgrad_s2d_matches_autograd(lib, per_sample, c, cout, hw):
    """Data gradient through the fused space-to-depth conv (DySOEM_SimFPN.py:71-91) incl. the skip-path
    residual and the per-sample pooled-attention shift, against autograd of the materialised formulation."""
    ops = _ops(lib)
    n, k = 3, 3
    g = torch.Generator().manual_seed(90 + c)
    x = bf16_round(torch.randn(n, c, hw, hw, generator=g)).requires_grad_(True)
    wts = bf16_round(torch.randn(n if per_sample else 1, cout, 4 * c, k, k, generator=g) / math.sqrt(36 * c))
    dy = bf16_round(torch.randn(n, cout, hw // 2, hw // 2, generator=g))
    res = bf16_round(torch.randn(n, c, hw, hw, generator=g))
    shift = torch.randn(n, 4 * c, generator=g) * 0.1
    f = _s2d(x)
    y = torch.cat([F.conv2d(f[i:i + 1], wts[i if per_sample else 0], None, 1, 1) for i in range(n)])
    extra = (f * shift.view(n, 4 * c, 1, 1)).sum()          # d/dx = shift broadcast over each parity class
    ((y * dy).sum() + extra).backward()
    ref = x.grad + res
    wt = torch.stack([ops.pack_weight(w.to(DEV), transposed=True) for w in wts])
    got = ops.conv_dgrad_s2d(nhwc(dy), wt if per_sample else wt[0], c, k, 1, w_batch=n if per_sample else 1,
                             res=nhwc(res), shift=shift.to(DEV))
    ops.check_device()
    assert_close_bf16(to_nchw(got), ref, "dgrad s2d")


def test_conv_dgrad_per_sample_shift(lib):
    ops = _ops(lib)
    n, cin, cout, hw = 3, 64, 96, 20
    x, wt = _conv_case(n, cin, cout, 3, 2, hw, hw, seed=17)
    g = torch.Generator().manual_seed(18)
    dy = bf16_round(torch.randn(n, cout, hw // 2, hw // 2, generator=g))
    shift = torch.randn(n, cin, generator=g)
    x = x.requires_grad_(True)
    (F.conv2d(x, wt, None, 2, 1) * dy).sum().backward()
    ref = x.grad + shift.view(n, cin, 1, 1)
    got = ops.conv_dgrad(nhwc(dy), ops.pack_weight(wt.to(DEV), transposed=True), cin, 3, 2, 1, (hw, hw),
                         shift=shift.to(DEV), shift_per_sample=True)
    ops.check_device()
    assert_close_bf16(to_nchw(got), ref, "dgrad per-sample shift")


@pytest.mark.parametrize("packed", [False, True])
@pytest.mark.parametrize("K,O,I,k,n", [(4, 64, 32, 3, 5), (3, 32, 128, 3, 4), (4, 512, 1024, 1, 32), (4, 32, 3, 3, 6)])
def test_dyn_bwd_contract(lib, packed, K, O, I, k, n):
    """d_bank[k] = sum_b a[b,k] dW_b, d_attn[b,k] = <dW_b, bank[k]> (autograd of _base.py:65-66)."""
    ops = _ops(lib)
    g = torch.Generator().manual_seed(K * 100 + k)
    bank = torch.randn(K, O, I, k, k, generator=g)
    attn = torch.softmax(torch.randn(n, K, generator=g), dim=1)
    dwb = torch.randn(n, O, I, k, k, generator=g)                         # OIHW per sample
    want_bank = torch.einsum("bk,boihw->koihw", attn, dwb)
    want_attn = torch.einsum("boihw,koihw->bk", dwb, bank)
    src = dwb.permute(0, 1, 3, 4, 2).contiguous() if packed else dwb      # packed: [O][kh][kw][I]
    d_bank = torch.zeros_like(bank).to(DEV)
    d_attn = torch.zeros(n, K, device=DEV)
    ops.dyn_bwd_contract(src.reshape(n, -1).to(DEV), attn.to(DEV), bank.to(DEV), d_bank, d_attn, packed=packed)
    torch.testing.assert_close(d_bank.cpu(), want_bank, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(d_attn.cpu(), want_attn, rtol=1e-3, atol=1e-2)


def test_pack_weight_layouts_and_batched(lib):
    """bf16 weight pack from OIHW and channels-last fp32 storage, both orientations, single and batched launch."""
    ops = _ops(lib)
    g = torch.Generator().manual_seed(77)
    shapes = [(64, 32, 3), (128, 64, 1), (96, 160, 3), (32, 64, 5), (1024, 512, 3)]
    ws = [torch.randn(o, i, k, k, generator=g).to(DEV) for o, i, k in shapes]
    ws_cl = [w.contiguous(memory_format=torch.channels_last) if w.shape[-1] > 1 else w for w in ws]
    jobs = []
    for w, wcl in zip(ws, ws_cl):
        o, i, k, _ = w.shape
        ref_n = w.permute(0, 2, 3, 1).reshape(o, -1).contiguous().to(torch.bfloat16)
        ref_t = w.permute(1, 2, 3, 0).reshape(i, -1).contiguous().to(torch.bfloat16)
        for src in (w, wcl):
            assert torch.equal(ops.pack_weight(src), ref_n)
            assert torch.equal(ops.pack_weight(src, transposed=True), ref_t)
        jobs.append((wcl, torch.empty_like(ref_n), False, ref_n))
        jobs.append((w, torch.empty_like(ref_t), True, ref_t))
    table, n, chunks = ops.build_pack_table([(w, out, t) for w, out, t, _ in jobs])
    ops.pack_weights_batched(table, n, chunks)
    for _, out, _, ref in jobs:
        assert torch.equal(out, ref)


@pytest.mark.parametrize("cin,k,stride,pad,h,w", [(3, 3, 1, 1, 40, 56), (1, 3, 2, 1, 32, 32), (3, 1, 1, 0, 16, 24),
                                                  (1, 5, 2, 1, 33, 47)])
def test_stem_im2col_then_gemm_equals_conv(lib, cin, k, stride, pad, h, w):
    """im2col of the cin<=3 input (bit-exact vs F.unfold in bf16) and the stem conv as a 1x1 tcgen05 GEMM over the
    32 patch channels, incl. its weight gradient, against F.conv2d / conv2d_weight."""
    ops = _ops(lib)
    n, cout = 3, 32
    g = torch.Generator().manual_seed(300 + k)
    x = torch.rand(n, cin, h, w, generator=g)
    wt = bf16_round(torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k))
    cols = ops.stem_im2col(x.to(DEV), k, stride, pad)
    ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    kk = cin * k * k
    ref_cols = F.unfold(x, k, padding=pad, stride=stride).view(n, kk, ho, wo).permute(0, 2, 3, 1)
    assert torch.equal(cols[..., :kk].float().cpu(), bf16_round(ref_cols))
    assert torch.all(cols[..., kk:] == 0)
    wp = F.pad(wt.flatten(1), (0, 32 - kk)).to(torch.bfloat16).to(DEV).contiguous()
    y = ops.conv_fwd(cols, wp, cout, 1, 1, 0)
    ref = F.conv2d(bf16_round(x), wt, None, stride, pad)
    assert_close_bf16(to_nchw(y), ref, "stem as gemm")
    dy = bf16_round(torch.randn(n, cout, ho, wo, generator=g))
    dwp = ops.conv_wgrad(cols, nhwc(dy), 1, 1, 0)
    ops.check_device()
    ref_dw = torch.nn.grad.conv2d_weight(bf16_round(x), wt.shape, dy, stride, pad)
    assert rel_l2(dwp[:, :kk].reshape(wt.shape).cpu(), ref_dw) < 2e-3


@pytest.mark.parametrize("cin,stride,pad,h,w,n", [(3, 1, 1, 40, 56, 3), (3, 1, 1, 64, 64, 2), (1, 2, 1, 32, 32, 3),
                                                  (2, 1, 0, 21, 19, 2), (3, 2, 1, 33, 47, 5), (3, 1, 1, 128, 160, 6)])
def test_stem_mma_fwd_and_wgrad_equal_conv(lib, cin, stride, pad, h, w, n):
    """The 3x3 cin<=3 stem with its patch rows built in shared memory (stem_mma.cu): raw output + batch sums, the
    affine/activation epilogue, shared and per-sample kernels, output into a channel slice, and the weight gradient
    (shared and per-sample) against F.conv2d / conv2d_weight on the bf16-rounded operands.  Sizes cover tiles that end
    inside an image row, a ragged last tile, stride 2 and more tiles than CTAs."""
    ops = _ops(lib)
    from multimodal_uav_det_b200._lib import EPI_STATS
    k, cout = 3, 32
    assert ops.stem_mma_supported(cin, cout, k) and not ops.stem_mma_supported(cin, 64, k) and not ops.stem_mma_supported(3, 32, 5)
    g = torch.Generator().manual_seed(900 + h + cin)
    x = torch.rand(n, cin, h, w, generator=g) - 0.3
    kk = cin * k * k
    wt = bf16_round(torch.randn(cout, cin, k, k, generator=g) / math.sqrt(kk))
    wp = F.pad(wt.flatten(1), (0, 32 - kk)).to(torch.bfloat16).to(DEV).contiguous()
    xb = bf16_round(x)
    ref = F.conv2d(xb, wt, None, stride, pad)
    s1 = torch.zeros(32, device=DEV)
    s2 = torch.zeros(32, device=DEV)
    y = ops.stem_mma_fwd(x.to(DEV), wp, k, stride, pad, epi=EPI_STATS, sum_=s1, sumsq=s2)
    ops.check_device()
    assert_close_bf16(to_nchw(y), ref, "stem_mma raw")
    yr = to_nchw(y)
    torch.testing.assert_close(s1.cpu(), yr.sum(dim=(0, 2, 3)), rtol=1e-3, atol=2e-2)
    torch.testing.assert_close(s2.cpu(), (yr * yr).sum(dim=(0, 2, 3)), rtol=1e-3, atol=2e-2)
    # affine + activation epilogue, written into a channel slice of a wider buffer (pixel stride 64)
    scale = (torch.rand(32, generator=g) + 0.5).to(DEV)
    shift = (torch.rand(32, generator=g) - 0.5).to(DEV)
    wide = torch.full((n, ref.shape[2], ref.shape[3], 64), 7.0, dtype=torch.bfloat16, device=DEV)
    y2 = ops.stem_mma_fwd(x.to(DEV), wp, k, stride, pad, act="leaky", scale=scale, shift=shift, out=wide[..., 32:])
    ref2 = F.leaky_relu(ref * scale.cpu().view(1, -1, 1, 1) + shift.cpu().view(1, -1, 1, 1), 0.1)
    assert_close_bf16(to_nchw(y2), ref2, "stem_mma affine slice")
    assert torch.all(wide[..., :32] == 7.0)
    y3 = ops.stem_mma_fwd(x.to(DEV), wp, k, stride, pad, act="leaky", scale=scale, shift=shift)
    assert torch.equal(y3, y2)
    # per-sample kernels
    wts = bf16_round(torch.randn(n, cout, cin, k, k, generator=g) / math.sqrt(kk))
    wps = F.pad(wts.flatten(2), (0, 32 - kk)).to(torch.bfloat16).to(DEV).contiguous()
    y4 = ops.stem_mma_fwd(x.to(DEV), wps, k, stride, pad)
    ref4 = torch.cat([F.conv2d(xb[i:i + 1], wts[i], None, stride, pad) for i in range(n)])
    assert_close_bf16(to_nchw(y4), ref4, "stem_mma per-sample kernels")
    # weight gradient
    dy = bf16_round(torch.randn(ref.shape, generator=g))
    dwp = ops.stem_mma_wgrad(x.to(DEV), nhwc(dy), k, stride, pad)
    ops.check_device()
    ref_dw = torch.nn.grad.conv2d_weight(xb, wt.shape, dy, stride, pad)
    assert rel_l2(dwp[:, :kk].reshape(wt.shape).cpu(), ref_dw) < 2e-3
    assert torch.all(dwp[:, kk:] == 0)
    dws = ops.stem_mma_wgrad(x.to(DEV), nhwc(dy), k, stride, pad, per_sample=True)
    ops.check_device()
    for i in range(n):
        ref_i = torch.nn.grad.conv2d_weight(xb[i:i + 1], wt.shape, dy[i:i + 1], stride, pad)
        assert rel_l2(dws[i, :, :kk].reshape(wt.shape).cpu(), ref_i) < 2e-3, f"per-sample wgrad {i}"
    # accumulation into a caller-provided buffer
    acc = dwp.clone()
    ops.stem_mma_wgrad(x.to(DEV), nhwc(dy), k, stride, pad, out=acc)
    assert rel_l2(acc.cpu(), 2 * dwp.cpu()) < 1e-5


@pytest.mark.parametrize("loss_fn", ["ciou", "mse"])
@pytest.mark.parametrize("grid", [20, 37])
def test_fused_yolo_head_loss_matches_batched_torch_loss(lib, loss_fn, grid):
    """csrc/loss.cu (value + analytic gradient, 3 launches) against utils.metrics.yolo_head_loss under torch
    autograd — which tests/test_host.py pins to the reference's per-sample loop and the golden fixture."""
    from oracle import oracle as O
    from multimodal_uav_det_b200.utils import metrics as M
    b, a = 6, 3
    anchors = [[[199, 73], [315, 92], [268, 182]], [[91, 54], [120, 75], [157, 60]], [[29, 23], [48, 30], [67, 38]]]
    g = torch.Generator().manual_seed(500 + grid)
    tg = []
    for i in range(b):
        nbox = 1 + i % 3                                   # several targets in some samples
        cxy = torch.rand(nbox, 2, generator=g) * 400 + 120
        wh = torch.rand(nbox, 2, generator=g) * 60 + 20
        boxes = torch.cat([cxy - wh / 2, cxy + wh / 2], 1)
        per = None
        for k in range(nbox):
            t = O.encode_targets(boxes[k:k + 1], anchors, [32, 16, 8], 640, grids=[grid, grid, grid])[2]
            per = t if per is None else torch.where(t[..., :1] == 1.0, t, per)
        tg.append(per)
    tgt = torch.stack(tg).to(DEV)
    assert tgt.shape == (b, a, grid, grid, 5) and (tgt[..., 0] == 1).sum() >= b
    sa = torch.tensor(anchors[2]).float() / 8
    p_bbox = (torch.randn(b, a, grid, grid, 4, generator=g) * 1.5).to(DEV).requires_grad_(True)
    p_obj = (torch.randn(b, a, grid, grid, 1, generator=g) * 2).to(DEV).requires_grad_(True)
    weights = (4.0, 1.0, 4.0)
    bl_ref, ol_ref, nt_ref = M.yolo_head_loss(p_bbox, p_obj, tgt, sa.to(DEV), 2.0, weights, loss_fn)
    (bl_ref * 0.7 + ol_ref * 1.3).backward()
    gb_ref, go_ref = p_bbox.grad.clone(), p_obj.grad.clone()
    p_bbox.grad = None
    p_obj.grad = None
    bl, ol, nt = M.yolo_head_loss_fused(p_bbox, p_obj, tgt, sa, 2.0, weights, loss_fn)
    (bl * 0.7 + ol * 1.3).backward()
    torch.testing.assert_close(bl, bl_ref, rtol=2e-5, atol=1e-5)
    torch.testing.assert_close(ol, ol_ref, rtol=2e-5, atol=1e-5)
    torch.testing.assert_close(nt, nt_ref, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(p_obj.grad, go_ref, rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(p_bbox.grad, gb_ref, rtol=2e-4, atol=1e-6)
    assert (p_bbox.grad != 0).sum() > 0


# ------------------------------------------------------------------------------------------------
# target encoder
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("grids,anchors", [
    ([20, 40, 80], [[[116, 90], [156, 198], [373, 326]], [[30, 61], [62, 45], [59, 119]], [[10, 13], [16, 30], [33, 23]]]),
    ([320, 160, 80], [[[10, 13], [16, 30], [33, 23]], [[30, 61], [62, 45], [59, 119]], [[116, 90], [156, 198], [373, 326]]]),
])
def test_encode_targets_matches_cpu_encoder_bit_exact(lib, grids, anchors):
    """GPU target encoder vs the oracle's restatement of AntiUAVDataset.__generate_yolo_bboxes (itself pinned to
    the reference encoder by tests/test_oracle.py), bit for bit, on boxes that cover every branch: tiny
    targets (best anchor only), targets matching one or several anchors with IoU >= 0.5, centres on cell edges."""
    from oracle import oracle as O
    ops = _ops(lib)
    g = torch.Generator().manual_seed(3)
    n = 256
    c = torch.rand(n, 2, generator=g) * 600 + 20
    wh = torch.cat([torch.rand(n // 2, 2, generator=g) * 60 + 4, torch.rand(n // 2, 2, generator=g) * 380 + 20])
    boxes = torch.cat([c - wh / 2, c + wh / 2], 1).clamp(0, 639.5)
    boxes[:8, :2] = torch.tensor([32.0, 64.0]); boxes[:8, 2:] = torch.tensor([96.0, 128.0])   # centre exactly on a cell edge
    boxes[8:16] = torch.tensor([100.0, 100.0, 216.0, 190.0])                                   # == an anchor's size
    valid = torch.ones(n, dtype=torch.bool); valid[5::17] = False
    got = ops.encode_targets(boxes.to(DEV), anchors, grids, 640, valid=valid.to(DEV))
    ops.check_device()
    multi = 0
    for i in range(n):
        want = O.encode_targets(boxes[i:i + 1], anchors, None, 640, grids=grids) if valid[i] else \
            [torch.zeros(3, s, s, 5) for s in grids]
        for h in range(len(grids)):
            assert torch.equal(got[h][i].cpu(), want[h]), f"frame {i} head {h}"
            multi += int(want[h][..., 0].sum() > 1)
    assert multi > 0, "no target matched several anchors; the test lost a branch"
    # out-of-grid centre: the reference raises IndexError
    bad = torch.tensor([[630.0, 10.0, 660.0, 40.0]])
    with pytest.raises(IndexError):
        ops.encode_targets(bad.to(DEV), anchors, grids, 640)


# ------------------------------------------------------------------------------------------------
# hardware probe: shifted / non-atom-strided UMMA descriptors into a TMA-written swizzled tile
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("bk", [64, 32])
@pytest.mark.parametrize("shift,sbo_rows", [(0, 8), (1, 8), (3, 8), (8, 8), (0, 10), (1, 10), (11, 10), (22, 10), (0, 9), (4, 12)])
def test_umma_descriptor_row_shift_and_group_stride(lib, bk, shift, sbo_rows):
    """D[m] = A[shift + (m // 8) * sbo_rows + m % 8] @ B^T for a K-major swizzled A tile written by ONE TMA box: a start
    address of whole rows (not whole 8-row atoms) and a group stride that is not a multiple of the atom.  This is what
    a 'halo tile + nine descriptors' 3x3 implicit GEMM needs from the tensor core (csrc/umma_probe.cu)."""
    import ctypes as C
    rows_total = 192
    g = torch.Generator().manual_seed(1000 + shift * 16 + sbo_rows)
    a = bf16_round(torch.randn(rows_total, bk, generator=g))
    b = bf16_round(torch.randn(64, bk, generator=g))
    ad, bd = a.to(DEV).to(torch.bfloat16).contiguous(), b.to(DEV).to(torch.bfloat16).contiguous()
    out = torch.full((128, 64), float("nan"), dtype=torch.float32, device=DEV)
    fn = lib.uavdet_debug_umma_probe
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    rc = fn(ad.data_ptr(), bd.data_ptr(), rows_total, bk, shift, sbo_rows, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.uavdet_last_error()
    torch.cuda.synchronize()
    rows = torch.tensor([shift + (m // 8) * sbo_rows + m % 8 for m in range(128)])
    want = a[rows] @ b.t()
    err = (out.cpu() - want).abs().max().item()
    print(f"bk={bk} shift={shift} sbo_rows={sbo_rows}: max |err| = {err:.3e}")
    assert err < 1e-3
