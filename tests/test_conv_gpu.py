"""GPU parity of the tensor-core convolution (premvos_conv2d_forward) against a plain PyTorch fp32/fp64
CPU convolution of the same op.  Tolerance: the split-bf16 x3 scheme keeps ~16 mantissa bits per operand,
so 1e-4 relative (||d||_inf / ||ref||_inf) is required here -- ten times tighter than the 1e-3 contract."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err
from premvos_b200 import _lib, ops

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _ref(x, w, b, stride, dil, pad, slope, res):
    pt, pl, pb, pr = pad
    xp = F.pad(x.double(), (pl, pr, pt, pb))
    y = F.conv2d(xp, w.double(), None if b is None else b.double(), stride=stride, dilation=dil)
    if res is not None:
        y = y + res.double()
    return torch.where(y > 0, y, y * slope).float()


CASES = [
    # (N, Cin, H, W, Cout, k, stride, dil, pad(t,l,b,r), slope, residual)   -- what each case exercises
    (1, 16, 16, 8, 16, 3, 1, 1, (1, 1, 1, 1), 0.1, False),      # one exact tile, halo mode
    (2, 40, 37, 29, 72, 3, 1, 1, (1, 1, 1, 1), 0.1, False),     # ragged tiles, K tail (40 = 5 chunks), N tail
    (1, 565, 28, 64, 128, 3, 1, 1, (1, 1, 1, 1), 0.1, False),   # PWC dc_conv1 shape (K = 565)
    (1, 128, 40, 48, 128, 3, 1, 2, (2, 2, 2, 2), 0.1, False),   # dilation 2 (halo)
    (1, 128, 40, 48, 96, 3, 1, 4, (4, 4, 4, 4), 0.1, False),    # dilation 4 (halo)
    (1, 96, 40, 48, 64, 3, 1, 8, (8, 8, 8, 8), 0.1, False),     # dilation 8
    (1, 64, 40, 48, 32, 3, 1, 16, (16, 16, 16, 16), 0.1, False),  # dilation 16 (tap mode)
    (2, 3, 64, 64, 16, 3, 2, 1, (1, 1, 1, 1), 0.1, False),      # PWC conv1a: Cin 3, stride 2
    (2, 32, 33, 47, 64, 3, 2, 1, (1, 1, 1, 1), 0.1, False),     # stride 2, odd sizes
    (1, 64, 31, 45, 64, 3, 2, 1, (0, 0, 1, 1), 0.0, False),     # ResNet stride-2 3x3: pad (0,1) + VALID, ReLU
    (1, 64, 30, 44, 256, 1, 1, 1, (0, 0, 0, 0), 1.0, False),    # 1x1, two N tiles, no activation
    (1, 256, 30, 44, 64, 1, 1, 1, (0, 0, 0, 0), 0.0, True),     # 1x1 + residual + ReLU (bottleneck tail)
    (1, 256, 29, 43, 512, 1, 2, 1, (0, 0, 0, 0), 1.0, False),   # 1x1 stride-2 shortcut
    (1, 196, 7, 16, 196, 3, 1, 1, (1, 1, 1, 1), 0.1, False),    # PWC level 6: 196 channels (24.5 chunks)
    (1, 529, 7, 16, 10, 3, 1, 1, (1, 1, 1, 1), 1.0, False),     # flow head: Cout 10 -> N = 16
    (4, 117, 112, 256, 128, 3, 1, 1, (1, 1, 1, 1), 0.1, False),  # bench-size level 2 layer: MT = 2 tiles
    (1, 3, 75, 131, 64, 7, 2, 1, (2, 2, 3, 3), 0.0, False),     # ResNet stem 7x7 s2, pad (2,3)
    (2, 661, 14, 32, 96, 3, 1, 1, (1, 1, 1, 1), 0.1, False),    # PWC level 5: 8 CTAs -> split-K + finish kernel
    (1, 2048, 9, 9, 256, 1, 1, 1, (0, 0, 0, 0), 0.0, True),     # tiny grid 1x1 with residual through the split-K path
    (3, 728, 25, 25, 728, 1, 1, 1, (0, 0, 0, 0), 1.0, True),    # Xception middle flow: flattened 1x1, 625 px (not a multiple of 8)
    (2, 1024, 46, 83, 256, 1, 1, 1, (0, 0, 0, 0), 0.0, False),  # ResNet group2 conv1: flattened, 3818 px
    (20, 728, 25, 25, 728, 1, 1, 1, (0, 0, 0, 0), 1.0, True),   # bench-size Xception middle flow: 256 x 256 tiles (BN 256, MT 2,
                                                                #   8 epilogue warps, lockstep ring), N tail 216, residual
    (6, 264, 97, 97, 256, 1, 1, 1, (0, 0, 0, 0), 0.0, False),   # wide tile, one N tile, K = 33 chunks (k-block tail), ragged M tail
    (16, 128, 49, 49, 1024, 1, 1, 1, (0, 0, 0, 0), 0.0, False), # wide tile, 4 N tiles, KC from a 128-channel input
    (1, 64, 480, 160, 128, 3, 1, 1, (1, 1, 1, 1), 0.1, True),   # 300 items on 296 CTA slots: the last 4 are K-split (tail split),
                                                                #   halo mode, residual through conv_finish_tail_kernel
    (2, 96, 240, 168, 72, 3, 1, 2, (2, 2, 2, 2), 0.0, False),   # tail split with ragged tiles / N tail, dilation 2
    # CTA-pair kernel (cta_group::2, stream-K; the wide cases above run on it too: items split over 2-3 pairs, k-block tail)
    (15, 256, 25, 25, 512, 1, 1, 1, (0, 0, 0, 0), 0.1, True),   # 75 pixel tiles: the last pair has a dead peer tile
    (40, 728, 25, 25, 728, 1, 1, 1, (0, 0, 0, 0), 0.0, True),   # the benchmarked launch plan of the middle flow (40 crops)
    (5, 1536, 25, 25, 2048, 1, 1, 1, (0, 0, 0, 0), 0.0, False), # exit flow: 8 channel tiles, K = 48 k-blocks, 13 pairs x 8 = 104 items
    (1, 1024, 46, 83, 256, 1, 1, 1, (0, 0, 0, 0), 0.0, True),   # stream-K with 6.5 k-blocks per pair: every item is finished from 5-6 partial sums
    # folded small maps: several whole images per 128-row tile, per-image zero padding from the TMA's out-of-bounds fill
    (9, 512, 7, 7, 512, 3, 1, 1, (1, 1, 1, 1), 0.0, True),      # RoI head 3x3 at 7x7: stacked halo tile, 2 images at a row pitch of 9, odd image count
    (40, 1024, 4, 4, 2048, 3, 1, 1, (1, 1, 1, 1), 0.0, False),  # ReID res16: 8 images per tile, 5 groups x 16 channel tiles
    (13, 1024, 14, 14, 512, 1, 2, 1, (0, 0, 0, 0), 0.0, False), # RoI head entry: 1x1 stride 2 onto 7x7 (element strides in the folded box)
    (21, 2048, 4, 4, 4096, 1, 1, 1, (0, 0, 0, 0), 1.0, True),   # 1x1 on 4x4 maps + residual: last group holds 5 of 8 images
    (6, 512, 8, 8, 1024, 3, 2, 1, (0, 0, 1, 1), 1.0, False),    # TF SAME stride 2 (pad after only) onto 4x4: 6 images in one tile, split-K
    (5, 64, 8, 8, 64, 3, 1, 1, (1, 1, 1, 1), 0.1, False),       # 8x8 maps: stacked halo tile with two sub-tiles, 3 images at a row pitch of 10
    (3, 96, 5, 11, 40, 3, 1, 1, (1, 1, 1, 1), 0.1, True),       # 55-pixel maps wider than a tile row: 2 per tile, ragged channels
    (7, 40, 6, 5, 24, 3, 1, 1, (1, 1, 1, 1), 0.1, True),        # stacked halo tile: 2 images of 6x5 at a row pitch of 8, odd image count
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "n%d_c%d_%dx%d_o%d_k%d_s%d_d%d" % c[:8])
def test_conv2d_matches_torch(case):
    N, Cin, H, W, Cout, k, stride, dil, pad, slope, use_res = case
    g = torch.Generator().manual_seed(hash(case) % (2 ** 31))
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / np.sqrt(Cin * k * k)
    b = torch.randn(Cout, generator=g) * 0.1
    Ho = (H + pad[0] + pad[2] - dil * (k - 1) - 1) // stride + 1
    Wo = (W + pad[1] + pad[3] - dil * (k - 1) - 1) // stride + 1
    res = torch.randn(N, Cout, Ho, Wo, generator=g) if use_res else None
    ref = _ref(x, w, b, stride, dil, pad, slope, res)
    got = ops.conv2d(x.cuda(), w, b, stride, dil, pad, slope, None if res is None else res.cuda()).cpu()
    assert got.shape == ref.shape
    assert rel_err(got.numpy(), ref.numpy()) < TOL


def test_conv2d_folded_tiles_for_halo_capable_layers(monkeypatch):
    """3x3 stride-1 layers on 7x7 / 8x8 maps keep halo mode by default (measured faster); PREMVOS_FOLD=2 forces the folded
    tiling for them too, so that path stays covered for every geometry it accepts."""
    monkeypatch.setenv("PREMVOS_FOLD", "2")
    for case in CASES:
        N, Cin, H, W, Cout, k, stride, dil, pad, slope, use_res = case
        if not (k == 3 and stride == 1 and H * W <= 64 and N >= 2):
            continue
        test_conv2d_matches_torch(case)


def test_conv2d_linearity_and_zero_padding_at_full_size():
    # size-independent properties at the bench resolution
    g = torch.Generator().manual_seed(3)
    x = torch.randn(1, 32, 112, 256, generator=g).cuda()
    w = torch.randn(32, 32, 3, 3, generator=g) / 17.0
    y1 = ops.conv2d(x, w, None, 1, 1, (1, 1, 1, 1), 1.0)
    y2 = ops.conv2d(2.0 * x, w, None, 1, 1, (1, 1, 1, 1), 1.0)
    assert torch.allclose(y2, 2.0 * y1, rtol=1e-4, atol=1e-5)          # exact up to the hi/lo split rounding
    # an explicitly padded input with VALID geometry gives the same answer as implicit zero padding
    xp = torch.nn.functional.pad(x, (1, 1, 1, 1))
    y3 = ops.conv2d(xp, w, None, 1, 1, (0, 0, 0, 0), 1.0)
    assert torch.equal(y3, y1)


def test_conv2d_errors_are_loud():
    x = torch.zeros(1, 8, 8, 8, device="cuda")
    with pytest.raises(_lib.PremvosError):
        ops.conv2d(x, np.zeros((8, 8, 3, 3), np.float32), None, stride=3)
    with pytest.raises(ValueError):
        ops.conv2d(x, np.zeros((8, 4, 3, 3), np.float32))


SEP_CASES = [
    # (N, C, H, W, Cout, relu_in, relu_mid, slope)
    (1, 32, 16, 8, 32, False, False, 1.0),       # one exact tile, one k-block
    (2, 64, 37, 29, 128, True, False, 0.0),      # ragged tiles, two k-blocks, ReLU in front and behind (Xception entry flow)
    (3, 128, 45, 52, 128, False, False, 0.0),    # entry-flow block1 separable_conv2 geometry (4 k-blocks), several tiles per CTA
    (1, 40, 20, 21, 96, False, True, 1.0),       # channel tail inside the last k-block (40 = 5 chunks), ReLU between dw and pw, 96 outputs
    (4, 128, 193, 193, 128, False, False, 0.0),  # full size: 4 crops of the benchmarked layer (persistent CTAs over 1 300 tiles)
]


@pytest.mark.parametrize("case", SEP_CASES, ids=lambda c: "n%d_c%d_%dx%d_o%d" % c[:5])
def test_fused_separable_conv_matches_torch(case):
    N, C, H, W, Cout, relu_in, relu_mid, slope = case
    g = torch.Generator().manual_seed(hash(case) % (2 ** 31))
    x = torch.randn(N, C, H, W, generator=g)
    dw = torch.randn(C, 1, 3, 3, generator=g) / 3.0
    db = torch.randn(C, generator=g) * 0.1
    pw = torch.randn(Cout, C, 1, 1, generator=g) / np.sqrt(C)
    pb = torch.randn(Cout, generator=g) * 0.1
    xin = F.relu(x) if relu_in else x
    mid = F.conv2d(xin.double(), dw.double(), db.double(), padding=1, groups=C)
    if relu_mid:
        mid = F.relu(mid)
    y = F.conv2d(mid, pw.double(), pb.double())
    ref = torch.where(y > 0, y, y * slope).float()
    got = ops.sepconv2d(x.cuda(), dw, db, pw, pb, relu_in, relu_mid, slope).cpu()
    assert got.shape == ref.shape
    assert rel_err(got.numpy(), ref.numpy()) < TOL
