"""CPU restatement (torch fp32 / numpy) of the ReID network's forward path -- TEST INFRASTRUCTURE ONLY.

Follows the reference (paths relative to /root/reference/code):

* call surface ..................... MergeTrack/ReID_net_functions.py:19-45 (ReID_net_init, add_ReID)
* network spec ..................... ReID_net/configs/run:33-65 (conv0, res0..res16, conv1, fc1, fc2, outputTriplet; 128-d output)
* layers ........................... ReID_net/network/NetworkLayers.py:106-155 (Conv: [BN] -> activation -> conv -> max pool),
                                     :157-210 (ResidualUnit2: BN0 -> ReLU -> [1x1 projection W0] ; W1 -> BN -> ReLU -> W2 ... ; + res),
                                     :236-252 (FullyConnected: [BN] -> matmul + b -> activation),
                                     ReID_net/network/NetworkOutputLayers.py:253-272 (FullyConnectedWithTripletLoss: the same, linear),
                                     ReID_net/network/Util_Network.py:12-30 (conv2d / max_pool: padding SAME), :87-96 (BN variables),
                                     BATCH_NORM_EPSILON = 1e-5 (NetworkLayers.py:12), inference BN = tf.nn.batch_normalization on the
                                     moving statistics (:59-61)
* crop path ........................ ReID_net/datasets/Similarity/DAVIS_Forward_Feed.py:34-120: image / 255; box context region
                                     x1.2, tf.round (half to even), clip -- including the reference's `maximum(excess, 1)`, which
                                     always shrinks a box by at least one pixel --; crop; bilinear resize to 128 x 128
                                     (tf.image.resize_images, TF1 legacy: in = out * in_size / out_size) unless min(h, w) <= 10
                                     (zeros); normalize with the ImageNet mean / std (datasets/Util/Normalization.py:9-22)

PARITY UNPINNED for the network: the reference ships no vectors for it and TensorFlow cannot be imported here; the restated
third-party arithmetic (SAME padding of strided convolutions / max pooling, legacy resize) is the published TensorFlow 1.x
behaviour, the legacy resize shared with oracle/refnet_oracle.py (pinned there by hand-derived cases).  PINNED
(tests/golden/reid_reference_golden.npz, made by tests/golden/make_reid_reference_goldens.py): `apply_context_region` against the
reference method's own source executed with a numpy stand-in for its six TensorFlow ops (80 boxes, bit-exact), `normalize`
against the reference's numpy function, the layer table against the reference's configs/live.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

INPUT_SIZE = 128
CONTEXT_REGION_FACTOR = 1.2
BN_EPS = 1e-5
IMAGENET_RGB_MEAN = np.array([0.485, 0.456, 0.406], dtype="float32")
IMAGENET_RGB_STD = np.array([0.229, 0.224, 0.225], dtype="float32")
EMBEDDING_DIM = 128

# (name, n_features per conv, filter sizes, strides per conv)  -- configs/run:36-59
RESIDUAL_UNITS = (
    [("res0", [128, 128], [3, 3], [2, 1])] + [("res%d" % i, [128, 128], [3, 3], [1, 1]) for i in (1, 2)] +
    [("res3", [256, 256], [3, 3], [2, 1])] + [("res%d" % i, [256, 256], [3, 3], [1, 1]) for i in (4, 5)] +
    [("res6", [512, 512], [3, 3], [2, 1])] + [("res%d" % i, [512, 512], [3, 3], [1, 1]) for i in range(7, 12)] +
    [("res12", [512, 1024], [3, 3], [1, 2]), ("res13", [512, 1024], [3, 3], [1, 1]), ("res14", [512, 1024], [3, 3], [1, 1]),
     ("res15", [512, 1024, 2048], [1, 3, 1], [1, 2, 1]), ("res16", [1024, 2048, 4096], [1, 3, 1], [1, 1, 1])])


def reid_param_shapes() -> "OrderedDict[str, tuple]":
    t = OrderedDict()

    def bn(scope, c):
        for v in ("beta", "gamma", "mean_ema", "var_ema"):
            t["%s/%s" % (scope, v)] = (c,)
    t["conv0/W"] = (3, 3, 3, 64)
    cin = 64
    for name, feats, ks, strides in RESIDUAL_UNITS:
        bn(name + "/bn0", cin)
        stride_res = int(np.prod(strides))
        if feats[-1] != cin or stride_res != 1:
            t[name + "/W0"] = (1, 1, cin, feats[-1])
        c = cin
        for i, (f, k) in enumerate(zip(feats, ks)):
            if i > 0:
                bn("%s/bn%d" % (name, i + 1), c)
            t["%s/W%d" % (name, i + 1)] = (k, k, c, f)
            c = f
        cin = feats[-1]
    bn("conv1/bn", cin)
    t["conv1/W"] = (3, 3, cin, 500)
    for name, fin, fout in (("fc1", 2 * 2 * 500, 500), ("fc2", 500, 500), ("outputTriplet", 500, EMBEDDING_DIM)):
        bn(name + "/bn", fin)
        t[name + "/W"] = (fin, fout)
        t[name + "/b"] = (fout,)
    return t


# ---------------------------------------------------------------------------------------------
# crop path
# ---------------------------------------------------------------------------------------------
def apply_context_region(boxes_xywh, H, W, factor=CONTEXT_REGION_FACTOR):
    """DAVIS_Forward_Feed.py:36-58 on float32 [n,4] (x, y, w, h) -> int32 [n,4] crop boxes (x, y, w, h)."""
    b = np.asarray(boxes_xywh, dtype=np.float32).reshape(-1, 4).copy()
    f = np.float32(factor)
    xs, ys, ws, hs = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    fm1 = np.float32(factor - 1.0)            # the python double 0.19999999999999996 meets a float32 tensor: 0.2f, not 1.2f - 1.0f
    xs = xs - np.float32(0.5) * ws * fm1
    ys = ys - np.float32(0.5) * hs * fm1
    ws = ws * f
    hs = hs * f
    xs, ys, ws, hs = (np.rint(v).astype(np.int32) for v in (xs, ys, ws, hs))      # tf.round: half to even
    xs = np.maximum(xs, 0)
    ys = np.maximum(ys, 0)
    ws = ws - np.maximum(xs + ws - W, 1)                                            # sic: at least one pixel is always taken off
    hs = hs - np.maximum(ys + hs - H, 1)
    return np.stack([xs, ys, ws, hs], 1).astype(np.int32)


def legacy_resize_bilinear(img_hwc: torch.Tensor, out_h: int, out_w: int) -> torch.Tensor:
    """tf.image.resize_images(img, size) of TF 1.x (align_corners=False, no half-pixel centres) on [h,w,c] float32."""
    h, w = img_hwc.shape[:2]

    def axis(out_n, in_n):
        scale = np.float32(in_n) / np.float32(out_n)
        src = np.arange(out_n, dtype=np.float32) * scale
        lo = np.floor(src).astype(np.int64)
        hi = np.minimum(lo + 1, in_n - 1)
        return torch.from_numpy(lo), torch.from_numpy(hi), torch.from_numpy((src - lo.astype(np.float32)).astype(np.float32))
    ylo, yhi, yl = axis(out_h, h)
    xlo, xhi, xl = axis(out_w, w)
    top_l, top_r = img_hwc[ylo][:, xlo], img_hwc[ylo][:, xhi]
    bot_l, bot_r = img_hwc[yhi][:, xlo], img_hwc[yhi][:, xhi]
    top = top_l + (top_r - top_l) * xl.view(1, -1, 1)
    bot = bot_l + (bot_r - bot_l) * xl.view(1, -1, 1)
    return top + (bot - top) * yl.view(-1, 1, 1)


def make_crops(image_rgb_uint8: np.ndarray, boxes_xywh, size=INPUT_SIZE) -> torch.Tensor:
    """-> float32 [n, size, size, 3] normalised crops (the network's `inputs`)."""
    H, W = image_rgb_uint8.shape[:2]
    conv_image = torch.from_numpy(np.asarray(image_rgb_uint8, dtype=np.float32) / np.float32(255))
    mean, std = torch.from_numpy(IMAGENET_RGB_MEAN), torch.from_numpy(IMAGENET_RGB_STD)
    out = []
    for x, y, w, h in apply_context_region(boxes_xywh, H, W):
        crop = conv_image[max(y, 0):max(y + h, 0), max(x, 0):max(x + w, 0)] if (h > 0 and w > 0) else conv_image[:0, :0]
        if min(h, w) > 10:
            img = legacy_resize_bilinear(crop, size, size)
        else:
            img = torch.zeros(size, size, 3)
        out.append((img - mean) / std)
    return torch.stack(out) if out else torch.zeros(0, size, size, 3)


# ---------------------------------------------------------------------------------------------
# network
# ---------------------------------------------------------------------------------------------
def _t(P, name):
    return torch.as_tensor(P[name])


def _bn(P, scope, x):
    """tf.nn.batch_normalization on the moving statistics; x is [N,C,H,W] or [N,C]."""
    g, b, m, v = (_t(P, "%s/%s" % (scope, k)) for k in ("gamma", "beta", "mean_ema", "var_ema"))
    shp = (1, -1, 1, 1) if x.dim() == 4 else (1, -1)
    return (x - m.view(shp)) * (g * torch.rsqrt(v + BN_EPS)).view(shp) + b.view(shp)


def _same_pad(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def conv2d_same(x, w_hwio, stride=1):
    """tf.nn.conv2d(padding='SAME') on NCHW x with an HWIO kernel: the extra padding pixel goes to the bottom / right."""
    w = w_hwio.permute(3, 2, 0, 1).contiguous()
    k = w.shape[2]
    pt, pb = _same_pad(x.shape[2], k, stride)
    pl, pr = _same_pad(x.shape[3], k, stride)
    return F.conv2d(F.pad(x, (pl, pr, pt, pb)), w, None, stride=stride)


def max_pool_same(x, k, s):
    pt, pb = _same_pad(x.shape[2], k, s)
    pl, pr = _same_pad(x.shape[3], k, s)
    return F.max_pool2d(F.pad(x, (pl, pr, pt, pb), value=float("-inf")), k, s)


def residual_unit2(P, name, x, feats, ks, strides):
    cin = x.shape[1]
    curr = F.relu(_bn(P, name + "/bn0", x))
    res = x
    stride_res = int(np.prod(strides))
    if feats[-1] != cin or stride_res != 1:
        res = conv2d_same(curr, _t(P, name + "/W0"), stride_res)
    curr = conv2d_same(curr, _t(P, name + "/W1"), strides[0])
    for i in range(1, len(feats)):
        curr = F.relu(_bn(P, "%s/bn%d" % (name, i + 1), curr))
        curr = conv2d_same(curr, _t(P, "%s/W%d" % (name, i + 1)), strides[i])
    return curr + res


def reid_forward(P, crops_nhwc, return_intermediates=False):
    """crops float32 [n,128,128,3] (make_crops) -> embeddings float32 [n,128]."""
    inter = OrderedDict()
    x = torch.as_tensor(crops_nhwc, dtype=torch.float32).permute(0, 3, 1, 2).contiguous()
    x = conv2d_same(x, _t(P, "conv0/W"))                                   # activation "linear", no BN, no bias
    inter["conv0"] = x
    for name, feats, ks, strides in RESIDUAL_UNITS:
        x = residual_unit2(P, name, x, feats, ks, strides)
        if name in ("res0", "res2", "res5", "res11", "res14", "res16"):
            inter[name] = x
    x = conv2d_same(F.relu(_bn(P, "conv1/bn", x)), _t(P, "conv1/W"))        # BN -> ReLU -> conv 3x3 -> max pool 3x3 / 3
    x = max_pool_same(x, 3, 3)
    inter["conv1"] = x
    h = x.permute(0, 2, 3, 1).reshape(x.shape[0], -1)                        # NHWC flatten (prepare_collapsed_input_and_dropout)
    h = F.relu(_bn(P, "fc1/bn", h) @ _t(P, "fc1/W") + _t(P, "fc1/b"))
    h = F.relu(_bn(P, "fc2/bn", h) @ _t(P, "fc2/W") + _t(P, "fc2/b"))
    out = _bn(P, "outputTriplet/bn", h) @ _t(P, "outputTriplet/W") + _t(P, "outputTriplet/b")
    inter["embedding"] = out
    return (out, inter) if return_intermediates else out


def add_ReID(P, proposals, image_rgb_uint8):
    """MergeTrack/ReID_net_functions.py:26-45 on an already decoded RGB frame."""
    boxes = [p["bbox"] for p in proposals]
    if not boxes:
        return proposals
    emb = reid_forward(P, make_crops(image_rgb_uint8, boxes)).numpy()
    for p, e in zip(proposals, emb):
        p["ReID"] = e.tolist()
    return proposals
