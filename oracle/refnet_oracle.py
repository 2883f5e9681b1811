"""CPU restatement (torch fp32 / numpy) of the refinement network's inference path -- TEST INFRASTRUCTURE ONLY.

Follows the reference (paths relative to /root/reference/code/refinement_net unless noted):

* per-proposal data set-up ......... datasets/few_shot_segmentation/FewShotFeedSegmentationDataset.py:35-51
* guidance mask, crop, resize ...... datasets/Dataset.py:48-56,141-186; datasets/Resize.py:150-193;
                                     datasets/util/{BoundingBox.py:15-19, Util.py:23-29, Normalization.py:9-36}
* un-normalise*255, [-1,1] map ..... network/deeplab/DeepLabV3Plus.py:13; deeplab/core/feature_extractor.py:114-116
* Xception-65 (output stride 16) ... deeplab/core/xception.py:70-89 (fixed_padding), 92-190 (separable_conv2d_same),
                                     193-293 (xception_module), 296-363 (stack_blocks_dense), 430-433 (root convs),
                                     496-560 (xception_65), 563-613 (arg scope, BN eps 1e-3 via feature_extractor.py:202)
* ASPP, decoder, logits ............ deeplab/model.py:328-435 (dynamic-shape image pooling :384-396), 503-598, 601-661,
                                     664-707; BN eps 1e-5 (:364-369, :530-536)
* output layer ..................... network/SegmentationOutputLayers.py:35-61, 106-135
* result packing ................... forwarding/FewShotSegmentationForwarder.py:139-149 and
                                     ../MergeTrack/refinement_net_functions.py:38-65 (mask*255 -> COCO RLE, conf_score)

Third-party arithmetic that is not vendored is restated from its published behaviour (TensorFlow 1.8):
  tf.image.resize_images / resize_bilinear (align_corners=False): in = out * (in_size / out_size) (no half-pixel
      centres), lower = floor(in), upper = min(lower + 1, in_size - 1), lerp.
  tf.image.resize_bilinear(align_corners=True): in = out * (in_size - 1) / (out_size - 1).
  tf.image.resize_nearest_neighbor (align_corners=False): in = min(floor(out * scale), in_size - 1).
  slim.conv2d / separable_conv2d 'SAME' and slim.batch_norm inference ((x - mean) * rsqrt(var + eps) * gamma + beta).
  pycocotools RLE (maskApi.c rleEncode + rleToString): column-major runs starting with a zero-run, LEB128-like
      string with differences against the run two positions earlier.

PARITY UNPINNED for the network's TensorFlow graph: the reference ships no golden vectors for it and TensorFlow cannot be
imported here; the restated resize kernels are pinned by hand-derived cases in tests/test_oracle_refnet.py.  Pinned against
vectors produced by the reference's own numpy functions executed from their source in the build container
(tests/golden/make_reference_function_goldens.py): encode_bbox_as_mask_np (BoundingBox.py:15-19) and normalize / the ImageNet
constants (Normalization.py), bit-exactly.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

INPUT_SIZE = 385
MARGIN = 50
IMAGENET_RGB_MEAN = np.array([0.485, 0.456, 0.406], dtype="float32")
IMAGENET_RGB_STD = np.array([0.229, 0.224, 0.225], dtype="float32")
XCEPTION_BN_EPS = 1e-3
ASPP_BN_EPS = 1e-5

# (scope, depth_list, skip, activation_fn_in_separable_conv, num_units, stride)   -- xception.py:507-548
XCEPTION_65_BLOCKS = [
    ("entry_flow/block1", [128, 128, 128], "conv", False, 1, 2),
    ("entry_flow/block2", [256, 256, 256], "conv", False, 1, 2),
    ("entry_flow/block3", [728, 728, 728], "conv", False, 1, 2),
    ("middle_flow/block1", [728, 728, 728], "sum", False, 16, 1),
    ("exit_flow/block1", [728, 1024, 1024], "conv", False, 1, 2),
    ("exit_flow/block2", [1536, 1536, 2048], "none", True, 1, 1),
]


def blocks_with_middle_units(units):
    """Xception-65 with a shorter middle flow (tests only; the reference always uses 16)."""
    return [(s, d, k, a, (units if s == "middle_flow/block1" else u), st) for s, d, k, a, u, st in XCEPTION_65_BLOCKS
            if not (s == "middle_flow/block1" and units == 0)]


LOW_LEVEL_FEATURE = "entry_flow/block2/unit_1/xception_module/separable_conv2_pointwise"


# ---------------------------------------------------------------------------------------------
# parameters (slim variable names)
# ---------------------------------------------------------------------------------------------
def refnet_param_shapes(blocks=XCEPTION_65_BLOCKS, n_classes=2):
    t = OrderedDict()

    def bn(scope, c):
        for v in ("gamma", "beta", "moving_mean", "moving_variance"):
            t[scope + "/BatchNorm/" + v] = (c,)

    def conv(scope, k, cin, cout):
        t[scope + "/weights"] = (k, k, cin, cout)
        bn(scope, cout)

    def sep(scope, cin, cout):
        t[scope + "_depthwise/depthwise_weights"] = (3, 3, cin, 1)
        bn(scope + "_depthwise", cin)
        conv(scope + "_pointwise", 1, cin, cout)

    x = "xception_65/"
    conv(x + "entry_flow/conv1_1", 3, 4, 32)
    conv(x + "entry_flow/conv1_2", 3, 32, 64)
    cin = 64
    for scope, depths, skip, _, units, _ in blocks:
        for u in range(units):
            s = "%s%s/unit_%d/xception_module" % (x, scope, u + 1)
            c = cin
            for i, d in enumerate(depths):
                sep("%s/separable_conv%d" % (s, i + 1), c, d)
                c = d
            if skip == "conv":
                conv(s + "/shortcut", 1, cin, depths[-1])
            cin = depths[-1]
    conv("image_pooling", 1, cin, 256)
    conv("aspp0", 1, cin, 256)
    for i in (1, 2, 3):
        sep("aspp%d" % i, cin, 256)
    conv("concat_projection", 1, 1280, 256)
    conv("decoder/feature_projection0", 1, 256, 48)
    sep("decoder/decoder_conv0", 304, 256)
    sep("decoder/decoder_conv1", 256, 256)
    t["logits/features/weights"] = (1, 1, 256, n_classes)
    t["logits/features/biases"] = (n_classes,)
    return t


# ---------------------------------------------------------------------------------------------
# TensorFlow resize kernels restated (single image, torch [C,H,W])
# ---------------------------------------------------------------------------------------------
def tf_resize_bilinear(img, out_h, out_w, align_corners=False):
    C, H, W = img.shape
    f = np.float32

    def axis(in_size, out_size):
        if align_corners and out_size > 1:
            scale = f(in_size - 1) / f(out_size - 1)
        else:
            scale = f(in_size) / f(out_size)
        src = (np.arange(out_size, dtype=np.float32) * scale).astype(np.float32)
        lo = np.floor(src).astype(np.int64)
        hi = np.minimum(lo + 1, in_size - 1)
        lerp = (src - lo.astype(np.float32)).astype(np.float32)
        return lo, hi, torch.from_numpy(lerp)

    ylo, yhi, yl = axis(H, out_h)
    xlo, xhi, xl = axis(W, out_w)
    tl = img[:, ylo][:, :, xlo]
    tr = img[:, ylo][:, :, xhi]
    bl = img[:, yhi][:, :, xlo]
    br = img[:, yhi][:, :, xhi]
    xl = xl.view(1, 1, -1)
    yl = yl.view(1, -1, 1)
    top = tl + (tr - tl) * xl
    bot = bl + (br - bl) * xl
    return top + (bot - top) * yl


def tf_resize_nearest(img, out_h, out_w):
    C, H, W = img.shape
    f = np.float32
    ys = np.minimum(np.floor(np.arange(out_h, dtype=np.float32) * (f(H) / f(out_h))).astype(np.int64), H - 1)
    xs = np.minimum(np.floor(np.arange(out_w, dtype=np.float32) * (f(W) / f(out_w))).astype(np.int64), W - 1)
    return img[:, ys][:, :, xs]


# ---------------------------------------------------------------------------------------------
# data path
# ---------------------------------------------------------------------------------------------
def crop_box(bbox_y0x0y1x1, h, w):
    """Resize.py:155-164 (tf.round = round half to even)"""
    y0, x0, y1, x1 = [int(v) for v in np.round(np.asarray(bbox_y0x0y1x1, dtype=np.float32))]
    return max(y0 - MARGIN, 0), max(x0 - MARGIN, 0), min(y1 + MARGIN, h), min(x1 + MARGIN, w)


def encode_bbox_as_mask_np(bbox_y0x0y1x1, shape):
    """datasets/util/BoundingBox.py:15-19: 1 inside round(bbox) (numpy rounding: half to even), uint8 [H,W,1]."""
    encoded = np.zeros(tuple(shape[:2]) + (1,), np.uint8)
    y0, x0, y1, x1 = np.round(bbox_y0x0y1x1).astype(np.int64)
    encoded[y0:y1, x0:x1] = 1
    return encoded


def normalize(img, img_mean=None, img_std=None):
    """datasets/util/Normalization.py:9-21 on a numpy image [..., 3]."""
    img_mean = IMAGENET_RGB_MEAN if img_mean is None else img_mean
    img_std = IMAGENET_RGB_STD if img_std is None else img_std
    return ((np.asarray(img, np.float32) - img_mean) / img_std).astype(np.float32)


def make_network_input(image, bbox_xywh, size=INPUT_SIZE):
    """image: float32 [H,W,3] RGB in 0..1 (already /255, FewShotFeedSegmentationDataset.py:37); bbox: x,y,w,h.
    Returns (inputs [385,385,4] float32 = what the `inputs` placeholder path feeds the network, crop box)."""
    image = np.asarray(image, dtype=np.float32)
    H, W = image.shape[:2]
    x0, y0, bw, bh = [np.float32(v) for v in bbox_xywh]
    bbox = np.array([y0, x0, y0 + bh, x0 + bw], dtype=np.float32)      # FewShotFeedSegmentationDataset.py:40-43
    guidance = encode_bbox_as_mask_np(bbox, (H, W))
    cy0, cx0, cy1, cx1 = crop_box(bbox, H, W)
    img_c = torch.from_numpy(image[cy0:cy1, cx0:cx1]).permute(2, 0, 1)
    g_c = torch.from_numpy(guidance[cy0:cy1, cx0:cx1].astype(np.float32)).permute(2, 0, 1)
    img_r = tf_resize_bilinear(img_c, size, size)            # Util.py:23-26
    g_r = tf_resize_nearest(g_c, size, size)                  # Util.py:27-28
    img_n = (img_r - torch.from_numpy(IMAGENET_RGB_MEAN).view(3, 1, 1)) / torch.from_numpy(IMAGENET_RGB_STD).view(3, 1, 1)
    inputs = torch.cat([img_n, g_r], 0).permute(1, 2, 0).contiguous().numpy()
    return inputs, (cy0, cx0, cy1, cx1)


# ---------------------------------------------------------------------------------------------
# network
# ---------------------------------------------------------------------------------------------
def _t(P, name):
    return torch.as_tensor(P[name])


def _bn(P, scope, x, eps):
    g, b = _t(P, scope + "/BatchNorm/gamma"), _t(P, scope + "/BatchNorm/beta")
    m, v = _t(P, scope + "/BatchNorm/moving_mean"), _t(P, scope + "/BatchNorm/moving_variance")
    scale = g * torch.rsqrt(v + eps)
    return (x - m.view(1, -1, 1, 1)) * scale.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)


def _conv_bn(P, scope, x, eps, stride=1, padding=0, relu=True):
    w = _t(P, scope + "/weights").permute(3, 2, 0, 1).contiguous()
    y = _bn(P, scope, F.conv2d(x, w, None, stride=stride, padding=padding), eps)
    return F.relu(y) if relu else y


def _depthwise_bn(P, scope, x, eps, stride=1, rate=1, relu=False):
    w = _t(P, scope + "/depthwise_weights")            # [3,3,C,1]
    C = w.shape[2]
    wt = w.permute(2, 3, 0, 1).contiguous()            # [C,1,3,3]
    if stride == 1:
        y = F.conv2d(x, wt, None, stride=1, padding=rate, dilation=rate, groups=C)      # SAME
    else:
        # fixed_padding (xception.py:70-89) + VALID
        k_eff = 3 + 2 * (rate - 1)
        pad_total = k_eff - 1
        pb = pad_total // 2
        pe = pad_total - pb
        y = F.conv2d(F.pad(x, (pb, pe, pb, pe)), wt, None, stride=stride, padding=0, dilation=rate, groups=C)
    y = _bn(P, scope, y, eps)
    return F.relu(y) if relu else y


def _separable(P, scope, x, cout_unused, eps, stride=1, rate=1, relu_inside=False):
    """split separable conv: depthwise(+BN[+ReLU]) then pointwise(+BN[+ReLU])  (xception.py:163-178, model.py:690-707)"""
    y = _depthwise_bn(P, scope + "_depthwise", x, eps, stride, rate, relu_inside)
    return _conv_bn(P, scope + "_pointwise", y, eps, relu=relu_inside)


def xception_module(P, scope, x, depths, skip, act_in_sep, stride, rate, end_points):
    """xception.py:193-293"""
    residual = x
    for i in range(3):
        if not act_in_sep:
            residual = F.relu(residual)
        s = "%s/separable_conv%d" % (scope, i + 1)
        residual = _separable(P, s, residual, depths[i], XCEPTION_BN_EPS, stride if i == 2 else 1, rate, act_in_sep)
        end_points[s + "_pointwise"] = residual
    if skip == "conv":
        # slim.conv2d 1x1 stride s 'SAME' == sampling every s-th pixel (no padding needed for a 1x1 kernel)
        shortcut = _conv_bn(P, scope + "/shortcut", x, XCEPTION_BN_EPS, stride=stride, relu=False)
        return residual + shortcut
    if skip == "sum":
        return residual + x
    return residual


def xception_65(P, x, end_points, blocks=XCEPTION_65_BLOCKS, output_stride=16):
    """xception.py:366-453 + 296-363 with output_stride handling"""
    px = "xception_65/"
    # resnet_utils.conv2d_same: stride 2 -> explicit pad (1,1) + VALID; stride 1 -> SAME
    net = _conv_bn(P, px + "entry_flow/conv1_1", F.pad(x, (1, 1, 1, 1)), XCEPTION_BN_EPS, stride=2)
    net = _conv_bn(P, px + "entry_flow/conv1_2", net, XCEPTION_BN_EPS, padding=1)
    target = output_stride // 2
    current_stride, rate = 1, 1
    for scope, depths, skip, act, units, stride in blocks:
        for u in range(units):
            s = "%s%s/unit_%d/xception_module" % (px, scope, u + 1)
            if current_stride == target:
                net = xception_module(P, s, net, depths, skip, act, 1, rate, end_points)
                rate *= stride
            else:
                net = xception_module(P, s, net, depths, skip, act, stride, 1, end_points)
                current_stride *= stride
    return net


def deeplab_logits(P, inputs_nhwc, blocks=XCEPTION_65_BLOCKS, return_intermediates=False):
    """DeepLabV3Plus.__init__ + multi_scale_logits for a batch [N,385,385,4] of assembled inputs -> logits [N,h/4,w/4,2]"""
    inter = OrderedDict()
    x = torch.as_tensor(inputs_nhwc, dtype=torch.float32)
    mean = torch.from_numpy(np.concatenate([IMAGENET_RGB_MEAN, np.zeros(1, np.float32)]))
    std = torch.from_numpy(np.concatenate([IMAGENET_RGB_STD, np.ones(1, np.float32)]))
    x = (x * std + mean) * 255                                   # unnormalize(inputs) * 255 (DeepLabV3Plus.py:13)
    x = np.float32(2.0 / 255.0) * x - 1.0                        # feature_extractor.py:114-116
    inter["net_input"] = x
    x = x.permute(0, 3, 1, 2).contiguous()
    ep = {}
    feat = xception_65(P, x, ep, blocks)
    inter["xception_out"] = feat
    low = ep["xception_65/" + LOW_LEVEL_FEATURE]
    inter["low_level"] = low
    h, w = feat.shape[2:]
    # ASPP (model.py:361-435)
    pooled = feat.mean(dim=(2, 3), keepdim=True)
    img_feat = _conv_bn(P, "image_pooling", pooled, ASPP_BN_EPS).expand(-1, -1, h, w)
    branches = [img_feat, _conv_bn(P, "aspp0", feat, ASPP_BN_EPS)]
    for i, r in enumerate((6, 12, 18), 1):
        branches.append(_separable(P, "aspp%d" % i, feat, 256, ASPP_BN_EPS, 1, r, True))
    cat = torch.cat(branches, 1)
    inter["aspp_concat"] = cat
    aspp = _conv_bn(P, "concat_projection", cat, ASPP_BN_EPS)
    inter["aspp_out"] = aspp
    # decoder (model.py:503-598): both features resized to (H-1)/4+1 with align_corners=True
    dh = int((float(x.shape[2]) - 1.0) * 0.25 + 1.0)
    dw = int((float(x.shape[3]) - 1.0) * 0.25 + 1.0)
    low48 = _conv_bn(P, "decoder/feature_projection0", low, ASPP_BN_EPS)
    up = torch.stack([tf_resize_bilinear(a, dh, dw, True) for a in aspp])
    low48 = torch.stack([tf_resize_bilinear(a, dh, dw, True) for a in low48])
    dec = torch.cat([up, low48], 1)
    inter["decoder_in"] = dec
    dec = _separable(P, "decoder/decoder_conv0", dec, 256, ASPP_BN_EPS, 1, 1, True)
    dec = _separable(P, "decoder/decoder_conv1", dec, 256, ASPP_BN_EPS, 1, 1, True)
    inter["decoder_out"] = dec
    wl = _t(P, "logits/features/weights").permute(3, 2, 0, 1).contiguous()
    logits = F.conv2d(dec, wl, _t(P, "logits/features/biases"))
    logits = logits.permute(0, 2, 3, 1).contiguous()             # NHWC [N,97,97,2]
    inter["logits"] = logits
    if return_intermediates:
        return logits, inter
    return logits


def segmentation_output(logits_hw2, crop, frame_h, frame_w, size=INPUT_SIZE):
    """SegmentationSoftmax extraction path for one proposal (SegmentationOutputLayers.py:35-61, 106-135).
    logits_hw2: torch [h,w,2].  Returns (mask int64 [H,W], posterior float32 [H,W])."""
    cy0, cx0, cy1, cx1 = crop
    lg = tf_resize_bilinear(logits_hw2.permute(2, 0, 1), size, size)     # resize_logits (:36)
    prob = torch.softmax(lg, dim=0)
    class_pred = torch.argmax(lg, dim=0)
    ch, cw = cy1 - cy0, cx1 - cx0
    mask_c = tf_resize_nearest(class_pred.unsqueeze(0).to(torch.float32), ch, cw)[0].to(torch.int64)
    post_c = tf_resize_bilinear(prob[1:2], ch, cw)[0]
    mask = torch.zeros((frame_h, frame_w), dtype=torch.int64)
    post = torch.zeros((frame_h, frame_w), dtype=torch.float32)
    mask[cy0:cy1, cx0:cx1] = mask_c
    post[cy0:cy1, cx0:cx1] = post_c
    return mask.numpy(), post.numpy()


def conf_score(mask, posterior):
    """refinement_net_functions.py:58-62"""
    c = posterior.copy()
    c[mask == 0] = 1 - posterior[mask == 0]
    c = 2 * c - 1
    return c[:].mean()


# ---------------------------------------------------------------------------------------------
# COCO RLE (pycocotools maskApi.c)
# ---------------------------------------------------------------------------------------------
def rle_encode(mask):
    """pycocotools.mask.encode(np.asfortranarray(mask)) -> {'size': [h, w], 'counts': str}"""
    m = np.asarray(mask)
    h, w = m.shape
    flat = (m.T.reshape(-1) != 0).astype(np.uint8)           # column-major
    counts = []
    prev, run = 0, 0
    for v in flat:
        if v != prev:
            counts.append(run)
            run = 0
            prev = v
        run += 1
    counts.append(run)
    s = []
    for i, x in enumerate(counts):
        x = int(x)
        if i > 2:
            x -= int(counts[i - 2])
        more = True
        while more:
            c = x & 0x1f
            x >>= 5
            more = (x != -1) if (c & 0x10) else (x != 0)
            if more:
                c |= 0x20
            s.append(chr(c + 48))
    return {"size": [h, w], "counts": "".join(s)}


def rle_decode(rle):
    h, w = rle["size"]
    s = rle["counts"]
    counts = []
    p = 0
    while p < len(s):
        x, k, more = 0, 0, True
        while more:
            c = ord(s[p]) - 48
            x |= (c & 0x1f) << (5 * k)
            more = bool(c & 0x20)
            p += 1
            k += 1
            if not more and (c & 0x10):
                x |= -1 << (5 * k)
        if len(counts) > 2:
            x += counts[-2]
        counts.append(x)
    flat = np.zeros(h * w, np.uint8)
    pos, v = 0, 0
    for c in counts:
        flat[pos:pos + c] = v
        pos += c
        v = 1 - v
    return flat.reshape(w, h).T


def do_refinement(P, proposals, image_rgb_uint8, blocks=XCEPTION_65_BLOCKS, size=INPUT_SIZE):
    """MergeTrack/refinement_net_functions.py:38-65 with the network replaced by this oracle."""
    image = (np.asarray(image_rgb_uint8) / 255).astype(np.float32)    # numpy true division (float64), fed as float32
    H, W = image.shape[:2]
    for prop in proposals:
        inputs, crop = make_network_input(image, prop["bbox"], size)
        logits = deeplab_logits(P, inputs[None], blocks)[0]
        mask, post = segmentation_output(logits, crop, H, W, size)
        enc = rle_encode(mask.astype("uint8") * 255)
        prop["segmentation"] = enc
        prop["conf_score"] = str(conf_score(mask, post))
    return proposals
