"""Seeded synthetic weights and DAVIS/Sintel-shaped frames (there is no network for real
checkpoints or datasets).  Everything is generated with numpy's PCG64 so the
same seed gives the same bytes in the build container and on the GPU box."""
from __future__ import annotations

from collections import OrderedDict

import numpy as np


def pwc_param_shapes() -> "OrderedDict[str, tuple]":
    """Reference state_dict keys/shapes (code/optical_flow_net-PWC-Net/models/PWCNet.py:50-131),
    in registration order: 128 tensors, 9 374 340 parameters."""
    t = OrderedDict()

    def conv(name, cin, cout, seq=True):
        key = name + (".0" if seq else "")
        t[key + ".weight"] = (cout, cin, 3, 3)
        t[key + ".bias"] = (cout,)

    def deconv(name, cin, cout):
        t[name + ".weight"] = (cin, cout, 4, 4)
        t[name + ".bias"] = (cout,)

    chans = [3, 16, 32, 64, 96, 128, 196]
    for lvl in range(1, 7):
        ci, co = chans[lvl - 1], chans[lvl]
        names = ("conv%da" % lvl, "conv%daa" % lvl, "conv%db" % lvl)
        if lvl == 6:  # the reference applies conv6aa (stride 2) first, PWCNet.py:65-67,193
            names = ("conv6aa", "conv6a", "conv6b")
        conv(names[0], ci, co)
        conv(names[1], co, co)
        conv(names[2], co, co)
    dd = [128, 256, 352, 416, 448]
    for lvl, extra in ((6, 0), (5, 132), (4, 100), (3, 68), (2, 36)):
        od = 81 + extra
        for i, (cin, cout) in enumerate(((od, 128), (od + dd[0], 128), (od + dd[1], 96),
                                         (od + dd[2], 64), (od + dd[3], 32))):
            conv("conv%d_%d" % (lvl, i), cin, cout)
        conv("predict_flow%d" % lvl, od + dd[4], 2, seq=False)
        deconv("deconv%d" % lvl, 2, 2)
        if lvl != 2:
            deconv("upfeat%d" % lvl, od + dd[4], 2)
    for name, cin, cout in (("dc_conv1", 565, 128), ("dc_conv2", 128, 128), ("dc_conv3", 128, 128),
                            ("dc_conv4", 128, 96), ("dc_conv5", 96, 64), ("dc_conv6", 64, 32)):
        conv(name, cin, cout)
    conv("dc_conv7", 32, 2, seq=False)
    return t


def pwc_synthetic_state_dict(seed: int = 0) -> "OrderedDict[str, np.ndarray]":
    """He-normal (fan_in) weights like the reference's own init (PWCNet.py:133-137) plus small
    non-zero biases so the bias path is exercised.  Flow heads are scaled x2 (up-features x0.5) so the
    synthetic flows move the warping layer by a few pixels per level."""
    rng = np.random.default_rng(seed)
    sd = OrderedDict()
    for name, shape in pwc_param_shapes().items():
        if name.endswith(".weight"):
            if "deconv" in name or "upfeat" in name:
                fan_in = shape[1] * shape[2] * shape[3]  # torch fan_in of a ConvTranspose2d weight
            else:
                fan_in = shape[1] * shape[2] * shape[3]
            std = np.sqrt(2.0 / fan_in)
            if name.startswith("predict_flow"):
                std *= 2.0
            elif name.startswith("dc_conv7") or name.startswith("upfeat"):
                std *= 0.5
            sd[name] = (rng.standard_normal(shape) * std).astype(np.float32)
        else:
            sd[name] = (rng.standard_normal(shape) * 0.02).astype(np.float32)
    return sd


def _smooth_texture(rng, h, w, c=3, octaves=(2, 4, 8, 16, 32, 64)):
    """Band-limited random texture in [0,255]: sum of bilinearly up-sampled noise octaves."""
    img = np.zeros((h, w, c), dtype=np.float64)
    for o in octaves:
        gh, gw = h // o + 3, w // o + 3
        g = rng.random((gh, gw, c))
        ys = np.arange(h) / o
        xs = np.arange(w) / o
        y0 = ys.astype(np.int64)
        x0 = xs.astype(np.int64)
        fy = (ys - y0)[:, None, None]
        fx = (xs - x0)[None, :, None]
        a = g[y0][:, x0]
        b = g[y0][:, x0 + 1]
        cc = g[y0 + 1][:, x0]
        d = g[y0 + 1][:, x0 + 1]
        img += (a * (1 - fy) * (1 - fx) + b * (1 - fy) * fx + cc * fy * (1 - fx) + d * fy * fx) * o
    img -= img.min()
    img /= img.max()
    return img * 255.0


def synthetic_frame_pair(h: int = 436, w: int = 1024, seed: int = 1, shift=(3.7, -2.2), noise_sigma=2.0):
    """Two uint8 RGB frames [h,w,3]; frame 2 = frame 1 translated by ``shift`` (x,y) px + noise."""
    rng = np.random.default_rng(seed)
    m = 16
    big = _smooth_texture(rng, h + 2 * m, w + 2 * m)
    f1 = big[m:m + h, m:m + w]
    dx, dy = shift
    ys = np.arange(h) + m - dy
    xs = np.arange(w) + m - dx
    y0 = np.floor(ys).astype(np.int64)
    x0 = np.floor(xs).astype(np.int64)
    fy = (ys - y0)[:, None, None]
    fx = (xs - x0)[None, :, None]
    f2 = (big[y0][:, x0] * (1 - fy) * (1 - fx) + big[y0][:, x0 + 1] * (1 - fy) * fx
          + big[y0 + 1][:, x0] * fy * (1 - fx) + big[y0 + 1][:, x0 + 1] * fy * fx)
    f2 = f2 + rng.standard_normal(f2.shape) * noise_sigma
    to_u8 = lambda a: np.clip(np.rint(a), 0, 255).astype(np.uint8)
    return to_u8(f1), to_u8(f2)


def synthetic_pwc_input(batch: int, h_: int, w_: int, seed: int = 1) -> np.ndarray:
    """float32 [batch,6,h_,w_] network input (BGR/255), h_/w_ already multiples of 64."""
    out = np.empty((batch, 6, h_, w_), dtype=np.float32)
    for b in range(batch):
        f1, f2 = synthetic_frame_pair(h_, w_, seed=seed + b, shift=(3.7 - 0.9 * b, -2.2 + 0.7 * b))
        for k, f in enumerate((f1, f2)):
            out[b, 3 * k:3 * k + 3] = np.transpose(f[:, :, ::-1].astype(np.float32) / np.float32(255.0), (2, 0, 1))
    return out


# ---- proposal network (tensorpack variable names, proposal_net/basemodel.py + model.py) -----------------
def propnet_param_shapes(num_blocks=(3, 4, 23, 3), num_class=2, second_num_class=81, mode_mask=False):
    t = OrderedDict()

    def conv_bn(scope, k, cin, cout):
        t[scope + "/W"] = (k, k, cin, cout)  # HWIO
        for v in ("gamma", "beta", "mean/EMA", "variance/EMA"):
            t[scope + "/bn/" + v] = (cout,)

    conv_bn("conv0", 7, 3, 64)
    cin = 64
    for g, (ch, nb) in enumerate(zip((64, 128, 256, 512), num_blocks)):
        for b in range(nb):
            s = "group%d/block%d" % (g, b)
            conv_bn(s + "/conv1", 1, cin, ch)
            conv_bn(s + "/conv2", 3, ch, ch)
            conv_bn(s + "/conv3", 1, ch, ch * 4)
            if cin != ch * 4:
                conv_bn(s + "/convshortcut", 1, cin, ch * 4)
            cin = ch * 4
    t["rpn/conv0/W"] = (3, 3, 1024, 1024)
    t["rpn/conv0/b"] = (1024,)
    t["rpn/class/W"] = (1, 1, 1024, 15)
    t["rpn/class/b"] = (15,)
    t["rpn/box/W"] = (1, 1, 1024, 60)
    t["rpn/box/b"] = (60,)
    t["fastrcnn/class/W"] = (2048, num_class)
    t["fastrcnn/class/b"] = (num_class,)
    t["fastrcnn/box/W"] = (2048, (num_class - 1) * 4)
    t["fastrcnn/box/b"] = ((num_class - 1) * 4,)
    if second_num_class:
        t["secondclassification/class/W"] = (2048, second_num_class)
        t["secondclassification/class/b"] = (second_num_class,)
    if mode_mask:   # model.py:495-509
        t["maskrcnn/deconv/W"] = (2, 2, 256, 2048)
        t["maskrcnn/deconv/b"] = (256,)
        t["maskrcnn/conv/W"] = (1, 1, 256, num_class - 1)
        t["maskrcnn/conv/b"] = (num_class - 1,)
    return t


def maskrcnn_synthetic_params(seed=0, num_class=2):
    """Seeded weights of the Mask R-CNN mask head (proposal_net/model.py:495-509): 'maskrcnn/deconv/{W [2,2,256,2048], b}',
    'maskrcnn/conv/{W [1,1,256,num_class-1], b}'; the 1x1 layer is wide enough that sigmoid outputs spread over (0,1)."""
    rng = np.random.default_rng(seed + 7777)
    P = OrderedDict()
    P["maskrcnn/deconv/W"] = (rng.standard_normal((2, 2, 256, 2048)) * np.sqrt(2.0 / 2048)).astype(np.float32)
    P["maskrcnn/deconv/b"] = (rng.standard_normal((256,)) * 0.05).astype(np.float32)
    P["maskrcnn/conv/W"] = (rng.standard_normal((1, 1, 256, num_class - 1)) * (4.0 / np.sqrt(256))).astype(np.float32)
    P["maskrcnn/conv/b"] = (rng.standard_normal((num_class - 1,)) * 0.05).astype(np.float32)
    return P


def propnet_synthetic_params(seed=0, num_blocks=(3, 4, 23, 3), num_class=2, second_num_class=81):
    """Seeded weights with trained-net-like statistics: He-normal convolutions, BatchNorm gamma/variance near 1
    (+-10 %), small beta/mean; the last BN of every bottleneck is damped (gamma ~0.25) so that activations neither
    explode nor vanish through 33 residual blocks; head weights wide enough that scores spread over (0,1) and few
    decisions sit on a threshold."""
    rng = np.random.default_rng(seed)
    P = OrderedDict()
    for name, shape in propnet_param_shapes(num_blocks, num_class, second_num_class).items():
        if name.endswith("/W") and len(shape) == 4:
            fan_in = shape[0] * shape[1] * shape[2]
            std = np.sqrt(2.0 / fan_in)
            if name.startswith("rpn/class"):
                std = 3.0 / np.sqrt(fan_in)
            elif name.startswith("rpn/box"):
                std = 0.3 / np.sqrt(fan_in)
            P[name] = (rng.standard_normal(shape) * std).astype(np.float32)
        elif name.endswith("/W"):
            std = {"fastrcnn/class/W": 4.0, "fastrcnn/box/W": 1.0}.get(name, 3.0) / np.sqrt(shape[0])
            P[name] = (rng.standard_normal(shape) * std).astype(np.float32)
        elif name.endswith("/b"):
            P[name] = (rng.standard_normal(shape) * 0.05).astype(np.float32)
        elif name.endswith("bn/gamma"):
            base = 0.25 if "/conv3/" in name else 1.0
            P[name] = (base * (1.0 + 0.1 * rng.uniform(-1, 1, shape))).astype(np.float32)
        elif name.endswith("bn/beta"):
            P[name] = (0.05 * rng.standard_normal(shape)).astype(np.float32)
        elif name.endswith("bn/mean/EMA"):
            P[name] = (0.05 * rng.standard_normal(shape)).astype(np.float32)
        elif name.endswith("bn/variance/EMA"):
            P[name] = (1.0 + 0.1 * rng.uniform(-1, 1, shape)).astype(np.float32)
        else:
            raise AssertionError(name)
    return P


def synthetic_bgr_frame(h, w, seed=2):
    """uint8 BGR frame [h,w,3]: low-pass filtered noise (smooth texture) plus a few bright rectangles."""
    rng = np.random.default_rng(seed)
    coarse = rng.uniform(0, 255, (h // 8 + 2, w // 8 + 2, 3))
    img = np.kron(coarse, np.ones((8, 8, 1)))[:h, :w]
    k = np.ones(5) / 5.0
    for ax in (0, 1):
        img = np.apply_along_axis(lambda v: np.convolve(v, k, mode="same"), ax, img)
    for _ in range(6):
        y0, x0 = int(rng.integers(0, max(1, h - 20))), int(rng.integers(0, max(1, w - 20)))
        hh, ww = int(rng.integers(10, max(11, h // 3))), int(rng.integers(10, max(11, w // 3)))
        img[y0:y0 + hh, x0:x0 + ww] = rng.uniform(0, 255, 3)
    return np.clip(img + rng.normal(0, 2, img.shape), 0, 255).astype(np.uint8)


# ---- refinement network (slim variable names, refinement_net/network/deeplab) -----------------------------
def refnet_param_shapes(middle_units=16, n_classes=2):
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

    blocks = [("entry_flow/block1", [128, 128, 128], "conv", 1), ("entry_flow/block2", [256, 256, 256], "conv", 1),
              ("entry_flow/block3", [728, 728, 728], "conv", 1), ("middle_flow/block1", [728, 728, 728], "sum", middle_units),
              ("exit_flow/block1", [728, 1024, 1024], "conv", 1), ("exit_flow/block2", [1536, 1536, 2048], "none", 1)]
    x = "xception_65/"
    conv(x + "entry_flow/conv1_1", 3, 4, 32)
    conv(x + "entry_flow/conv1_2", 3, 32, 64)
    cin = 64
    for scope, depths, skip, units in blocks:
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


def refnet_synthetic_params(seed=0, middle_units=16, n_classes=2):
    """Seeded DeepLabv3+/Xception-65 variables with trained-net-like statistics (He-normal kernels, BN gamma and
    variance near 1, small beta / mean; the BN closing a residual unit is damped so 20 stacked units stay O(1))."""
    rng = np.random.default_rng(seed)
    P = OrderedDict()
    for name, shape in refnet_param_shapes(middle_units, n_classes).items():
        if name.endswith("depthwise_weights"):
            # variance-preserving: a ReLU follows the depthwise conv only in exit_flow/block2, ASPP and the decoder
            relu_after = ("exit_flow/block2" in name) or name.startswith("aspp") or name.startswith("decoder")
            P[name] = (rng.standard_normal(shape) * np.sqrt((2.0 if relu_after else 1.0) / 9.0)).astype(np.float32)
        elif name.endswith("/weights"):
            fan_in = shape[0] * shape[1] * shape[2]
            std = np.sqrt(2.0 / fan_in)
            if name.startswith("logits/"):
                std = 4.0 / np.sqrt(fan_in)
            P[name] = (rng.standard_normal(shape) * std).astype(np.float32)
        elif name.endswith("/biases"):
            P[name] = (rng.standard_normal(shape) * 0.1).astype(np.float32)
        elif name.endswith("/gamma"):
            base = 0.3 if "separable_conv3_pointwise" in name and "exit_flow/block2" not in name else 1.0
            P[name] = (base * (1.0 + 0.1 * rng.uniform(-1, 1, shape))).astype(np.float32)
        elif name.endswith("/beta") or name.endswith("/moving_mean"):
            P[name] = (0.05 * rng.standard_normal(shape)).astype(np.float32)
        elif name.endswith("/moving_variance"):
            P[name] = (1.0 + 0.1 * rng.uniform(-1, 1, shape)).astype(np.float32)
        else:
            raise AssertionError(name)
    return P


def synthetic_boxes(n, h, w, seed=3, min_size=40, max_size=300):
    """n proposal boxes [x, y, w, h] (float, like the 1-decimal JSON boxes): sizes ~U(min,max), clipped to the frame."""
    rng = np.random.default_rng(seed)
    bw = rng.uniform(min_size, min(max_size, w - 2), n)
    bh = rng.uniform(min_size, min(max_size, h - 2), n)
    x0 = rng.uniform(0, w - bw)
    y0 = rng.uniform(0, h - bh)
    return np.round(np.stack([x0, y0, bw, bh], 1), 1).astype(np.float32)


def synthetic_masks(n, h, w, seed=4):
    """`n` object masks uint8 [n,h,w] (0/1): unions of ellipses, the shape of MergeTrack's selected proposals; the last mask
    of a set of 3 or more is empty (an object that left the frame)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[:h, :w]
    masks = np.zeros((n, h, w), np.uint8)
    for i in range(n):
        if n >= 3 and i == n - 1:
            break
        for _ in range(3):
            cy, cx = rng.uniform(0, h), rng.uniform(0, w)
            ry, rx = rng.uniform(0.05, 0.3) * h, rng.uniform(0.05, 0.3) * w
            masks[i] |= ((((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2) < 1).astype(np.uint8)
    return masks


# ---- ReID network (MergeTrack/ReID_net_functions.py, ReID_net/configs/run:33-65) -------------------------------------------------
REID_UNITS = (
    [("res0", [128, 128], [3, 3], [2, 1])] + [("res%d" % i, [128, 128], [3, 3], [1, 1]) for i in (1, 2)] +
    [("res3", [256, 256], [3, 3], [2, 1])] + [("res%d" % i, [256, 256], [3, 3], [1, 1]) for i in (4, 5)] +
    [("res6", [512, 512], [3, 3], [2, 1])] + [("res%d" % i, [512, 512], [3, 3], [1, 1]) for i in range(7, 12)] +
    [("res12", [512, 1024], [3, 3], [1, 2]), ("res13", [512, 1024], [3, 3], [1, 1]), ("res14", [512, 1024], [3, 3], [1, 1]),
     ("res15", [512, 1024, 2048], [1, 3, 1], [1, 2, 1]), ("res16", [1024, 2048, 4096], [1, 3, 1], [1, 1, 1])])


def reid_param_shapes():
    """Variable names and shapes of the ReID network (TensorFlow scopes of ReID_net/network/NetworkLayers.py: conv kernels HWIO
    `<layer>/W[i]`, BatchNorm `<layer>/bn[i]/{beta,gamma,mean_ema,var_ema}`, fully connected `<layer>/{W,b}`)."""
    t = OrderedDict()

    def bn(scope, c):
        for v in ("beta", "gamma", "mean_ema", "var_ema"):
            t["%s/%s" % (scope, v)] = (c,)
    t["conv0/W"] = (3, 3, 3, 64)
    cin = 64
    for name, feats, ks, strides in REID_UNITS:
        bn(name + "/bn0", cin)
        if feats[-1] != cin or int(np.prod(strides)) != 1:
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
    for name, fin, fout in (("fc1", 2000, 500), ("fc2", 500, 500), ("outputTriplet", 500, 128)):
        bn(name + "/bn", fin)
        t[name + "/W"] = (fin, fout)
        t[name + "/b"] = (fout,)
    return t


def reid_synthetic_params(seed=0):
    """Seeded ReID-network variables (He-normal kernels; the convolution closing a residual unit is damped so that 17 stacked
    pre-activation units with near-identity BatchNorm statistics keep their activations within a few orders of magnitude)."""
    rng = np.random.default_rng(seed)
    P = OrderedDict()
    last_conv = {"%s/W%d" % (name, len(feats)) for name, feats, _, _ in REID_UNITS}
    for name, shape in reid_param_shapes().items():
        leaf = name.rsplit("/", 1)[1]
        if leaf.startswith("W") and len(shape) == 4:
            std = np.sqrt(2.0 / (shape[0] * shape[1] * shape[2]))
            if name in last_conv:
                std *= 0.5
            P[name] = (rng.standard_normal(shape, dtype=np.float32) * np.float32(std))
        elif leaf == "W":
            P[name] = (rng.standard_normal(shape, dtype=np.float32) * np.float32(np.sqrt(2.0 / shape[0])))
        elif leaf == "b":
            P[name] = (rng.standard_normal(shape) * 0.1).astype(np.float32)
        elif leaf == "gamma":
            P[name] = (1.0 + 0.1 * rng.uniform(-1, 1, shape)).astype(np.float32)
        elif leaf in ("beta", "mean_ema"):
            P[name] = (0.05 * rng.standard_normal(shape)).astype(np.float32)
        elif leaf == "var_ema":
            P[name] = (1.0 + 0.1 * rng.uniform(-1, 1, shape)).astype(np.float32)
        else:
            raise AssertionError(name)
    return P
