"""CPU restatement (torch fp32 / numpy) of the proposal network's inference path -- TEST INFRASTRUCTURE ONLY.

Follows the reference (paths relative to /root/reference/code/proposal_net):

* image normalisation ............ basemodel.py:12-26 (BGR, ImageNet mean/std reversed), train.py:87-90
* ResNet-101 C4 backbone ......... basemodel.py:36-89 (explicit pads (2,3)/(0,1) + VALID, stride 2 inside
                                   the 3x3, shortcut 1x1 s2 on l[:,:,:-1,:-1], frozen BN), config.py:61
* conv5 head ..................... basemodel.py:93-99
* RPN head ....................... model.py:31-51
* anchors ........................ data.py:35-74, utils/generate_anchors.py:40-99, train.py:92-105
* box decode / clip .............. model.py:18-27, 114-139, config.py:77
* proposal generation ............ model.py:170-217 (top-k 1000 -> clip -> drop empty -> NMS 0.7 keep <= 100)
* RoIAlign ....................... model.py:301-374 (tf.image.crop_and_resize 28x28 + 2x2 avg pool)
* Fast R-CNN / second head ....... model.py:378-395, 552-565, train.py:157-189
* inference tail ................. train.py:275-295, model.py:439-491 (incl. the `final_posterior` quirk:
                                   label_probs is gathered by CATEGORY index, train.py:287-288)
* host glue ...................... eval.py:61-110, common.py:35-62,107-119, train.py:388-428

Third-party kernels that are not vendored in the reference are restated from their published
behaviour (TensorFlow 1.8, tensorpack @6fdde15 per proposal_net/README:9):
  tf.image.non_max_suppression ... greedy, descending score, suppress when IoU > thr, IoU formula of
                                   tensorflow/core/kernels/non_max_suppression_op.cc (areas from
                                   min/max corners, 0 if an area <= 0).  Ties: lower index first.
  tf.image.crop_and_resize ....... tensorflow/core/kernels/crop_and_resize_op.cc: in_y = y1*(H-1) + i*
                                   (y2-y1)*(H-1)/(crop-1); outside [0,H-1] -> extrapolation value 0;
                                   lerp of floor/ceil neighbours.
  tensorpack BatchNorm ........... inference form (x - mean) * gamma / sqrt(var + 1e-5) + beta.

PARITY UNPINNED for the network's TensorFlow graph: the reference ships no golden vectors for it and
TensorFlow / tensorpack cannot be imported here.  Pinned pieces: the anchor table in
utils/generate_anchors.py:20-38 (tests/test_oracle_propnet.py), and -- against vectors produced by the
REFERENCE'S OWN functions executed from their source in the build container
(tests/golden/make_reference_function_goldens.py, tests/test_reference_function_goldens.py) --
clip_boxes and CustomResize (common.py), fill_full_mask (eval.py) bit-exactly, and the NMS IoU against
utils/np_box_ops.py to 1e-6.
"""
from __future__ import annotations

import math
from collections import OrderedDict, namedtuple

import numpy as np
import torch
import torch.nn.functional as F

# config.py:61-123
RESNET_NUM_BLOCK = [3, 4, 23, 3]
SHORT_EDGE_SIZE, MAX_SIZE = 800, 1333
ANCHOR_STRIDE = 16
ANCHOR_SIZES = (32, 64, 128, 256, 512)
ANCHOR_RATIOS = (0.5, 1.0, 2.0)
NUM_ANCHOR = 15
BBOX_DECODE_CLIP = float(np.log(MAX_SIZE / 16.0))
RPN_MIN_SIZE = 0
RPN_PROPOSAL_NMS_THRESH = 0.7
TEST_PRE_NMS_TOPK = 1000
TEST_POST_NMS_TOPK = 100
FASTRCNN_BBOX_REG_WEIGHTS = np.array([10, 10, 5, 5], dtype="float32")
FASTRCNN_NMS_THRESH = 0.5
RESULT_SCORE_THRESH = 0.5
RESULTS_PER_IM = 20
BN_EPS = 1e-5
NUM_CLASS = 2            # --agnostic (train.py:622-623)
SECOND_NUM_CLASS = 81    # --second_head


# ---------------------------------------------------------------------------------------------
# parameters (tensorpack variable names)
# ---------------------------------------------------------------------------------------------
def maskrcnn_param_shapes(num_class=NUM_CLASS):
    """model.py:495-509 (tensorpack Deconv2D 'deconv': W [kh, kw, out, in]; Conv2D 'conv': W HWIO)."""
    return OrderedDict([("maskrcnn/deconv/W", (2, 2, 256, 2048)), ("maskrcnn/deconv/b", (256,)),
                        ("maskrcnn/conv/W", (1, 1, 256, num_class - 1)), ("maskrcnn/conv/b", (num_class - 1,))])


def propnet_param_shapes(num_blocks=RESNET_NUM_BLOCK, num_class=NUM_CLASS, second_num_class=SECOND_NUM_CLASS):
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
        if g == 2:
            t["rpn/conv0/W"] = (3, 3, 1024, 1024)
            t["rpn/conv0/b"] = (1024,)
            t["rpn/class/W"] = (1, 1, 1024, NUM_ANCHOR)
            t["rpn/class/b"] = (NUM_ANCHOR,)
            t["rpn/box/W"] = (1, 1, 1024, 4 * NUM_ANCHOR)
            t["rpn/box/b"] = (4 * NUM_ANCHOR,)
    t["fastrcnn/class/W"] = (2048, num_class)
    t["fastrcnn/class/b"] = (num_class,)
    t["fastrcnn/box/W"] = (2048, (num_class - 1) * 4)
    t["fastrcnn/box/b"] = ((num_class - 1) * 4,)
    if second_num_class:
        t["secondclassification/class/W"] = (2048, second_num_class)
        t["secondclassification/class/b"] = (second_num_class,)
    return t


# ---------------------------------------------------------------------------------------------
# host glue
# ---------------------------------------------------------------------------------------------
def custom_resize_shape(h, w, size=SHORT_EDGE_SIZE, max_size=MAX_SIZE):
    """common.py:49-62"""
    scale = size * 1.0 / min(h, w)
    if h < w:
        newh, neww = size, scale * w
    else:
        newh, neww = scale * h, size
    if max(newh, neww) > max_size:
        scale = max_size * 1.0 / max(newh, neww)
        newh = newh * scale
        neww = neww * scale
    return int(newh + 0.5), int(neww + 0.5)


def clip_boxes_np(boxes, shape):
    """common.py:107-119"""
    orig_shape = boxes.shape
    boxes = boxes.reshape([-1, 4])
    h, w = shape
    boxes[:, [0, 1]] = np.maximum(boxes[:, [0, 1]], 0)
    boxes[:, 2] = np.minimum(boxes[:, 2], w)
    boxes[:, 3] = np.minimum(boxes[:, 3], h)
    return boxes.reshape(orig_shape)


SecondDetectionResult = namedtuple(
    "SecondDetectionResult",
    ["box", "score", "class_id", "posterior", "mask", "second_class_id", "second_posterior", "feature_fastrcnn_pooled"])


def detect_one_image(img, model_func, size=SHORT_EDGE_SIZE, max_size=MAX_SIZE):
    """eval.py:61-110 with USE_SECOND_HEAD, no feature extraction; masks (a seventh model output) are pasted by fill_full_mask."""
    import cv2
    orig_shape = img.shape[:2]
    newh, neww = custom_resize_shape(orig_shape[0], orig_shape[1], size, max_size)
    resized_img = cv2.resize(img, (neww, newh), interpolation=cv2.INTER_LINEAR)
    scale = (resized_img.shape[0] * 1.0 / img.shape[0] + resized_img.shape[1] * 1.0 / img.shape[1]) / 2
    boxes, probs, labels, posteriors, second_labels, second_posteriors, *masks = model_func(resized_img)
    boxes = boxes / scale
    boxes = clip_boxes_np(boxes, orig_shape)
    masks = [fill_full_mask(b, m, orig_shape) for b, m in zip(boxes, masks[0])] if masks else [None] * len(boxes)
    features = [None for _ in range(labels.size)]
    return [SecondDetectionResult(*args) for args in
            zip(boxes, probs, labels, posteriors, masks, second_labels, second_posteriors, features)]


def convert_results_to_json(results):
    """train.py:388-428 (the emitted record has only bbox xywh (1 dp) and score (2 dp))."""
    out = []
    for r in results:
        box = np.array(r.box, dtype=np.float64).copy()
        box[2] -= box[0]
        box[3] -= box[1]
        out.append({"bbox": [float(round(x, 1)) for x in box], "score": float(round(float(r.score), 2))})
    return out


# ---------------------------------------------------------------------------------------------
# anchors
# ---------------------------------------------------------------------------------------------
def _whctrs(a):
    w = a[2] - a[0] + 1
    h = a[3] - a[1] + 1
    return w, h, a[0] + 0.5 * (w - 1), a[1] + 0.5 * (h - 1)


def _mkanchors(ws, hs, xc, yc):
    ws = ws[:, np.newaxis]
    hs = hs[:, np.newaxis]
    return np.hstack((xc - 0.5 * (ws - 1), yc - 0.5 * (hs - 1), xc + 0.5 * (ws - 1), yc + 0.5 * (hs - 1)))


def generate_anchors(base_size=16, ratios=(0.5, 1, 2), scales=2 ** np.arange(3, 6)):
    """utils/generate_anchors.py:40-99"""
    ratios = np.asarray(ratios, dtype=np.float64)
    scales = np.asarray(scales, dtype=np.float64)
    base = np.array([1, 1, base_size, base_size], dtype="float32") - 1
    w, h, xc, yc = _whctrs(base)
    size = w * h
    ws = np.round(np.sqrt(size / ratios))
    hs = np.round(ws * ratios)
    ratio_anchors = _mkanchors(ws, hs, xc, yc)
    out = []
    for i in range(ratio_anchors.shape[0]):
        w, h, xc, yc = _whctrs(ratio_anchors[i])
        out.append(_mkanchors(w * scales, h * scales, xc, yc))
    return np.vstack(out)


def get_all_anchors(stride=ANCHOR_STRIDE, sizes=ANCHOR_SIZES, max_size=MAX_SIZE):
    """data.py:35-74 -> [FS, FS, 15, 4] float32, x2/y2 + 1"""
    cell = generate_anchors(stride, scales=np.array(sizes, dtype=np.float64) / stride,
                            ratios=np.array(ANCHOR_RATIOS, dtype=np.float64))
    fs = max_size // stride
    shifts = np.arange(0, fs) * stride
    sx, sy = np.meshgrid(shifts, shifts)
    sx, sy = sx.flatten(), sy.flatten()
    shifts = np.vstack((sx, sy, sx, sy)).transpose()
    K, A = shifts.shape[0], cell.shape[0]
    field = (cell.reshape((1, A, 4)) + shifts.reshape((1, K, 4)).transpose((1, 0, 2))).reshape((fs, fs, A, 4))
    assert np.all(field == field.astype("int32"))
    field = field.astype("float32")
    field[:, :, :, [2, 3]] += 1
    return field


# ---------------------------------------------------------------------------------------------
# TensorFlow kernels restated
# ---------------------------------------------------------------------------------------------
def tf_iou(boxes, i, j):
    """ComputeIOU of non_max_suppression_op.cc in fp32 (box layout irrelevant: min/max of the corner pairs)."""
    f = np.float32
    b = boxes
    ymin_i, xmin_i = min(b[i, 0], b[i, 2]), min(b[i, 1], b[i, 3])
    ymax_i, xmax_i = max(b[i, 0], b[i, 2]), max(b[i, 1], b[i, 3])
    ymin_j, xmin_j = min(b[j, 0], b[j, 2]), min(b[j, 1], b[j, 3])
    ymax_j, xmax_j = max(b[j, 0], b[j, 2]), max(b[j, 1], b[j, 3])
    area_i = f(f(ymax_i - ymin_i) * f(xmax_i - xmin_i))
    area_j = f(f(ymax_j - ymin_j) * f(xmax_j - xmin_j))
    if area_i <= 0 or area_j <= 0:
        return f(0.0)
    iy0, ix0 = max(ymin_i, ymin_j), max(xmin_i, xmin_j)
    iy1, ix1 = min(ymax_i, ymax_j), min(xmax_i, xmax_j)
    inter = f(max(f(iy1 - iy0), f(0.0)) * max(f(ix1 - ix0), f(0.0)))
    return f(inter / f(f(area_i + area_j) - inter))


def tf_non_max_suppression(boxes, scores, max_output_size, iou_threshold):
    """Greedy NMS: candidates by descending score (ties: lower index first), a candidate is kept when its
    IoU with every already kept box is <= iou_threshold.  Returns int32 indices into boxes."""
    boxes = np.asarray(boxes, dtype=np.float32).reshape(-1, 4)
    scores = np.asarray(scores, dtype=np.float32).reshape(-1)
    order = np.lexsort((np.arange(scores.size), -scores.astype(np.float64)))
    thr = np.float32(iou_threshold)
    keep = []
    for idx in order:
        if len(keep) >= max_output_size:
            break
        ok = True
        for k in reversed(keep):
            if tf_iou(boxes, idx, k) > thr:
                ok = False
                break
        if ok:
            keep.append(int(idx))
    return np.asarray(keep, dtype=np.int32)


def tf_crop_and_resize(image, boxes, crop):
    """tf.image.crop_and_resize(image[1,H,W,C], boxes[n,4] normalised y1x1y2x2, zeros, [crop,crop]), bilinear,
    extrapolation_value 0.  image: torch [C,H,W] float32.  Returns [n,C,crop,crop]."""
    C, H, W = image.shape
    n = boxes.shape[0]
    f = np.float32
    out = torch.zeros((n, C, crop, crop), dtype=torch.float32)
    for b in range(n):
        y1, x1, y2, x2 = (f(v) for v in boxes[b])
        hs = f(f(f(y2 - y1) * f(H - 1)) / f(crop - 1)) if crop > 1 else f(0)
        ws = f(f(f(x2 - x1) * f(W - 1)) / f(crop - 1)) if crop > 1 else f(0)
        ys = [f(f(y1 * f(H - 1)) + f(f(i) * hs)) if crop > 1 else f(0.5) * f(y1 + y2) * f(H - 1) for i in range(crop)]
        xs = [f(f(x1 * f(W - 1)) + f(f(i) * ws)) if crop > 1 else f(0.5) * f(x1 + x2) * f(W - 1) for i in range(crop)]
        yv = [(0 <= y <= H - 1) for y in ys]
        xv = [(0 <= x <= W - 1) for x in xs]
        ty = np.array([math.floor(y) if v else 0 for y, v in zip(ys, yv)], dtype=np.int64)
        by = np.array([math.ceil(y) if v else 0 for y, v in zip(ys, yv)], dtype=np.int64)
        lx = np.array([math.floor(x) if v else 0 for x, v in zip(xs, xv)], dtype=np.int64)
        rx = np.array([math.ceil(x) if v else 0 for x, v in zip(xs, xv)], dtype=np.int64)
        yl = torch.tensor([f(y - f(t)) for y, t in zip(ys, ty)], dtype=torch.float32).view(1, crop, 1)
        xl = torch.tensor([f(x - f(l)) for x, l in zip(xs, lx)], dtype=torch.float32).view(1, 1, crop)
        tl = image[:, ty][:, :, lx]
        tr = image[:, ty][:, :, rx]
        bl = image[:, by][:, :, lx]
        br = image[:, by][:, :, rx]
        top = tl + (tr - tl) * xl
        bot = bl + (br - bl) * xl
        val = top + (bot - top) * yl
        mask = torch.tensor(yv, dtype=torch.bool).view(1, crop, 1) & torch.tensor(xv, dtype=torch.bool).view(1, 1, crop)
        out[b] = torch.where(mask, val, torch.zeros_like(val))
    return out


# ---------------------------------------------------------------------------------------------
# network pieces
# ---------------------------------------------------------------------------------------------
def image_preprocess(img_hwc):
    """basemodel.py:12-26 + train.py:87-90: float32 HWC BGR (0..255) -> [1,3,H,W] normalised"""
    x = torch.as_tensor(np.asarray(img_hwc), dtype=torch.float32)
    x = x * np.float32(1.0 / 255)
    mean = torch.tensor([0.485, 0.456, 0.406][::-1], dtype=torch.float32)
    std = torch.tensor([0.229, 0.224, 0.225][::-1], dtype=torch.float32)
    x = (x - mean) / std
    return x.permute(2, 0, 1).unsqueeze(0).contiguous()


def _w(P, name):
    """HWIO -> OIHW"""
    return torch.as_tensor(P[name]).permute(3, 2, 0, 1).contiguous()


def _bn(P, scope, x):
    g, b = torch.as_tensor(P[scope + "/bn/gamma"]), torch.as_tensor(P[scope + "/bn/beta"])
    m, v = torch.as_tensor(P[scope + "/bn/mean/EMA"]), torch.as_tensor(P[scope + "/bn/variance/EMA"])
    scale = g / torch.sqrt(v + BN_EPS)
    return (x - m.view(1, -1, 1, 1)) * scale.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)


def _conv_bn(P, scope, x, stride=1, padding=0, relu=True):
    y = _bn(P, scope, F.conv2d(x, _w(P, scope + "/W"), None, stride=stride, padding=padding))
    return F.relu(y) if relu else y


def resnet_bottleneck(P, scope, l, ch_out, stride):
    """basemodel.py:36-60"""
    shortcut = l
    l = _conv_bn(P, scope + "/conv1", l)
    if stride == 2:
        l = F.pad(l, (0, 1, 0, 1))
        l = _conv_bn(P, scope + "/conv2", l, stride=2, padding=0)
    else:
        l = _conv_bn(P, scope + "/conv2", l, stride=1, padding=1)
    l = _conv_bn(P, scope + "/conv3", l, relu=False)
    if shortcut.shape[1] != ch_out * 4:
        if stride == 2:
            shortcut = shortcut[:, :, :-1, :-1]
        shortcut = _conv_bn(P, scope + "/convshortcut", shortcut, stride=stride, relu=False)
    return F.relu(l + shortcut)


def resnet_group(P, name, l, features, count, stride):
    for i in range(count):
        l = resnet_bottleneck(P, "%s/block%d" % (name, i), l, features, stride if i == 0 else 1)
    return l


def pretrained_resnet_conv4(P, image, num_blocks):
    """basemodel.py:74-89"""
    l = F.pad(image, (2, 3, 2, 3))
    l = _conv_bn(P, "conv0", l, stride=2)
    l = F.pad(l, (0, 1, 0, 1))
    l = F.max_pool2d(l, 3, 2)
    l = resnet_group(P, "group0", l, 64, num_blocks[0], 1)
    l = resnet_group(P, "group1", l, 128, num_blocks[1], 2)
    l = resnet_group(P, "group2", l, 256, num_blocks[2], 2)
    return l


def rpn_head(P, featuremap):
    """model.py:31-51 -> label_logits [fH,fW,NA], box_logits [fH,fW,NA,4]"""
    hidden = F.relu(F.conv2d(featuremap, _w(P, "rpn/conv0/W"), torch.as_tensor(P["rpn/conv0/b"]), padding=1))
    label = F.conv2d(hidden, _w(P, "rpn/class/W"), torch.as_tensor(P["rpn/class/b"]))
    box = F.conv2d(hidden, _w(P, "rpn/box/W"), torch.as_tensor(P["rpn/box/b"]))
    label = label.permute(0, 2, 3, 1)[0]
    fh, fw = box.shape[2], box.shape[3]
    box = box.permute(0, 2, 3, 1).reshape(fh, fw, NUM_ANCHOR, 4)
    return label, box


def decode_bbox_target(box_predictions, anchors):
    """model.py:114-139 (torch float32 tensors [...,4])"""
    shp = anchors.shape
    p = box_predictions.reshape(-1, 2, 2)
    txty, twth = p[:, 0], p[:, 1]
    a = anchors.reshape(-1, 2, 2)
    a1, a2 = a[:, 0], a[:, 1]
    waha = a2 - a1
    xaya = (a2 + a1) * 0.5
    wbhb = torch.exp(torch.minimum(twth, torch.tensor(BBOX_DECODE_CLIP, dtype=torch.float32))) * waha
    xbyb = txty * waha + xaya
    x1y1 = xbyb - wbhb * 0.5
    x2y2 = xbyb + wbhb * 0.5
    return torch.cat([x1y1, x2y2], dim=1).reshape(shp)


def clip_boxes_t(boxes, h, w):
    """model.py:18-27: max(.,0) then min with [w,h,w,h]"""
    boxes = torch.clamp(boxes, min=0.0)
    m = torch.tensor([w, h, w, h], dtype=torch.float32)
    return torch.minimum(boxes, m)


def generate_rpn_proposals(boxes, scores, h, w, pre_topk=TEST_PRE_NMS_TOPK, post_topk=TEST_POST_NMS_TOPK):
    """model.py:170-217.  top_k(sorted=False) order is unspecified in the reference; the candidates are kept in
    (score desc, index asc) order here, which is also what the NMS consumes."""
    scores_np = scores.numpy()
    k = min(pre_topk, scores_np.size)
    order = np.lexsort((np.arange(scores_np.size), -scores_np.astype(np.float64)))[:k]
    topk_scores = scores[order]
    topk_boxes = clip_boxes_t(boxes[order], h, w)
    wbhb = topk_boxes[:, 2:] - topk_boxes[:, :2]
    valid = (wbhb > RPN_MIN_SIZE).all(dim=1)
    vb = topk_boxes[valid]
    vs = topk_scores[valid]
    y1x1y2x2 = vb[:, [1, 0, 3, 2]]
    keep = tf_non_max_suppression(y1x1y2x2.numpy(), vs.numpy(), post_topk, RPN_PROPOSAL_NMS_THRESH)
    keep_t = torch.as_tensor(keep, dtype=torch.long)
    return vb[keep_t], vs[keep_t], {"topk_indices": order[valid.numpy()], "nms_keep": keep}


def roi_align(featuremap, boxes, output_shape=14):
    """model.py:301-374: crop_and_resize at 2x resolution, then 2x2 average pooling."""
    crop = output_shape * 2
    C, H, W = featuremap.shape[1:]
    f = np.float32
    b = boxes.numpy().astype(np.float32)
    x0, y0, x1, y1 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    sw = (x1 - x0) / f(crop)
    sh = (y1 - y0) / f(crop)
    nx0 = (x0 + sw / f(2) - f(0.5)) / f(W - 1)
    ny0 = (y0 + sh / f(2) - f(0.5)) / f(H - 1)
    nw = sw * f(crop - 1) / f(W - 1)
    nh = sh * f(crop - 1) / f(H - 1)
    tfb = np.stack([ny0, nx0, ny0 + nh, nx0 + nw], axis=1).astype(np.float32)
    ret = tf_crop_and_resize(featuremap[0], tfb, crop)
    return F.avg_pool2d(ret, 2, 2)


def resnet_conv5(P, x, num_block):
    return resnet_group(P, "group3", x, 512, num_block, 2)


def fastrcnn_predictions(boxes, probs):
    """model.py:439-491.  boxes [n, #cat, 4], probs [n, #class] -> (pred_indices [m,2] = (box, cat), final_probs [m]).
    Selected entries are returned in (cat, box) order, top_k(sorted=False) order being unspecified."""
    ncat = boxes.shape[1]
    n = boxes.shape[0]
    masks = np.zeros((ncat, n), dtype=bool)
    for c in range(ncat):
        prob = probs[:, c + 1].numpy()
        box = boxes[:, c].numpy()
        ids = np.nonzero(prob > np.float32(RESULT_SCORE_THRESH))[0]
        sel = tf_non_max_suppression(box[ids], prob[ids], RESULTS_PER_IM, FASTRCNN_NMS_THRESH)
        masks[c, ids[sel]] = True
    cat_ids, box_ids = np.nonzero(masks)
    p = probs[:, 1:].numpy().T[masks]
    k = min(RESULTS_PER_IM, p.size)
    order = np.lexsort((np.arange(p.size), -p.astype(np.float64)))[:k]
    order = np.sort(order)
    pred = np.stack([box_ids[order], cat_ids[order]], axis=1).astype(np.int64).reshape(-1, 2)
    return pred, p[order].astype(np.float32)


def propnet_forward(P, img_hwc, num_blocks=RESNET_NUM_BLOCK, return_intermediates=False, max_rois=None):
    """Inference branch of Model._build_graph (train.py:107-189, 274-295) for one already resized image.
    Returns the six `get_model_output_names()` arrays (train.py:52-62)."""
    inter = OrderedDict()
    h, w = img_hwc.shape[:2]
    image = image_preprocess(img_hwc)
    fm = pretrained_resnet_conv4(P, image, num_blocks[:3])
    inter["featuremap"] = fm
    label_logits, box_logits = rpn_head(P, fm)
    inter["rpn_label_logits"], inter["rpn_box_logits"] = label_logits, box_logits
    fh, fw = h // ANCHOR_STRIDE, w // ANCHOR_STRIDE
    assert (fh, fw) == tuple(fm.shape[2:]), ((fh, fw), fm.shape)
    anchors = torch.as_tensor(get_all_anchors()[:fh, :fw])
    decoded = decode_bbox_target(box_logits, anchors)
    inter["rpn_decoded_boxes"] = decoded
    prop_boxes, prop_scores, dbg = generate_rpn_proposals(decoded.reshape(-1, 4), label_logits.reshape(-1), h, w)
    inter["proposal_boxes"], inter["proposal_scores"] = prop_boxes, prop_scores
    inter["topk_indices"], inter["nms_keep"] = dbg["topk_indices"], dbg["nms_keep"]
    n = prop_boxes.shape[0]
    ncat = NUM_CLASS - 1
    if n > 0:
        roi = roi_align(fm, prop_boxes * np.float32(1.0 / ANCHOR_STRIDE), 14)
        inter["roi_resized"] = roi
        feat = resnet_conv5(P, roi, num_blocks[-1])
        inter["feature_fastrcnn"] = feat
        pooled = feat.mean(dim=(2, 3))
        cls = pooled @ torch.as_tensor(P["fastrcnn/class/W"]) + torch.as_tensor(P["fastrcnn/class/b"])
        box = (pooled @ torch.as_tensor(P["fastrcnn/box/W"]) + torch.as_tensor(P["fastrcnn/box/b"])).reshape(-1, ncat, 4)
        second = pooled @ torch.as_tensor(P["secondclassification/class/W"]) + torch.as_tensor(P["secondclassification/class/b"])
    else:
        cls, box, second = torch.zeros(0, NUM_CLASS), torch.zeros(0, ncat, 4), torch.zeros(0, SECOND_NUM_CLASS)
    inter["fastrcnn_label_logits"], inter["fastrcnn_box_logits"], inter["second_logits"] = cls, box, second
    label_probs = torch.softmax(cls, dim=1)
    anchors2 = prop_boxes.unsqueeze(1).repeat(1, ncat, 1)
    dec = decode_bbox_target(box / torch.as_tensor(FASTRCNN_BBOX_REG_WEIGHTS), anchors2)
    dec = clip_boxes_t(dec, h, w)
    inter["fastrcnn_all_probs"], inter["fastrcnn_all_boxes"] = label_probs, dec
    pred, final_probs = fastrcnn_predictions(dec, label_probs)
    final_boxes = dec.numpy()[pred[:, 0], pred[:, 1]].reshape(-1, 4)
    final_labels = (pred[:, 1] + 1).astype(np.int64)
    only = pred[:, 1]                                   # train.py:287 -- the CATEGORY index, as in the reference
    final_posterior = label_probs.numpy()[only].reshape(-1, NUM_CLASS)
    second_probs = torch.softmax(second, dim=1).numpy()
    second_final_posterior = second_probs[only].reshape(-1, SECOND_NUM_CLASS)
    second_final_labels = (np.argmax(final_posterior, axis=-1) + 1).astype(np.int64).reshape(-1)
    out = (final_boxes.astype(np.float32), final_probs, final_labels, final_posterior.astype(np.float32),
           second_final_labels, second_final_posterior.astype(np.float32))
    inter["pred_indices"] = pred
    if "maskrcnn/deconv/W" in P:      # config.MODE_MASK: `final_masks` is the seventh output (train.py:52-62, 297-309)
        out = out + (final_masks(P, fm, out[0], final_labels, num_blocks),)
    if return_intermediates:
        return out, inter
    return out


# ---- Mask R-CNN mask head (MODE_MASK, inactive under simple_run.sh's --forward) --------------------------
def maskrcnn_head(P, feature):
    """model.py:495-509: Deconv2D(256, 2, stride 2) + ReLU, Conv2D(num_class - 1, 1).  feature [N,2048,7,7] -> [N,#cat,14,14].
    tf conv2d_transpose: out[2y+dy, 2x+dx, co] += in[y, x, ci] * W[dy, dx, co, ci]  (== torch conv_transpose2d, weight [ci,co,dy,dx])."""
    wt = torch.as_tensor(P["maskrcnn/deconv/W"]).permute(3, 2, 0, 1).contiguous()
    l = F.relu(F.conv_transpose2d(feature, wt, torch.as_tensor(P["maskrcnn/deconv/b"]), stride=2))
    w2 = torch.as_tensor(P["maskrcnn/conv/W"]).permute(3, 2, 0, 1).contiguous()
    return F.conv2d(l, w2, torch.as_tensor(P["maskrcnn/conv/b"]))


def final_masks(P, featuremap, final_boxes, final_labels, num_blocks=RESNET_NUM_BLOCK):
    """train.py:297-309: RoIAlign on the FINAL boxes, conv5 again, mask head, logits of each box's own category, sigmoid.
    -> float32 [n, 14, 14]."""
    n = len(final_boxes)
    if n == 0:
        return np.zeros((0, 14, 14), np.float32)
    roi = roi_align(featuremap, torch.as_tensor(np.asarray(final_boxes, np.float32)) * np.float32(1.0 / ANCHOR_STRIDE), 14)
    logits = maskrcnn_head(P, resnet_conv5(P, roi, num_blocks[-1]))
    idx = torch.as_tensor(np.asarray(final_labels, np.int64) - 1)
    return torch.sigmoid(logits[torch.arange(n), idx]).numpy().astype(np.float32)


def fill_full_mask(box, mask, shape):
    """eval.py:35-58: paste the MxM mask into its box on a zero canvas of `shape` = (h, w); cv2.resize (float32 INTER_LINEAR;
    INTER_AREA for an exact 2x down-scale, as OpenCV substitutes) restated by oracle/cv_resize_oracle.py, threshold 0.5."""
    from oracle import cv_resize_oracle as RZ
    box = np.asarray(box, np.float32)
    x0, y0 = (int(v) for v in box[:2] + np.float32(0.5))
    x1, y1 = (int(v) for v in box[2:] - np.float32(0.5))
    x1, y1 = max(x0, x1), max(y0, y1)
    w, h = x1 + 1 - x0, y1 + 1 - y0
    m = np.asarray(mask, np.float32)
    M = m.shape[0]
    if (w, h) == (M, M):
        r = m
    elif M == 2 * w and M == 2 * h:
        r = ((m[0::2, 0::2] + m[0::2, 1::2] + m[1::2, 0::2] + m[1::2, 1::2]) * np.float32(0.25)).astype(np.float32)
    else:
        r = RZ.resize_linear_f32(m, h, w)
    ret = np.zeros(shape, np.uint8)
    ret[y0:y1 + 1, x0:x1 + 1] = (r > 0.5).astype(np.uint8)
    return ret
