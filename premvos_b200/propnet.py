"""Host-side mirror of the reference's proposal call surface, backed by the CUDA library.

Reference surface reproduced here (paths relative to code/proposal_net):
  pred_func(img) -> (final_boxes, final_probs, final_labels, final_posterior, second_final_labels,
                     second_final_posterior)                      train.py:52-62, 653-657 (OfflinePredictor)
  detect_one_image(img, model_func) -> [SecondDetectionResult]     eval.py:24-26, 61-110
  CustomResize(800, 1333)                                          common.py:35-62
  convert_results_to_json(results, img_idx)                        train.py:388-428

All tensor arithmetic runs in libpremvos_b200.so; there is no TensorFlow, PyTorch-op or CPU fallback.
"""
from __future__ import annotations

import ctypes
from collections import OrderedDict, namedtuple

import numpy as np

from . import _lib
from .synth import propnet_param_shapes

SHORT_EDGE_SIZE, MAX_SIZE, RESULTS_PER_IM = 800, 1333, 20

SecondDetectionResult = namedtuple(
    "SecondDetectionResult",
    ["box", "score", "class_id", "posterior", "mask", "second_class_id", "second_posterior", "feature_fastrcnn_pooled"])


class ProposalNet:
    """The `--forward --agnostic --second_head` graph.  `load_params(dict)` takes tensorpack variable names
    (as `get_model_loader(path)` would restore them); device handles are created per resized-image shape."""

    def __init__(self, num_blocks=(3, 4, 23, 3), num_class=2, second_num_class=81, mode_mask=False):
        """mode_mask=True (config.MODE_MASK): the graph also has the Mask R-CNN mask head (model.py:495-509) and pred_func
        returns `final_masks` [n,14,14] as a seventh output (train.py:52-62)."""
        self.num_blocks = tuple(num_blocks)
        self.num_class = num_class
        self.second_num_class = second_num_class
        self.mode_mask = bool(mode_mask)
        self._shapes = propnet_param_shapes(self.num_blocks, num_class, second_num_class, self.mode_mask)
        self._params = OrderedDict()
        self._handles = {}

    def load_params(self, params):
        missing = [k for k in self._shapes if k not in params]
        unexpected = [k for k in params if k not in self._shapes]
        if missing or unexpected:
            raise RuntimeError("proposal_net variables: missing %s, unexpected %s" % (missing[:5], unexpected[:5]))
        for k, shp in self._shapes.items():
            v = np.ascontiguousarray(params[k], dtype=np.float32)
            if tuple(v.shape) != tuple(shp):
                raise RuntimeError("size mismatch for %s: got %s, expected %s" % (k, tuple(v.shape), tuple(shp)))
            self._params[k] = v
        self._drop_handles()
        return self

    def _drop_handles(self):
        for h in self._handles.values():
            _lib.lib().premvos_propnet_destroy(h)
        self._handles = {}

    def __del__(self):
        try:
            self._drop_handles()
        except Exception:
            pass

    def _handle(self, H, W, batch=1):
        """Device handle for resized images of H x W, `batch` frames per forward (premvos_propnet_set_option "batch")."""
        import torch
        key = (torch.cuda.current_device(), H, W, int(batch))   # per device: a handle owns device buffers and a captured graph
        if key in self._handles:
            return self._handles[key]
        if not self._params:
            raise RuntimeError("ProposalNet: load_params() first")
        L = _lib.lib()
        h = ctypes.c_void_p()
        _lib.check(L.premvos_propnet_create(ctypes.byref(h), H, W, self.num_class, self.second_num_class))
        try:
            for g, nb in enumerate(self.num_blocks):
                _lib.check(L.premvos_propnet_set_option(h, b"num_blocks%d" % g, int(nb)))
            if self.mode_mask:
                _lib.check(L.premvos_propnet_set_option(h, b"mode_mask", 1))
            if batch != 1:
                _lib.check(L.premvos_propnet_set_option(h, b"batch", int(batch)))
            for k, v in self._params.items():
                _lib.check(L.premvos_propnet_set_param(h, k.encode(), v.ctypes.data_as(ctypes.c_void_p), v.size))
            _lib.check(L.premvos_propnet_finalize(h))
        except Exception:
            L.premvos_propnet_destroy(h)
            raise
        self._handles[key] = h
        return h

    def launches_per_forward(self, H, W, batch=1):
        return int(_lib.lib().premvos_propnet_launches_per_forward(self._handle(H, W, batch)))

    # -- pred_func -----------------------------------------------------------------------------------
    def __call__(self, img):
        """img: [h,w,3] BGR uint8 or float32 (0..255), already resized (eval.py:75-78)."""
        img = np.ascontiguousarray(img, dtype=np.float32)
        if img.ndim != 3 or img.shape[2] != 3:
            raise ValueError("expected an [h,w,3] BGR image, got %s" % (img.shape,))
        H, W = img.shape[:2]
        h = self._handle(H, W)
        n = ctypes.c_int()
        boxes = np.zeros((RESULTS_PER_IM, 4), np.float32)
        probs = np.zeros((RESULTS_PER_IM,), np.float32)
        labels = np.zeros((RESULTS_PER_IM,), np.int64)
        post = np.zeros((RESULTS_PER_IM, self.num_class), np.float32)
        slabels = np.zeros((RESULTS_PER_IM,), np.int64)
        spost = np.zeros((RESULTS_PER_IM, max(self.second_num_class, 1)), np.float32)
        vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        _lib.check(_lib.lib().premvos_propnet_forward_host(h, vp(img), ctypes.byref(n), vp(boxes), vp(probs), vp(labels),
                                                           vp(post), vp(slabels), vp(spost)))
        m = n.value
        out = (boxes[:m].copy(), probs[:m].copy(), labels[:m].copy(), post[:m].copy(), slabels[:m].copy(),
               spost[:m, :self.second_num_class].copy())
        return out + (self.read_masks(H, W, m),) if self.mode_mask else out

    def read_masks(self, H, W, rows, batch=1, image=0):
        """`final_masks` of the last forward: float32 [rows,14,14] (sigmoid outputs, row i belongs to final_boxes[i])."""
        import torch
        masks = np.zeros((int(rows), 14, 14), np.float32)
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().premvos_propnet_read_masks(self._handle(H, W, batch), st, int(image),
                                                         masks.ctypes.data_as(ctypes.c_void_p), int(rows)))
        return masks

    # -- resident-pipeline entry points (device tensors, no synchronisation until read_results) -----------
    def forward_device(self, img):
        """img: CUDA tensor [h,w,3] BGR, uint8 or float32 (0..255), already resized -- or [B,h,w,3]: B frames in one launch
        group (a batched handle).  Enqueues the whole graph on the current torch stream; fetch with read_results(h, w[, B, b])
        or copy_results_device."""
        import torch
        if not isinstance(img, torch.Tensor) or not img.is_cuda or img.dtype not in (torch.uint8, torch.float32):
            raise TypeError("img must be a CUDA uint8 / float32 tensor (this build has no CPU path)")
        if img.dim() not in (3, 4) or img.shape[-1] != 3 or not img.is_contiguous():
            raise ValueError("expected a contiguous [h,w,3] or [B,h,w,3] BGR image, got %s" % (tuple(img.shape),))
        H, W = int(img.shape[-3]), int(img.shape[-2])
        batch = 1 if img.dim() == 3 else int(img.shape[0])
        with torch.cuda.device(img.device):
            h = self._handle(H, W, batch)
            st = torch.cuda.current_stream().cuda_stream
            fn = _lib.lib().premvos_propnet_forward_u8 if img.dtype == torch.uint8 else _lib.lib().premvos_propnet_forward
            _lib.check(fn(h, img.data_ptr(), st))

    def copy_results_device(self, H, W, count, boxes, probs=None, batch=1):
        """Device-to-device copy of the last forward_device's results on the current torch stream (no synchronisation):
        count CUDA int32 [B], boxes CUDA float32 [B,20,4], probs CUDA float32 [B,20] (contiguous; B = batch, may be squeezed
        for batch 1)."""
        import torch
        for t, numel in ((count, batch), (boxes, batch * RESULTS_PER_IM * 4), (probs, batch * RESULTS_PER_IM)):
            if t is not None and (not t.is_cuda or not t.is_contiguous() or t.numel() != numel):
                raise ValueError("copy_results_device: outputs must be contiguous CUDA tensors for %d image(s)" % batch)
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().premvos_propnet_copy_results(self._handle(H, W, batch), st, count.data_ptr(), boxes.data_ptr(),
                                                           None if probs is None else probs.data_ptr(), None))

    def read_results(self, H, W, batch=1, image=0):
        """Synchronises the current torch stream and returns the six arrays of pred_func for the last forward_device
        (image `image` of a batched forward)."""
        import torch
        h = self._handle(H, W, batch)
        n = ctypes.c_int()
        boxes = np.zeros((RESULTS_PER_IM, 4), np.float32)
        probs = np.zeros((RESULTS_PER_IM,), np.float32)
        labels = np.zeros((RESULTS_PER_IM,), np.int64)
        post = np.zeros((RESULTS_PER_IM, self.num_class), np.float32)
        slabels = np.zeros((RESULTS_PER_IM,), np.int64)
        spost = np.zeros((RESULTS_PER_IM, max(self.second_num_class, 1)), np.float32)
        vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().premvos_propnet_read_results_image(h, st, int(image), ctypes.byref(n), vp(boxes), vp(probs), vp(labels),
                                                                 vp(post), vp(slabels), vp(spost)))
        m = n.value
        return (boxes[:m].copy(), probs[:m].copy(), labels[:m].copy(), post[:m].copy(), slabels[:m].copy(),
                spost[:m, :self.second_num_class].copy())

    def get_tensor(self, name, H, W, batch=1):
        L = _lib.lib()
        h = self._handle(H, W, batch)
        n = ctypes.c_int64()
        _lib.check(L.premvos_propnet_get_tensor(h, name.encode(), None, ctypes.byref(n)))
        buf = np.empty(n.value, dtype=np.float32)
        if n.value:
            _lib.check(L.premvos_propnet_get_tensor(h, name.encode(), buf.ctypes.data_as(ctypes.c_void_p), ctypes.byref(n)))
        return buf


def custom_resize_shape(h, w, size=SHORT_EDGE_SIZE, max_size=MAX_SIZE):
    """CustomResize._get_augment_params (common.py:49-62)"""
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


def clip_boxes(boxes, shape):
    """common.py:107-119"""
    orig_shape = boxes.shape
    boxes = boxes.reshape([-1, 4])
    h, w = shape
    boxes[:, [0, 1]] = np.maximum(boxes[:, [0, 1]], 0)
    boxes[:, 2] = np.minimum(boxes[:, 2], w)
    boxes[:, 3] = np.minimum(boxes[:, 3], h)
    return boxes.reshape(orig_shape)


def fill_full_masks(boxes, masks, shape):
    """eval.py:35-58 for all boxes of an image, on the device: boxes [n,4] x1y1x2y2 (original-image coordinates, clipped),
    masks float32 [n,M,M] -> uint8 [n,h,w] (premvos_fill_full_masks_host)."""
    boxes = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 4)
    masks = np.ascontiguousarray(masks, dtype=np.float32)
    n = boxes.shape[0]
    if masks.ndim != 3 or masks.shape[0] != n or masks.shape[1] != masks.shape[2]:
        raise ValueError("masks must be [n,M,M] with n = len(boxes), got %s" % (masks.shape,))
    out = np.zeros((n, int(shape[0]), int(shape[1])), np.uint8)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    _lib.check(_lib.lib().premvos_fill_full_masks_host(vp(masks), vp(boxes), n, int(masks.shape[1]), int(shape[0]), int(shape[1]), vp(out)))
    return out


def fill_full_mask(box, mask, shape):
    """eval.py:35-58 (one box)."""
    return fill_full_masks(np.asarray(box)[None], np.asarray(mask)[None], shape)[0]


def detect_one_image(img, model_func, size=SHORT_EDGE_SIZE, max_size=MAX_SIZE):
    """eval.py:61-110 for USE_SECOND_HEAD without feature extraction; with a seventh `final_masks` output (MODE_MASK) every
    result carries its full-image binary mask (fill_full_mask)."""
    import cv2
    orig_shape = img.shape[:2]
    newh, neww = custom_resize_shape(orig_shape[0], orig_shape[1], size, max_size)
    resized_img = cv2.resize(img, (neww, newh), interpolation=cv2.INTER_LINEAR)
    scale = (resized_img.shape[0] * 1.0 / img.shape[0] + resized_img.shape[1] * 1.0 / img.shape[1]) / 2
    boxes, probs, labels, posteriors, second_labels, second_posteriors, *masks = model_func(resized_img)
    boxes = boxes / scale
    boxes = clip_boxes(boxes, orig_shape)
    if masks:
        masks = list(fill_full_masks(boxes, masks[0], orig_shape)) if len(boxes) else []
    else:
        masks = [None] * len(boxes)
    features = [None for _ in range(labels.size)]
    return [SecondDetectionResult(*args) for args in
            zip(boxes, probs, labels, posteriors, masks, second_labels, second_posteriors, features)]


def convert_results_to_json(results, img_idx=None):
    """train.py:388-428: [{'bbox': [x, y, w, h] (1 decimal), 'score': (2 decimals)}]"""
    img_res = []
    for r in results:
        box = np.array(r.box, dtype=np.float32)      # the reference subtracts in place on the float32 box and rounds np.float32 values
        box[2] -= box[0]
        box[3] -= box[1]
        res = {"bbox": list(map(lambda x: float(round(x, 1)), box)), "score": float(round(r.score, 2))}
        if r.mask is not None:                        # train.py:421-426: COCO RLE of the pasted mask
            from .refnet import rle_encode
            res["segmentation"] = rle_encode(np.asarray(r.mask))
        img_res.append(res)
    return img_res
