"""Host-side mirror of the reference's ReID call surface, backed by the CUDA library.

Reference surface reproduced here:
  ReID_net_init() -> engine                                        MergeTrack/ReID_net_functions.py:19-25
  add_ReID(proposals, image_fn, ReID_net) -> proposals             MergeTrack/ReID_net_functions.py:26-45
     = session.run([test_network.get_output_layer().outputs, valid_data.crop_list],
                   feed_dict={valid_data.image: image, valid_data.boxes: boxes})      one run per frame, all boxes
  graph                                                            ReID_net/configs/live ("network"), ReID_net/network/NetworkLayers.py
  crop pipeline                                                    ReID_net/datasets/Similarity/DAVIS_Forward_Feed.py:34-120

All tensor arithmetic (crops, resize, normalisation, the residual network, the embedding head) runs in libpremvos_b200.so.
No TensorFlow, no CPU fallback.
"""
from __future__ import annotations

import ctypes
from collections import OrderedDict

import numpy as np

from . import _lib
from .synth import reid_param_shapes

EMBEDDING_DIM = 128
LIVE_CONFIG = "ReID_net/configs/live"      # ReID_net_functions.py:20, relative to the reference's cwd `code/`


class ReIDNet:
    def __init__(self, max_batch=64):
        self.max_batch = int(max_batch)
        self._shapes = reid_param_shapes()
        self._handle = None
        self._params = None

    def load_params(self, params):
        missing = [k for k in self._shapes if k not in params]
        unexpected = [k for k in params if k not in self._shapes]
        if missing or unexpected:
            raise RuntimeError("ReID_net variables: missing %s, unexpected %s" % (missing[:5], unexpected[:5]))
        for k, shp in self._shapes.items():
            if tuple(np.shape(params[k])) != tuple(shp):
                raise RuntimeError("size mismatch for %s: got %s, expected %s" % (k, tuple(np.shape(params[k])), tuple(shp)))
        self._params = OrderedDict((k, np.ascontiguousarray(params[k], dtype=np.float32)) for k in self._shapes)
        self.close()
        return self

    def close(self):
        if self._handle is not None:
            _lib.lib().premvos_reidnet_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _h(self):
        import torch
        if self._handle is not None:
            if torch.cuda.current_device() != self._handle_device:
                raise RuntimeError("ReIDNet handle lives on cuda:%d, called with cuda:%d current (one net object per device)"
                                   % (self._handle_device, torch.cuda.current_device()))
            return self._handle
        if self._params is None:
            raise RuntimeError("ReIDNet: load_params() first")
        L = _lib.lib()
        h = ctypes.c_void_p()
        _lib.check(L.premvos_reidnet_create(ctypes.byref(h), self.max_batch))
        try:
            for k, v in self._params.items():
                _lib.check(L.premvos_reidnet_set_param(h, k.encode(), v.ctypes.data_as(ctypes.c_void_p), v.size))
            _lib.check(L.premvos_reidnet_finalize(h))
        except Exception:
            L.premvos_reidnet_destroy(h)
            raise
        self._handle = h
        self._handle_device = torch.cuda.current_device()
        return h

    def embed(self, image_rgb_uint8, boxes_xywh):
        """-> float32 [n, 128]: one embedding per box of the frame (`ys`, ReID_net_functions.py:37-38)"""
        img = np.ascontiguousarray(image_rgb_uint8)
        if img.dtype != np.uint8 or img.ndim != 3 or img.shape[2] != 3:
            raise ValueError("expected a uint8 RGB image [H,W,3], got %s %s" % (img.dtype, img.shape))
        boxes = np.ascontiguousarray(boxes_xywh, dtype=np.float32).reshape(-1, 4)
        n, (H, W) = boxes.shape[0], img.shape[:2]
        out = np.zeros((n, EMBEDDING_DIM), np.float32)
        vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        _lib.check(_lib.lib().premvos_reidnet_forward_host(self._h(), vp(img), H, W, vp(boxes), n, vp(out)))
        return out

    def embed_device(self, frame, boxes_xywh, out=None):
        """Resident-pipeline entry point: `frame` CUDA uint8 RGB [H,W,3], `boxes_xywh` CUDA float32 [n,4]; enqueues on the
        current torch stream, never synchronises.  -> CUDA float32 [n, 128]"""
        import torch
        if not (isinstance(frame, torch.Tensor) and frame.is_cuda and frame.dtype == torch.uint8 and frame.is_contiguous()):
            raise TypeError("frame must be a contiguous CUDA uint8 tensor (this build has no CPU path)")
        if frame.dim() != 3 or frame.shape[2] != 3:
            raise ValueError("expected an RGB frame [H,W,3], got %s" % (tuple(frame.shape),))
        if not (isinstance(boxes_xywh, torch.Tensor) and boxes_xywh.is_cuda and boxes_xywh.dtype == torch.float32
                and boxes_xywh.is_contiguous() and boxes_xywh.dim() == 2 and boxes_xywh.shape[1] == 4):
            raise TypeError("boxes_xywh must be a contiguous CUDA float32 tensor [n,4]")
        n, H, W = int(boxes_xywh.shape[0]), int(frame.shape[0]), int(frame.shape[1])
        with torch.cuda.device(frame.device):
            if out is None:
                out = torch.empty((n, EMBEDDING_DIM), dtype=torch.float32, device=frame.device)
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().premvos_reidnet_forward(self._h(), frame.data_ptr(), H, W, boxes_xywh.data_ptr(), n,
                                                          out.data_ptr(), st))
        return out

    def launches_per_forward(self):
        return int(_lib.lib().premvos_reidnet_launches_per_forward(self._h()))

    def get_tensor(self, name):
        L = _lib.lib()
        n = ctypes.c_int64()
        _lib.check(L.premvos_reidnet_get_tensor(self._h(), name.encode(), None, ctypes.byref(n)))
        buf = np.empty(n.value, dtype=np.float32)
        _lib.check(L.premvos_reidnet_get_tensor(self._h(), name.encode(), buf.ctypes.data_as(ctypes.c_void_p), ctypes.byref(n)))
        return buf


class Engine:
    """ReID_net/Engine.py as far as `add_ReID` uses it: an object that owns the loaded network.  `Engine(config)` reads the
    config's "load" entry (TensorFlow checkpoint prefix or .npz / .npy dump) and checks that the config describes the network
    this library implements; `Engine(net)` wraps an already loaded ReIDNet."""

    def __init__(self, net_or_config, params=None, **kw):
        self.config = None
        if isinstance(net_or_config, ReIDNet):
            self.net = net_or_config
            return
        import json
        if isinstance(net_or_config, dict):
            self.config = dict(net_or_config)
        else:
            with open(net_or_config) as f:
                self.config = json.load(f)
        size = [int(v) for v in self.config.get("input_size", [128, 128])]
        if size != [128, 128] or float(self.config.get("context_region_factor_val", 1.2)) != 1.2 \
                or int(self.config.get("num_classes", EMBEDDING_DIM)) != EMBEDDING_DIM:
            raise ValueError("ReID_net config: only the shipped geometry (128 x 128 crops, context region 1.2, 128-d output) is built")
        kw.setdefault("max_batch", min(int(self.config.get("batch_size_eval", 64)), 256))
        if params is None:
            from . import weights
            load = self.config.get("load")
            if not load:
                raise KeyError("ReID_net config has no 'load' entry")
            try:
                params = weights.load_reid_net_variables(load)
            except (OSError, IOError) as e:
                raise FileNotFoundError("ReID_net weights '%s' cannot be read: %s" % (load, e)) from e
        self.net = ReIDNet(**kw).load_params(params)


def ReID_net_init(params=None, config_path=None, **kw) -> Engine:
    """ReID_net_functions.py:19-25: no arguments -- reads `ReID_net/configs/live` relative to the current directory.
    Extensions: `params` (an already loaded variable dict, no file access), `config_path`, ReIDNet keyword arguments."""
    if params is not None and config_path is None:
        return Engine(ReIDNet(**kw).load_params(params))
    return Engine(config_path or LIVE_CONFIG, params, **kw)


def add_ReID(proposals, image_fn, ReID_net: Engine):
    """ReID_net_functions.py:26-45: every proposal gets 'ReID' = its 128 embedding values as a list of Python floats.
    `image_fn`: file name or RGB uint8 array."""
    if isinstance(image_fn, np.ndarray):
        image = image_fn
    else:
        import cv2
        image = cv2.imread(image_fn, cv2.IMREAD_COLOR)
        if image is None:
            raise FileNotFoundError(image_fn)
        image = image[:, :, ::-1]
    boxes = [prop["bbox"] for prop in proposals]
    if not boxes:
        return proposals
    emb = ReID_net.net.embed(image, boxes)
    for prop, e in zip(proposals, emb):
        prop["ReID"] = e.tolist()
    return proposals
