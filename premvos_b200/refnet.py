"""Host-side mirror of the reference's refinement call surface, backed by the CUDA library.

Reference surface reproduced here:
  refinement_net_init() -> engine                                  MergeTrack/refinement_net_functions.py:19-24
  do_refinement(proposals, image_fn, engine) -> proposals          MergeTrack/refinement_net_functions.py:38-65
  engine.valid_data.set_up_data_for_image(image, boxes)            refinement_net/datasets/few_shot_segmentation/
  engine.valid_data.get_feed_dict_for_next_step(image_data, idx)     FewShotFeedSegmentationDataset.py:35-51, FeedDataset.py
  engine.trainer.validation_step(feed_dict=..., extraction_keys=[...]) -> {'extractions': {...}}
                                                                    refinement_net/core/Trainer.py:128-133,150-169
  COCO RLE 'segmentation' records                                  pycocotools.mask.encode as used at :52-56

All tensor arithmetic (crop/resize, DeepLabv3+, output layer, conf_score) runs in libpremvos_b200.so; the proposals of a
frame are run as one batch instead of one session.run per proposal.  No TensorFlow, no CPU fallback.
"""
from __future__ import annotations

import ctypes
from collections import OrderedDict

import numpy as np

from . import _lib
from .synth import refnet_param_shapes

# refinement_net/core/Extractions.py:1-6, datasets/DataKeys.py
EXTRACTIONS = "extractions"
SEGMENTATION_POSTERIORS_ORIGINAL_SIZE = "segmentation_posteriors_original_size"
SEGMENTATION_MASK_ORIGINAL_SIZE = "segmentation_mask_original_size"
OBJ_TAGS = "obj_tags"
IMAGES = "images"
BBOXES_y0x0y1x1 = "bboxes_y0x0y1x1"


class RefinementNet:
    def __init__(self, max_batch=16, input_size=385, middle_units=16):
        self.max_batch, self.input_size, self.middle_units = max_batch, input_size, middle_units
        self._shapes = refnet_param_shapes(middle_units)
        self._handle = None
        self._params = None

    def load_params(self, params):
        missing = [k for k in self._shapes if k not in params]
        unexpected = [k for k in params if k not in self._shapes]
        if missing or unexpected:
            raise RuntimeError("refinement_net variables: missing %s, unexpected %s" % (missing[:5], unexpected[:5]))
        for k, shp in self._shapes.items():
            if tuple(np.shape(params[k])) != tuple(shp):
                raise RuntimeError("size mismatch for %s: got %s, expected %s" % (k, tuple(np.shape(params[k])), tuple(shp)))
        self._params = OrderedDict((k, np.ascontiguousarray(params[k], dtype=np.float32)) for k in self._shapes)
        self.close()
        return self

    def close(self):
        if self._handle is not None:
            _lib.lib().premvos_refnet_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _h(self):
        if self._handle is not None:
            import torch
            if torch.cuda.current_device() != self._handle_device:
                raise RuntimeError("RefinementNet handle lives on cuda:%d, called with cuda:%d current (one net object per device)"
                                   % (self._handle_device, torch.cuda.current_device()))
            return self._handle
        if self._params is None:
            raise RuntimeError("RefinementNet: load_params() first")
        L = _lib.lib()
        h = ctypes.c_void_p()
        _lib.check(L.premvos_refnet_create(ctypes.byref(h), self.max_batch, self.input_size, self.middle_units))
        try:
            for k, v in self._params.items():
                _lib.check(L.premvos_refnet_set_param(h, k.encode(), v.ctypes.data_as(ctypes.c_void_p), v.size))
            _lib.check(L.premvos_refnet_finalize(h))
        except Exception:
            L.premvos_refnet_destroy(h)
            raise
        import torch
        self._handle = h
        self._handle_device = torch.cuda.current_device()
        return h

    def refine(self, image_rgb_uint8, boxes_xywh, want_posteriors=False):
        """-> masks uint8 [n,H,W] (0/1), conf_scores float32 [n], posteriors float32 [n,H,W] or None"""
        img = np.ascontiguousarray(image_rgb_uint8)
        if img.dtype != np.uint8 or img.ndim != 3 or img.shape[2] != 3:
            raise ValueError("expected a uint8 RGB image [H,W,3], got %s %s" % (img.dtype, img.shape))
        boxes = np.ascontiguousarray(boxes_xywh, dtype=np.float32).reshape(-1, 4)
        n, (H, W) = boxes.shape[0], img.shape[:2]
        masks = np.zeros((n, H, W), np.uint8)
        conf = np.zeros((n,), np.float32)
        post = np.zeros((n, H, W), np.float32) if want_posteriors else None
        vp = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
        _lib.check(_lib.lib().premvos_refnet_forward_host(self._h(), vp(img), H, W, vp(boxes), n, vp(masks), vp(conf), vp(post)))
        return masks, conf, post

    def refine_device(self, frame, boxes_xywh, masks=None, conf=None):
        """Resident-pipeline entry point: `frame` CUDA uint8 RGB [H,W,3], `boxes_xywh` CUDA float32 [n,4]; enqueues on the
        current torch stream, never synchronises.  -> (masks CUDA uint8 [n,H,W], conf_scores CUDA float32 [n])"""
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
            if masks is None:
                masks = torch.empty((n, H, W), dtype=torch.uint8, device=frame.device)
            if conf is None:
                conf = torch.empty((n,), dtype=torch.float32, device=frame.device)
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().premvos_refnet_forward(self._h(), frame.data_ptr(), H, W, boxes_xywh.data_ptr(), n,
                                                         masks.data_ptr(), conf.data_ptr(), None, st))
        return masks, conf

    def get_tensor(self, name):
        L = _lib.lib()
        n = ctypes.c_int64()
        _lib.check(L.premvos_refnet_get_tensor(self._h(), name.encode(), None, ctypes.byref(n)))
        buf = np.empty(n.value, dtype=np.float32)
        _lib.check(L.premvos_refnet_get_tensor(self._h(), name.encode(), buf.ctypes.data_as(ctypes.c_void_p), ctypes.byref(n)))
        return buf


# ---- COCO RLE (pycocotools maskApi.c: rleEncode + rleToString / rleFrString) ---------------------------------
def rle_encode(mask):
    m = np.asarray(mask)
    h, w = m.shape
    flat = np.ascontiguousarray(m.T).reshape(-1) != 0            # column-major
    change = np.flatnonzero(flat[1:] != flat[:-1]) + 1
    bounds = np.concatenate([[0], change, [flat.size]])
    counts = np.diff(bounds).tolist()
    if flat.size and flat[0]:
        counts = [0] + counts                                      # runs start with a zero-run
    out = []
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
            out.append(chr(c + 48))
    return {"size": [h, w], "counts": "".join(out)}


def rle_decode(rle):
    h, w = rle["size"]
    s = rle["counts"]
    if isinstance(s, bytes):
        s = s.decode("ascii")
    counts, p = [], 0
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


# ---- the MergeTrack surface -----------------------------------------------------------------------------------
class _ValidData:
    """engine.valid_data (FewShotFeedSegmentationDataset): dict-of-dicts protocol kept as is."""

    def set_up_data_for_image(self, image, boxes):
        obj_data = {}
        for box_id, box in enumerate(boxes):
            x0, y0, x1, y1 = box
            obj_data[box_id] = {IMAGES: image, BBOXES_y0x0y1x1: [y0, x0, y1 + y0, x1 + x0], OBJ_TAGS: str(box_id), "_xywh": list(box)}
        return obj_data if obj_data else None

    def get_feed_dict_for_next_step(self, image_data, idx):
        return image_data[idx]


class _Trainer:
    def __init__(self, net):
        self.net = net

    def validation_step(self, feed_dict=None, extraction_keys=()):
        """One proposal, like the reference's session.run (core/Trainer.py:128-133): arrays shaped [1, ...] in lists."""
        masks, conf, post = self.net.refine(feed_dict[IMAGES], [feed_dict["_xywh"]], want_posteriors=True)
        ex = {SEGMENTATION_MASK_ORIGINAL_SIZE: [masks.astype(np.int64)], SEGMENTATION_POSTERIORS_ORIGINAL_SIZE: [post],
              OBJ_TAGS: [np.array([feed_dict[OBJ_TAGS].encode("utf-8")])]}
        return {EXTRACTIONS: {k: v for k, v in ex.items() if k in extraction_keys}, "measures": {}, "loss": 0.0}


class Engine:
    """refinement_net/core/Engine.py as far as inference goes: `Engine(config)` builds the network and restores the weights the
    config names, `.run()` is the stage-5 forwarder; `Engine(net)` wraps an already loaded RefinementNet."""

    def __init__(self, net_or_config, params=None, **kw):
        self.config = None
        if isinstance(net_or_config, RefinementNet):
            self.net = net_or_config
        else:
            self.config = net_or_config if isinstance(net_or_config, Config) else Config(net_or_config)
            self.net = _net_from_config(self.config, params, **kw)
        self.valid_data = _ValidData()
        self.trainer = _Trainer(self.net)

    def run(self, log=None):
        if self.config is None:
            raise RuntimeError("Engine.run() needs the config the engine was built from (bb_input_dir / output_dir)")
        return run_forwarder(self.config, engine=self, log=log)


class Config:
    """refinement_net/core/Config.py: the JSON config files of the reference (`configs/live`, `configs/run`), typed getters."""

    def __init__(self, path_or_dict):
        import json
        if isinstance(path_or_dict, dict):
            self._d, self.path = dict(path_or_dict), None
        else:
            self.path = path_or_dict
            with open(path_or_dict) as f:
                self._d = json.load(f)

    def string(self, key, default=None):
        v = self._d.get(key, default)
        if v is None and default is None and key not in self._d:
            raise KeyError("config has no '%s'" % key)
        return v

    def int(self, key, default=None):
        return int(self.string(key, default))

    def bool(self, key, default=None):
        return bool(self.string(key, default))

    def int_list(self, key, default=None):
        return [int(v) for v in self.string(key, default)]


LIVE_CONFIG = "refinement_net/configs/live"      # refinement_net_functions.py:20, relative to the reference's cwd `code/`


def _net_from_config(config: "Config", params=None, **kw) -> RefinementNet:
    """Network from the config's input size, weights from config["load"] (a TensorFlow checkpoint prefix, or an .npz / .npy dump of
    the same variables) unless `params` are given."""
    size = config.int_list("input_size_train", [385, 385])
    if size[0] != size[1]:
        raise ValueError("input_size_train must be square, got %s" % (size,))
    kw.setdefault("input_size", size[0])
    net = RefinementNet(**kw)
    if params is None:
        from . import weights
        load = config.string("load")
        try:
            params = weights.load_refinement_net_variables(load, middle_units=net.middle_units)
        except (OSError, IOError) as e:
            raise FileNotFoundError("refinement_net weights '%s' (config %s) cannot be read: %s" % (load, config.path, e)) from e
    return net.load_params(params)


def refinement_net_init(params=None, config_path=None, **kw) -> Engine:
    """refinement_net_functions.py:19-24: no arguments -- reads `refinement_net/configs/live` relative to the current directory
    (the reference runs from `code/`) and restores the checkpoint its "load" entry names.  Extensions: `params` (an already loaded
    variable dict, no file access), `config_path`, and RefinementNet keyword arguments (max_batch, middle_units)."""
    if params is not None and config_path is None:
        net = RefinementNet(**kw).load_params(params)
        return Engine(net)
    return Engine(Config(config_path or LIVE_CONFIG), params, **kw)


def run_forwarder(config, engine: Engine = None, params=None, log=None, **kw):
    """The stage-5 batch job `Engine(Config(path)).run()` with "task": "few_shot_segmentation" and need_train false
    (refinement_net/main.py:19-32 -> forwarding/FewShotSegmentationForwarder.py:85-155): for every
    `<bb_input_dir>/<video>/<frame>.json` read the proposals, refine them on `<image_input_dir>/<video>/<frame>.jpg` and write the
    list -- every proposal now with 'segmentation' and 'conf_score' -- to `<output_dir>/<video>/<frame>.json`.
    `config`: path, dict or Config.  Returns the number of frames written."""
    import glob
    import json
    import os
    if not isinstance(config, Config):
        config = Config(config)
    if engine is None:
        engine = Engine(config, params, **kw)
    img_dir, bb_dir, out_dir = config.string("image_input_dir"), config.string("bb_input_dir"), config.string("output_dir")
    frames = 0
    for fn in sorted(glob.glob(os.path.join(bb_dir, "*", "*.json"))):
        video, name = os.path.basename(os.path.dirname(fn)), os.path.basename(fn)
        with open(fn) as f:
            proposals = json.load(f)
        image_fn = os.path.join(img_dir, video, name[:-5] + ".jpg")
        if not os.path.exists(image_fn):
            raise FileNotFoundError("frame %s of the proposals %s is missing" % (image_fn, fn))
        proposals = do_refinement(proposals, image_fn, engine)
        os.makedirs(os.path.join(out_dir, video), exist_ok=True)
        with open(os.path.join(out_dir, video, name), "w") as f:
            json.dump(proposals, f)
        frames += 1
        if log:
            log("refined %s/%s: %d proposals" % (video, name, len(proposals)))
    return frames


def do_refinement(proposals, image_fn, refinement_net: Engine):
    """refinement_net_functions.py:38-65, all proposals of the frame in one batched call.  `image_fn`: file name or RGB array."""
    if isinstance(image_fn, np.ndarray):
        image = image_fn
    else:
        import cv2
        image = cv2.imread(image_fn, cv2.IMREAD_COLOR)[:, :, ::-1]
    boxes = [prop["bbox"] for prop in proposals]
    if not boxes:
        return proposals
    masks, conf, _ = refinement_net.net.refine(image, boxes)
    for prop, m, c in zip(proposals, masks, conf):
        prop["segmentation"] = rle_encode(m * 255)
        prop["conf_score"] = str(c)
    return proposals
