"""Pins the TensorFlow-side box arithmetic of the proposal and refinement graphs WITHOUT TensorFlow (build container only: needs
/root/reference): the reference functions' UNMODIFIED source is extracted from the files where they lie and exec'd with a numpy
stand-in for the handful of element-wise / shape TensorFlow ops they call.  What is pinned is the functions' own arithmetic and
control flow (operation order, constants, which operand is cast where); the stand-in supplies only op semantics that are not in
question (tf.maximum, tf.split, tf.round = half to even, python scalars adopting the tensor's dtype, ...).

  proposal_net/model.py   decode_bbox_target (:114-139), clip_boxes (:18-27),
                          roi_align -> crop_and_resize -> transform_fpcoor_for_tf (:301-374): the normalised boxes handed to
                          tf.image.crop_and_resize are captured by the stand-in (the op itself is not run)
  refinement_net/datasets/Resize.py   bbox_crop_and_resize_fixed_size (:150-193): crop boxes (CROP_BOXES_y0x0y1x1) and the
                          slices taken from the image / guidance tensors

    python tests/golden/make_tf_shim_goldens.py   ->  tests/golden/tf_shim_golden.npz
"""
import ast
import importlib.util
import os
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/code"


def _as(x, like):
    """python / numpy scalars adopt the tensor operand's dtype, as TensorFlow converts constants"""
    if isinstance(x, np.ndarray) and x.ndim > 0:
        return x
    return np.asarray(x, dtype=like.dtype if isinstance(like, np.ndarray) else None)


class T(np.ndarray):
    """ndarray whose binary operators cast scalar operands to its own dtype first"""
    def _b(self, other, op):
        o = other if isinstance(other, np.ndarray) and other.ndim > 0 else np.asarray(other, dtype=self.dtype)
        return getattr(np.ndarray, op)(self, o).view(T)
    def __add__(self, o): return self._b(o, "__add__")
    def __radd__(self, o): return self._b(o, "__radd__")
    def __sub__(self, o): return self._b(o, "__sub__")
    def __rsub__(self, o): return self._b(o, "__rsub__")
    def __mul__(self, o): return self._b(o, "__mul__")
    def __rmul__(self, o): return self._b(o, "__rmul__")
    def __truediv__(self, o): return self._b(o, "__truediv__")


def t(x, dtype=None):
    return np.asarray(x, dtype=dtype).view(T)


def make_tf(captured):
    img = types.SimpleNamespace(crop_and_resize=lambda image, boxes, box_ind, crop_size: (
        captured.append(np.array(boxes)), t(np.zeros((len(boxes), crop_size[0], crop_size[1], image.shape[-1]), np.float32)))[1])
    nn = types.SimpleNamespace(avg_pool=lambda x, *a, **k: x)
    return types.SimpleNamespace(
        int32=np.int32, float32=np.float32, image=img, nn=nn,
        shape=lambda x: t(np.array(np.shape(x), np.int32)),
        maximum=lambda a, b: t(np.maximum(a, _as(b, a))), minimum=lambda a, b, name=None: t(np.minimum(a, _as(b, a))),
        exp=lambda a: t(np.exp(a)), reshape=lambda a, s: t(np.reshape(a, tuple(int(v) for v in s))),
        split=lambda a, n, axis=0: [t(v) for v in np.split(a, n, axis=axis)], concat=lambda xs, axis=0: t(np.concatenate(xs, axis=axis)),
        tile=lambda a, m: t(np.tile(a, m)), reverse=lambda a, axes: t(np.flip(a, axes)), to_float=lambda a: t(np.asarray(a), np.float32),
        cast=lambda a, d: t(np.asarray(a).astype(d)), round=lambda a: t(np.rint(a)), unstack=lambda a: [v for v in np.asarray(a)],
        stack=lambda xs, axis=0: t(np.stack([np.asarray(v) for v in xs], axis)), transpose=lambda a, p: t(np.transpose(a, p)),
        zeros=lambda s, dtype=np.float32: t(np.zeros([int(v) for v in s], dtype)), stop_gradient=lambda a: a,
        constant=lambda v, dtype=None: t(np.asarray(v, dtype=dtype)))


def extract(path, names, ns):
    tree = ast.parse(open(path).read())
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            node.decorator_list = []          # @under_name_scope() / @layer_register: graph naming only
            for sub in ast.walk(node):
                if isinstance(sub, ast.FunctionDef):
                    sub.decorator_list = []
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    missing = [n for n in names if n not in ns]
    assert not missing, missing
    return ns


def main():
    rng = np.random.default_rng(23)
    out = {}
    os.environ.setdefault("USER", "premvos")
    spec = importlib.util.spec_from_file_location("ref_config", os.path.join(REF, "proposal_net/config.py"))
    config = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(config)
    captured = []
    tf = make_tf(captured)
    ns = extract(os.path.join(REF, "proposal_net/model.py"), ["decode_bbox_target", "clip_boxes", "crop_and_resize", "roi_align"],
                 {"tf": tf, "config": config, "np": np})
    # ---- decode_bbox_target: logits incl. values beyond BBOX_DECODE_CLIP ----
    xy = rng.uniform(-50, 1300, (200, 2))
    anchors = np.concatenate([xy, xy + rng.uniform(8, 600, (200, 2))], 1).astype(np.float32)
    logits = (rng.standard_normal((200, 4)) * [0.5, 0.5, 1.5, 1.5]).astype(np.float32)
    logits[:5, 2:] = [[6.0, 4.5], [4.42, 4.43], [10, -10], [0, 0], [-3, 5]]
    out["dec_anchors"], out["dec_logits"] = anchors, logits
    out["dec_out"] = np.asarray(ns["decode_bbox_target"](t(logits), t(anchors)), np.float32)
    out["dec_clip"] = np.float64(config.BBOX_DECODE_CLIP)
    # ---- clip_boxes (TF version) ----
    cb = rng.uniform(-80, 1500, (120, 4)).astype(np.float32)
    out["clip_in"], out["clip_hw"] = cb, np.array([568, 1333], np.int32)
    out["clip_out"] = np.asarray(ns["clip_boxes"](t(cb), t(out["clip_hw"])), np.float32)
    # ---- roi_align -> crop_and_resize -> transform_fpcoor_for_tf: the boxes handed to tf.image.crop_and_resize ----
    fm = t(np.zeros((1, 8, 36, 84), np.float32))
    b0 = rng.uniform(-10, 1200, (150, 2))
    rois = np.concatenate([b0, b0 + rng.uniform(0.5, 500, (150, 2))], 1).astype(np.float32) * np.float32(1.0 / 16)
    ns["roi_align"](fm, t(rois), 14)
    out["roi_boxes"], out["roi_fm_hw"], out["roi_crop"] = rois, np.array([36, 84]), np.int64(28)
    out["roi_tf_boxes"] = captured[-1].astype(np.float32)
    # ---- refinement_net/datasets/Resize.py bbox_crop_and_resize_fixed_size ----
    keys = types.SimpleNamespace(IMAGES="images", BBOXES_y0x0y1x1="bboxes_y0x0y1x1", SEGMENTATION_LABELS="segmentation_labels",
                                 BBOX_GUIDANCE="bbox_guidance", RAW_SEGMENTATION_LABELS="raw_segmentation_labels",
                                 LASER_GUIDANCE="laser_guidance", SEGMENTATION_LABELS_ORIGINAL_SIZE="segmentation_labels_original_size",
                                 CROP_BOXES_y0x0y1x1="crop_boxes_y0x0y1x1")
    seen = []
    rz = extract(os.path.join(REF, "refinement_net/datasets/Resize.py"), ["bbox_crop_and_resize_fixed_size"],
                 {"tf": tf, "DataKeys": keys, "resize_image": lambda res, size, bilinear: (seen.append(np.shape(res)), res)[1]})
    H, W = 480, 854
    image = t(np.zeros((H, W, 3), np.float32))
    guidance = t(np.zeros((H, W, 1), np.uint8))
    xywh = np.concatenate([rng.uniform(-20, W - 5, (60, 1)), rng.uniform(-20, H - 5, (60, 1)), rng.uniform(1, 500, (60, 2))], 1)
    xywh[:6] = [[10.5, 20.5, 99.5, 50.5], [0, 0, 854, 480], [800.5, 400.5, 100, 100], [2.5, 3.5, 1, 1], [-12.5, -3.5, 90, 70], [400, 200, 0.4, 0.4]]
    boxes = np.stack([xywh[:, 1], xywh[:, 0], xywh[:, 1] + xywh[:, 3], xywh[:, 0] + xywh[:, 2]], 1).astype(np.float32)
    crops, shapes = [], []
    for b in boxes:
        seen.clear()
        r = rz["bbox_crop_and_resize_fixed_size"]({keys.IMAGES: image, keys.BBOXES_y0x0y1x1: t(b), keys.BBOX_GUIDANCE: guidance}, (385, 385))
        crops.append(np.asarray(r[keys.CROP_BOXES_y0x0y1x1], np.int64))
        shapes.append(list(seen[0][:2]))
    out["crop_hw"], out["crop_bboxes"], out["crop_out"], out["crop_slice_shapes"] = np.array([H, W]), boxes, np.stack(crops), np.array(shapes)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "tf_shim_golden.npz"), **out)
    print("wrote tf_shim_golden.npz:", {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
