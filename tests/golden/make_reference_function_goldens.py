"""Runs the REFERENCE'S OWN pure numpy / OpenCV functions (unmodified source, extracted from the files where they lie under
/root/reference and exec'd in an isolated namespace -- their modules import TensorFlow / tensorpack / pycocotools at the top and
cannot be imported whole) on seeded inputs and stores inputs + outputs as tests/golden/reference_functions_golden.npz.
tests/test_reference_function_goldens.py holds the oracle restatements and the host mirrors against these vectors.

    python tests/golden/make_reference_function_goldens.py          (build container only: needs /root/reference and cv2)

Functions executed:
  proposal_net/eval.py              fill_full_mask
  proposal_net/common.py            clip_boxes, CustomResize._get_augment_params (with a stub `transform` namespace)
  proposal_net/utils/np_box_ops.py  iou (whole module: pure numpy)
  MergeTrack/merge_functions.py     warp_flow, get_flow
  optical_flow_net-PWC-Net/script_pwc_multi.py   writeFlowFile
  proposal_net/data.py              get_all_anchors (with the reference's config.py and utils/generate_anchors.py)
  refinement_net/datasets/util/BoundingBox.py    encode_bbox_as_mask_np (numpy namespace with the removed alias np.int = int)
  refinement_net/datasets/util/Normalization.py  normalize, unnormalize (whole module: pure numpy)
"""
import ast
import collections
import importlib.util
import os
import sys
import tempfile

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/code"
sys.path.insert(0, ROOT)
from premvos_b200 import synth  # noqa: E402


def extract(path, names, extra_globals=None):
    """exec only the named top-level function / class definitions of a reference file, source text untouched."""
    src = open(path).read()
    tree = ast.parse(src)
    ns = {"np": np, "cv2": cv2, "sys": sys}
    ns.update(extra_globals or {})
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    missing = [n for n in names if n not in ns]
    assert not missing, missing
    return ns


def main():
    cv2.ipp.setUseIPP(False)     # OpenCV's own code path for the float resize inside fill_full_mask
    rng = np.random.default_rng(7)
    out = {}
    # ---- proposal_net/eval.py fill_full_mask ----
    ns = extract(os.path.join(REF, "proposal_net/eval.py"), ["fill_full_mask"])
    H, W = 96, 128
    boxes = np.array([[3.2, 5.4, 17.4, 19.3], [0, 0, 7, 7], [10.2, 10.1, 10.9, 10.4], [0, 0, 128, 96], [100.5, 60.5, 127.6, 95.7],
                      [20, 30, 34, 44], [5.5, 5.5, 6.4, 90.0], [0.4, 0.4, 2.6, 2.6], [50, 50, 57, 57], [1, 1, 29, 15]], np.float32)
    masks = rng.uniform(0, 1, (len(boxes), 14, 14)).astype(np.float32)
    out["ffm_boxes"], out["ffm_masks"] = boxes, masks
    out["ffm_full"] = np.stack([ns["fill_full_mask"](b.copy(), m.copy(), (H, W)) for b, m in zip(boxes, masks)])
    # ---- proposal_net/common.py clip_boxes, CustomResize ----
    Resize = collections.namedtuple("ResizeTransform", "h w newh neww interp")

    class Base:
        def _init(self, params):
            for k, v in params.items():
                if k != "self":
                    setattr(self, k, v)

    transform = type("transform", (), {"TransformAugmentorBase": Base, "ResizeTransform": Resize})
    ns = extract(os.path.join(REF, "proposal_net/common.py"), ["clip_boxes", "CustomResize"], {"transform": transform})
    cb = (rng.uniform(-30, 160, (40, 4))).astype(np.float32)
    out["clip_in"] = cb
    out["clip_out"] = ns["clip_boxes"](cb.copy(), (H, W))
    shapes = np.array([[480, 854], [436, 1024], [1080, 1920], [720, 1280], [100, 140], [854, 480], [800, 800], [333, 2000], [64, 64],
                       [799, 1333], [801, 1334]])
    res = []
    for h, w in shapes:
        t = ns["CustomResize"](800, 1333)._get_augment_params(np.zeros((h, w, 3), np.uint8))
        res.append([t.newh, t.neww])
    out["resize_in"], out["resize_out"] = shapes, np.array(res)
    # ---- proposal_net/utils/np_box_ops.py iou ----
    spec = importlib.util.spec_from_file_location("np_box_ops", os.path.join(REF, "proposal_net/utils/np_box_ops.py"))
    box_ops = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(box_ops)
    xy = rng.uniform(0, 100, (30, 2))
    wh = rng.uniform(1, 60, (30, 2))
    bx = np.concatenate([xy, xy + wh], 1).astype(np.float32)     # [y_min, x_min, y_max, x_max] order does not matter for IoU
    out["iou_boxes"] = bx
    out["iou_out"] = box_ops.iou(bx, bx).astype(np.float64)
    # ---- MergeTrack/merge_functions.py warp_flow, get_flow + script_pwc_multi.py writeFlowFile ----
    mt = extract(os.path.join(REF, "MergeTrack/merge_functions.py"), ["warp_flow", "get_flow"])
    pw = extract(os.path.join(REF, "optical_flow_net-PWC-Net/script_pwc_multi.py"), ["writeFlowFile"])
    h, w = 48, 70
    m = synth.synthetic_masks(3, h, w, seed=31)
    flow = (rng.standard_normal((h, w, 2)) * 5).astype(np.float32)
    flow[:6] = np.round(flow[:6] * 64) / 64
    out["wf_masks"], out["wf_flow"] = m, flow
    out["wf_warped"] = np.stack([mt["warp_flow"](x, flow.copy()) for x in m])
    gray = rng.integers(0, 256, (h, w), dtype=np.uint8)
    out["wf_gray"] = gray
    out["wf_remapped"] = mt["warp_flow"](gray, flow.copy(), binarize=False)
    with tempfile.TemporaryDirectory() as d:
        fn = os.path.join(d, "x.flo")
        pw["writeFlowFile"](fn, flow)
        out["flo_bytes"] = np.frombuffer(open(fn, "rb").read(), dtype=np.uint8)
        out["flo_read"] = mt["get_flow"](fn)
    # ---- refinement_net/datasets/util: guidance mask, normalisation ----
    class NpCompat:   # numpy 2 dropped the alias `np.int` the reference uses; everything else is numpy itself
        int = int

        def __getattr__(self, name):
            return getattr(np, name)

    # ---- proposal_net/data.py get_all_anchors with the reference's own config.py and generate_anchors.py ----
    os.environ.setdefault("USER", "premvos")          # config.py reads it at import time
    spec = importlib.util.spec_from_file_location("ref_config", os.path.join(REF, "proposal_net/config.py"))
    ref_config = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_config)
    spec = importlib.util.spec_from_file_location("ref_generate_anchors", os.path.join(REF, "proposal_net/utils/generate_anchors.py"))
    ga = importlib.util.module_from_spec(spec)
    ga.np = None
    spec.loader.exec_module(ga)

    class NpCompat0:   # numpy 2 dropped the aliases np.float / np.int the reference uses; everything else is numpy itself
        float = float
        int = int

        def __getattr__(self, name):
            return getattr(np, name)

    ga.np = NpCompat0()
    dat = extract(os.path.join(REF, "proposal_net/data.py"), ["get_all_anchors"],
                  {"np": NpCompat0(), "config": ref_config, "memoized": lambda f: f, "generate_anchors": ga.generate_anchors})
    field = dat["get_all_anchors"]()
    import hashlib
    out["anchors_shape"] = np.array(field.shape)
    out["anchors_sha1"] = np.frombuffer(hashlib.sha1(np.ascontiguousarray(field).tobytes()).digest(), dtype=np.uint8)
    out["anchors_cell00"], out["anchors_cell57"] = field[0, 0], field[5, 7]
    out["config_consts"] = np.array([ref_config.ANCHOR_STRIDE, ref_config.MAX_SIZE, ref_config.SHORT_EDGE_SIZE, ref_config.TEST_PRE_NMS_TOPK,
                                     ref_config.TEST_POST_NMS_TOPK, ref_config.RESULTS_PER_IM], np.int64)
    out["config_thresh"] = np.array([ref_config.RPN_PROPOSAL_NMS_THRESH, ref_config.FASTRCNN_NMS_THRESH, ref_config.RESULT_SCORE_THRESH], np.float64)
    bb = extract(os.path.join(REF, "refinement_net/datasets/util/BoundingBox.py"), ["encode_bbox_as_mask_np"], {"np": NpCompat()})
    gboxes = np.array([[3.5, 4.5, 20.5, 30.49], [0.2, 0.7, 9.5, 11.5], [10, 12, 10.4, 40], [2.5, 2.5, 3.5, 3.5], [-0.4, 0.0, 47.6, 69.9]], np.float32)
    out["guid_boxes"] = gboxes
    out["guid_masks"] = np.stack([bb["encode_bbox_as_mask_np"](b, (48, 70, 3)) for b in gboxes])
    spec = importlib.util.spec_from_file_location("Normalization", os.path.join(REF, "refinement_net/datasets/util/Normalization.py"))
    norm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(norm)
    img = rng.uniform(0, 1, (9, 11, 3)).astype(np.float32)
    out["norm_in"] = img
    out["norm_out"] = norm.normalize(img.copy())
    out["norm_back"] = norm.unnormalize(out["norm_out"].copy())
    out["norm_mean"], out["norm_std"] = norm.IMAGENET_RGB_MEAN, norm.IMAGENET_RGB_STD
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "reference_functions_golden.npz"), cv2_version=cv2.__version__, **out)
    print("wrote reference_functions_golden.npz:", {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
