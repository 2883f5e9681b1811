"""Pins what CAN be pinned of the ReID path without TensorFlow (build container only: needs /root/reference):

  * ReID_net/datasets/Similarity/DAVIS_Forward_Feed.py `apply_contex_region` -- the method's UNMODIFIED source, extracted from
    the file where it lies and exec'd with a numpy stand-in for the six TensorFlow ops it calls (tf.cast, tf.round = round half
    to even, tf.maximum, tf.stack, tf.int32, tf.float32).  The arithmetic (python-double `factor - 1.0` meeting float32 tensors,
    in-place `-=` / `*=`, the `maximum(excess, 1)`) is the reference's; only the op semantics are the stand-in's.
  * ReID_net/datasets/Util/Normalization.py `normalize` -- pure numpy, executed as is.
  * ReID_net/configs/live -- the "network" section and the crop parameters, stored verbatim (JSON) so that the tests can hold the
    restated layer table against the reference's own configuration without reading /root/reference at test time.

    python tests/golden/make_reid_reference_goldens.py   ->  tests/golden/reid_reference_golden.npz
"""
import ast
import importlib.util
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/code"


def extract_method(path, cls, name, ns):
    tree = ast.parse(open(path).read())
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for sub in node.body:
                if isinstance(sub, ast.FunctionDef) and sub.name == name:
                    exec(compile(ast.Module(body=[sub], type_ignores=[]), path, "exec"), ns)
                    return ns[name]
    raise KeyError(name)


def main():
    tf = types.SimpleNamespace(
        int32=np.int32, float32=np.float32,
        cast=lambda x, t: np.asarray(x).astype(t),
        round=np.rint,                                  # tf.round: "rounds half to even (banker's rounding)"
        maximum=np.maximum,
        stack=lambda xs, axis=0: np.stack(xs, axis))
    fn = extract_method(os.path.join(REF, "ReID_net/datasets/Similarity/DAVIS_Forward_Feed.py"), "DAVISForwardFeedDataset",
                        "apply_contex_region", {"tf": tf})
    rng = np.random.default_rng(11)
    H, W = 480, 854
    xy = rng.uniform(-40, 1, (6, 2)) * [-1, -1]
    boxes = np.concatenate([
        np.array([[10, 20, 100, 50], [800, 400, 100, 100], [2.5, 3.5, 0, 0], [-30, -10, 50, 40], [0, 0, 854, 480], [200, 100, 6, 30],
                  [12.5, 7.5, 25, 35], [0.5, 0.5, 2.5, 7.5], [853, 479, 10, 10], [100.25, 50.75, 333.3, 444.4]], np.float64),
        np.concatenate([rng.uniform(-20, 860, (40, 1)), rng.uniform(-20, 480, (40, 1)), rng.uniform(1, 500, (40, 2))], 1),
        np.round(np.concatenate([rng.uniform(0, 800, (30, 2)), rng.uniform(5, 300, (30, 2))], 1) * 4) / 4]).astype(np.float32)
    self = types.SimpleNamespace(context_region_factor=1.2)
    out = {"ctx_boxes": boxes.copy(), "ctx_dims": np.array([H, W, 3]),
           "ctx_out": fn(self, boxes.copy(), np.array([H, W, 3], np.int32))}
    spec = importlib.util.spec_from_file_location("ReidNormalization", os.path.join(REF, "ReID_net/datasets/Util/Normalization.py"))
    norm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(norm)
    img = rng.uniform(0, 1, (7, 9, 3)).astype(np.float32)
    out["norm_in"], out["norm_out"] = img, norm.normalize(img.copy())
    cfg = json.load(open(os.path.join(REF, "ReID_net/configs/live")))
    keep = {k: cfg[k] for k in ("network", "input_size", "context_region_factor_val", "num_classes", "batch_size_eval",
                                "output_embedding_layer", "dataset", "task")}
    out["config_live_json"] = np.frombuffer(json.dumps(keep, sort_keys=True).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "reid_reference_golden.npz"), **out)
    print("wrote reid_reference_golden.npz:", {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
