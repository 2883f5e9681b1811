"""Oracle restatements of the TensorFlow-side box arithmetic against tests/golden/tf_shim_golden.npz -- vectors made by executing
the reference functions' own source with a numpy stand-in for their TensorFlow ops (tests/golden/make_tf_shim_goldens.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import propnet_oracle as PO, refnet_oracle as RO


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "tf_shim_golden.npz"))


def test_decode_bbox_target_equals_reference_source(g):
    assert abs(float(g["dec_clip"]) - PO.BBOX_DECODE_CLIP) < 1e-12
    got = PO.decode_bbox_target(torch.from_numpy(g["dec_logits"]), torch.from_numpy(g["dec_anchors"])).numpy()
    # torch.exp and numpy.exp may differ in the last bit; everything else is the same float32 arithmetic in the same order
    assert np.abs(got - g["dec_out"]).max() <= 2e-7 * np.abs(g["dec_out"]).max()
    assert np.isfinite(got).all()


def test_clip_boxes_tf_equals_reference_source(g):
    h, w = [int(v) for v in g["clip_hw"]]
    got = PO.clip_boxes_t(torch.from_numpy(g["clip_in"]), h, w).numpy()
    assert np.array_equal(got, g["clip_out"])


def test_roi_align_box_transform_equals_reference_source(g, monkeypatch):
    seen = {}

    def fake(image, boxes, crop):
        seen["boxes"], seen["crop"] = np.array(boxes), crop
        return torch.zeros(len(boxes), image.shape[0], crop, crop)
    monkeypatch.setattr(PO, "tf_crop_and_resize", fake)
    h, w = [int(v) for v in g["roi_fm_hw"]]
    PO.roi_align(torch.zeros(1, 8, h, w), torch.from_numpy(g["roi_boxes"]), 14)
    assert seen["crop"] == int(g["roi_crop"])
    assert np.array_equal(seen["boxes"].astype(np.float32), g["roi_tf_boxes"])


def test_refinement_crop_boxes_equal_reference_source(g):
    H, W = [int(v) for v in g["crop_hw"]]
    for bbox, want, shp in zip(g["crop_bboxes"], g["crop_out"], g["crop_slice_shapes"]):
        got = RO.crop_box(bbox, H, W)
        assert list(got) == want.tolist()
        # the slice image[y0:y1, x0:x1] the reference takes (numpy / TF slice semantics agree for these non-negative bounds)
        assert [max(got[2] - got[0], 0), max(got[3] - got[1], 0)] == shp.tolist()
