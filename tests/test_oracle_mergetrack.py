"""Pins the MergeTrack restatements (oracle/mergetrack_oracle.py, oracle/cv_resize_oracle.resize_linear_f32) against the
committed golden vectors made with OpenCV (tests/golden/make_mergetrack_golden.py) and, when cv2 is importable, against
cv2.remap / cv2.resize themselves run here -- bit-exact for the 8-bit remap."""
import os

import numpy as np
import pytest

from oracle import cv_resize_oracle as RZ
from oracle import mergetrack_oracle as M
from oracle import pwc_oracle
from premvos_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mergetrack_golden.npz")


def test_warp_flow_matches_golden():
    g = np.load(GOLD)
    for m, want in zip(g["masks"], g["warped"]):
        np.testing.assert_array_equal(M.warp_flow(m, g["flow"]), want)
    np.testing.assert_array_equal(M.warp_flow(g["gray"], g["flow"], binarize=False), g["remapped"])
    assert g["warped"][0].sum() > 0 and g["warped"][2].sum() == 0


def test_flow_postprocess_matches_golden():
    g = np.load(GOLD)
    H, W = g["post"].shape[:2]
    flo = (g["flow2"] * np.float32(20.0)).astype(np.float32)
    u = RZ.resize_linear_f32(flo[0], H, W) * np.float32(W / 128.0)
    v = RZ.resize_linear_f32(flo[1], H, W) * np.float32(H / 64.0)
    np.testing.assert_array_equal(np.dstack((u, v)), g["post"])          # OpenCV's own code path (IPP off): bit-exact


@pytest.mark.parametrize("h,w,scale", [(480, 854, 3.0), (436, 1024, 12.0), (37, 53, 40.0), (5, 7, 0.4)])
def test_remap_restatement_is_bit_exact_with_cv2(h, w, scale):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(h + w)
    flow = (rng.standard_normal((h, w, 2)) * scale).astype(np.float32)
    flow[: h // 4] = np.round(flow[: h // 4] * 64) / 64
    for img in (synth.synthetic_masks(2, h, w, seed=h)[0], rng.integers(0, 256, (h, w), dtype=np.uint8)):
        mp = M.flow_to_map(flow)
        np.testing.assert_array_equal(M.remap_linear_u8(img, mp), cv2.remap(img, mp.copy(), None, cv2.INTER_LINEAR))


def test_float_resize_restatement_vs_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    a = (rng.standard_normal((112, 256)) * 20).astype(np.float32)
    got = RZ.resize_linear_f32(a, 436, 1024)
    # the IPP path of opencv-python differs in the last bits; OpenCV's own code is what the restatement follows
    assert np.abs(cv2.resize(a, (1024, 436)) - got).max() <= 2e-5 * np.abs(got).max()
    was = cv2.ipp.useIPP()
    try:
        cv2.ipp.setUseIPP(False)
        np.testing.assert_array_equal(cv2.resize(a, (1024, 436)), got)
    finally:
        cv2.ipp.setUseIPP(was)
    want = pwc_oracle.postprocess_flow(a[None].repeat(2, 0) / 20, 436, 1024, 448, 1024)
    assert np.abs(want[:, :, 0] - got * np.float32(1.0)).max() <= 2e-5 * np.abs(got).max()


def test_to_bbox_and_warp_proposals():
    m = np.zeros((10, 12), np.uint8)
    np.testing.assert_array_equal(M.to_bbox(m), [0, 0, 0, 0])
    m[3:6, 4:9] = 1
    m[8, 1] = 1
    np.testing.assert_array_equal(M.to_bbox(m), [1, 3, 8, 6])
    flow = np.zeros((10, 12, 2), np.float32)
    flow[..., 0] = 2                                  # content moves 2 px to the right
    props = [{"mask": m, "final_score": 0.5, "object_score": 0.25, "id": 7}]
    out = M.warp_proposals(props, flow)
    np.testing.assert_array_equal(out[0]["mask"][:, 2:], m[:, :-2])
    np.testing.assert_array_equal(out[0]["bbox"], [3, 3, 8, 6])
    assert out[0]["score"] == 0.75 and out[0]["id"] == 7
    # a half-pixel shift: a pixel stays 1 iff the weights on its set taps reach 1/2 (remap rounds 0.5 up to 1)
    flow[..., 0] = 0.5
    half = M.warp_flow(m, flow)
    np.testing.assert_array_equal(half[3:6, 3:11], np.array([[0, 1, 1, 1, 1, 1, 1, 0]] * 3))
