"""Oracle restatements and CPU-side host mirrors against vectors produced by the REFERENCE'S OWN functions
(tests/golden/make_reference_function_goldens.py executes their unmodified source in the build container): fill_full_mask,
clip_boxes, CustomResize, np_box_ops.iou, warp_flow, writeFlowFile / get_flow."""
import os

import numpy as np
import pytest

from oracle import mergetrack_oracle as MO
from oracle import propnet_oracle as PO
from oracle import pwc_oracle
from oracle import refnet_oracle as RO
from premvos_b200 import propnet, pwc

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_functions_golden.npz"))


def test_fill_full_mask_equals_reference_function():
    H, W = G["ffm_full"].shape[1:]
    for b, m, want in zip(G["ffm_boxes"], G["ffm_masks"], G["ffm_full"]):
        np.testing.assert_array_equal(PO.fill_full_mask(b, m, (H, W)), want)
    assert G["ffm_full"].sum() > 0


def test_clip_boxes_and_custom_resize_equal_reference_functions():
    H, W = G["ffm_full"].shape[1:]
    np.testing.assert_array_equal(PO.clip_boxes_np(G["clip_in"].copy(), (H, W)), G["clip_out"])
    np.testing.assert_array_equal(propnet.clip_boxes(G["clip_in"].copy(), (H, W)), G["clip_out"])
    for (h, w), want in zip(G["resize_in"], G["resize_out"]):
        assert PO.custom_resize_shape(int(h), int(w)) == tuple(want)
        assert propnet.custom_resize_shape(int(h), int(w)) == tuple(want)
    assert tuple(G["resize_out"][0]) == (749, 1333) and tuple(G["resize_out"][1]) == (568, 1333)     # the bench sizes


def test_tf_iou_agrees_with_reference_np_box_ops():
    b = G["iou_boxes"]
    got = np.array([[PO.tf_iou(b, i, j) for j in range(len(b))] for i in range(len(b))], np.float64)
    assert np.abs(got - G["iou_out"]).max() < 1e-6


def test_warp_flow_equals_reference_function():
    for m, want in zip(G["wf_masks"], G["wf_warped"]):
        np.testing.assert_array_equal(MO.warp_flow(m, G["wf_flow"]), want)
    np.testing.assert_array_equal(MO.warp_flow(G["wf_gray"], G["wf_flow"], binarize=False), G["wf_remapped"])
    assert G["wf_warped"].sum() > 0


def test_flo_file_format_equals_reference_writer_and_reader(tmp_path):
    flow = G["wf_flow"]
    np.testing.assert_array_equal(G["flo_read"], flow)                     # reference writer -> reference reader
    assert pwc_oracle.flo_bytes(flow) == G["flo_bytes"].tobytes()          # oracle bytes == reference writer's bytes
    fn = str(tmp_path / "a.flo")
    pwc.writeFlowFile(fn, flow)                                            # host mirror: same bytes, and reads the reference's file
    assert open(fn, "rb").read() == G["flo_bytes"].tobytes()
    ref_fn = str(tmp_path / "ref.flo")
    open(ref_fn, "wb").write(G["flo_bytes"].tobytes())
    np.testing.assert_array_equal(pwc.readFlowFile(ref_fn), flow)


def test_guidance_mask_and_normalisation_equal_reference_functions():
    for b, want in zip(G["guid_boxes"], G["guid_masks"]):
        np.testing.assert_array_equal(RO.encode_bbox_as_mask_np(b, (48, 70)), want)
    assert G["guid_masks"][3].sum() == 4          # 2.5 -> 2, 3.5 -> 4: numpy rounds half to even
    np.testing.assert_array_equal(RO.IMAGENET_RGB_MEAN, G["norm_mean"])
    np.testing.assert_array_equal(RO.IMAGENET_RGB_STD, G["norm_std"])
    np.testing.assert_array_equal(RO.normalize(G["norm_in"]), G["norm_out"])
    assert np.abs(G["norm_back"] - G["norm_in"]).max() < 1e-6


def test_anchor_field_and_config_constants_equal_the_reference():
    import hashlib
    field = PO.get_all_anchors()
    assert tuple(G["anchors_shape"]) == field.shape == (83, 83, 15, 4) and field.dtype == np.float32
    assert hashlib.sha1(np.ascontiguousarray(field).tobytes()).digest() == G["anchors_sha1"].tobytes()
    np.testing.assert_array_equal(field[0, 0], G["anchors_cell00"])
    np.testing.assert_array_equal(field[5, 7], G["anchors_cell57"])
    # config.py constants the oracle (and the CUDA library) hard-code
    want = [PO.ANCHOR_STRIDE, PO.MAX_SIZE, PO.SHORT_EDGE_SIZE, PO.TEST_PRE_NMS_TOPK, PO.TEST_POST_NMS_TOPK, PO.RESULTS_PER_IM]
    assert list(G["config_consts"]) == want
    assert list(G["config_thresh"]) == [PO.RPN_PROPOSAL_NMS_THRESH, PO.FASTRCNN_NMS_THRESH, PO.RESULT_SCORE_THRESH]
