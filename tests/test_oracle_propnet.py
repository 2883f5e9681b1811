"""CPU checks that pin the proposal-network oracle to what the reference itself states, plus hand-derived cases
for the TensorFlow kernels it restates (the reference ships no golden vectors for this network)."""
import numpy as np
import pytest
import torch

from oracle import propnet_oracle as O
from premvos_b200 import propnet, synth


def test_anchor_table_from_reference_docstring():
    # proposal_net/utils/generate_anchors.py:20-38 (the reference's own known answer)
    want = np.array([[-84., -40., 99., 55.], [-176., -88., 191., 103.], [-360., -184., 375., 199.],
                     [-56., -56., 71., 71.], [-120., -120., 135., 135.], [-248., -248., 263., 263.],
                     [-36., -80., 51., 95.], [-80., -168., 95., 183.], [-168., -344., 183., 359.]])
    np.testing.assert_array_equal(O.generate_anchors(), want)


def test_all_anchors_field():
    a = O.get_all_anchors()
    assert a.shape == (83, 83, 15, 4) and a.dtype == np.float32        # MAX_SIZE // 16 = 83
    cell = O.generate_anchors(16, scales=np.array(O.ANCHOR_SIZES) / 16.0, ratios=np.array(O.ANCHOR_RATIOS))
    np.testing.assert_array_equal(a[0, 0, :, :2], cell[:, :2])
    np.testing.assert_array_equal(a[0, 0, :, 2:], cell[:, 2:] + 1)       # data.py:73
    np.testing.assert_array_equal(a[2, 5] - a[0, 0], np.tile([80, 32, 80, 32], (15, 1)))


def test_custom_resize_matches_survey_shape():
    assert O.custom_resize_shape(480, 854) == (749, 1333)
    assert O.custom_resize_shape(480, 640) == (800, 1067)
    assert propnet.custom_resize_shape(480, 854) == (749, 1333)


def test_nms_hand_cases():
    boxes = np.array([[0, 0, 10, 10], [1, 1, 11, 11], [20, 20, 30, 30], [0, 0, 10, 10.5]], np.float32)
    scores = np.array([0.9, 0.8, 0.7, 0.95], np.float32)
    # IoU(3,0) = 100/105 > 0.7 ; IoU(3,1) = 9*9.5/(105+100-85.5)=0.715 > 0.7 ; box 2 is disjoint
    np.testing.assert_array_equal(O.tf_non_max_suppression(boxes, scores, 10, 0.7), [3, 2])
    np.testing.assert_array_equal(O.tf_non_max_suppression(boxes, scores, 1, 0.7), [3])
    # strictly greater: IoU exactly equal to the threshold is kept
    b2 = np.array([[0, 0, 2, 2], [0, 0, 2, 1]], np.float32)             # IoU = 0.5
    np.testing.assert_array_equal(O.tf_non_max_suppression(b2, [1.0, 0.5], 10, 0.5), [0, 1])
    # ties -> lower index first; degenerate (zero-area) boxes never suppress nor get suppressed
    b3 = np.array([[0, 0, 4, 4], [0, 0, 4, 4], [1, 1, 1, 3]], np.float32)
    np.testing.assert_array_equal(O.tf_non_max_suppression(b3, [0.5, 0.5, 0.9], 10, 0.3), [2, 0])


def test_crop_and_resize_hand_cases():
    img = torch.arange(12, dtype=torch.float32).reshape(1, 3, 4)        # [C=1, H=3, W=4]
    # identity crop: box covering the whole image sampled at the pixel centres
    out = O.tf_crop_and_resize(img, np.array([[0, 0, 1, 1]], np.float32), 3)
    np.testing.assert_allclose(out[0, 0, :, 0].numpy(), [0, 4, 8])
    np.testing.assert_allclose(out[0, 0, 0].numpy(), [0, 1.5, 3])
    # extrapolation: samples outside [0, H-1] x [0, W-1] are 0
    out = O.tf_crop_and_resize(img, np.array([[-0.5, 0, 0.5, 1]], np.float32), 3)
    np.testing.assert_allclose(out[0, 0, 0].numpy(), [0, 0, 0])
    np.testing.assert_allclose(out[0, 0, 1].numpy(), [0, 1.5, 3])


def test_roi_align_constant_and_linear_images():
    fm = torch.ones(1, 2, 12, 16)
    fm[0, 1] = torch.arange(16, dtype=torch.float32).view(1, 16).expand(12, 16)   # f(x) = x
    boxes = torch.tensor([[2.0, 3.0, 9.0, 8.0]])
    out = O.roi_align(fm, boxes, 14)
    assert tuple(out.shape) == (1, 2, 14, 14)
    np.testing.assert_allclose(out[0, 0].numpy(), 1.0, rtol=1e-6)
    # bin centres of a linear image: x0 + (j + 0.5) * w/14 - 0.5
    want = 2.0 + (np.arange(14) + 0.5) * 7.0 / 14 - 0.5
    np.testing.assert_allclose(out[0, 1, 0].numpy(), want, rtol=1e-5)


def test_decode_and_clip():
    anchors = torch.tensor([[0., 0., 16., 16.]])
    d = O.decode_bbox_target(torch.tensor([[0.5, -0.25, np.log(2.0), 100.0]]), anchors)
    w, h = 32.0, np.float32(np.exp(np.float32(O.BBOX_DECODE_CLIP))) * 16           # th clipped at log(1333/16)
    np.testing.assert_allclose(d.numpy(), [[16 - w / 2, 4 - h / 2, 16 + w / 2, 4 + h / 2]], rtol=1e-6)
    c = O.clip_boxes_t(d, 100, 50)
    np.testing.assert_allclose(c.numpy(), [[0, 0, 32, 100]], rtol=1e-6)


def test_forward_contract_small_net():
    nb = (1, 1, 1, 1)
    P = synth.propnet_synthetic_params(3, nb)
    img = synth.synthetic_bgr_frame(96, 128, seed=4).astype(np.float32)
    out, inter = O.propnet_forward(P, img, list(nb), True)
    boxes, probs, labels, post, slabels, spost = out
    n = boxes.shape[0]
    assert 0 < n <= 20 and probs.shape == (n,) and labels.dtype == np.int64 and post.shape == (n, 2) and spost.shape == (n, 81)
    assert (probs > 0.5).all() and (labels == 1).all()
    assert tuple(inter["featuremap"].shape) == (1, 1024, 6, 8)
    # the reference's posterior quirk: every row is label_probs[0]
    np.testing.assert_array_equal(post, np.tile(inter["fastrcnn_all_probs"].numpy()[0], (n, 1)))
    res = O.detect_one_image(synth.synthetic_bgr_frame(60, 80, seed=4), lambda im: O.propnet_forward(P, im.astype(np.float32), list(nb)),
                             size=96, max_size=128)
    js = O.convert_results_to_json(res)
    assert all(set(r) == {"bbox", "score"} and len(r["bbox"]) == 4 for r in js)


def test_mask_head_oracle_shapes_and_fill_full_mask():
    # model.py:495-509 / train.py:297-309 / eval.py:35-58 restatements: shapes, the deconvolution's phase structure, the paste
    import torch
    from premvos_b200 import synth
    P = synth.maskrcnn_synthetic_params(0)
    assert {k: v.shape for k, v in P.items()} == dict(O.maskrcnn_param_shapes())
    feat = torch.randn(3, 2048, 7, 7, generator=torch.Generator().manual_seed(0))
    logits = O.maskrcnn_head(P, feat)
    assert tuple(logits.shape) == (3, 1, 14, 14)
    # out[2y+dy, 2x+dx] depends only on in[y, x] and W[dy, dx]
    d = torch.relu(torch.einsum("nchw,oc->nohw", feat, torch.as_tensor(P["maskrcnn/deconv/W"][1, 0])) +
                   torch.as_tensor(P["maskrcnn/deconv/b"])[None, :, None, None])
    want = torch.einsum("nohw,o->nhw", d, torch.as_tensor(P["maskrcnn/conv/W"][0, 0, :, 0])) + float(P["maskrcnn/conv/b"][0])
    assert torch.allclose(logits[:, 0, 1::2, 0::2], want, rtol=1e-4, atol=1e-4)
    # fill_full_mask: integer rectangle int(x0 + .5) .. int(x1 - .5), identity when the box is 14 x 14, 2x2 mean when it is 7 x 7
    m = np.random.default_rng(1).uniform(0, 1, (14, 14)).astype(np.float32)
    full = O.fill_full_mask(np.array([3.2, 5.4, 17.4, 19.3], np.float32), m, (30, 40))
    assert full.shape == (30, 40) and full[:5].sum() == 0 and full[:, :3].sum() == 0
    np.testing.assert_array_equal(full[5:19, 3:17], (m > 0.5).astype(np.uint8))
    half = O.fill_full_mask(np.array([0, 0, 7, 7], np.float32), m, (10, 10))
    np.testing.assert_array_equal(half[:7, :7], (m.reshape(7, 2, 7, 2).mean(axis=(1, 3)) > 0.5).astype(np.uint8))
    cv2 = pytest.importorskip("cv2")
    for box in ([1.0, 2.0, 30.7, 22.2], [0.0, 0.0, 3.4, 25.0], [10.2, 10.1, 10.9, 10.4]):
        b = np.array(box, np.float32)
        x0, y0 = list(map(int, b[:2] + 0.5)); x1, y1 = list(map(int, b[2:] - 0.5))
        x1, y1 = max(x0, x1), max(y0, y1)
        ref = np.zeros((30, 40), np.uint8)
        ref[y0:y1 + 1, x0:x1 + 1] = (cv2.resize(m, (x1 + 1 - x0, y1 + 1 - y0)) > 0.5).astype(np.uint8)
        np.testing.assert_array_equal(O.fill_full_mask(b, m, (30, 40)), ref)
