"""CPU checks of the refinement-network oracle: hand-derived cases for the TF1 resize kernels it restates, the crop /
guidance bookkeeping, the COCO RLE codec and the tensor sizes the reference's graph is known to produce
(SURVEY.md section 8a, rows R2-R8).  The reference ships no golden vectors for this network."""
import numpy as np
import torch

from oracle import refnet_oracle as O
from premvos_b200 import refnet, synth


def test_legacy_bilinear_resize_has_no_half_pixel_offset():
    img = torch.arange(4, dtype=torch.float32).reshape(1, 1, 4)          # 0 1 2 3
    out = O.tf_resize_bilinear(img, 1, 8)                                 # src = x * 0.5
    np.testing.assert_allclose(out[0, 0].numpy(), [0, 0.5, 1, 1.5, 2, 2.5, 3, 3])   # last sample clamps (hi = min(lo+1, 3))
    out = O.tf_resize_bilinear(img, 1, 2)                                 # src = x * 2 -> pixels 0 and 2
    np.testing.assert_allclose(out[0, 0].numpy(), [0, 2])
    out = O.tf_resize_bilinear(img, 1, 7, align_corners=True)             # src = x * 3/6
    np.testing.assert_allclose(out[0, 0].numpy(), [0, 0.5, 1, 1.5, 2, 2.5, 3])


def test_nearest_resize():
    img = torch.arange(5, dtype=torch.float32).reshape(1, 1, 5)
    np.testing.assert_array_equal(O.tf_resize_nearest(img, 1, 10)[0, 0].numpy(), [0, 0, 1, 1, 2, 2, 3, 3, 4, 4])
    np.testing.assert_array_equal(O.tf_resize_nearest(img, 1, 3)[0, 0].numpy(), [0, 1, 3])   # floor(x * 5/3)


def test_crop_box_and_guidance():
    # Resize.py:151-164: rounded box +- 50 px clipped to the frame; BoundingBox.py:15-19: mask of the rounded box
    assert O.crop_box([100.4, 200.6, 180.5, 300.5], 480, 854) == (50, 151, 230, 350)       # 180.5 -> 180, 300.5 -> 300 (half to even)
    assert O.crop_box([10, 20, 470, 840], 480, 854) == (0, 0, 480, 854)
    img = np.zeros((100, 120, 3), np.float32)
    inputs, crop = O.make_network_input(img, [30.0, 20.0, 40.0, 50.0], size=65)              # x,y,w,h
    assert crop == (0, 0, 100, 120) and inputs.shape == (65, 65, 4)
    g = inputs[..., 3]
    assert set(np.unique(g)) == {0.0, 1.0}
    ys, xs = np.nonzero(g)
    # nearest resize: pixel (y,x) of the 65-grid looks at floor(y*100/65), floor(x*120/65)
    assert ys.min() == int(np.ceil(20 * 65 / 100)) and xs.min() == int(np.ceil(30 * 65 / 120))
    # RGB of a black frame after (v - mean) / std
    np.testing.assert_allclose(inputs[0, 0, :3], -O.IMAGENET_RGB_MEAN / O.IMAGENET_RGB_STD, rtol=1e-6)


def test_rle_codec():
    m = np.zeros((5, 4), np.uint8)
    m[1:3, 1] = 1
    m[4, 3] = 1
    enc = O.rle_encode(m)
    assert enc["size"] == [5, 4]
    # column-major runs: 6 zeros, 2 ones, 11 zeros, 1 one  -> counts [6,2,11,1]; chars: 6->'6', 2->'2', 11->';', 1-2=-1 -> 'O'
    assert enc["counts"] == "62;O"
    np.testing.assert_array_equal(O.rle_decode(enc), m)
    rng = np.random.default_rng(0)
    for _ in range(5):
        big = (rng.uniform(size=(37, 53)) > 0.6).astype(np.uint8)
        e = refnet.rle_encode(big * 255)
        assert e == O.rle_encode(big)
        np.testing.assert_array_equal(refnet.rle_decode(e), big)
    ones = np.ones((3, 3), np.uint8)                                     # starts with a foreground pixel -> leading zero-run
    assert O.rle_encode(ones)["counts"] == "09" and refnet.rle_encode(ones)["counts"] == "09"


def test_graph_sizes_at_385():
    # xception.py sizes for 385: 193 -> 97 -> 49 -> 25 (output stride 16), decoder at 97
    shapes = O.refnet_param_shapes()
    assert shapes["xception_65/entry_flow/conv1_1/weights"] == (3, 3, 4, 32)
    assert shapes["xception_65/middle_flow/block1/unit_16/xception_module/separable_conv3_pointwise/weights"] == (1, 1, 728, 728)
    assert shapes["concat_projection/weights"] == (1, 1, 1280, 256)
    assert shapes["decoder/decoder_conv0_depthwise/depthwise_weights"] == (3, 3, 304, 1)
    n_params = sum(int(np.prod(s)) for k, s in shapes.items())
    assert 40.0e6 < n_params < 42.5e6                                    # SURVEY: 40.8 M parameters
    assert list(shapes.items()) == list(synth.refnet_param_shapes(16).items())
    assert int((385 - 1.0) * 0.25 + 1.0) == 97


def test_forward_small_and_conf_score():
    P = synth.refnet_synthetic_params(1, middle_units=0)
    blocks = O.blocks_with_middle_units(0)
    frame = synth.synthetic_bgr_frame(90, 110, seed=7)
    props = O.do_refinement(P, [{"bbox": [20.0, 10.0, 50.0, 60.0]}], frame, blocks, size=65)
    m = O.rle_decode(props[0]["segmentation"])
    assert m.shape == (90, 110) and 0 < m.sum() < 90 * 110
    assert -1.0 <= float(props[0]["conf_score"]) <= 1.0
    # outside the crop the mask is 0 and the posterior 0 -> contributes +1 to conf_score
    image = (frame / 255).astype(np.float32)
    inputs, crop = O.make_network_input(image, [20.0, 10.0, 50.0, 60.0], 65)
    logits = O.deeplab_logits(P, inputs[None], blocks)[0]
    mask, post = O.segmentation_output(logits, crop, 90, 110, 65)
    cy0, cx0, cy1, cx1 = crop
    outside = np.ones((90, 110), bool)
    outside[cy0:cy1, cx0:cx1] = False
    assert (mask[outside] == 0).all() and (post[outside] == 0).all()
