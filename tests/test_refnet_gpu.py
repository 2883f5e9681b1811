"""GPU parity of the refinement network (premvos_refnet_*) against the CPU oracle.  Tolerance (BASELINE.json
north_star): <= 1e-3 relative (||d||_inf / ||ref||_inf) on every floating-point tensor; masks must agree except on pixels
whose two logits are within the tolerance of each other (reported)."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import refnet_oracle as O
from premvos_b200 import _lib, refnet, synth

pytestmark = pytest.mark.gpu
TOL = 1e-3


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    _lib.lib()


def _check(P, blocks, net, frame, boxes, S, check_inter=True):
    n = len(boxes)
    masks, conf, post = net.refine(frame, boxes, want_posteriors=True)
    image = (frame / 255).astype(np.float32)
    H, W = frame.shape[:2]
    NB = net.max_batch
    report = []
    last = (n - 1) // NB * NB                                  # the test hook shows the LAST batch
    for i in range(n):
        inputs, crop = O.make_network_input(image, boxes[i], S)
        logits, inter = O.deeplab_logits(P, inputs[None], blocks, True)
        if check_inter and i >= last:
            j = i - last
            g = lambda name, shp: net.get_tensor(name).reshape((NB,) + shp)[j]
            report.append(("net_input[%d]" % i, rel_err(g("net_input", (8, S, S))[:4], inter["net_input"][0].permute(2, 0, 1).numpy())))
            np.testing.assert_array_equal(net.get_tensor("crops").reshape(NB, 4)[j], crop)
            for name in ("low_level", "xception_out", "aspp_concat", "aspp_out", "decoder_in", "decoder_out"):
                t = inter[name][0].numpy()
                report.append(("%s[%d]" % (name, i), rel_err(g(name, t.shape), t)))
            lg = inter["logits"][0].numpy()
            got_lg = net.get_tensor("logits").reshape(NB, lg.shape[0], lg.shape[1], 16)[j][..., :2]
            report.append(("logits[%d]" % i, rel_err(got_lg, lg)))
        mask_ref, post_ref = O.segmentation_output(logits[0], crop, H, W, S)
        report.append(("posterior[%d]" % i, rel_err(post[i], post_ref)))
        diff = masks[i] != mask_ref
        if diff.any():   # only pixels whose foreground posterior is within tolerance of 0.5 may flip
            lgS = O.tf_resize_bilinear(logits[0].permute(2, 0, 1), S, S)
            # the flipped pixels must be ones whose two logits are closer than the tolerance allows to resolve
            margin = float((lgS[1] - lgS[0]).abs().min())
            assert diff.sum() <= 1e-4 * diff.size and margin < 4 * TOL * float(lgS.abs().max()), (int(diff.sum()), margin)
        cs = O.conf_score(mask_ref, post_ref)
        report.append(("conf_score[%d]" % i, abs(float(conf[i]) - float(cs)) / max(abs(float(cs)), 1e-6)))
    print("\n".join("%-20s %.3e" % r for r in report))
    bad = [r for r in report if not (r[1] < TOL)]
    assert not bad, bad
    return masks, conf


def test_refnet_small_net_matches_oracle():
    S, mu = 129, 1
    P = synth.refnet_synthetic_params(0, mu)
    blocks = O.blocks_with_middle_units(mu)
    net = refnet.RefinementNet(max_batch=2, input_size=S, middle_units=mu).load_params(P)
    frame = synth.synthetic_bgr_frame(120, 160, seed=5)
    boxes = synth.synthetic_boxes(3, 120, 160, seed=3, min_size=30, max_size=100)      # 3 boxes, batches of 2: ragged last batch
    boxes[2] = [0.0, 0.0, 160.0, 120.0]                                                # whole frame
    _check(P, blocks, net, frame, boxes, S)


def test_guidance_of_boxes_with_negative_origin_follows_numpy_slicing():
    # util/BoundingBox.py:15-19 writes encoded[y0:y1, x0:x1] = 1: a negative start counts from the far edge (numpy), so a box that
    # begins left of / above the frame selects a different (usually empty) region -- reproduced, not "fixed"
    S, mu = 129, 0
    P = synth.refnet_synthetic_params(3, mu)
    blocks = O.blocks_with_middle_units(mu)
    net = refnet.RefinementNet(max_batch=4, input_size=S, middle_units=mu).load_params(P)
    frame = synth.synthetic_bgr_frame(120, 160, seed=6)
    boxes = np.array([[-12.0, 10.0, 70.0, 60.0], [20.0, -8.0, 50.0, 70.0], [-30.0, -20.0, 100.0, 90.0], [10.0, 10.0, 60.0, 60.0]], np.float32)
    _check(P, blocks, net, frame, boxes, S, check_inter=True)


def test_refnet_full_xception65_at_385():
    # BASELINE config C4 geometry: 385x385 crops, Xception-65 (16 middle units), DAVIS-shaped 480x854 frame
    S, mu = 385, 16
    P = synth.refnet_synthetic_params(2, mu)
    blocks = O.blocks_with_middle_units(mu)
    net = refnet.RefinementNet(max_batch=4, input_size=S, middle_units=mu).load_params(P)
    frame = synth.synthetic_bgr_frame(480, 854, seed=3)
    boxes = synth.synthetic_boxes(2, 480, 854, seed=3)
    torch.set_num_threads(__import__("os").cpu_count() or 1)
    masks, conf = _check(P, blocks, net, frame, boxes, S)
    # batched == one at a time (the reference's batch-1 protocol).  The CTA-pair kernel balances a launch by cutting the K loop of
    # a few items between two pairs (stream-K); where it cuts depends on the number of crops in the launch, so the fp32 summation
    # order -- not the result beyond rounding -- differs between a batched and a batch-1 launch: equal to ~1e-6, masks equal except
    # for pixels whose two logits tie to that precision
    m1, c1, _ = net.refine(frame, boxes[1:2])
    assert (m1[0] != masks[1]).mean() < 1e-4
    assert abs(float(c1[0]) - float(conf[1])) < 1e-5
    # the same launch twice: bit for bit (fixed summation order for a given launch shape)
    m2, c2, _ = net.refine(frame, boxes[1:2])
    np.testing.assert_array_equal(m2[0], m1[0])
    assert c2[0] == c1[0]


def test_refnet_benchmarked_plan_40_crops():
    # the launch plan bench.py times: one launch group of 40 crops of 385x385 through Xception-65 (CTA-pair GEMMs with the stream-K
    # schedule of 300-item launches, F8 depthwise path).  Crops 0 and 39 against the oracle; every crop against a batch-1 forward of
    # the same network (another launch plan: single-CTA kernels) within rounding.
    S, mu, K = 385, 16, 40
    P = synth.refnet_synthetic_params(2, mu)
    blocks = O.blocks_with_middle_units(mu)
    net = refnet.RefinementNet(max_batch=K, input_size=S, middle_units=mu).load_params(P)
    frame = synth.synthetic_bgr_frame(436, 1024, seed=3)
    boxes = synth.synthetic_boxes(K, 436, 1024, seed=103)
    masks, conf, post = net.refine(frame, boxes, want_posteriors=True)
    torch.set_num_threads(__import__("os").cpu_count() or 1)
    image = (frame / 255).astype(np.float32)
    for i in (0, K - 1):
        inputs, crop = O.make_network_input(image, boxes[i], S)
        logits = O.deeplab_logits(P, inputs[None], blocks)
        mask_ref, post_ref = O.segmentation_output(logits[0], crop, 436, 1024, S)
        assert rel_err(post[i], post_ref) < TOL
        assert (masks[i] != mask_ref).mean() < 1e-4
        cs = O.conf_score(mask_ref, post_ref)
        assert abs(float(conf[i]) - float(cs)) / max(abs(float(cs)), 1e-6) < TOL
    single = refnet.RefinementNet(max_batch=1, input_size=S, middle_units=mu).load_params(P)
    for i in range(0, K, 3):
        m1, c1, p1 = single.refine(frame, boxes[i:i + 1], want_posteriors=True)
        assert rel_err(post[i], p1[0]) < 0.5 * TOL          # two launch plans, each within TOL of the oracle (measured: 2e-4)
        assert (m1[0] != masks[i]).mean() < 1e-4
        assert abs(float(c1[0]) - float(conf[i])) < 0.5 * TOL


def test_do_refinement_surface_and_rle():
    S, mu = 129, 0
    P = synth.refnet_synthetic_params(4, mu)
    engine = refnet.refinement_net_init(P, max_batch=8, input_size=S, middle_units=mu)
    frame = synth.synthetic_bgr_frame(100, 140, seed=8)
    props = [{"bbox": b.tolist()} for b in synth.synthetic_boxes(5, 100, 140, seed=9, min_size=20, max_size=90)]
    out = refnet.do_refinement([dict(p) for p in props], frame, engine)
    ref = O.do_refinement(P, [dict(p) for p in props], frame, O.blocks_with_middle_units(mu), S)
    for a, b in zip(out, ref):
        assert set(a) == {"bbox", "segmentation", "conf_score"} and a["segmentation"]["size"] == [100, 140]
        ma, mb = refnet.rle_decode(a["segmentation"]), O.rle_decode(b["segmentation"])
        assert (ma != mb).mean() < 1e-3
        assert abs(float(a["conf_score"]) - float(b["conf_score"])) < TOL
    # the reference's inner protocol (one validation_step per proposal) gives the same masks
    data = engine.valid_data.set_up_data_for_image(frame, [p["bbox"] for p in props])
    step = engine.trainer.validation_step(feed_dict=engine.valid_data.get_feed_dict_for_next_step(data, 2),
                                          extraction_keys=[refnet.SEGMENTATION_POSTERIORS_ORIGINAL_SIZE,
                                                           refnet.SEGMENTATION_MASK_ORIGINAL_SIZE, refnet.OBJ_TAGS])
    ex = step[refnet.EXTRACTIONS]
    assert ex[refnet.SEGMENTATION_MASK_ORIGINAL_SIZE][0].shape == (1, 100, 140)
    np.testing.assert_array_equal(ex[refnet.SEGMENTATION_MASK_ORIGINAL_SIZE][0][0], refnet.rle_decode(out[2]["segmentation"]))
    assert int(ex[refnet.OBJ_TAGS][0][0].decode("utf-8")) == 2
    assert refnet.do_refinement([], frame, engine) == []


def test_errors_are_loud():
    with pytest.raises(RuntimeError):
        refnet.RefinementNet(max_batch=1, input_size=65, middle_units=0).refine(np.zeros((50, 50, 3), np.uint8), [[1, 1, 10, 10]])
    net = refnet.RefinementNet(max_batch=1, input_size=65, middle_units=0).load_params(synth.refnet_synthetic_params(0, 0))
    with pytest.raises(ValueError):
        net.refine(np.zeros((50, 50, 3), np.float32), [[1, 1, 10, 10]])
