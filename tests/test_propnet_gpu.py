"""GPU parity of the proposal network (premvos_propnet_* / premvos_topk_host / premvos_nms_host) against the CPU
oracle.  Tolerance (BASELINE.json north_star): <= 1e-3 relative (||d||_inf / ||ref||_inf) on every floating-point
tensor; index tensors (top-k, NMS keep lists) must be bit-exact -- checked on identical inputs for the single ops, and
end to end on seeds whose decisions are not on a threshold."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import propnet_oracle as O
from premvos_b200 import _lib, ops, propnet, synth

pytestmark = pytest.mark.gpu
TOL = 1e-3


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    _lib.lib()


@pytest.mark.parametrize("n,k,seed", [(57270, 1000, 0), (1000, 1000, 1), (5000, 7, 2), (3, 10, 3), (70000, 1024, 4)])
def test_topk_indices_exact(n, k, seed):
    rng = np.random.default_rng(seed)
    s = rng.standard_normal(n).astype(np.float32)
    s[rng.integers(0, n, n // 10)] = s[0]          # heavy ties, some on the k-th value
    if seed == 2:
        s[:] = 1.0                                  # all equal: lowest indices win
    got = ops.top_k(s, k)
    want = np.lexsort((np.arange(n), -s.astype(np.float64)))[:min(k, n)]
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("n,thr,max_out,seed", [(1000, 0.7, 100, 0), (1000, 0.5, 20, 1), (100, 0.5, 20, 2), (1, 0.7, 100, 3),
                                                (1024, 0.3, 2000, 4), (37, 0.7, 5, 5)])
def test_nms_indices_exact(n, thr, max_out, seed):
    rng = np.random.default_rng(seed)
    c = rng.uniform(0, 600, (n, 2))
    wh = rng.uniform(5, 250, (n, 2))
    boxes = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
    boxes[::17] = boxes[1::17][:boxes[::17].shape[0]]           # exact duplicates (IoU = 1)
    boxes[5 % n, 2:] = boxes[5 % n, :2]                         # zero-area box
    scores = rng.standard_normal(n).astype(np.float32)
    scores[::13] = scores[0]                                    # ties
    got = ops.non_max_suppression(boxes, scores, max_out, thr)
    want = O.tf_non_max_suppression(boxes, scores, max_out, thr)
    np.testing.assert_array_equal(got, want)


def _run_pair(nb, H, W, wseed, iseed):
    P = synth.propnet_synthetic_params(wseed, nb)
    img = synth.synthetic_bgr_frame(H, W, seed=iseed).astype(np.float32)
    net = propnet.ProposalNet(nb).load_params(P)
    got = net(img)
    ref, inter = O.propnet_forward(P, img, list(nb), True)
    return net, got, ref, inter


def test_propnet_intermediates_match_oracle():
    nb = (1, 2, 2, 1)
    H, W = 160, 224
    net, got, ref, inter = _run_pair(nb, H, W, 5, 6)
    g = lambda name: net.get_tensor(name, H, W)
    report = []
    fm = inter["featuremap"].numpy()
    report.append(("featuremap", rel_err(g("featuremap").reshape(fm.shape), fm)))
    fh, fw = fm.shape[2:]
    rpn = g("rpn_out").reshape(fh, fw, 80)
    report.append(("rpn_label_logits", rel_err(rpn[..., :15], inter["rpn_label_logits"].numpy())))
    report.append(("rpn_box_logits", rel_err(rpn[..., 15:75].reshape(fh, fw, 15, 4), inter["rpn_box_logits"].numpy())))
    np.testing.assert_array_equal(g("cell_anchors").reshape(15, 4) + np.array([0, 0, 1, 1], np.float32), O.get_all_anchors()[0, 0])
    report.append(("rpn_decoded_boxes", rel_err(g("rpn_decoded_boxes").reshape(-1, 4), inter["rpn_decoded_boxes"].reshape(-1, 4).numpy())))
    print("\n".join("%-22s %.3e" % r for r in report))
    # index tensors: exact
    np.testing.assert_array_equal(g("topk_indices").astype(np.int64), inter["topk_indices"])
    np.testing.assert_array_equal(g("nms_keep").astype(np.int64), inter["nms_keep"])
    n = inter["proposal_boxes"].shape[0]
    report2 = [("proposal_boxes", rel_err(g("proposal_boxes").reshape(n, 4), inter["proposal_boxes"].numpy())),
               ("proposal_scores", rel_err(g("proposal_scores"), inter["proposal_scores"].numpy()))]
    roi = g("roi_resized").reshape(100, 1024, 14, 14)[:n]
    report2.append(("roi_resized", rel_err(roi, inter["roi_resized"].numpy())))
    feat = g("feature_fastrcnn").reshape(100, 2048, 7, 7)[:n]
    report2.append(("feature_fastrcnn", rel_err(feat, inter["feature_fastrcnn"].numpy())))
    logits = g("head_logits").reshape(100, 87)[:n]
    report2.append(("fastrcnn_label_logits", rel_err(logits[:, :2], inter["fastrcnn_label_logits"].numpy())))
    report2.append(("fastrcnn_box_logits", rel_err(logits[:, 2:6], inter["fastrcnn_box_logits"].reshape(n, 4).numpy())))
    report2.append(("second_logits", rel_err(logits[:, 6:], inter["second_logits"].numpy())))
    report2.append(("fastrcnn_all_probs", rel_err(g("fastrcnn_all_probs").reshape(n, 2), inter["fastrcnn_all_probs"].numpy())))
    report2.append(("fastrcnn_all_boxes", rel_err(g("fastrcnn_all_boxes").reshape(n, 4), inter["fastrcnn_all_boxes"].reshape(n, 4).numpy())))
    print("\n".join("%-22s %.3e" % r for r in report2))
    bad = [r for r in report + report2 if not (r[1] < TOL)]
    assert not bad, bad
    # final outputs: same selection (exact indices), values within tolerance
    np.testing.assert_array_equal(g("final_box_index").astype(np.int64), inter["pred_indices"][:, 0])
    for a, b in zip(got, ref):
        assert a.shape == b.shape and a.dtype == b.dtype
        if a.dtype == np.int64:
            np.testing.assert_array_equal(a, b)
        elif a.size:
            assert rel_err(a, b) < TOL


def test_propnet_second_seed_and_determinism():
    nb = (1, 1, 1, 1)
    H, W = 128, 128
    net, got, ref, inter = _run_pair(nb, H, W, 8, 9)
    got2 = net(synth.synthetic_bgr_frame(H, W, seed=9).astype(np.float32))
    for a, b in zip(got, got2):
        np.testing.assert_array_equal(a, b)
    assert got[0].shape == ref[0].shape
    np.testing.assert_array_equal(net.get_tensor("nms_keep", H, W).astype(np.int64), inter["nms_keep"])
    for a, b in zip(got, ref):
        if a.dtype != np.int64 and a.size:
            assert rel_err(a, b) < TOL


def test_detect_one_image_end_to_end():
    nb = (1, 1, 1, 1)
    P = synth.propnet_synthetic_params(10, nb)
    net = propnet.ProposalNet(nb).load_params(P)
    frame = synth.synthetic_bgr_frame(120, 160, seed=11)
    res = propnet.detect_one_image(frame, net, size=192, max_size=256)
    ref = O.detect_one_image(frame, lambda im: O.propnet_forward(P, im.astype(np.float32), list(nb)), size=192, max_size=256)
    assert len(res) == len(ref)
    for a, b in zip(res, ref):
        assert rel_err(a.box, b.box) < TOL and abs(float(a.score) - float(b.score)) < TOL
    assert propnet.convert_results_to_json(res) == O.convert_results_to_json(ref)


def test_errors_are_loud():
    net = propnet.ProposalNet((1, 1, 1, 1))
    with pytest.raises(RuntimeError):
        net(np.zeros((64, 64, 3), np.float32))       # no parameters loaded
    with pytest.raises(ValueError):
        propnet.ProposalNet((1, 1, 1, 1)).load_params(synth.propnet_synthetic_params(0, (1, 1, 1, 1)))(np.zeros((64, 64), np.float32))
    with pytest.raises(_lib.PremvosError):
        ops.non_max_suppression(np.zeros((2000, 4), np.float32), np.zeros(2000, np.float32), 10, 0.5)
