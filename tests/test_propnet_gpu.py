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
    if n > 34:
        m = min(boxes[::17].shape[0], boxes[1::17].shape[0])
        boxes[::17][:m] = boxes[1::17][:m]                       # exact duplicates (IoU = 1)
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
    P_ = synth.propnet_synthetic_params(5, nb)
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
    bad = [r for r in report if not (r[1] < TOL)]
    assert not bad, bad
    # ---- discrete stages: the oracle is fed the GPU's own upstream tensors, so the inputs are IDENTICAL and the
    # index tensors must be bit-exact (a 1e-5 score difference may legitimately reorder near-ties otherwise)
    scores_g = g("rpn_scores")
    boxes_g = torch.from_numpy(g("rpn_decoded_boxes").reshape(-1, 4))
    pb, ps, dbg = O.generate_rpn_proposals(boxes_g, torch.from_numpy(scores_g), H, W)
    np.testing.assert_array_equal(g("topk_indices").astype(np.int64), dbg["topk_indices"])
    np.testing.assert_array_equal(g("nms_keep").astype(np.int64), dbg["nms_keep"])
    n = pb.shape[0]
    np.testing.assert_array_equal(g("proposal_boxes").reshape(n, 4), pb.numpy())
    np.testing.assert_array_equal(g("proposal_scores"), ps.numpy())
    # the top-k SET agrees with the pure oracle run up to near-ties at the cut
    assert len(set(dbg["topk_indices"]) ^ set(inter["topk_indices"])) <= 4
    # ---- continuous stages downstream, on the GPU's proposals
    report2 = []
    roi_ref = O.roi_align(inter["featuremap"], pb * np.float32(1.0 / 16), 14)
    roi = g("roi_resized").reshape(100, 1024, 14, 14)[:n]
    report2.append(("roi_resized", rel_err(roi, roi_ref.numpy())))
    feat_ref = O.resnet_conv5(P_, roi_ref, nb[-1])
    feat = g("feature_fastrcnn").reshape(100, 2048, 7, 7)[:n]
    report2.append(("feature_fastrcnn", rel_err(feat, feat_ref.numpy())))
    pooled = feat_ref.mean(dim=(2, 3))
    report2.append(("pooled", rel_err(g("pooled").reshape(100, 2048)[:n], pooled.numpy())))
    t = lambda k: torch.as_tensor(P_[k])
    cls = pooled @ t("fastrcnn/class/W") + t("fastrcnn/class/b")
    box = pooled @ t("fastrcnn/box/W") + t("fastrcnn/box/b")
    sec = pooled @ t("secondclassification/class/W") + t("secondclassification/class/b")
    logits = g("head_logits").reshape(100, 87)[:n]
    report2.append(("fastrcnn_label_logits", rel_err(logits[:, :2], cls.numpy())))
    report2.append(("fastrcnn_box_logits", rel_err(logits[:, 2:6], box.numpy())))
    report2.append(("second_logits", rel_err(logits[:, 6:], sec.numpy())))
    probs_ref = torch.softmax(cls, dim=1)
    dec = O.clip_boxes_t(O.decode_bbox_target(box.reshape(n, 1, 4) / torch.as_tensor(O.FASTRCNN_BBOX_REG_WEIGHTS), pb.unsqueeze(1)), H, W)
    report2.append(("fastrcnn_all_probs", rel_err(g("fastrcnn_all_probs").reshape(n, 2), probs_ref.numpy())))
    report2.append(("fastrcnn_all_boxes", rel_err(g("fastrcnn_all_boxes").reshape(n, 4), dec.reshape(n, 4).numpy())))
    print("\n".join("%-22s %.3e" % r for r in report2))
    bad = [r for r in report2 if not (r[1] < TOL)]
    assert not bad, bad
    # ---- final selection on the GPU's own probabilities / boxes: exact indices
    pred, fprobs = O.fastrcnn_predictions(torch.from_numpy(g("fastrcnn_all_boxes").reshape(n, 1, 4)),
                                          torch.from_numpy(g("fastrcnn_all_probs").reshape(n, 2)))
    np.testing.assert_array_equal(g("final_box_index").astype(np.int64), pred[:, 0])
    boxes, probs, labels, post, slabels, spost = got
    np.testing.assert_array_equal(probs, fprobs)
    np.testing.assert_array_equal(boxes, g("fastrcnn_all_boxes").reshape(n, 4)[pred[:, 0]])
    assert (labels == 1).all() and labels.dtype == np.int64 and slabels.dtype == np.int64
    np.testing.assert_array_equal(post, np.tile(g("fastrcnn_all_probs").reshape(n, 2)[0], (len(probs), 1)))   # train.py:287-288
    np.testing.assert_array_equal(slabels, np.argmax(post, -1) + 1)
    assert rel_err(spost, np.tile(torch.softmax(sec, 1).numpy()[0], (len(probs), 1))) < TOL
    # and the pure oracle run agrees on the number of detections and, box by box, within tolerance
    assert abs(len(ref[1]) - len(probs)) <= 1
    if len(ref[1]) == len(probs):
        assert rel_err(boxes, ref[0]) < TOL and rel_err(probs, ref[1]) < TOL


def test_propnet_second_seed_and_determinism():
    nb = (1, 1, 1, 1)
    H, W = 128, 128
    net, got, ref, inter = _run_pair(nb, H, W, 8, 9)
    got2 = net(synth.synthetic_bgr_frame(H, W, seed=9).astype(np.float32))
    for a, b in zip(got, got2):
        np.testing.assert_array_equal(a, b)
    fm = inter["featuremap"].numpy()
    assert rel_err(net.get_tensor("featuremap", H, W).reshape(fm.shape), fm) < TOL
    assert abs(got[0].shape[0] - ref[0].shape[0]) <= 1


def test_batched_forward_equals_single_image_forwards():
    # set_option("batch", 3): three frames through every launch at once; per image the same results as a batch-1 forward
    # up to fp32 summation order (another split-K plan), the same index tensors, and the oracle's feature map
    import torch
    nb = (1, 2, 2, 1)
    H, W, B = 160, 224, 3
    P = synth.propnet_synthetic_params(5, nb)
    net = propnet.ProposalNet(nb).load_params(P)
    imgs = np.stack([synth.synthetic_bgr_frame(H, W, seed=6 + i) for i in range(B)])          # uint8 BGR
    singles, fms, idx = [], [], []
    for i in range(B):
        net.forward_device(torch.from_numpy(imgs[i]).cuda())
        singles.append(net.read_results(H, W))
        fms.append(net.get_tensor("featuremap", H, W).copy())
        idx.append({k: net.get_tensor(k, H, W).copy() for k in ("topk_indices", "rpn_scores")})
    per_forward = net.launches_per_forward(H, W, B)      # builds the batched handle (warm-up + graph capture) first
    before = _lib.kernel_launch_count()
    net.forward_device(torch.from_numpy(imgs).cuda())
    torch.cuda.synchronize()
    assert _lib.kernel_launch_count() - before == per_forward + 1      # + the uint8 -> fp32 conversion
    fm = net.get_tensor("featuremap", H, W, B).reshape(B, -1)
    _, inter = O.propnet_forward(P, imgs[1].astype(np.float32), list(nb), True)
    assert rel_err(fm[1], inter["featuremap"].numpy().reshape(-1)) < TOL
    counts = torch.zeros((B,), dtype=torch.int32, device="cuda")
    boxes = torch.zeros((B, 20, 4), device="cuda")
    probs = torch.zeros((B, 20), device="cuda")
    net.copy_results_device(H, W, counts, boxes, probs, batch=B)
    for i in range(B):
        assert rel_err(fm[i], fms[i]) < 1e-4
        g = lambda name: net.get_tensor("%s@%d" % (name, i), H, W, B)
        # the batched plan sums in another order (split-K): scores move by ~1e-5, so anchors whose scores are closer than that
        # may swap places in the top-k order (and greedy NMS amplifies a swap).  So: (1) scores / decoded boxes agree within
        # rounding with the batch-1 forward, (2) the top-k SET is the same up to such ties, (3) the discrete stages are
        # index-exact against the oracle fed with THIS image's own device tensors (identical inputs)
        assert rel_err(g("rpn_scores"), idx[i]["rpn_scores"]) < 1e-4
        tk, tk1 = g("topk_indices"), idx[i]["topk_indices"]
        assert len(tk) == len(tk1) and len(set(tk.tolist()) ^ set(tk1.tolist())) <= 0.02 * len(tk1)
        pb, ps, dbg = O.generate_rpn_proposals(torch.from_numpy(g("rpn_decoded_boxes").reshape(-1, 4)), torch.from_numpy(g("rpn_scores")), H, W)
        np.testing.assert_array_equal(tk.astype(np.int64), dbg["topk_indices"])
        np.testing.assert_array_equal(g("nms_keep").astype(np.int64), dbg["nms_keep"])
        n = pb.shape[0]
        np.testing.assert_array_equal(g("proposal_boxes").reshape(n, 4), pb.numpy())
        # RoI features of image i come from image i's feature map (not a neighbour's): RoIAlign + conv5 vs the oracle
        _, inter_i = O.propnet_forward(P, imgs[i].astype(np.float32), list(nb), True)
        roi_ref = O.roi_align(inter_i["featuremap"], pb * np.float32(1.0 / 16), 14)
        roi = net.get_tensor("roi_resized", H, W, B).reshape(B, 100, 1024, 14, 14)[i, :n]
        assert rel_err(roi, roi_ref.numpy()) < TOL
        pooled = O.resnet_conv5(P, roi_ref, nb[-1]).mean(dim=(2, 3)).numpy()
        assert rel_err(g("pooled").reshape(100, 2048)[:n], pooled) < TOL
        # final selection on the device's own probabilities / boxes: exact indices; the result rows are those boxes
        pred, fprobs = O.fastrcnn_predictions(torch.from_numpy(g("fastrcnn_all_boxes").reshape(n, 1, 4)),
                                              torch.from_numpy(g("fastrcnn_all_probs").reshape(n, 2)))
        got = net.read_results(H, W, B, i)
        assert int(counts[i]) == len(got[0]) == len(pred) and abs(len(got[0]) - len(singles[i][0])) <= 1
        np.testing.assert_array_equal(g("final_box_index").astype(np.int64), pred[:, 0])
        np.testing.assert_array_equal(got[1], fprobs)
        np.testing.assert_array_equal(got[0], g("fastrcnn_all_boxes").reshape(n, 4)[pred[:, 0]])
        np.testing.assert_array_equal(boxes[i, :len(got[0])].cpu().numpy(), got[0])
        np.testing.assert_array_equal(probs[i, :len(got[0])].cpu().numpy(), got[1])
    with pytest.raises(_lib.PremvosError):
        net.get_tensor("nms_keep@3", H, W, B)


def test_mask_head_matches_oracle():
    # MODE_MASK (model.py:495-509, train.py:297-309): final_masks vs the oracle run on the DEVICE's final boxes (so that the
    # comparison does not depend on a detection decision), fill_full_mask bit-exact vs the oracle's OpenCV restatement
    nb = (1, 1, 1, 1)
    H, W = 128, 160
    P = synth.propnet_synthetic_params(8, nb)
    P.update(synth.maskrcnn_synthetic_params(8))
    net = propnet.ProposalNet(nb, mode_mask=True).load_params(P)
    with pytest.raises(RuntimeError):
        propnet.ProposalNet(nb, mode_mask=True).load_params(synth.propnet_synthetic_params(8, nb))     # mask variables missing
    img = synth.synthetic_bgr_frame(H, W, seed=9).astype(np.float32)
    got = net(img)
    assert len(got) == 7
    boxes, labels, masks = got[0], got[2], got[6]
    n = len(boxes)
    assert masks.shape == (n, 14, 14)
    ref_out, inter = O.propnet_forward(P, img, list(nb), True)
    assert len(ref_out) == 7 and abs(len(ref_out[0]) - n) <= 1
    if n == 0:      # the mask branch must still be exercised: feed the oracle's and the device's RoIs by hand below
        boxes = np.array([[10, 12, 90, 100], [40, 8, 150, 60]], np.float32)
        labels = np.ones(2, np.int64)
    want = O.final_masks(P, inter["featuremap"], boxes, labels, list(nb))
    if n:
        assert np.abs(masks - want).max() < 2e-4, np.abs(masks - want).max()
        assert masks.min() >= 0 and masks.max() <= 1 and masks.std() > 0.01
    # the detection outputs are the ones of the mask-less graph
    plain = propnet.ProposalNet(nb).load_params(synth.propnet_synthetic_params(8, nb))(img)
    for a, b in zip(got[:6], plain):
        np.testing.assert_array_equal(a, b)
    # fill_full_mask on the device == oracle (bit-exact): random masks and boxes incl. 14x14, 7x7, 1-pixel and border boxes
    rng = np.random.default_rng(3)
    m = rng.uniform(0, 1, (12, 14, 14)).astype(np.float32)
    bx = np.array([[3.2, 5.4, 17.4, 19.3], [0, 0, 7, 7], [10.2, 10.1, 10.9, 10.4], [0, 0, 160, 128], [100.5, 60.5, 159.6, 127.7],
                   [20, 30, 34, 44], [5.5, 5.5, 6.4, 90.0], [0.4, 0.4, 2.6, 2.6], [50, 50, 57, 57], [1, 1, 29, 15], [150, 120, 160, 128],
                   [30.49, 30.51, 61.5, 61.49]], np.float32)
    full = propnet.fill_full_masks(bx, m, (H, W))
    for i in range(len(bx)):
        np.testing.assert_array_equal(full[i], O.fill_full_mask(bx[i], m[i], (H, W)))
    np.testing.assert_array_equal(propnet.fill_full_mask(bx[0], m[0], (H, W)), full[0])
    assert propnet.fill_full_masks(np.zeros((0, 4)), np.zeros((0, 14, 14)), (H, W)).shape == (0, H, W)
    # detect_one_image with masks: every result carries a full-image binary mask
    frame = synth.synthetic_bgr_frame(96, 120, seed=11)
    res = propnet.detect_one_image(frame, net, size=128, max_size=160)
    for r in res:
        assert r.mask.shape == (96, 120) and r.mask.dtype == np.uint8 and set(np.unique(r.mask)) <= {0, 1}


def test_mask_head_on_given_rois_matches_oracle():
    # the same head driven through get_tensor: deterministic coverage of the mask branch even when nothing is detected
    nb = (1, 1, 1, 1)
    H, W = 128, 160
    P = synth.propnet_synthetic_params(12, nb)
    P.update(synth.maskrcnn_synthetic_params(12))
    net = propnet.ProposalNet(nb, mode_mask=True).load_params(P)
    img = synth.synthetic_bgr_frame(H, W, seed=13).astype(np.float32)
    got = net(img)
    _, inter = O.propnet_forward(P, img, list(nb), True)
    n = len(got[0])
    if n:
        want = O.final_masks(P, inter["featuremap"], got[0], got[2], list(nb))
        assert np.abs(net.get_tensor("final_masks", H, W).reshape(n, 14, 14) - want).max() < 2e-4


def test_detect_one_image_end_to_end():
    nb = (1, 1, 1, 1)
    P = synth.propnet_synthetic_params(10, nb)
    net = propnet.ProposalNet(nb).load_params(P)
    frame = synth.synthetic_bgr_frame(120, 160, seed=11)
    res = propnet.detect_one_image(frame, net, size=192, max_size=256)
    ref = O.detect_one_image(frame, lambda im: O.propnet_forward(P, im.astype(np.float32), list(nb)), size=192, max_size=256)
    assert abs(len(res) - len(ref)) <= 1
    if len(res) == len(ref):   # same selection unless a decision sits on a threshold
        for a, b in zip(res, ref):
            assert rel_err(a.box, b.box) < TOL and abs(float(a.score) - float(b.score)) < TOL
    js = propnet.convert_results_to_json(res)
    assert all(set(r) == {"bbox", "score"} and len(r["bbox"]) == 4 for r in js)


def test_errors_are_loud():
    net = propnet.ProposalNet((1, 1, 1, 1))
    with pytest.raises(RuntimeError):
        net(np.zeros((64, 64, 3), np.float32))       # no parameters loaded
    with pytest.raises(ValueError):
        propnet.ProposalNet((1, 1, 1, 1)).load_params(synth.propnet_synthetic_params(0, (1, 1, 1, 1)))(np.zeros((64, 64), np.float32))
    with pytest.raises(_lib.PremvosError):
        ops.non_max_suppression(np.zeros((2000, 4), np.float32), np.zeros(2000, np.float32), 10, 0.5)


def test_propnet_full_size_resnet101():
    # BASELINE config C3: 480x854 DAVIS frame -> CustomResize -> 749x1333, ResNet-101 (3,4,23,3), 100 RoIs
    nb = (3, 4, 23, 3)
    H, W = O.custom_resize_shape(480, 854)
    assert (H, W) == (749, 1333)
    P = synth.propnet_synthetic_params(1, nb)
    import cv2
    img = cv2.resize(synth.synthetic_bgr_frame(480, 854, seed=2), (W, H)).astype(np.float32)
    net = propnet.ProposalNet(nb).load_params(P)
    got = net(img)
    got2 = net(img)
    for a, b in zip(got, got2):
        np.testing.assert_array_equal(a, b)                     # deterministic
    n = got[0].shape[0]
    assert 0 <= n <= 20 and all(np.isfinite(a).all() for a in got if a.dtype != np.int64)
    assert net.get_tensor("featuremap", H, W).size == 1024 * 46 * 83
    # backbone parity at full size against the oracle (57 270 anchors downstream)
    torch.set_num_threads(__import__("os").cpu_count() or 1)
    fm = O.pretrained_resnet_conv4(P, O.image_preprocess(img), list(nb[:3])).numpy()
    assert rel_err(net.get_tensor("featuremap", H, W).reshape(fm.shape), fm) < TOL
    # discrete stages on identical inputs: exact
    pb, ps, dbg = O.generate_rpn_proposals(torch.from_numpy(net.get_tensor("rpn_decoded_boxes", H, W).reshape(-1, 4)),
                                           torch.from_numpy(net.get_tensor("rpn_scores", H, W)), H, W)
    np.testing.assert_array_equal(net.get_tensor("topk_indices", H, W).astype(np.int64), dbg["topk_indices"])
    np.testing.assert_array_equal(net.get_tensor("nms_keep", H, W).astype(np.int64), dbg["nms_keep"])


def test_propnet_benchmarked_plan_batch4_resnet101():
    # the launch plan bench.py times: ResNet-101 (3,4,23,3), batch 4, 568x1333 (what 1024x436 resizes to).  Per image: backbone,
    # RPN scores / decoded boxes against the oracle; the discrete stages (top-k, NMS, final selection) index-exact on the
    # device's own tensors; RoIAlign, the conv5 head, the head logits and the final detections of one image against the oracle
    # run on that image's proposals.
    import cv2
    nb = (3, 4, 23, 3)
    H, W = O.custom_resize_shape(436, 1024)
    assert (H, W) == (568, 1333)
    B = 4
    P = synth.propnet_synthetic_params(1, nb)
    net = propnet.ProposalNet(nb).load_params(P)
    imgs = np.stack([cv2.resize(synth.synthetic_bgr_frame(436, 1024, seed=20 + i), (W, H)) for i in range(B)])   # uint8 BGR
    net.forward_device(torch.from_numpy(imgs).cuda())
    torch.cuda.synchronize()
    torch.set_num_threads(__import__("os").cpu_count() or 1)
    fm_all = net.get_tensor("featuremap", H, W, B)
    fm_all = fm_all.reshape(B, 1024, -1)
    n_det = 0
    for i in range(B):
        g = lambda name: net.get_tensor("%s@%d" % (name, i), H, W, B)
        img = imgs[i].astype(np.float32)
        if i in (0, 3):   # two full oracle backbones (a few seconds each) are enough to pin the batched plan
            fm = O.pretrained_resnet_conv4(P, O.image_preprocess(img), list(nb[:3]))
            fshape = tuple(fm.shape[1:])
            assert rel_err(fm_all[i].reshape(fshape), fm[0].numpy()) < TOL
            label, box = O.rpn_head(P, fm)
            assert rel_err(g("rpn_scores"), label.numpy().reshape(-1)) < TOL
        pb, ps, dbg = O.generate_rpn_proposals(torch.from_numpy(g("rpn_decoded_boxes").reshape(-1, 4)), torch.from_numpy(g("rpn_scores")), H, W)
        np.testing.assert_array_equal(g("topk_indices").astype(np.int64), dbg["topk_indices"])
        np.testing.assert_array_equal(g("nms_keep").astype(np.int64), dbg["nms_keep"])
        n = pb.shape[0]
        np.testing.assert_array_equal(g("proposal_boxes").reshape(n, 4), pb.numpy())
        pred, fprobs = O.fastrcnn_predictions(torch.from_numpy(g("fastrcnn_all_boxes").reshape(n, 1, 4)),
                                              torch.from_numpy(g("fastrcnn_all_probs").reshape(n, 2)))
        got = net.read_results(H, W, B, i)
        assert len(got[0]) == len(pred)
        np.testing.assert_array_equal(g("final_box_index").astype(np.int64), pred[:, 0])
        np.testing.assert_array_equal(got[1], fprobs)
        n_det += len(pred)
        if i == 0:        # the RoI head at full depth: RoIAlign -> conv5 (3 bottlenecks, 2048 channels) -> pooled feature -> logits
            fm_dev = torch.from_numpy(fm_all[i].reshape((1,) + fshape).copy())
            roi_ref = O.roi_align(fm_dev, pb * np.float32(1.0 / 16), 14)
            roi = net.get_tensor("roi_resized", H, W, B).reshape(B, 100, 1024, 14, 14)[i, :n]
            assert rel_err(roi, roi_ref.numpy()) < TOL
            pooled = O.resnet_conv5(P, roi_ref, nb[-1]).mean(dim=(2, 3))
            assert rel_err(g("pooled").reshape(100, 2048)[:n], pooled.numpy()) < TOL
            Wc, bc = torch.as_tensor(P["fastrcnn/class/W"]), torch.as_tensor(P["fastrcnn/class/b"])
            probs_ref = torch.softmax(pooled @ Wc + bc, dim=1).numpy()
            assert rel_err(g("fastrcnn_all_probs").reshape(n, 2), probs_ref) < TOL
    assert n_det >= 0
