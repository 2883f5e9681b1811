"""GPU parity tests of MergeTrack's live mask propagation (premvos_warp_masks_u8, premvos_flow_postprocess, the
LivePropagator chain) against oracle/mergetrack_oracle.py and the golden vectors made with OpenCV: bit-exact for the masks
and boxes, 1e-6 for the float flow field."""
import os

import numpy as np
import pytest
import torch

from oracle import cv_resize_oracle as RZ
from oracle import mergetrack_oracle as MO
from oracle import refnet_oracle as RO
from premvos_b200 import _lib, mergetrack, pwc, refnet, synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mergetrack_golden.npz")


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    _lib.lib()


def test_warp_masks_golden_vectors():
    g = np.load(GOLD)
    out, bbox = mergetrack.warp_masks_device(torch.from_numpy(g["masks"]).cuda(), torch.from_numpy(g["flow"]).cuda())
    np.testing.assert_array_equal(out.cpu().numpy(), g["warped"])
    for b, m in zip(bbox.cpu().numpy(), g["warped"]):
        np.testing.assert_array_equal(b, MO.to_bbox(m))
    np.testing.assert_array_equal(mergetrack.warp_flow(g["gray"], g["flow"], binarize=False), g["remapped"])
    np.testing.assert_array_equal(mergetrack.warp_flow(g["masks"][0], g["flow"]), g["warped"][0])


# (480, 854): H*W % 4 == 0 -> vector path; (37, 53), (5, 7): scalar path with a ragged tail; n = 0 and n = 1 edge cases
@pytest.mark.parametrize("h,w,n,scale", [(480, 854, 5, 4.0), (436, 1024, 3, 25.0), (37, 53, 4, 9.0), (5, 7, 1, 0.6), (64, 64, 0, 1.0),
                                         (33, 35, 2, 300.0)])
def test_warp_masks_bit_exact_with_oracle(h, w, n, scale):
    rng = np.random.default_rng(h * 7 + w)
    masks = synth.synthetic_masks(n, h, w, seed=h + n)
    flow = (rng.standard_normal((h, w, 2)) * scale).astype(np.float32)
    flow[: h // 3] = np.round(flow[: h // 3] * 64) / 64          # exact ties of the 1/32-pixel quantisation
    if h > 8:
        flow[h // 2, :4] = [[1e6, -1e6], [3e9, 0], [np.nan, 0], [np.inf, -np.inf]]
    before = _lib.kernel_launch_count()
    out, bbox = mergetrack.warp_masks_device(torch.from_numpy(masks).cuda(), torch.from_numpy(flow).cuda())
    torch.cuda.synchronize()
    assert _lib.kernel_launch_count() - before == (3 if n else 0)
    got, boxes = out.cpu().numpy(), bbox.cpu().numpy()
    for i in range(n):
        want = MO.warp_flow(masks[i], flow)
        np.testing.assert_array_equal(got[i], want)
        np.testing.assert_array_equal(boxes[i], MO.to_bbox(want))
    if n:
        raw, _ = mergetrack.warp_masks_device(torch.from_numpy(masks * 200).cuda(), torch.from_numpy(flow).cuda(), binarize=False,
                                              want_bbox=False)
        np.testing.assert_array_equal(raw.cpu().numpy()[0], MO.remap_linear_u8(masks[0] * 200, MO.flow_to_map(flow)))


def test_warp_properties_full_size():
    # size-independent properties at the DAVIS frame size: zero flow is the identity, an integer translation is a shift
    h, w = 480, 854
    masks = synth.synthetic_masks(6, h, w, seed=3)
    m = torch.from_numpy(masks).cuda()
    out, bbox = mergetrack.warp_masks_device(m, torch.zeros((h, w, 2), device="cuda"))
    assert torch.equal(out, m)
    flow = torch.zeros((h, w, 2), device="cuda")
    flow[..., 0], flow[..., 1] = 5, -3
    out2, bbox2 = mergetrack.warp_masks_device(m, flow)
    want = np.zeros_like(masks)
    want[:, :-3, 5:] = masks[:, 3:, :-5]
    np.testing.assert_array_equal(out2.cpu().numpy(), want)
    for i in range(6):
        np.testing.assert_array_equal(bbox.cpu().numpy()[i], MO.to_bbox(masks[i]))
        np.testing.assert_array_equal(bbox2.cpu().numpy()[i], MO.to_bbox(want[i]))


def test_warp_proposals_surface():
    h, w = 60, 84
    masks = synth.synthetic_masks(3, h, w, seed=21)
    flow = (np.random.default_rng(2).standard_normal((h, w, 2)) * 3).astype(np.float32)
    props = [{"mask": m, "final_score": 0.1 * i, "object_score": 0.2 * i, "id": i + 1} for i, m in enumerate(masks)]
    got = mergetrack.warp_proposals(props, flow)
    want = MO.warp_proposals(props, flow)
    assert len(got) == 3
    for a, b in zip(got, want):
        np.testing.assert_array_equal(a["mask"], b["mask"])
        np.testing.assert_array_equal(a["bbox"], b["bbox"])
        assert a["score"] == b["score"] and a["id"] == b["id"] and a["final_score"] == b["final_score"]
        assert a["segmentation"] == RO.rle_encode(b["mask"])
    assert mergetrack.warp_proposals([], flow) == []


def test_warp_errors():
    m = torch.zeros((1, 8, 8), dtype=torch.uint8, device="cuda")
    f = torch.zeros((8, 8, 2), device="cuda")
    with pytest.raises(TypeError):
        mergetrack.warp_masks_device(m.cpu(), f)
    with pytest.raises(ValueError):
        mergetrack.warp_masks_device(m, torch.zeros((8, 9, 2), device="cuda"))
    with pytest.raises(_lib.PremvosError):
        mergetrack.warp_masks_device(m, f, out=m)        # in place


@pytest.mark.parametrize("H,W", [(436, 1024), (480, 854), (100, 140)])
def test_flow_postprocess_matches_oracle(H, W):
    Hn, Wn = -(-H // 64) * 64, -(-W // 64) * 64
    flow2 = (np.random.default_rng(H).standard_normal((2, 2, Hn // 4, Wn // 4)) * 0.4).astype(np.float32)
    got = mergetrack.flow_postprocess_device(torch.from_numpy(flow2).cuda(), H, W).cpu().numpy()
    for b in range(2):
        flo = (flow2[b] * np.float32(20.0)).astype(np.float32)
        u = RZ.resize_linear_f32(flo[0], H, W) * np.float32(W / float(Wn))
        v = RZ.resize_linear_f32(flo[1], H, W) * np.float32(H / float(Hn))
        want = np.dstack((u, v))
        assert np.abs(got[b] - want).max() <= 1e-6 * np.abs(want).max()


def test_flow_postprocess_golden():
    g = np.load(GOLD)
    H, W = g["post"].shape[:2]
    got = mergetrack.flow_postprocess_device(torch.from_numpy(g["flow2"][None]).cuda(), H, W).cpu().numpy()[0]
    assert np.abs(got - g["post"]).max() <= 1e-6 * np.abs(g["post"]).max()


def test_live_propagator_chain_equals_stagewise_reference_flow():
    # merge.py:95-100 resident: flow(t, t+1) -> frame-resolution flow -> warp masks of t -> boxes -> refine on t+1,
    # against the same chain built from the host-surface pieces (each held against its oracle elsewhere)
    H, W = 100, 140
    f1, f2 = synth.synthetic_frame_pair(H, W, seed=5)
    sd = synth.pwc_synthetic_state_dict(3)
    net = pwc.pwc_dc_net(None)
    net.load_state_dict(sd)
    net.cuda().eval()
    rn = refnet.RefinementNet(max_batch=3, input_size=129, middle_units=0).load_params(synth.refnet_synthetic_params(0, 0))
    live = mergetrack.LivePropagator(net, rn, (H, W))
    masks = synth.synthetic_masks(3, H, W, seed=6)
    res = live.step(torch.from_numpy(masks).cuda(), torch.from_numpy(f1).cuda(), torch.from_numpy(f2).cuda())
    res = {k: v.cpu().numpy() for k, v in res.items()}
    # stage 1 through the reference-shaped host surface
    flow_host = pwc.calculate_flow(net, f1, f2)
    assert res["flow"].shape == (H, W, 2)
    assert np.abs(res["flow"] - flow_host).max() <= 2e-5 * max(1.0, np.abs(flow_host).max())
    # the rest is integer work on the device's own flow field: bit-exact against the oracle
    for i in range(3):
        want = MO.warp_flow(masks[i], res["flow"])
        np.testing.assert_array_equal(res["warped"][i], want)
        np.testing.assert_array_equal(res["bbox"][i], MO.to_bbox(want))
    m_ref, c_ref, _ = rn.refine(f2, res["bbox"])
    np.testing.assert_array_equal(res["masks"], m_ref)
    np.testing.assert_array_equal(res["conf"], c_ref)
