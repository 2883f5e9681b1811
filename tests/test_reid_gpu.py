"""ReID network on the GPU (through the C ABI: premvos_reidnet_*) against oracle/reid_oracle.py: crop boxes bit-exact, the
normalised crops to float rounding, every recorded unit output and the 128-d embedding within 1e-3 (||.||_inf relative,
BASELINE.md 4.5: the tolerance of the bf16x3 tensor-core convolutions)."""
import numpy as np
import pytest
import torch

from oracle import reid_oracle as RO
from premvos_b200 import reid, synth
from conftest import rel_err

pytestmark = pytest.mark.gpu

H, W = 240, 427
BOXES = np.array([[40, 30, 120, 90],        # ordinary box
                  [380, 200, 80, 60],       # runs over the right / bottom border
                  [-12.5, -3.5, 90, 70],    # negative origin, .5 ties
                  [200, 100, 6, 30],        # narrower than 10 px after the context region: blank crop
                  [0, 0, 427, 240],         # whole frame
                  [150.3, 60.7, 33.3, 140.2]], dtype=np.float32)


@pytest.fixture(scope="module")
def setup():
    P = synth.reid_synthetic_params(0)
    net = reid.ReIDNet(max_batch=4).load_params(P)       # 6 boxes -> two launch groups (4 + 2)
    frame = synth.synthetic_bgr_frame(H, W, seed=11)
    return P, net, frame


def test_embeddings_and_intermediates_match_oracle(setup):
    P, net, frame = setup
    emb = net.embed(frame, BOXES)
    crops = RO.make_crops(frame, BOXES)
    ref, inter = RO.reid_forward(P, crops, True)
    assert emb.shape == (6, 128)
    assert rel_err(emb, ref.numpy()) < 1e-3
    # state of the LAST group = boxes 4, 5
    got_crops = net.get_tensor("crops").reshape(4, 4)[:2].astype(np.int32)
    assert np.array_equal(got_crops, RO.apply_context_region(BOXES, H, W)[4:6])
    x = net.get_tensor("net_input").reshape(4, 8, 128, 128)[:2, :3]
    assert np.abs(x - crops[4:6].permute(0, 3, 1, 2).numpy()).max() < 2e-5      # split-bf16 storage of O(2) values
    for name in ("conv0", "res0", "res2", "res5", "res11", "res14", "res16"):
        want = inter[name][4:6].numpy()
        got = net.get_tensor(name).reshape((4,) + want.shape[1:])[:2]
        assert rel_err(got, want) < 1e-3, name
    conv1 = net.get_tensor("conv1").reshape(4, 4, 4, 512)[:2, :, :, :500]
    pooled = conv1.reshape(2, 2, 2, 2, 2, 500).max(axis=(2, 4))                 # [n, qy, qx, c]
    assert rel_err(pooled.transpose(0, 3, 1, 2), inter["conv1"][4:6].numpy()) < 1e-3


def test_blank_crop_and_group_independence(setup):
    P, net, frame = setup
    one = net.embed(frame, BOXES[3:4])                      # a single blank crop
    ref = RO.reid_forward(P, RO.make_crops(frame, BOXES[3:4])).numpy()
    assert rel_err(one, ref) < 1e-3
    # the same box gives the same embedding wherever it sits in a launch group (no cross-crop state)
    a = net.embed(frame, BOXES[[0, 1, 2]])
    b = net.embed(frame, BOXES[[2, 0]])
    assert np.array_equal(a[0], b[1]) and np.array_equal(a[2], b[0])
    assert net.embed(frame, np.zeros((0, 4), np.float32)).shape == (0, 128)


def test_device_entry_point_and_surface(setup):
    P, net, frame = setup
    host = net.embed(frame, BOXES)
    dev = net.embed_device(torch.from_numpy(frame).cuda(), torch.from_numpy(BOXES).cuda())
    torch.cuda.synchronize()
    assert np.array_equal(dev.cpu().numpy(), host)
    with pytest.raises(TypeError):
        net.embed_device(torch.from_numpy(frame), torch.from_numpy(BOXES).cuda())
    eng = reid.ReID_net_init(params=P, max_batch=4)
    props = reid.add_ReID([{"bbox": b.tolist()} for b in BOXES[:2]], frame, eng)
    ref = RO.add_ReID(P, [{"bbox": b.tolist()} for b in BOXES[:2]], frame)
    for p, r in zip(props, ref):
        assert isinstance(p["ReID"], list) and len(p["ReID"]) == 128 and isinstance(p["ReID"][0], float)
        assert rel_err(p["ReID"], r["ReID"]) < 1e-3
    assert net.launches_per_forward() > 50


def test_live_propagator_attaches_embeddings(setup):
    """merge.py:95-101 resident: flow -> warp -> boxes -> refinement AND the ReID embeddings of the same boxes on frame t+1."""
    from premvos_b200 import mergetrack, pwc, refnet
    P, net, _ = setup
    h, w = 100, 140
    f1, f2 = synth.synthetic_frame_pair(h, w, seed=5)
    flow_net = pwc.pwc_dc_net(None)
    flow_net.load_state_dict(synth.pwc_synthetic_state_dict(3))
    flow_net.cuda().eval()
    rn = refnet.RefinementNet(max_batch=3, input_size=129, middle_units=0).load_params(synth.refnet_synthetic_params(0, 0))
    live = mergetrack.LivePropagator(flow_net, rn, (h, w), reid_net=net)
    masks = synth.synthetic_masks(3, h, w, seed=6)
    res = live.step(torch.from_numpy(masks).cuda(), torch.from_numpy(f1).cuda(), torch.from_numpy(f2).cuda())
    torch.cuda.synchronize()
    bbox = res["bbox"].cpu().numpy()
    assert res["reid"].shape == (3, 128)
    want = RO.reid_forward(P, RO.make_crops(f2, bbox)).numpy()
    assert rel_err(res["reid"].cpu().numpy(), want) < 1e-3
    assert np.array_equal(res["reid"].cpu().numpy(), net.embed(f2, bbox))
