"""GPU tests of the resident pipeline (premvos_b200/pipeline.py) and of the device entry points it uses
(premvos_pwc_forward_u8, premvos_propnet_forward_u8 + copy_results, premvos_refnet_forward): every stage must give the
same bits as the host entry point of the same network, which the per-network tests hold against the CPU oracle."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from premvos_b200 import _lib, pipeline, propnet, pwc, refnet, synth

pytestmark = pytest.mark.gpu

NB = (1, 1, 1, 1)
H, W = 100, 140


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    _lib.lib()


@pytest.fixture(scope="module")
def parts():
    sd = synth.pwc_synthetic_state_dict(3)
    G = synth.propnet_synthetic_params(5, NB)
    S = synth.propnet_synthetic_params(6, NB)
    R = synth.refnet_synthetic_params(0, 0)
    pipe = pipeline.FramePipeline({k: torch.from_numpy(v) for k, v in sd.items()}, G, S, R, (H, W), pairs_per_step=2,
                                  boxes_per_frame=3, refine_batch=2, num_blocks=NB, middle_units=0, refine_input_size=129)
    frames = [synth.synthetic_bgr_frame(H, W, seed=20 + i)[:, :, ::-1].copy() for i in range(4)]   # RGB uint8
    return pipe, sd, G, S, R, frames


def _stage_refs(sd, G, S, R, f0, f1, boxes):
    """The same unit through the host entry points of the three networks."""
    pair, prop, frame = pipeline.prepare_unit(f0, f1)
    net = pwc.pwc_dc_net(None)
    net.load_state_dict(sd)
    net.cuda().eval()
    flow = net.forward_host_u8(pair[None].copy())[0]
    dets = [propnet.ProposalNet(NB).load_params(P)(prop) for P in (G, S)]
    rn = refnet.RefinementNet(max_batch=2, input_size=129, middle_units=0).load_params(R)
    masks, conf, _ = rn.refine(frame, boxes)
    return flow, dets, masks, conf


def test_device_entry_points_equal_host_entry_points(parts):
    pipe, sd, G, S, R, frames = parts
    boxes = synth.synthetic_boxes(3, H, W, seed=9, min_size=20, max_size=90)
    flow, dets, masks, conf = _stage_refs(sd, G, S, R, frames[0], frames[1], boxes)
    pair, prop, frame = pipeline.prepare_unit(frames[0], frames[1])
    # proposal net: uint8 device image -> read_results
    pn = propnet.ProposalNet(NB).load_params(G)
    pn.forward_device(torch.from_numpy(prop).cuda())
    got = pn.read_results(*prop.shape[:2])
    for a, b in zip(got, dets[0]):
        np.testing.assert_array_equal(a, b)
    # refinement net: device frame + boxes (3 boxes with groups of 2: one graph replay + one partial group)
    rn = refnet.RefinementNet(max_batch=2, input_size=129, middle_units=0).load_params(R)
    m, c = rn.refine_device(torch.from_numpy(frame).cuda(), torch.from_numpy(boxes).cuda())
    np.testing.assert_array_equal(m.cpu().numpy(), masks)
    np.testing.assert_array_equal(c.cpu().numpy(), conf)
    with pytest.raises(TypeError):
        rn.refine_device(torch.from_numpy(frame), torch.from_numpy(boxes).cuda())
    with pytest.raises(TypeError):
        pn.forward_device(torch.from_numpy(prop))


def test_pipeline_step_fixed_boxes_and_host_call(parts):
    pipe, sd, G, S, R, frames = parts
    units = [(frames[0], frames[1]), (frames[1], frames[2])]
    prep = [pipeline.prepare_unit(a, b) for a, b in units]
    boxes = np.stack([synth.synthetic_boxes(3, H, W, seed=30 + i, min_size=20, max_size=90) for i in range(2)])
    ff = torch.from_numpy(np.stack([p[0] for p in prep]))
    pi = torch.from_numpy(np.stack([p[1] for p in prep]))
    fr = torch.from_numpy(np.stack([p[2] for p in prep]))
    before = _lib.kernel_launch_count()
    out = pipe.run_device(ff.cuda(), pi.cuda(), fr.cuda(), torch.from_numpy(boxes).cuda())
    torch.cuda.synchronize()
    assert _lib.kernel_launch_count() - before == pipe.launches_per_step()
    out = {k: v.cpu().numpy().copy() for k, v in out.items()}
    for b, (f0, f1) in enumerate(units):
        flow, dets, masks, conf = _stage_refs(sd, G, S, R, f0, f1, boxes[b])
        # a batch-2 PWC handle may plan another (deterministic) split-K factor than the batch-1 handle: equal within rounding
        assert rel_err(out["flow"][b], flow) < 1e-4
        for which in range(2):
            n = len(dets[which][0])
            assert out["det_count"][which, b] == n
            # the pipeline runs the B frames as one batched forward: equal to the batch-1 results up to fp32 summation order
            np.testing.assert_allclose(out["det_boxes"][which, b, :n], dets[which][0], rtol=1e-4, atol=1e-3)
            np.testing.assert_allclose(out["det_probs"][which, b, :n], dets[which][1], rtol=1e-4, atol=1e-5)
        np.testing.assert_array_equal(out["masks"][b], masks)
        np.testing.assert_array_equal(out["conf"][b], conf)
    # the end-to-end call on pinned host buffers gives the same bits, twice (staging is reused)
    for _ in range(2):
        host = pipe.run_host(ff.pin_memory(), pi.pin_memory(), fr.pin_memory(), torch.from_numpy(boxes).pin_memory())
        for k in out:
            np.testing.assert_array_equal(host[k].numpy(), out[k])
    with pytest.raises(ValueError):
        pipe.run_device(ff.cuda()[:1], pi.cuda(), fr.cuda(), None)


def test_pipeline_refines_detected_boxes_and_shards_a_video(parts):
    pipe, sd, G, S, R, frames = parts
    # one rank
    whole = pipeline.run_video(pipe, frames)
    assert sorted(whole) == [0, 1, 2]
    # two ranks, run one after the other on this GPU: the union is the same result, unit by unit
    merged = {}
    for r in range(2):
        part = pipeline.run_video(pipe, frames, rank=r, world=2)
        assert sorted(part) == list(range(r, 3, 2))
        merged.update(part)
    for t in whole:
        for k in whole[t]:
            np.testing.assert_array_equal(np.asarray(merged[t][k]), np.asarray(whole[t][k]))
    # the refined boxes are the combined detections of the two passes (stage 4)
    for t in whole:
        r = whole[t]
        prop = pipeline.prepare_unit(frames[t], frames[t + 1])[1]
        g = r["det_boxes"][0][:r["det_count"][0]]
        s = r["det_boxes"][1][:r["det_count"][1]]
        bx = pipeline.combine_proposals(g, s, (H, W), prop.shape[:2])
        assert len(bx) == r["det_count"].sum()
        bx = bx[:pipe.K]
        assert r["num_boxes"] == len(bx)
        if len(bx):
            rn = refnet.RefinementNet(max_batch=2, input_size=129, middle_units=0).load_params(R)
            masks, conf, _ = rn.refine(frames[t + 1], bx)
            np.testing.assert_array_equal(r["masks"][:len(bx)], masks)
            np.testing.assert_array_equal(r["conf"][:len(bx)], conf)


@pytest.mark.parametrize("sh,sw,dh,dw", [(436, 1024, 448, 1024), (436, 1024, 568, 1333), (480, 854, 512, 896), (480, 854, 749, 1333),
                                         (100, 140, 800, 1120), (37, 53, 64, 64), (64, 64, 37, 53), (5, 7, 64, 128)])
def test_device_resize_is_bit_exact_with_cv2(sh, sw, dh, dw):
    from oracle import cv_resize_oracle as RZ   # pinned against cv2 itself in tests/test_oracle_resize.py
    from premvos_b200 import ops

    class cv2:   # the checker: the restatement (cv2 may be absent on a GPU box)
        INTER_LINEAR = 1
        resize = staticmethod(lambda im, size, interpolation=1: RZ.resize_linear_u8(im, size[1], size[0]))
    rng = np.random.default_rng(sh + dw)
    img = rng.integers(0, 256, (2, sh, sw, 3), dtype=np.uint8)
    got = ops.resize_linear_u8(torch.from_numpy(img).cuda(), dh, dw).cpu().numpy()
    for b in range(2):
        np.testing.assert_array_equal(got[b], cv2.resize(img[b], (dw, dh), interpolation=cv2.INTER_LINEAR))
    # RGB -> resized BGR in one pass (the proposal stage's input), and a single-channel image
    rev = ops.resize_linear_u8(torch.from_numpy(img[0]).cuda(), dh, dw, reverse_channels=True).cpu().numpy()
    np.testing.assert_array_equal(rev, cv2.resize(np.ascontiguousarray(img[0][:, :, ::-1]), (dw, dh)))
    g1 = ops.resize_linear_u8(torch.from_numpy(np.ascontiguousarray(img[0][:, :, :1])).cuda(), dh, dw).cpu().numpy()
    np.testing.assert_array_equal(g1[:, :, 0], cv2.resize(img[0][:, :, 0], (dw, dh)))


def test_device_resize_errors():
    from premvos_b200 import ops
    x = torch.zeros((8, 8, 3), dtype=torch.uint8, device="cuda")
    with pytest.raises(_lib.PremvosError):
        ops.resize_linear_u8(x, 4, 4)            # exact 2x down-scale = INTER_AREA in OpenCV
    with pytest.raises(TypeError):
        ops.resize_linear_u8(x.cpu(), 16, 16)
    with pytest.raises(_lib.PremvosError):
        ops.resize_linear_u8(torch.zeros((8, 8, 2), dtype=torch.uint8, device="cuda"), 16, 16)


def test_pipeline_from_original_frames_equals_host_prepared_inputs(parts):
    pipe, sd, G, S, R, frames = parts
    units = [(frames[0], frames[1]), (frames[2], frames[3])]
    prep = [pipeline.prepare_unit(a, b) for a, b in units]
    boxes = torch.from_numpy(np.stack([synth.synthetic_boxes(3, H, W, seed=40 + i, min_size=20, max_size=90) for i in range(2)]))
    ff = torch.from_numpy(np.stack([p[0] for p in prep])).cuda()
    pi = torch.from_numpy(np.stack([p[1] for p in prep])).cuda()
    fr = torch.from_numpy(np.stack([p[2] for p in prep])).cuda()
    want = {k: v.cpu().numpy().copy() for k, v in pipe.run_device(ff, pi, fr, boxes.cuda()).items()}
    prev = torch.from_numpy(np.stack([u[0] for u in units]))
    cur = torch.from_numpy(np.stack([u[1] for u in units]))
    f2, p2 = pipe.prepare_device(prev.cuda(), cur.cuda())
    np.testing.assert_array_equal(f2.cpu().numpy(), ff.cpu().numpy())       # device resize == cv2.resize, bit for bit
    np.testing.assert_array_equal(p2.cpu().numpy(), pi.cpu().numpy())
    before = _lib.kernel_launch_count()
    got = pipe.run_frames_device(prev.cuda(), cur.cuda(), boxes.cuda())
    torch.cuda.synchronize()
    assert _lib.kernel_launch_count() - before == pipe.launches_per_step_from_frames()
    for k in want:
        np.testing.assert_array_equal(got[k].cpu().numpy(), want[k])
    host = pipe.run_frames_host(prev.pin_memory(), cur.pin_memory(), boxes.pin_memory())
    for k in want:
        np.testing.assert_array_equal(host[k].numpy(), want[k])


def test_pack_mask_bits_round_trip_and_flat_step_results(parts):
    # the transport format of the masks: 8 pixels per byte, any non-zero byte = 1, masks of a size that is not a multiple of 64
    from premvos_b200 import ops
    rng = np.random.default_rng(5)
    for shape in ((3, 2, 37, 53), (5, 100, 140), (1, 1, 8, 8)):
        m = (rng.random(shape) < 0.4).astype(np.uint8) * rng.integers(1, 255, shape).astype(np.uint8)
        packed = ops.pack_mask_bits(torch.from_numpy(m).cuda())
        assert tuple(packed.shape) == shape[:-2] + (ops.packed_mask_bytes(shape[-2] * shape[-1]),)
        np.testing.assert_array_equal(ops.unpack_mask_bits(packed.cpu(), shape[-2], shape[-1]), (m != 0).astype(np.uint8))
    with pytest.raises(TypeError):
        ops.pack_mask_bits(torch.zeros(4, 4, dtype=torch.uint8))
    # one flat buffer per step == the separate outputs
    pipe, sd, G, S, R, frames = parts
    B, K = pipe.B, pipe.K
    prev = torch.from_numpy(np.stack([frames[0]] * B)).cuda()
    cur = torch.from_numpy(np.stack([frames[1]] * B)).cuda()
    boxes = torch.from_numpy(np.stack([synth.synthetic_boxes(K, pipe.H, pipe.W, seed=3, min_size=20, max_size=80)] * B)).cuda()
    out = pipe.run_frames_device(prev, cur, boxes)
    flat = pipe.pack_step_results()
    torch.cuda.synchronize()
    assert flat.numel() < 0.5 * sum(v.numel() * v.element_size() for v in out.values())   # masks 8x smaller
    back = pipe.unpack_step_results(flat)
    for k in ("flow", "det_count", "det_boxes", "det_probs", "conf", "masks"):
        np.testing.assert_array_equal(back[k], out[k].cpu().numpy())
