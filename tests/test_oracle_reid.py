"""ReID oracle (oracle/reid_oracle.py) against hand-derived cases of the arithmetic it restates, and the host logic of
premvos_b200/reid.py that needs no GPU.  The reference ships no vectors for this network (oracle header: PARITY UNPINNED); what CAN
be pinned without TensorFlow is pinned here: the context-region integer arithmetic of DAVIS_Forward_Feed.py:36-58 (worked by
hand below), TF1's legacy bilinear sampling grid, TensorFlow's SAME padding for strided windows, and the variable table of
configs/run."""
import json

import numpy as np
import pytest
import torch

from oracle import reid_oracle as RO
from premvos_b200 import reid, synth


def test_context_region_hand_cases():
    H, W = 480, 854
    boxes = [[10, 20, 100, 50],      # xs = 10 - 10.000002 -> -0 -> 0 ; ys = 20 - 5.000001 -> 15 ; 120 x 60 ; minus the "at least 1"
             [800, 400, 100, 100],   # runs over the right / bottom border: clipped to the frame exactly
             [2.5, 3.5, 0, 0],       # ties round half to even: 2.5 -> 2, 3.5 -> 4 ; empty box -> size -1
             [-30, -10, 50, 40]]     # negative origin: clamped to 0 AFTER rounding, the size keeps the full 1.2 x
    got = RO.apply_context_region(boxes, H, W)
    assert got.tolist() == [[0, 15, 119, 59], [790, 390, 64, 90], [2, 4, -1, -1], [0, 0, 59, 47]]


def test_legacy_resize_grid():
    img = torch.tensor([[0., 1.], [2., 3.]]).view(2, 2, 1)
    out = RO.legacy_resize_bilinear(img, 4, 4)[:, :, 0]
    # src = out * 0.5 -> (0, .5, 1, 1.5): the last sample interpolates towards the clamped neighbour = itself
    want = torch.tensor([[0., .5, 1., 1.], [1., 1.5, 2., 2.], [2., 2.5, 3., 3.], [2., 2.5, 3., 3.]])
    assert torch.equal(out, want)
    # same sampling as the refinement oracle's (pinned there)
    from oracle import refnet_oracle as FO
    x = torch.from_numpy(np.random.default_rng(0).random((13, 17, 3), dtype=np.float32))
    if hasattr(FO, "legacy_resize_bilinear"):
        assert torch.allclose(FO.legacy_resize_bilinear(x, 32, 32), RO.legacy_resize_bilinear(x, 32, 32), atol=1e-6)


def test_same_padding_puts_the_odd_pixel_last():
    x = torch.ones(1, 1, 4, 4)
    w = torch.ones(3, 3, 1, 1)
    assert RO.conv2d_same(x, w, 2)[0, 0].tolist() == [[9., 6.], [6., 4.]]
    assert RO.conv2d_same(x, w, 1)[0, 0].tolist() == [[4., 6., 6., 4.], [6., 9., 9., 6.], [6., 9., 9., 6.], [4., 6., 6., 4.]]
    assert RO.conv2d_same(torch.arange(16.).view(1, 1, 4, 4), torch.ones(1, 1, 1, 1), 2)[0, 0].tolist() == [[0., 2.], [8., 10.]]
    p = RO.max_pool_same(torch.arange(16.).view(1, 1, 4, 4), 3, 3)[0, 0]      # 4 -> 2: one padding row on each side
    assert p.tolist() == [[5., 7.], [13., 15.]]


def test_variable_table_matches_config_run():
    shapes = RO.reid_param_shapes()
    assert shapes == synth.reid_param_shapes()
    assert shapes["conv0/W"] == (3, 3, 3, 64)
    assert shapes["res0/W0"] == (1, 1, 64, 128) and "res1/W0" not in shapes
    assert shapes["res12/W1"] == (3, 3, 512, 512) and shapes["res12/W2"] == (3, 3, 512, 1024) and shapes["res12/W0"] == (1, 1, 512, 1024)
    assert "res13/W0" not in shapes and shapes["res13/W1"] == (3, 3, 1024, 512)
    assert shapes["res15/W1"] == (1, 1, 1024, 512) and shapes["res15/W3"] == (1, 1, 1024, 2048)
    assert shapes["res16/W3"] == (1, 1, 2048, 4096) and shapes["conv1/W"] == (3, 3, 4096, 500)
    assert shapes["fc1/W"] == (2000, 500) and shapes["outputTriplet/W"] == (500, 128)
    assert shapes["res3/bn0/mean_ema"] == (128,) and shapes["res3/bn2/gamma"] == (256,)
    assert sum(int(np.prod(s)) for s in shapes.values()) > 100e6


@pytest.fixture(scope="module")
def params():
    return synth.reid_synthetic_params(0)


def test_forward_and_add_reid(params):
    frame = synth.synthetic_bgr_frame(120, 160, seed=5)
    props = [{"bbox": [20.0, 30.0, 60.0, 50.0]}, {"bbox": [100.0, 80.0, 5.0, 5.0]}]      # the second crop is below 10 px: zeros
    crops = RO.make_crops(frame, [p["bbox"] for p in props])
    assert crops.shape == (2, 128, 128, 3)
    blank = (torch.zeros(3) - torch.from_numpy(RO.IMAGENET_RGB_MEAN)) / torch.from_numpy(RO.IMAGENET_RGB_STD)
    assert torch.equal(crops[1], blank.expand(128, 128, 3))
    emb, inter = RO.reid_forward(params, crops, True)
    assert emb.shape == (2, 128) and torch.isfinite(emb).all()
    assert inter["conv0"].shape == (2, 64, 128, 128) and inter["res0"].shape == (2, 128, 64, 64)
    assert inter["res14"].shape == (2, 1024, 8, 8) and inter["res16"].shape == (2, 4096, 4, 4) and inter["conv1"].shape == (2, 500, 2, 2)
    assert float((emb[0] - emb[1]).abs().max()) > 1e-3       # the embedding depends on the crop
    out = RO.add_ReID(params, [dict(p) for p in props], frame)
    assert len(out[0]["ReID"]) == 128 and isinstance(out[0]["ReID"][0], float)
    assert np.allclose(out[0]["ReID"], emb[0].numpy())
    assert RO.add_ReID(params, [], frame) == []


def test_host_surface_checks(params, tmp_path):
    net = reid.ReIDNet(max_batch=4)
    bad = dict(params)
    del bad["res5/W2"]
    with pytest.raises(RuntimeError, match="missing"):
        net.load_params(bad)
    bad = dict(params)
    bad["extra/W"] = np.zeros(3, np.float32)
    with pytest.raises(RuntimeError, match="unexpected"):
        net.load_params(bad)
    bad = dict(params)
    bad["fc1/W"] = np.zeros((500, 2000), np.float32)
    with pytest.raises(RuntimeError, match="size mismatch for fc1/W"):
        net.load_params(bad)
    with pytest.raises(RuntimeError, match="load_params"):
        reid.ReIDNet()._h()
    # config handling of ReID_net_init (ReID_net_functions.py:19-25): geometry the library does not build, missing weights
    cfg = {"input_size": [96, 96], "load": "nowhere"}
    with pytest.raises(ValueError, match="128 x 128"):
        reid.Engine(cfg)
    path = tmp_path / "live"
    path.write_text(json.dumps({"input_size": [128, 128], "num_classes": 128, "load": str(tmp_path / "no_such_checkpoint.npz")}))
    with pytest.raises(FileNotFoundError):
        reid.ReID_net_init(config_path=str(path))
    with pytest.raises(KeyError):
        reid.Engine({"input_size": [128, 128]})
    # weights written in tensorpack's exchange format come back through the loader
    from premvos_b200 import weights
    small = {"tower0/%s:0" % k: v for k, v in params.items() if not k.startswith("res1")}
    np.savez(tmp_path / "w.npz", **small)
    with pytest.raises(KeyError, match="ReID_net"):
        weights.load_reid_net_variables(str(tmp_path / "w.npz"))
    # add_ReID on an engine whose network is a stand-in: proposals get python float lists, empty lists pass through
    class _Net:
        def embed(self, image, boxes):
            return np.arange(len(boxes) * 128, dtype=np.float32).reshape(len(boxes), 128)
    eng = reid.Engine.__new__(reid.Engine)
    eng.net, eng.config = _Net(), None
    out = reid.add_ReID([{"bbox": [1, 2, 30, 40]}, {"bbox": [5, 6, 70, 80]}], np.zeros((50, 60, 3), np.uint8), eng)
    assert out[1]["ReID"][:2] == [128.0, 129.0] and len(out[0]["ReID"]) == 128
    assert reid.add_ReID([], np.zeros((50, 60, 3), np.uint8), eng) == []


def test_pinned_against_reference_source_vectors(golden_dir):
    """tests/golden/reid_reference_golden.npz (make_reid_reference_goldens.py): the reference's apply_contex_region source run on a
    numpy stand-in for TensorFlow's ops, its numpy normalize, and its configs/live."""
    import os
    g = np.load(os.path.join(golden_dir, "reid_reference_golden.npz"))
    H, W = int(g["ctx_dims"][0]), int(g["ctx_dims"][1])
    assert np.array_equal(RO.apply_context_region(g["ctx_boxes"], H, W), g["ctx_out"].astype(np.int32))
    mean, std = RO.IMAGENET_RGB_MEAN, RO.IMAGENET_RGB_STD
    assert np.array_equal((g["norm_in"] - mean) / std, g["norm_out"])
    cfg = json.loads(bytes(g["config_live_json"]).decode())
    assert cfg["input_size"] == [RO.INPUT_SIZE, RO.INPUT_SIZE] and cfg["context_region_factor_val"] == RO.CONTEXT_REGION_FACTOR
    assert cfg["num_classes"] == RO.EMBEDDING_DIM and cfg["output_embedding_layer"] == "outputTriplet"
    net = cfg["network"]
    prev = "conv0"
    assert net["conv0"] == {"class": "Conv", "n_features": 64, "activation": "linear"}
    feats_now = None
    for name, feats, ks, strides in RO.RESIDUAL_UNITS:
        spec = net[name]
        assert spec["class"] == "ResidualUnit2" and spec["from"] == [prev]
        nf = spec.get("n_features", feats_now[-1] if feats_now else 64)   # ResidualUnit2: n_features=None -> the input's width
        feats_now = tuple(nf) if isinstance(nf, list) else (nf,) * spec.get("n_convs", 2)
        assert tuple(feats) == feats_now, name
        assert len(feats) == spec.get("n_convs", 2)
        want_strides = [s[0] for s in spec["strides"]] if "strides" in spec else [1] * len(feats)
        assert list(strides) == want_strides, name
        want_ks = [f[0] for f in spec["filter_size"]] if "filter_size" in spec else [3] * len(feats)
        assert list(ks) == want_ks, name
        prev = name
    assert net["conv1"]["from"] == [prev] and net["conv1"]["n_features"] == 500 and net["conv1"]["pool_size"] == [3, 3]
    assert net["conv1"]["batch_norm"] is True and net["conv1"]["filter_size"] == [3, 3]
    assert net["fc1"]["n_features"] == 500 and net["fc2"]["n_features"] == 500 and net["fc1"]["batch_norm"] and net["fc2"]["batch_norm"]
    assert net["outputTriplet"]["class"] == "FullyConnectedWithTripletLoss" and net["outputTriplet"]["from"] == ["fc2"]
