"""Host logic of the refinement surface that needs no GPU: the reference's config files, the no-argument
refinement_net_init() (MergeTrack/refinement_net_functions.py:19-24) and the stage-5 directory protocol
(refinement_net/forwarding/FewShotSegmentationForwarder.py:85-155: read combined_proposals/<video>/<frame>.json, write
refined_proposals/<video>/<frame>.json with 'segmentation' + 'conf_score' added to every proposal)."""
import json
import os

import numpy as np
import pytest

from premvos_b200 import refnet


class _FakeNet:
    """stands in for the device network: box -> filled rectangle mask, conf = area fraction"""
    max_batch, input_size, middle_units = 4, 385, 16

    def refine(self, image, boxes, want_posteriors=False):
        H, W = image.shape[:2]
        masks = np.zeros((len(boxes), H, W), np.uint8)
        conf = np.zeros(len(boxes), np.float32)
        for i, (x, y, w, h) in enumerate(boxes):
            masks[i, int(y):int(y + h), int(x):int(x + w)] = 1
            conf[i] = masks[i].mean()
        return masks, conf, None


def _write_inputs(root):
    import cv2
    img_dir, bb_dir = os.path.join(root, "JPEGImages"), os.path.join(root, "combined_proposals")
    rng = np.random.default_rng(0)
    want = {}
    for video, frames in (("bear", ("00000", "00001")), ("dog", ("00000",))):
        os.makedirs(os.path.join(img_dir, video)); os.makedirs(os.path.join(bb_dir, video))
        for fr in frames:
            cv2.imwrite(os.path.join(img_dir, video, fr + ".jpg"), rng.integers(0, 255, (40, 60, 3), dtype=np.uint8))
            props = [{"bbox": [5.0, 4.0, 20.0, 10.0], "score": 0.9}, {"bbox": [30.0, 10.0, 12.0, 25.0], "score": 0.4}]
            if fr == "00001":
                props = []                                        # a frame without proposals is passed through
            with open(os.path.join(bb_dir, video, fr + ".json"), "w") as f:
                json.dump(props, f)
            want[(video, fr)] = props
    return img_dir, bb_dir, want


def test_stage5_directory_protocol(tmp_path):
    img_dir, bb_dir, want = _write_inputs(str(tmp_path))
    out_dir = os.path.join(str(tmp_path), "refined_proposals")
    cfg = refnet.Config({"image_input_dir": img_dir + "/", "bb_input_dir": bb_dir + "/", "output_dir": out_dir + "/",
                         "input_size_train": [385, 385], "load": "unused"})
    eng = refnet.Engine(_FakeNet.__new__(refnet.RefinementNet))          # an Engine around a net object ...
    eng.net = _FakeNet()                                                   # ... whose device calls are faked
    eng.trainer.net = eng.net
    eng.config = cfg
    assert eng.run() == 3
    for (video, fr), props in want.items():
        with open(os.path.join(out_dir, video, fr + ".json")) as f:
            got = json.load(f)
        assert len(got) == len(props)
        for g, p in zip(got, props):
            assert g["bbox"] == p["bbox"] and g["score"] == p["score"]
            assert g["segmentation"]["size"] == [40, 60] and isinstance(g["segmentation"]["counts"], str)
            m = refnet.rle_decode(g["segmentation"])
            x, y, w, h = (int(v) for v in p["bbox"])
            assert m.sum() == w * h and m[y:y + h, x:x + w].all()
            assert abs(float(g["conf_score"]) - w * h / (40 * 60)) < 1e-6


def test_config_getters_and_live_config_defaults(tmp_path):
    p = os.path.join(str(tmp_path), "live")
    with open(p, "w") as f:     # the keys of refinement_net/configs/live that the inference path reads
        json.dump({"model": "live_model", "load": "../weights/x/refinement_specific_weights", "batch_size_eval": 1,
                   "input_size_train": [385, 385], "freeze_batchnorm": True}, f)
    c = refnet.Config(p)
    assert c.string("model") == "live_model" and c.int("batch_size_eval") == 1 and c.int_list("input_size_train") == [385, 385]
    assert c.bool("freeze_batchnorm") is True and c.string("absent", "dflt") == "dflt"
    with pytest.raises(KeyError):
        c.string("absent")
    assert refnet.LIVE_CONFIG == "refinement_net/configs/live"


def test_refinement_net_init_without_arguments_reads_the_live_config(tmp_path, monkeypatch):
    # cwd = the reference's `code/` directory: no config there -> the same failure mode as the reference (file not found)
    monkeypatch.chdir(tmp_path)
    with pytest.raises(FileNotFoundError):
        refnet.refinement_net_init()
    # with the config in place but no checkpoint, the error names the checkpoint the config points to
    os.makedirs("refinement_net/configs")
    with open(refnet.LIVE_CONFIG, "w") as f:
        json.dump({"load": "../weights/missing/refinement_specific_weights", "input_size_train": [385, 385]}, f)
    with pytest.raises(FileNotFoundError, match="refinement_specific_weights"):
        refnet.refinement_net_init()
