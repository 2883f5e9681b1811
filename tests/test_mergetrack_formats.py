"""Host-side on-disk formats of stage 7 (premvos_b200/mergetrack.py): VOC palette PNGs, first-frame annotations, proposal JSON.
The palette is pinned by the SHA-1 of `(np.array(pascal_colormap) * 255).round()` computed from the reference's own table
(MergeTrack/merge_functions.py:250-506) in the build container."""
import hashlib
import json

import numpy as np
import pytest

from oracle import mergetrack_oracle as MO
from oracle import refnet_oracle as RO
from premvos_b200 import mergetrack, synth


def test_pascal_colormap_is_the_reference_table():
    cm = mergetrack.pascal_colormap()
    assert cm.shape == (256, 3) and cm.dtype == np.uint8
    assert hashlib.sha1(cm.tobytes()).hexdigest() == "cd59c439a2ded6190058462a26428a894d711ea4"
    assert cm[:4].tolist() == [[0, 0, 0], [128, 0, 0], [0, 128, 0], [128, 128, 0]] and cm[255].tolist() == [224, 224, 192]


def test_save_pngs_and_read_ann_roundtrip(tmp_path):
    Image = pytest.importorskip("PIL.Image")
    masks = synth.synthetic_masks(3, 40, 60, seed=2)[:2]
    props = [{"mask": masks[0], "id": 1}, {"mask": masks[1], "id": 5}]
    fn = str(tmp_path / "seq" / "00001.png")
    mergetrack.save_pngs(props, fn)
    im = Image.open(fn)
    assert im.mode == "P"
    ids = np.array(im)
    want = np.zeros((40, 60), np.uint8)
    want[masks[0] == 1] = 1
    want[masks[1] == 1] = 5            # later proposals win where masks overlap
    np.testing.assert_array_equal(ids, want)
    np.testing.assert_array_equal(np.array(im.getpalette()[:18]).reshape(6, 3), mergetrack.pascal_colormap()[:6])
    mergetrack.save_pngs(props, str(tmp_path / "empty.png"), empty=True)
    assert np.array(Image.open(str(tmp_path / "empty.png"))).sum() == 0
    # read_ann: one template per id, tight box, COCO RLE that decodes back to the mask
    anns = mergetrack.read_ann(fn)
    assert [a["id"] for a in anns] == [1, 5]
    for a in anns:
        m = (want == a["id"]).astype(np.uint8)
        np.testing.assert_array_equal(a["bbox"], MO.to_bbox(m))
        np.testing.assert_array_equal(RO.rle_decode(a["segmentation"]), m)
        assert a["segmentation"] == RO.rle_encode(m) and a["score"] == 1.0 and a["conf_score"] == "1.0"


def test_read_props(tmp_path):
    fn = tmp_path / "p.json"
    fn.write_text(json.dumps([{"bbox": [1, 2, 3, 4], "score": 0.5}, {"bbox": [0, 0, 1, 1], "score": 0.1, "ReID": [0.0] * 128}]))
    props = mergetrack.read_props(str(fn))
    assert len(props) == 2 and np.isinf(props[0]["ReID"]).all() and len(props[0]["ReID"]) == 128 and props[1]["ReID"] == [0.0] * 128
    assert mergetrack.read_props(str(tmp_path / "missing.json")) == []


def test_get_flow_bad_magic_prints_and_returns_none(tmp_path, capsys):
    # merge_functions.py:197-207: no exception, the reference prints and falls through (returns None)
    from premvos_b200 import mergetrack
    p = tmp_path / "bad.flo"
    p.write_bytes(np.array([1.0], np.float32).tobytes() + np.array([2, 2], np.int32).tobytes() + np.zeros(8, np.float32).tobytes())
    assert mergetrack.get_flow(str(p)) is None
    assert "Magic number incorrect" in capsys.readouterr().out
