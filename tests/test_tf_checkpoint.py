"""TensorFlow-free checkpoint reader (premvos_b200/tf_checkpoint.py).  The container format is restated from TensorFlow's
published layout and cannot be pinned against TensorFlow here (absent); pinned are the pieces with public known answers
(CRC-32C check value, snappy element encoding, LevelDB table magic) and the round trip against the writer."""
import struct

import numpy as np
import pytest

from premvos_b200 import synth, tf_checkpoint as T, weights


def test_known_answers():
    assert T.crc32c(b"123456789") == 0xE3069283                       # the CRC-32C check value
    assert T.crc32c(b"") == 0 and T.mask_crc(0) == 0xa282ead8
    assert T._varint(bytes([0xAC, 0x02]), 0) == (300, 2) and T._put_varint(300) == bytes([0xAC, 0x02])
    # snappy: length 11, literal "abcd" (tag (4-1)<<2), copy len 4 offset 4 (1-byte offset form), copy len 3 offset 8 (2-byte form)
    comp = bytes([11, (3 << 2) | 0]) + b"abcd" + bytes([((4 - 4) << 2) | 1 | (0 << 5), 4]) + bytes([((3 - 1) << 2) | 2]) + struct.pack("<H", 8)
    assert T.snappy_decompress(comp) == b"abcdabcdabc"
    # an overlapping copy (run-length): literal "x", copy len 5 offset 1
    assert T.snappy_decompress(bytes([6, 0]) + b"x" + bytes([((5 - 4) << 2) | 1, 1])) == b"xxxxxx"
    with pytest.raises(ValueError):
        T.snappy_decompress(bytes([3, 0]) + b"x")


def test_roundtrip_proposal_net_checkpoint(tmp_path):
    nb = (1, 1, 1, 1)
    P = synth.propnet_synthetic_params(4, nb)
    extra = {"global_step": np.array(12345, np.int64), "learning_rate": np.array(0.003, np.float32),
             "conv0/W/Momentum": np.zeros_like(P["conv0/W"]), "empty": np.zeros((0, 3), np.float32)}
    prefix = str(tmp_path / "ckpt" / "model-100")
    T.write_checkpoint(prefix, {**P, **extra}, block_size=512)        # small blocks: many data blocks, prefix compression
    listed = T.list_variables(prefix)
    assert set(listed) == set(P) | set(extra)
    assert listed["conv0/W"] == (np.dtype(np.float32), (7, 7, 3, 64)) and listed["global_step"] == (np.dtype(np.int64), ())
    got = T.read_checkpoint(prefix, verify_tensors=True)
    assert list(got) == sorted(got)                                    # table order
    for k, v in {**P, **extra}.items():
        np.testing.assert_array_equal(got[k], v)
        assert got[k].dtype == v.dtype and got[k].shape == v.shape
    sub = T.read_checkpoint(prefix, names=["conv0/W", "rpn/box/b"])
    assert list(sub) == ["conv0/W", "rpn/box/b"]
    with pytest.raises(KeyError):
        T.read_checkpoint(prefix, names=["nope"])
    # straight into the network's parameter dictionary (optimizer slots and bookkeeping dropped, shapes checked)
    sel = weights.select_proposal_net_variables(got, nb)
    assert list(sel) == list(synth.propnet_param_shapes(nb)) and all(np.array_equal(sel[k], P[k]) for k in P)
    sel2 = weights.load_proposal_net_variables(prefix, nb)             # a checkpoint prefix is accepted like a .npz
    assert all(np.array_equal(sel2[k], P[k]) for k in P)


def test_corruption_is_detected(tmp_path):
    prefix = str(tmp_path / "c")
    T.write_checkpoint(prefix, {"a": np.arange(6, dtype=np.float32).reshape(2, 3), "b/c": np.ones(4, np.int32)})
    raw = bytearray(open(prefix + ".index", "rb").read())
    raw[5] ^= 0xFF
    open(prefix + ".index", "wb").write(bytes(raw))
    with pytest.raises(ValueError):
        T.read_checkpoint(prefix)
    open(prefix + ".index", "wb").write(b"not a table")
    with pytest.raises(ValueError):
        T.read_checkpoint(prefix)
    T.write_checkpoint(prefix, {"a": np.arange(6, dtype=np.float32)})
    open(prefix + ".data-00000-of-00001", "wb").write(b"\x00" * 8)    # truncated shard
    with pytest.raises(ValueError):
        T.read_checkpoint(prefix)
