"""Reader (and a minimal writer) for TensorFlow V2 checkpoints ("tensor bundles") without TensorFlow -- SURVEY.md 8(f) N3.

The reference restores the proposal network (tensorpack SaverRestore, proposal_net/train.py:653-657) and the refinement
network (tf.train.Saver, refinement_net/core/Engine.py) from TensorFlow 1.x checkpoints: `<prefix>.index` +
`<prefix>.data-00000-of-0000N`.  TensorFlow is a third-party dependency that is absent from this image, so the file format is
RESTATED here from its published layout (tensorflow/core/util/tensor_bundle/tensor_bundle.{h,cc}, tensorflow/core/lib/io/
table_format.txt, tensor_bundle.proto) and is UNPINNED: there is no TensorFlow and no real checkpoint in the build container
to check it against; tests/test_tf_checkpoint.py round-trips it against the writer below, which follows the same description.

Layout:
  * `.index` is a LevelDB-style sorted string table: data blocks, a metaindex block, an index block and a 48-byte footer
    (two block handles as varint64 pairs, zero padding to 40 bytes, magic 0xdb4775248b80fb57 little-endian).  A block is
    followed by a 1-byte compression type (0 none, 1 snappy) and a masked CRC32C of contents + type.  Block contents are
    prefix-compressed entries (shared, non_shared, value_len as varint32, key suffix, value) followed by the restart array
    (uint32 offsets) and its length.
  * key "" holds a BundleHeaderProto (num_shards, endianness, version); every other key is a tensor name whose value is a
    BundleEntryProto: dtype (1), shape (2: TensorShapeProto with repeated dim {size}), shard_id (3), offset (4), size (5),
    crc32c (6, fixed32), slices (7, partitioned variables -- not supported here).
  * the tensor bytes (little-endian, C order) sit at [offset, offset + size) of shard file `.data-%05d-of-%05d`.
"""
from __future__ import annotations

import os
import struct
from collections import OrderedDict

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
# tensorflow/core/framework/types.proto
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
           17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
_DTYPE_IDS = {np.dtype(v): k for k, v in _DTYPES.items()}


# ---- primitives ----------------------------------------------------------------------------------------------------------
def _varint(buf, pos):
    out = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7
        if shift > 70:
            raise ValueError("malformed varint")


def _put_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


_CRC_TABLE = None


def crc32c(data, crc=0):
    """CRC-32C (Castagnoli, polynomial 0x1EDC6F41 reflected = 0x82F63B78)."""
    global _CRC_TABLE
    if _CRC_TABLE is None:
        t = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            t.append(c)
        _CRC_TABLE = t
    c = crc ^ 0xFFFFFFFF
    for b in data:
        c = _CRC_TABLE[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def mask_crc(c):
    return (((c >> 15) | (c << 17)) + 0xa282ead8) & 0xFFFFFFFF


def snappy_decompress(data):
    """Raw snappy block format: varint uncompressed length, then literal / copy elements."""
    n, pos = _varint(data, 0)
    out = bytearray()
    while pos < len(data):
        tag = data[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:   # literal
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(data[pos:pos + nb], "little")
                pos += nb
            ln += 1
            out += data[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | data[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = int.from_bytes(data[pos:pos + 2], "little")
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(data[pos:pos + 4], "little")
            pos += 4
        if off == 0 or off > len(out):
            raise ValueError("malformed snappy copy")
        for _ in range(ln):   # copies may overlap their own output
            out.append(out[-off])
    if len(out) != n:
        raise ValueError("snappy length mismatch: %d != %d" % (len(out), n))
    return bytes(out)


# ---- sorted string table -------------------------------------------------------------------------------------------------
def _read_block(buf, offset, size, verify):
    contents, ctype = buf[offset:offset + size], buf[offset + size]
    if verify:
        stored = struct.unpack_from("<I", buf, offset + size + 1)[0]
        if mask_crc(crc32c(buf[offset:offset + size + 1])) != stored:
            raise ValueError("table block at %d: checksum mismatch" % offset)
    if ctype == 1:
        contents = snappy_decompress(contents)
    elif ctype != 0:
        raise ValueError("table block at %d: unknown compression type %d" % (offset, ctype))
    return contents


def _block_entries(block):
    num_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * num_restarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def read_table(path, verify=True):
    """-> list of (key bytes, value bytes) of a LevelDB-format table file, in key order."""
    with open(path, "rb") as f:
        buf = f.read()
    if len(buf) < 48 or struct.unpack_from("<Q", buf, len(buf) - 8)[0] != TABLE_MAGIC:
        raise ValueError("%s is not a TensorFlow checkpoint index (bad table magic)" % path)
    footer = buf[len(buf) - 48:]
    pos = 0
    _, pos = _varint(footer, pos)
    _, pos = _varint(footer, pos)          # metaindex handle (unused)
    ioff, pos = _varint(footer, pos)
    isize, pos = _varint(footer, pos)
    out = []
    for _, handle in _block_entries(_read_block(buf, ioff, isize, verify)):
        off, p = _varint(handle, 0)
        size, _ = _varint(handle, p)
        out.extend(_block_entries(_read_block(buf, off, size, verify)))
    return out


# ---- protobuf (just the fields named in the module docstring) ------------------------------------------------------------
def _proto_fields(buf):
    pos = 0
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        field, wire = tag >> 3, tag & 7
        if wire == 0:
            v, pos = _varint(buf, pos)
        elif wire == 1:
            v, pos = buf[pos:pos + 8], pos + 8
        elif wire == 2:
            ln, pos = _varint(buf, pos)
            v, pos = buf[pos:pos + ln], pos + ln
        elif wire == 5:
            v, pos = struct.unpack_from("<I", buf, pos)[0], pos + 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wire)
        yield field, wire, v


def _parse_entry(buf):
    e = {"dtype": 0, "shape": [], "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "slices": 0}
    for field, _, v in _proto_fields(buf):
        if field == 1:
            e["dtype"] = v
        elif field == 2:
            for f2, _, v2 in _proto_fields(v):
                if f2 == 2:
                    size = 0
                    for f3, _, v3 in _proto_fields(v2):
                        if f3 == 1:
                            size = v3 - (1 << 64) if v3 >= (1 << 63) else v3
                    e["shape"].append(size)
                elif f2 == 3 and v2:
                    raise ValueError("tensor of unknown rank")
        elif field == 3:
            e["shard_id"] = v
        elif field == 4:
            e["offset"] = v
        elif field == 5:
            e["size"] = v
        elif field == 6:
            e["crc32c"] = v
        elif field == 7:
            e["slices"] += 1
    return e


def list_variables(prefix, verify=True):
    """-> OrderedDict name -> (numpy dtype, shape tuple) of a checkpoint `prefix` (as tf.train.list_variables)."""
    out = OrderedDict()
    for key, value in read_table(prefix + ".index", verify):
        if key == b"":
            continue
        e = _parse_entry(value)
        if e["dtype"] not in _DTYPES:
            continue   # strings / resources carry no weights
        out[key.decode("utf-8")] = (np.dtype(_DTYPES[e["dtype"]]), tuple(e["shape"]))
    return out


def read_checkpoint(prefix, names=None, verify=True, verify_tensors=False):
    """-> OrderedDict variable name -> ndarray for the checkpoint `prefix` (`prefix.index`, `prefix.data-*`).
    names: optional subset to load.  verify: check the index blocks' CRC32C; verify_tensors: also every tensor's (slow in
    pure Python).  Partitioned variables (slices) are not supported."""
    entries = read_table(prefix + ".index", verify)
    num_shards = 1
    for key, value in entries:
        if key == b"":
            for field, _, v in _proto_fields(value):
                if field == 1:
                    num_shards = v
                elif field == 2 and v != 0:
                    raise ValueError("big-endian checkpoints are not supported")
    shards = {}
    out = OrderedDict()
    try:
        for key, value in entries:
            if key == b"":
                continue
            name = key.decode("utf-8")
            if names is not None and name not in names:
                continue
            e = _parse_entry(value)
            if e["dtype"] not in _DTYPES:
                continue
            if e["slices"]:
                raise ValueError("%s is a partitioned variable (slices are not supported)" % name)
            dt = np.dtype(_DTYPES[e["dtype"]])
            count = int(np.prod(e["shape"], dtype=np.int64)) if e["shape"] else 1
            if count * dt.itemsize != e["size"]:
                raise ValueError("%s: %d bytes stored, shape %s of %s needs %d" % (name, e["size"], e["shape"], dt, count * dt.itemsize))
            sid = e["shard_id"]
            if sid not in shards:
                shards[sid] = open("%s.data-%05d-of-%05d" % (prefix, sid, num_shards), "rb")
            f = shards[sid]
            f.seek(e["offset"])
            raw = f.read(e["size"])
            if len(raw) != e["size"]:
                raise ValueError("%s: data shard %d is truncated" % (name, sid))
            if verify_tensors and e["crc32c"] is not None and mask_crc(crc32c(raw)) != e["crc32c"]:
                raise ValueError("%s: tensor checksum mismatch" % name)
            out[name] = np.frombuffer(raw, dtype=dt.newbyteorder("<")).astype(dt).reshape(e["shape"]).copy()
    finally:
        for f in shards.values():
            f.close()
    if names is not None:
        missing = [n for n in names if n not in out]
        if missing:
            raise KeyError("checkpoint %s lacks %d variable(s), e.g. %s" % (prefix, len(missing), missing[:5]))
    return out


# ---- minimal writer (tests, conversions): one shard, no compression ------------------------------------------------------
def _build_block(entries, restart_interval=16):
    out, restarts, last = bytearray(), [], b""
    for i, (key, value) in enumerate(entries):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(key), len(last)) and key[shared] == last[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value)) + key[shared:] + value
        last = key
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def _field(field, wire, payload):
    return _put_varint((field << 3) | wire) + payload


def write_checkpoint(prefix, variables, block_size=4096):
    """Writes `variables` ({name: ndarray}) as a one-shard tensor bundle."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    items = sorted(((k.encode("utf-8"), np.asarray(v)) for k, v in variables.items()), key=lambda kv: kv[0])   # 0-d stays 0-d
    header = _field(1, 0, _put_varint(1)) + _field(2, 0, _put_varint(0)) + _field(3, 2, _put_varint(2) + _field(1, 0, _put_varint(1)))
    kv = [(b"", header)]
    offset = 0
    with open("%s.data-00000-of-00001" % prefix, "wb") as f:
        for key, arr in items:
            if arr.dtype not in _DTYPE_IDS:
                raise TypeError("%s: dtype %s cannot be stored" % (key.decode(), arr.dtype))
            raw = arr.astype(arr.dtype.newbyteorder("<")).tobytes(order="C")
            f.write(raw)
            shape = b"".join(_field(2, 2, (lambda d: _put_varint(len(d)) + d)(_field(1, 0, _put_varint(int(s))))) for s in arr.shape)
            entry = _field(1, 0, _put_varint(_DTYPE_IDS[arr.dtype])) + _field(2, 2, _put_varint(len(shape)) + shape)
            entry += _field(4, 0, _put_varint(offset)) + _field(5, 0, _put_varint(len(raw)))
            if len(raw) <= (1 << 16):   # the pure-Python CRC is slow: big tensors go without (the reader then skips the check)
                entry += _field(6, 5, struct.pack("<I", mask_crc(crc32c(raw))))
            kv.append((key, entry))
            offset += len(raw)
    out = bytearray()
    index = []

    def emit(block):
        off = len(out)
        out.extend(block)
        out.append(0)
        out.extend(struct.pack("<I", mask_crc(crc32c(block + b"\x00"))))
        return _put_varint(off) + _put_varint(len(block))

    cur, cur_bytes = [], 0
    for key, value in kv:
        cur.append((key, value))
        cur_bytes += len(key) + len(value) + 3
        if cur_bytes >= block_size:
            index.append((cur[-1][0], emit(_build_block(cur))))
            cur, cur_bytes = [], 0
    if cur:
        index.append((cur[-1][0], emit(_build_block(cur))))
    meta = emit(_build_block([]))
    idx = emit(_build_block(index, restart_interval=1))
    footer = meta + idx
    out.extend(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC))
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(out))
