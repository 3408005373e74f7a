"""Readers for the two TF1 file formats the reference's GA3C-CADRL checkpoints come in -- without TensorFlow.

* ``<prefix>.meta``: a serialized ``MetaGraphDef`` protobuf.  Only its ``graph_def.node`` list is needed (name, op,
  inputs, attrs), decoded here straight from the protobuf wire format (varint / 64-bit / length-delimited / 32-bit).
* ``<prefix>.index`` + ``<prefix>.data-00000-of-00001``: a "tensor bundle".  The index is an SSTable in LevelDB's table
  format (prefix-compressed key/value blocks, a 48-byte footer with the metaindex / index block handles and the magic
  0xdb4775248b80fb57) whose values are ``BundleEntryProto`` messages (dtype, shape, offset, size); the data file holds the
  raw little-endian tensors.

The reference loads the same files with ``tf.train.import_meta_graph`` + ``Saver.restore``
(``policies/GA3C_CADRL/network.py:47-62``).
"""
import struct

import numpy as np

_NP = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64, 10: np.bool_}  # tensorflow DataType enum
_TABLE_MAGIC = 0xdb4775248b80fb57


def _varint(b, i):
    r = s = 0
    while True:
        c = b[i]
        i += 1
        r |= (c & 0x7f) << s
        s += 7
        if not c & 0x80:
            return r, i


def _signed(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def pb_fields(b):
    """[(field_number, wire_type, value)] of one protobuf message."""
    i, n, out = 0, len(b), []
    while i < n:
        key, i = _varint(b, i)
        f, w = key >> 3, key & 7
        if w == 0:
            v, i = _varint(b, i)
        elif w == 1:
            v, i = b[i:i + 8], i + 8
        elif w == 2:
            ln, i = _varint(b, i)
            v, i = b[i:i + ln], i + ln
        elif w == 5:
            v, i = b[i:i + 4], i + 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % w)
        out.append((f, w, v))
    return out


def _shape(b):  # TensorShapeProto: repeated Dim dim = 2 { int64 size = 1 }
    dims = []
    for f, _, v in pb_fields(b):
        if f == 2:
            size = 0
            for f2, _, v2 in pb_fields(v):
                if f2 == 1:
                    size = _signed(v2)
            dims.append(size)
    return dims


def _packed_varints(w, v):
    if w != 2:
        return [_signed(v)]
    out, i = [], 0
    while i < len(v):
        x, i = _varint(v, i)
        out.append(_signed(x))
    return out


def _tensor(b):  # TensorProto
    dt, shape, content, vals = None, [], None, []
    for f, w, v in pb_fields(b):
        if f == 1:
            dt = v
        elif f == 2:
            shape = _shape(v)
        elif f == 4:
            content = v
        elif f == 5:  # float_val
            vals += list(struct.unpack("<%df" % (len(v) // 4), v)) if w == 2 else [struct.unpack("<f", v)[0]]
        elif f in (7, 10):  # int_val, int64_val
            vals += _packed_varints(w, v)
        elif f == 11:  # bool_val
            vals += [bool(x) for x in v] if w == 2 else [bool(v)]
    if dt not in _NP:
        return None
    n = int(np.prod(shape)) if shape else 1
    if content:
        return np.frombuffer(content, dtype=_NP[dt]).reshape(shape).copy()
    if vals:
        a = np.array(vals, dtype=_NP[dt])
        return np.full(shape, a[0], dtype=_NP[dt]) if (a.size == 1 and n > 1) else a.reshape(shape)
    return np.zeros(shape, dtype=_NP[dt])


def _attr(b):  # AttrValue
    for f, _, v in pb_fields(b):
        if f == 2:
            return v
        if f == 3:
            return _signed(v)
        if f == 4:
            return struct.unpack("<f", v)[0]
        if f == 5:
            return bool(v)
        if f == 6:
            return ("type", v)
        if f == 7:
            return ("shape", _shape(v))
        if f == 8:
            return _tensor(v)
        if f == 1:
            return ("list", v)
    return None


def read_meta_graph(path):
    """{node name: {"op", "inputs", "attrs"}} of ``MetaGraphDef.graph_def`` plus the node order."""
    with open(path, "rb") as fh:
        b = fh.read()
    gd = [v for f, _, v in pb_fields(b) if f == 2]
    if not gd:
        raise ValueError("%s: no graph_def in the MetaGraphDef" % path)
    nodes, order = {}, []
    for f, _, nb in pb_fields(gd[0]):
        if f != 1:
            continue
        d = pb_fields(nb)
        name = [v for ff, _, v in d if ff == 1][0].decode()
        attrs = {}
        for ff, _, v in d:
            if ff == 5:
                kv = pb_fields(v)
                val = [x for k, _, x in kv if k == 2]
                attrs[[x for k, _, x in kv if k == 1][0].decode()] = _attr(val[0]) if val else None
        nodes[name] = {"op": [v for ff, _, v in d if ff == 2][0].decode(),
                       "inputs": [v.decode() for ff, _, v in d if ff == 3], "attrs": attrs}
        order.append(name)
    return nodes, order


def _handle(b, i):
    off, i = _varint(b, i)
    size, i = _varint(b, i)
    return off, size, i


def _table_block(buf, off, size):
    if buf[off + size] != 0:
        raise ValueError("compressed SSTable block (type %d) is not supported" % buf[off + size])
    blk = buf[off:off + size]
    num_restarts = struct.unpack("<I", blk[-4:])[0]
    end = len(blk) - 4 - 4 * num_restarts
    i, key, out = 0, b"", []
    while i < end:
        shared, i = _varint(blk, i)
        non_shared, i = _varint(blk, i)
        vlen, i = _varint(blk, i)
        key = key[:shared] + blk[i:i + non_shared]
        i += non_shared
        out.append((key, blk[i:i + vlen]))
        i += vlen
    return out


_CRC_TABLE = None


def crc32c(data):
    """CRC-32C (Castagnoli, reflected polynomial 0x82F63B78), table-driven."""
    global _CRC_TABLE
    if _CRC_TABLE is None:
        tab = []
        for n in range(256):
            c = n
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            tab.append(c)
        _CRC_TABLE = tab
    tab, c = _CRC_TABLE, 0xFFFFFFFF
    for byte in data:
        c = tab[(c ^ byte) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def _masked_crc(data):  # the bundle stores LevelDB's "masked" CRC of every tensor's bytes
    c = crc32c(data)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xa282ead8) & 0xFFFFFFFF


def read_checkpoint(prefix, verify=True):
    """{variable name: ndarray} of a single-shard TF1 tensor bundle ``<prefix>.index`` / ``.data-00000-of-00001``;
    ``verify`` checks every tensor's bytes against the CRC-32C the bundle index records for it."""
    with open(prefix + ".index", "rb") as fh:
        buf = fh.read()
    if len(buf) < 48 or struct.unpack("<Q", buf[-8:])[0] != _TABLE_MAGIC:
        raise ValueError("%s.index is not an SSTable" % prefix)
    foot = buf[-48:]
    _, _, i = _handle(foot, 0)  # metaindex handle
    ioff, isize, _ = _handle(foot, i)
    entries = {}
    for _, hv in _table_block(buf, ioff, isize):
        off, size, _ = _handle(hv, 0)
        for k, v in _table_block(buf, off, size):
            entries[k.decode()] = v
    with open(prefix + ".data-00000-of-00001", "rb") as fh:
        data = fh.read()
    out = {}
    for name, v in entries.items():
        if not name:  # BundleHeaderProto
            continue
        dt, shape, shard, off, size, crc = 1, [], 0, 0, 0, None
        for f, _, x in pb_fields(v):
            if f == 1:
                dt = x
            elif f == 2:
                shape = _shape(x)
            elif f == 3:
                shard = x
            elif f == 4:
                off = x
            elif f == 5:
                size = x
            elif f == 6:
                crc = struct.unpack("<I", x)[0]
        if shard != 0:
            raise ValueError("%s: multi-shard bundles are not supported" % prefix)
        if dt in _NP:
            raw = data[off:off + size]
            if len(raw) != size:
                raise ValueError("%s: tensor %s runs past the end of the data file" % (prefix, name))
            if verify and crc is not None and _masked_crc(raw) != crc:
                raise ValueError("%s: CRC mismatch for tensor %s" % (prefix, name))
            out[name] = np.frombuffer(raw, dtype=_NP[dt]).reshape(shape).copy()
    return out
