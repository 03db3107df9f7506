"""Reader (and a minimal writer) for TensorFlow V2 checkpoint bundles, without TensorFlow.

The reference restores `tf.train.latest_checkpoint(<checkpoint_dir>/<runname>)` with `tf.train.Saver().restore`
(sga.py:180-182, tf_boilerplate.py:47-48, 84-86): a text file `checkpoint` naming the prefix, `<prefix>.index`
(an SSTable of tensor name -> BundleEntryProto) and `<prefix>.data-00000-of-00001` (raw little-endian tensor bytes).

Formats restated from the TensorFlow sources (not vendored; **unpinned** against a real TF-written file in this
container -- PARITY_ASSUMPTIONS.md P13): tensorflow/core/lib/io/{table,block,format}.cc (LevelDB table: prefix-compressed
blocks with restart arrays, 1-byte compression tag + masked crc32c trailer, 48-byte footer with magic
0xdb4775248b80fb57), tensorflow/core/protobuf/tensor_bundle.proto (BundleHeaderProto / BundleEntryProto),
tensorflow/core/framework/{tensor_shape,types}.proto.  Only what Saver writes for dense variables is supported:
uncompressed blocks, one shard or several, no tensor slices.
"""
import os
import re
import struct

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
FOOTER_LEN = 48
# tensorflow/core/framework/types.proto
DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
          17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
DTYPE_IDS = {np.dtype(v): k for k, v in DTYPES.items()}


# ---- crc32c (Castagnoli), masked as in tensorflow/core/lib/hash/crc32c.h ------------------------------------------
def _crc_table():
    t = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        t.append(c)
    return t


_CRC = _crc_table()


def crc32c(data, crc=0):
    crc ^= 0xFFFFFFFF
    for b in bytes(data):
        crc = _CRC[(crc ^ b) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def mask_crc(crc):
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ---- varints / protobuf wire format --------------------------------------------------------------------------------
def _get_varint(buf, pos):
    shift = result = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


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


def _parse_message(buf):
    """-> {field number: [values]}; varints as ints, fixed32/64 as ints, length-delimited as bytes."""
    fields, pos = {}, 0
    while pos < len(buf):
        key, pos = _get_varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _get_varint(buf, pos)
        elif wt == 1:
            val = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            val = bytes(buf[pos:pos + n])
            pos += n
        elif wt == 5:
            val = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        fields.setdefault(num, []).append(val)
    return fields


def _signed64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def _parse_shape(buf):
    dims = []
    for d in _parse_message(buf).get(2, []):                 # TensorShapeProto.dim
        dims.append(_signed64(_parse_message(d).get(1, [0])[0]))
    return tuple(dims)


def _parse_entry(buf):
    f = _parse_message(buf)                                   # BundleEntryProto
    if 7 in f:
        raise NotImplementedError("sliced (partitioned) variables are not supported")
    return dict(dtype=f.get(1, [0])[0], shape=_parse_shape(f[2][0]) if 2 in f else (), shard_id=f.get(3, [0])[0],
                offset=f.get(4, [0])[0], size=f.get(5, [0])[0], crc32c=f.get(6, [0])[0])


# ---- table (SSTable) -------------------------------------------------------------------------------------------------
def _read_block(data, offset, size, verify):
    body = data[offset:offset + size]
    tag = data[offset + size]
    if verify:
        stored = struct.unpack_from("<I", data, offset + size + 1)[0]
        if mask_crc(crc32c(data[offset:offset + size + 1])) != stored:
            raise ValueError("checkpoint index: block checksum mismatch at offset %d" % offset)
    if tag != 0:
        raise NotImplementedError("compressed table blocks (tag %d) are not supported; Saver writes them uncompressed" % tag)
    return body


def _block_entries(block):
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _get_varint(block, pos)
        unshared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + unshared])
        pos += unshared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def read_index(path, verify=True):
    """-> (header fields, {tensor name: entry dict})."""
    with open(path, "rb") as fh:
        data = fh.read()
    if len(data) < FOOTER_LEN or struct.unpack_from("<Q", data, len(data) - 8)[0] != TABLE_MAGIC:
        raise ValueError("%s is not a TensorFlow checkpoint index (bad table magic)" % path)
    footer = data[len(data) - FOOTER_LEN:]
    pos = 0
    _, pos = _get_varint(footer, pos)            # metaindex handle
    _, pos = _get_varint(footer, pos)
    idx_off, pos = _get_varint(footer, pos)
    idx_size, pos = _get_varint(footer, pos)
    entries, header = {}, None
    for _, handle in _block_entries(_read_block(data, idx_off, idx_size, verify)):
        off, p = _get_varint(handle, 0)
        size, _ = _get_varint(handle, p)
        for key, val in _block_entries(_read_block(data, off, size, verify)):
            if key == b"":
                header = _parse_message(val)      # BundleHeaderProto: 1 num_shards, 2 endianness, 3 version
            else:
                entries[key.decode("utf-8")] = _parse_entry(val)
    if header is None:
        raise ValueError("%s: bundle header entry missing" % path)
    if header.get(2, [0])[0] != 0:
        raise NotImplementedError("big-endian bundles are not supported")
    return header, entries


def read_bundle(prefix, names=None, verify=False):
    """{tensor name: ndarray} of the bundle `<prefix>.index` / `<prefix>.data-*`.  names: optional filter (callable)."""
    header, entries = read_index(prefix + ".index")
    num_shards = header.get(1, [1])[0] or 1
    shards = {}
    out = {}
    for name, e in entries.items():
        if names is not None and not names(name):
            continue
        if e["dtype"] not in DTYPES:
            continue                               # strings / resources: nothing a model variable uses
        sid = e["shard_id"]
        if sid not in shards:
            shards[sid] = np.memmap("%s.data-%05d-of-%05d" % (prefix, sid, num_shards), dtype=np.uint8, mode="r")
        raw = shards[sid][e["offset"]:e["offset"] + e["size"]]
        if verify and mask_crc(crc32c(raw.tobytes())) != e["crc32c"]:
            raise ValueError("checkpoint data: checksum mismatch for %s" % name)
        out[name] = np.frombuffer(raw.tobytes(), dtype=DTYPES[e["dtype"]]).reshape(e["shape"]).copy()
    return out


def latest_checkpoint(directory):
    """tf.train.latest_checkpoint: the prefix named by the `checkpoint` state file, or None."""
    state = os.path.join(directory, "checkpoint")
    if not os.path.exists(state):
        return None
    m = re.search(r'^model_checkpoint_path:\s*"([^"]+)"', open(state).read(), re.M)
    if not m:
        return None
    p = m.group(1)
    return p if os.path.isabs(p) else os.path.join(directory, p)


# ---- writer (tests, and exporting synthetic weights in the reference's on-disk format) -----------------------------
def _msg_field(num, wt, payload):
    return _put_varint((num << 3) | wt) + payload


def _shape_proto(shape):
    out = b""
    for d in shape:
        dim = _msg_field(1, 0, _put_varint(int(d)))
        out += _msg_field(2, 2, _put_varint(len(dim)) + dim)
    return out


def _build_block(items, restart_interval):
    buf, restarts, last = bytearray(), [], b""
    for i, (key, val) in enumerate(items):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(buf))
        else:
            while shared < min(len(key), len(last)) and key[shared] == last[shared]:
                shared += 1
        buf += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(val)) + key[shared:] + val
        last = key
    if not restarts:
        restarts = [0]
    for r in restarts:
        buf += struct.pack("<I", r)
    buf += struct.pack("<I", len(restarts))
    return bytes(buf)


def write_bundle(prefix, tensors, restart_interval=16, block_entries=64):
    """Writes `<prefix>.index` and `<prefix>.data-00000-of-00001` (one shard, uncompressed) and the `checkpoint` file."""
    names = sorted(tensors)
    data = bytearray()
    items = [(b"", _msg_field(1, 0, _put_varint(1)) + _msg_field(3, 2, _put_varint(2) + _msg_field(1, 0, _put_varint(1))))]
    for name in names:
        a = np.asarray(tensors[name], order="C")
        raw = a.tobytes()
        shape = _shape_proto(a.shape)
        entry = (_msg_field(1, 0, _put_varint(DTYPE_IDS[a.dtype])) + _msg_field(2, 2, _put_varint(len(shape)) + shape) +
                 _msg_field(4, 0, _put_varint(len(data))) + _msg_field(5, 0, _put_varint(len(raw))) +
                 _msg_field(6, 5, struct.pack("<I", mask_crc(crc32c(raw)))))
        items.append((name.encode("utf-8"), entry))
        data += raw
    out = bytearray()

    def emit(block):
        off = len(out)
        trailer = b"\x00"
        out.extend(block + trailer + struct.pack("<I", mask_crc(crc32c(block + trailer))))
        return _put_varint(off) + _put_varint(len(block))

    index_items = []
    for i in range(0, len(items), block_entries):
        chunk = items[i:i + block_entries]
        index_items.append((chunk[-1][0], emit(_build_block(chunk, restart_interval))))
    meta = emit(_build_block([], 1))
    index = emit(_build_block(index_items, 1))
    footer = meta + index
    out.extend(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC))
    with open(prefix + ".index", "wb") as fh:
        fh.write(out)
    with open(prefix + ".data-00000-of-00001", "wb") as fh:
        fh.write(data)
    with open(os.path.join(os.path.dirname(prefix), "checkpoint"), "w") as fh:
        fh.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % ((os.path.basename(prefix),) * 2))
