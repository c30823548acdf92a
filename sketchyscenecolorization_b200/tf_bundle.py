"""TensorFlow "V2" checkpoint files (tensor bundle) without TensorFlow: writer and reader.

The reference snapshots with `tf.train.Saver(max_to_keep=100).save(sess, ckpt_dir/'model_{i}.ckpt', global_step=i)`
(main_procedure.py:141,235-237) and restores with `Saver().restore(sess, tf.train.latest_checkpoint(ckpt_dir))`
(:163-165,292-301,550,559).  With TF >= 1.x defaults that produces, per snapshot,

    model_<i>.ckpt-<i>.index                 an SSTable (LevelDB table format, tensorflow/core/lib/io/table*): key "" ->
                                             BundleHeaderProto, key <variable name> -> BundleEntryProto
    model_<i>.ckpt-<i>.data-00000-of-00001   the tensors' bytes, little endian, concatenated in key order
    checkpoint                               text CheckpointState (written by checkpoint.py)

This module restates that byte format (tensorflow/core/util/tensor_bundle/tensor_bundle.{h,cc},
tensorflow/core/protobuf/tensor_bundle.proto, tensorflow/core/lib/io/{format,block_builder,table_builder}.cc):

  table  = data blocks, metaindex block (empty), index block, 48-byte footer
  block  = entries [shared:varint32][non_shared:varint32][value_len:varint32][key suffix][value], restart offsets
           (uint32 LE each, one every 16 entries; every entry in an index block), restart count (uint32), then the block
           trailer: 1 byte compression type (0 = none, what BundleWriter uses) + masked CRC-32C of block + type
  index  = one entry per data block: key = a separator >= the block's last key, value = BlockHandle(offset, size) varint64s
  footer = metaindex handle + index handle, zero padded to 40 bytes, + magic 0xdb4775248b80fb57 (LE)
  masked crc = rotr15(crc32c) + 0xa282ead8  (crc32c::Mask)

TensorFlow is not available in this image and the published checkpoints are not in the reference repository, so the
format is pinned by what IS available (tests/test_tf_bundle_cpu.py): CRC-32C known answers (RFC 3720), the masked CRC
against tensorboard's independent implementation, the sub-messages against tensorboard's generated protobuf classes
(TensorShapeProto, VersionDef, DataType enum), and write -> read round trips including multi-block tables and
prefix-compressed keys.  Reading snappy-compressed blocks is not supported (BundleWriter never writes them).
"""
from __future__ import annotations

import ctypes
import struct

import numpy as np

MAGIC = 0xDB4775248B80FB57
BLOCK_SIZE = 262144            # table::Options().block_size in TensorFlow
RESTART_INTERVAL = 16
MASK_DELTA = 0xA282EAD8

# tensorflow/core/framework/types.proto
DT_FLOAT, DT_DOUBLE, DT_INT32, DT_UINT8, DT_INT64, DT_BOOL, DT_BFLOAT16, DT_HALF = 1, 2, 3, 4, 9, 10, 14, 19
_NP2DT = {np.dtype("<f4"): DT_FLOAT, np.dtype("<f8"): DT_DOUBLE, np.dtype("<i4"): DT_INT32, np.dtype("u1"): DT_UINT8,
          np.dtype("<i8"): DT_INT64, np.dtype("bool"): DT_BOOL, np.dtype("<f2"): DT_HALF}
_DT2NP = {v: k for k, v in _NP2DT.items()}


# ----------------------------------------------------------------------------------------------------------
# CRC-32C
# ----------------------------------------------------------------------------------------------------------
def crc32c(data, crc=0):
    """CRC-32C (Castagnoli) of a bytes-like / contiguous numpy array (libfgcolor's host routine, slicing-by-8)."""
    from . import _lib
    lib = _lib.load()
    if isinstance(data, np.ndarray):
        a = np.ascontiguousarray(data)
        return int(lib.fgc_crc32c(ctypes.c_void_p(a.ctypes.data), a.nbytes, crc)) & 0xFFFFFFFF
    b = bytes(data)
    return int(lib.fgc_crc32c(ctypes.c_char_p(b), len(b), crc)) & 0xFFFFFFFF


def mask_crc(c):
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + MASK_DELTA) & 0xFFFFFFFF


def unmask_crc(m):
    r = (m - MASK_DELTA) & 0xFFFFFFFF
    return ((r >> 17) | (r << 15)) & 0xFFFFFFFF


# ----------------------------------------------------------------------------------------------------------
# protobuf wire format (only what the two bundle messages need)
# ----------------------------------------------------------------------------------------------------------
def _varint(n):
    n &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _read_varint(buf, pos):
    shift = result = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _field_varint(num, val):
    return _varint(num << 3) + _varint(val)


def _field_bytes(num, payload):
    return _varint((num << 3) | 2) + _varint(len(payload)) + payload


def _field_fixed32(num, val):
    return _varint((num << 3) | 5) + struct.pack("<I", val)


def _parse(buf):
    """-> list of (field number, wire type, value); value is int (varint / fixed) or bytes (length delimited)."""
    out, pos = [], 0
    while pos < len(buf):
        key, pos = _read_varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _read_varint(buf, pos)
        elif wt == 2:
            ln, pos = _read_varint(buf, pos)
            v = bytes(buf[pos:pos + ln])
            pos += ln
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        out.append((num, wt, v))
    return out


def encode_shape(shape):
    """TensorShapeProto: repeated Dim dim = 2 { int64 size = 1 }.  (A scalar is the empty message.)"""
    return b"".join(_field_bytes(2, _field_varint(1, int(d)) if int(d) else b"") for d in shape)   # proto3 omits size = 0


def decode_shape(buf):
    dims = []
    for num, _, v in _parse(buf):
        if num == 2:
            size = 0
            for n2, _, v2 in _parse(v):
                if n2 == 1:
                    size = v2 if v2 < (1 << 63) else v2 - (1 << 64)
            dims.append(size)
    return tuple(dims)


def encode_header(num_shards=1):
    """BundleHeaderProto { int32 num_shards = 1; Endianness endianness = 2 (LITTLE = 0, default: omitted);
    VersionDef version = 3 { int32 producer = 1 } } -- kTensorBundleVersion = 1."""
    return _field_varint(1, num_shards) + _field_bytes(3, _field_varint(1, 1))


def encode_entry(dtype, shape, offset, size, crc_masked, shard_id=0):
    """BundleEntryProto { DataType dtype = 1; TensorShapeProto shape = 2; int32 shard_id = 3; int64 offset = 4;
    int64 size = 5; fixed32 crc32c = 6; repeated TensorSliceProto slices = 7 }  (proto3: zero fields are omitted)."""
    out = _field_varint(1, dtype) + _field_bytes(2, encode_shape(shape))
    if shard_id:
        out += _field_varint(3, shard_id)
    if offset:
        out += _field_varint(4, offset)
    if size:
        out += _field_varint(5, size)
    out += _field_fixed32(6, crc_masked)
    return out


def decode_entry(buf):
    e = dict(dtype=0, shape=(), shard_id=0, offset=0, size=0, crc32c=0, slices=0)
    for num, _, v in _parse(buf):
        if num == 1:
            e["dtype"] = v
        elif num == 2:
            e["shape"] = decode_shape(v)
        elif num == 3:
            e["shard_id"] = v
        elif num == 4:
            e["offset"] = v
        elif num == 5:
            e["size"] = v
        elif num == 6:
            e["crc32c"] = v
        elif num == 7:
            e["slices"] += 1
    return e


# ----------------------------------------------------------------------------------------------------------
# table (SSTable) writer / reader
# ----------------------------------------------------------------------------------------------------------
class _BlockBuilder:
    def __init__(self, restart_interval):
        self.ri = restart_interval
        self.buf = bytearray()
        self.restarts = [0]
        self.counter = 0
        self.last_key = b""

    def add(self, key, value):
        shared = 0
        if self.counter < self.ri:
            m = min(len(self.last_key), len(key))
            while shared < m and self.last_key[shared] == key[shared]:
                shared += 1
        else:
            self.restarts.append(len(self.buf))
            self.counter = 0
        self.buf += _varint(shared) + _varint(len(key) - shared) + _varint(len(value)) + key[shared:] + value
        self.last_key = key
        self.counter += 1

    def size_estimate(self):
        return len(self.buf) + 4 * len(self.restarts) + 4

    def empty(self):
        return not self.buf

    def finish(self):
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))


def _shortest_separator(start, limit):
    """leveldb BytewiseComparator::FindShortestSeparator: a short key in [start, limit)."""
    m = min(len(start), len(limit))
    d = 0
    while d < m and start[d] == limit[d]:
        d += 1
    if d < m and start[d] < 0xFF and start[d] + 1 < limit[d]:
        return start[:d] + bytes([start[d] + 1])
    return start


def _short_successor(key):
    """leveldb BytewiseComparator::FindShortSuccessor: a short key >= key."""
    for i, b in enumerate(key):
        if b != 0xFF:
            return key[:i] + bytes([b + 1])
    return key


def write_table(path, items, block_size=BLOCK_SIZE, restart_interval=RESTART_INTERVAL):
    """items: iterable of (key bytes, value bytes) in strictly increasing key order."""
    out = bytearray()
    index = _BlockBuilder(1)
    data = _BlockBuilder(restart_interval)
    pending = None           # (last key of the finished block, handle) waiting for the next key to pick a separator

    def emit(block_bytes):
        off = len(out)
        out.extend(block_bytes)
        trailer_type = b"\x00"                                   # kNoCompression
        out.extend(trailer_type + struct.pack("<I", mask_crc(crc32c(block_bytes + trailer_type))))
        return _varint(off) + _varint(len(block_bytes))

    last = None
    for key, value in items:
        if last is not None and not key > last:
            raise ValueError("table keys must be strictly increasing: %r after %r" % (key, last))
        if pending is not None:
            index.add(_shortest_separator(pending[0], key), pending[1])
            pending = None
        data.add(key, value)
        last = key
        if data.size_estimate() >= block_size:
            pending = (last, emit(data.finish()))
            data = _BlockBuilder(restart_interval)
    if not data.empty():
        pending = (last, emit(data.finish()))
    if pending is not None:
        index.add(_short_successor(pending[0]), pending[1])
    meta_handle = emit(_BlockBuilder(restart_interval).finish())   # empty metaindex block
    index_handle = emit(index.finish())
    footer = meta_handle + index_handle
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", MAGIC)
    out.extend(footer)
    with open(path, "wb") as f:
        f.write(out)


def _read_block(buf, handle_off, handle_size, verify=True):
    block = bytes(buf[handle_off:handle_off + handle_size])
    ctype = buf[handle_off + handle_size]
    crc = struct.unpack_from("<I", buf, handle_off + handle_size + 1)[0]
    if verify and unmask_crc(crc) != crc32c(block + bytes([ctype])):
        raise ValueError("table block checksum mismatch at offset %d" % handle_off)
    if ctype != 0:
        raise NotImplementedError("compressed table block (type %d): only uncompressed tables are supported" % ctype)
    nrestarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * nrestarts
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = _read_varint(block, pos)
        non_shared, pos = _read_varint(block, pos)
        vlen, pos = _read_varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        out.append((key, block[pos:pos + vlen]))
        pos += vlen
    return out


def read_table(path, verify=True):
    """-> list of (key, value) in file order."""
    buf = open(path, "rb").read()
    if len(buf) < 48 or struct.unpack_from("<Q", buf, len(buf) - 8)[0] != MAGIC:
        raise ValueError("%s is not a TensorFlow table file (bad magic)" % path)
    footer = buf[len(buf) - 48:]
    pos = 0
    _, pos = _read_varint(footer, pos)
    _, pos = _read_varint(footer, pos)
    ioff, pos = _read_varint(footer, pos)
    isize, pos = _read_varint(footer, pos)
    out = []
    for _, handle in _read_block(buf, ioff, isize, verify):
        off, p2 = _read_varint(handle, 0)
        size, _ = _read_varint(handle, p2)
        out.extend(_read_block(buf, off, size, verify))
    return out


# ----------------------------------------------------------------------------------------------------------
# bundle = index table + data file
# ----------------------------------------------------------------------------------------------------------
def write_bundle(prefix, tensors):
    """tensors: dict name -> numpy array (float32 / float64 / int32 / int64 / uint8 / bool / float16).
    Writes <prefix>.index and <prefix>.data-00000-of-00001 the way BundleWriter(prefix) + Add(name, tensor) per sorted
    name + Finish() does for a single shard."""
    names = sorted(tensors, key=lambda s: s.encode())
    items = [(b"", encode_header(1))]
    offset = 0
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        for name in names:
            a = np.asarray(tensors[name])
            dt = a.dtype.newbyteorder("<") if a.dtype.byteorder == ">" else a.dtype
            if np.dtype(dt) not in _NP2DT:
                raise TypeError("tensor %s: dtype %s has no TensorFlow bundle mapping here" % (name, a.dtype))
            a = np.asarray(a, dtype=dt)                  # (np.ascontiguousarray would turn a scalar into shape (1,))
            raw = a.tobytes()                            # C order
            f.write(raw)
            items.append((name.encode(), encode_entry(_NP2DT[np.dtype(dt)], a.shape, offset, len(raw), mask_crc(crc32c(raw)))))
            offset += len(raw)
    write_table(prefix + ".index", items)


def read_bundle_index(prefix, verify=True):
    """-> (header fields, {name: entry dict})"""
    rows = read_table(prefix + ".index", verify)
    if not rows or rows[0][0] != b"":
        raise ValueError("%s.index has no bundle header" % prefix)
    header = dict(num_shards=0, endianness=0, version=0)
    for num, _, v in _parse(rows[0][1]):
        if num == 1:
            header["num_shards"] = v
        elif num == 2:
            header["endianness"] = v
        elif num == 3:
            for n2, _, v2 in _parse(v):
                if n2 == 1:
                    header["version"] = v2
    return header, {k.decode(): decode_entry(v) for k, v in rows[1:]}


def read_bundle(prefix, names=None, verify=True):
    """-> {name: numpy array}.  Single- or multi-shard bundles, little endian, whole (unsliced) tensors."""
    header, entries = read_bundle_index(prefix, verify)
    if header["endianness"] != 0:
        raise NotImplementedError("big-endian tensor bundle")
    shards = {}
    out = {}
    for name, e in entries.items():
        if names is not None and name not in names:
            continue
        if e["slices"]:
            raise NotImplementedError("partitioned variable %s (tensor slices)" % name)
        if e["dtype"] not in _DT2NP:
            raise NotImplementedError("tensor %s: TensorFlow dtype %d" % (name, e["dtype"]))
        sid = e["shard_id"]
        if sid not in shards:
            shards[sid] = np.memmap("%s.data-%05d-of-%05d" % (prefix, sid, max(header["num_shards"], 1)), dtype=np.uint8, mode="r")
        raw = np.asarray(shards[sid][e["offset"]:e["offset"] + e["size"]])
        if verify and unmask_crc(e["crc32c"]) != crc32c(raw):
            raise ValueError("tensor %s: checksum mismatch" % name)
        out[name] = raw.view(_DT2NP[e["dtype"]]).reshape(e["shape"]).copy()
    return out
