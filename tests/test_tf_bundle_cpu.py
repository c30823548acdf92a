"""TensorFlow V2 checkpoint (tensor bundle) files without TensorFlow: the byte format is pinned by every independent witness
available in this image -- RFC 3720 CRC-32C vectors, tensorboard's own masked-CRC routine and generated protobuf classes,
the LevelDB table invariants (magic, block checksums, restart arrays, key order) -- and by write -> read round trips."""
import os
import struct

import numpy as np
import pytest

from sketchyscenecolorization_b200 import tf_bundle as tb


def test_crc32c_known_answers():
    # RFC 3720 (iSCSI) B.4 test vectors + the classic check value
    assert tb.crc32c(b"123456789") == 0xE3069283
    assert tb.crc32c(bytes(32)) == 0x8A9136AA
    assert tb.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert tb.crc32c(bytes(range(32))) == 0x46DD794E
    assert tb.crc32c(bytes(range(31, -1, -1))) == 0x113FDB5C
    # incremental == one shot, unaligned starts, numpy input
    data = np.random.default_rng(0).integers(0, 256, 100003, dtype=np.uint8)
    whole = tb.crc32c(data)
    assert tb.crc32c(data[37:].tobytes(), tb.crc32c(data[:37].tobytes())) == whole
    assert tb.crc32c(data.tobytes()) == whole


def test_masked_crc_matches_tensorboard():
    rw = pytest.importorskip("tensorboard.summary.writer.record_writer")
    for blob in (b"", b"123456789", bytes(range(256)) * 5, os.urandom(4097)):
        assert tb.mask_crc(tb.crc32c(blob)) == rw.masked_crc32c(blob)
        assert tb.unmask_crc(tb.mask_crc(tb.crc32c(blob))) == tb.crc32c(blob)


def test_protobuf_submessages_match_generated_classes():
    shape_pb2 = pytest.importorskip("tensorboard.compat.proto.tensor_shape_pb2")
    versions_pb2 = pytest.importorskip("tensorboard.compat.proto.versions_pb2")
    types_pb2 = pytest.importorskip("tensorboard.compat.proto.types_pb2")
    for shape in ((), (7,), (3, 3, 131, 128), (25, 512), (1, 768), (0, 4), (2 ** 33,)):
        ref = shape_pb2.TensorShapeProto(dim=[shape_pb2.TensorShapeProto.Dim(size=d) for d in shape])
        assert tb.encode_shape(shape) == ref.SerializeToString()
        assert tb.decode_shape(ref.SerializeToString()) == tuple(shape)
    # header: num_shards = 1, endianness LITTLE (0, omitted), version { producer: 1 }
    hdr = tb._parse(tb.encode_header(1))
    assert [(n, w) for n, w, _ in hdr] == [(1, 0), (3, 2)] and hdr[0][2] == 1
    assert hdr[1][2] == versions_pb2.VersionDef(producer=1).SerializeToString()
    assert (tb.DT_FLOAT, tb.DT_DOUBLE, tb.DT_INT32, tb.DT_UINT8, tb.DT_INT64, tb.DT_BOOL, tb.DT_BFLOAT16, tb.DT_HALF) == (
        types_pb2.DT_FLOAT, types_pb2.DT_DOUBLE, types_pb2.DT_INT32, types_pb2.DT_UINT8, types_pb2.DT_INT64, types_pb2.DT_BOOL,
        types_pb2.DT_BFLOAT16, types_pb2.DT_HALF)
    # entry: field numbers / wire types of tensor_bundle.proto, zero-valued scalars omitted, crc as fixed32
    e = tb._parse(tb.encode_entry(tb.DT_FLOAT, (3, 4), 0, 48, 0xDEADBEEF))
    assert [(n, w) for n, w, _ in e] == [(1, 0), (2, 2), (5, 0), (6, 5)]
    e = tb.decode_entry(tb.encode_entry(tb.DT_INT32, (), 123456789012, 4, 7, shard_id=2))
    assert e == dict(dtype=tb.DT_INT32, shape=(), shard_id=2, offset=123456789012, size=4, crc32c=7, slices=0)


def test_table_layout_and_roundtrip(tmp_path):
    # many keys with long shared prefixes and a small block size: several data blocks, prefix compression, restarts
    keys = sorted({("generator/mru_conv_unit_t_%d_layer_0/Conv_%d/%s" % (u, c, leaf)).encode()
                   for u in range(1, 9) for c in range(6) for leaf in ("weights", "biases", "weights/Adam", "weights/Adam_1")})
    items = [(b"", b"header")] + [(k, (b"v" + k) * (1 + i % 3)) for i, k in enumerate(keys)]
    path = str(tmp_path / "t.index")
    tb.write_table(path, items, block_size=1024)
    assert tb.read_table(path) == items
    buf = open(path, "rb").read()
    assert struct.unpack_from("<Q", buf, len(buf) - 8)[0] == 0xDB4775248B80FB57           # kTableMagicNumber
    footer = buf[-48:]
    pos = 0
    handles = []
    for _ in range(4):
        v, pos = tb._read_varint(footer, pos)
        handles.append(v)
    assert set(footer[pos:40]) <= {0}                                                       # zero padding up to 40 bytes
    moff, msize, ioff, isize = handles
    assert buf[moff:moff + msize] == struct.pack("<II", 0, 1)                               # empty metaindex block
    index_rows = tb._read_block(buf, ioff, isize)
    assert len(index_rows) > 3                                                              # really several data blocks
    prev_last = b""
    for sep, handle in index_rows:
        off, p2 = tb._read_varint(handle, 0)
        size, _ = tb._read_varint(handle, p2)
        assert buf[off + size] == 0                                                         # kNoCompression
        assert tb.unmask_crc(struct.unpack_from("<I", buf, off + size + 1)[0]) == tb.crc32c(buf[off:off + size + 1])
        rows = tb._read_block(buf, off, size)
        assert rows[0][0] > prev_last or prev_last == b"" and rows[0][0] == b""
        assert rows[-1][0] <= sep                                                           # separator >= last key of its block
        prev_last = rows[-1][0]
        block = buf[off:off + size]
        nrestarts = struct.unpack_from("<I", block, len(block) - 4)[0]
        assert nrestarts == (len(rows) + 15) // 16                                          # one restart every 16 entries
        first_restart = struct.unpack_from("<I", block, len(block) - 4 - 4 * nrestarts)[0]
        assert first_restart == 0 and block[0] == 0                                         # a restart entry shares 0 bytes
    # a flipped byte is caught by the block checksum
    bad = bytearray(buf)
    bad[10] ^= 0x40
    open(path, "wb").write(bad)
    with pytest.raises(ValueError):
        tb.read_table(path)
    # unsorted keys are refused
    with pytest.raises(ValueError):
        tb.write_table(path, [(b"b", b"1"), (b"a", b"2")])


def test_bundle_roundtrip(tmp_path):
    rng = np.random.default_rng(1)
    tensors = {
        "generator/Conv/weights": rng.standard_normal((7, 7, 3, 8)).astype(np.float32),
        "generator/Conv/weights/Adam_1": rng.random((7, 7, 3, 8)).astype(np.float32),
        "generator/TextLSTM/embedding": rng.standard_normal((58, 512)).astype(np.float32),
        "discriminator/Conv/discriminator/Conv/u": rng.standard_normal((1, 8)).astype(np.float32),
        "beta2_power": np.float32(0.81),
        "Variable": np.int32(41),
        "some/int64": np.arange(5, dtype=np.int64),
        "some/empty": np.zeros((0, 3), np.float32),
    }
    prefix = str(tmp_path / "model_41.ckpt-41")
    tb.write_bundle(prefix, tensors)
    assert sorted(os.listdir(tmp_path)) == ["model_41.ckpt-41.data-00000-of-00001", "model_41.ckpt-41.index"]
    header, entries = tb.read_bundle_index(prefix)
    assert header == dict(num_shards=1, endianness=0, version=1)
    names = sorted(tensors, key=lambda s: s.encode())
    assert list(entries) == names                                        # bytewise key order
    off = 0
    for n in names:                                                      # data file = tensors back to back in key order
        e = entries[n]
        assert (e["offset"], e["size"], e["shape"], e["shard_id"]) == (off, tensors[n].nbytes, tuple(tensors[n].shape), 0)
        off += tensors[n].nbytes
    assert os.path.getsize(prefix + ".data-00000-of-00001") == off
    got = tb.read_bundle(prefix)
    for n, a in tensors.items():
        assert got[n].dtype == a.dtype and got[n].shape == a.shape and np.array_equal(got[n], a), n
    assert list(tb.read_bundle(prefix, names={"Variable"})) == ["Variable"]
    # corrupt one tensor byte: the per-tensor checksum catches it
    with open(prefix + ".data-00000-of-00001", "r+b") as f:
        f.seek(entries["generator/Conv/weights"]["offset"] + 5)
        b = f.read(1)
        f.seek(-1, 1)
        f.write(bytes([b[0] ^ 1]))
    with pytest.raises(ValueError):
        tb.read_bundle(prefix)


def test_restore_accepts_the_other_spellings_of_the_provisional_keys(tmp_path):
    """The SN `u` vectors and the LSTM variables may sit under other names in a TensorFlow-written snapshot:
    `restore` finds the single-path `u`, the pre-1.2 `weights` / `biases` LSTM names, and whatever `key_map` says."""
    import torch
    from sketchyscenecolorization_b200 import checkpoint, tf_bundle
    from sketchyscenecolorization_b200.trainer import FgColorModel
    from torch_ops import TorchOps
    m = FgColorModel(TorchOps(torch.float32), "cpu", size=8, H=64, W=64)
    m.initialize(seed=3)
    prefix = checkpoint.save(m, str(tmp_path), 0, 1)
    t = tf_bundle.read_bundle(prefix)
    renamed = {}
    for k, v in t.items():
        k2 = k
        mm = __import__("re").match(r"^(.*)/\1/u$", k)
        if mm:
            k2 = mm.group(1) + "/u"
        k2 = k2.replace("basic_lstm_cell/kernel", "basic_lstm_cell/weights").replace("basic_lstm_cell/bias", "basic_lstm_cell/biases")
        if k2.startswith("generator/TextLSTM/embedding"):
            k2 = k2.replace("generator/TextLSTM/embedding", "generator/TextLSTM/my_embedding")
        renamed[k2] = v
    assert set(renamed) != set(t)
    other = os.path.join(str(tmp_path), "other", "model_0.ckpt-0")
    os.makedirs(os.path.dirname(other))
    tf_bundle.write_bundle(other, renamed)
    m2 = FgColorModel(TorchOps(torch.float32), "cpu", size=8, H=64, W=64)
    m2.initialize(seed=9)
    with pytest.raises(KeyError):
        checkpoint.restore(m2, other)                       # the embedding is under a name no built-in alias covers
    checkpoint.restore(m2, other, key_map={"generator/TextLSTM/embedding": "generator/TextLSTM/my_embedding"})
    assert torch.equal(m2.gstore.flat, m.gstore.flat) and torch.equal(m2.dstore.flat, m.dstore.flat)
    for k, v in m.dstore.state.items():
        assert torch.equal(m2.dstore.state[k], v)
    o = m.gstore.offsets["generator/TextLSTM/embedding"]
    assert torch.equal(m2.gstore.adam_v[o:o + 10], m.gstore.adam_v[o:o + 10])
