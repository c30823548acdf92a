"""The reference's real input pipeline without TensorFlow: TFRecord files of tf.train.Example -> training / validation batches.

Reference: Foreground_Instance_Colorization/data_preparation/data_preparation.py:21-32 writes, per category, a TFRecord of

    ImageName (bytes), cartoon_data (raw 384x384x3 uint8), sketch_data (raw 384x384x3 uint8), Category (bytes),
    Category_id (int64), Color_text (bytes), Text_vocab_indices (15 raw uint8)

and obj_lib/input_pipeline.py reads them back: `get_paired_input` (:45-126: decode, optional distance map, resize to
192 -- image BILINEAR, sketch AREA --, min-max normalise + dequantisation noise, [-1,1], NCHW), `build_input_queue_paired`
(:131-157: shuffle queue, min_after_dequeue 512) and `build_input_queue_paired_test` (:160-181: one ordered epoch).

File format (tensorflow/core/lib/io/record_writer.cc, example.proto, feature.proto):
    record  = uint64 length | uint32 masked_crc32c(length) | data | uint32 masked_crc32c(data)      (little endian)
    Example = { Features features = 1 { map<string, Feature> feature = 1 } }
    Feature = oneof { BytesList bytes_list = 1; FloatList float_list = 2 (packed floats); Int64List int64_list = 3 (packed varints) }

Split of the work: the record framing, the proto parsing, the shuffle queue and (only with --distance_map 1) scipy's
distance transform -- which the reference also runs on the host, through tf.py_func -- stay on host threads; everything
per pixel (resize, min-max normalisation, dequantisation noise, [-1,1] map, NCHW transpose) is one device call per batch,
`fgc_paired_input` (csrc/input.cu) behind `ops.paired_input`.  There is no host implementation of that step in this
package (the tests check the kernel against a numpy restatement kept outside it).

TF-1 resize semantics that matter at 384 -> 192 (legacy kernels, align_corners = False, no half-pixel centres): the source
coordinate of output pixel y is 2*y exactly, so BILINEAR picks the top-left pixel of each 2x2 block, while AREA averages
the block.  The framing and the proto encoding are pinned against tensorboard's independent TFRecord writer and its
generated protobuf classes (tests/test_tfrecord_cpu.py).
"""
from __future__ import annotations

import os
import struct

import numpy as np
import torch

from .tf_bundle import _field_bytes, _parse, _read_varint, _varint, crc32c, mask_crc

RAW = 384            # "cannot change" (input_pipeline.py:82,87)
TEXT_LEN = 15


# ----------------------------------------------------------------------------------------------------------
# TFRecord framing
# ----------------------------------------------------------------------------------------------------------
def read_tfrecord(path, verify=True):
    """Yields the payload bytes of every record."""
    with open(path, "rb") as f:
        while True:
            head = f.read(12)
            if not head:
                return
            if len(head) < 12:
                raise ValueError("%s: truncated record header" % path)
            length, lcrc = struct.unpack("<QI", head)
            if verify and mask_crc(crc32c(head[:8])) != lcrc:
                raise ValueError("%s: corrupt record length" % path)
            data = f.read(length)
            tail = f.read(4)
            if len(data) < length or len(tail) < 4:
                raise ValueError("%s: truncated record" % path)
            if verify and mask_crc(crc32c(data)) != struct.unpack("<I", tail)[0]:
                raise ValueError("%s: corrupt record data" % path)
            yield data


def write_tfrecord(path, records):
    with open(path, "wb") as f:
        for data in records:
            head = struct.pack("<Q", len(data))
            f.write(head + struct.pack("<I", mask_crc(crc32c(head))) + data + struct.pack("<I", mask_crc(crc32c(data))))


# ----------------------------------------------------------------------------------------------------------
# tf.train.Example
# ----------------------------------------------------------------------------------------------------------
def encode_example(features):
    """features: dict name -> bytes | str | int | list of ints | list of floats.  Keys are written sorted (map order is free)."""
    out = b""
    for name in sorted(features):
        v = features[name]
        if isinstance(v, str):
            v = v.encode()
        if isinstance(v, (bytes, bytearray)):
            feat = _field_bytes(1, _field_bytes(1, bytes(v)))                                   # BytesList { value = 1 }
        else:
            vals = list(v) if isinstance(v, (list, tuple, np.ndarray)) else [v]
            if vals and isinstance(vals[0], (float, np.floating)):
                feat = _field_bytes(2, _field_bytes(1, struct.pack("<%df" % len(vals), *vals)))  # FloatList, packed
            else:
                feat = _field_bytes(3, _field_bytes(1, b"".join(_varint(int(x)) for x in vals)))  # Int64List, packed
        entry = _field_bytes(1, name.encode()) + _field_bytes(2, feat)                          # map entry { key = 1, value = 2 }
        out += _field_bytes(1, entry)                                                           # Features.feature
    return _field_bytes(1, out)                                                                 # Example.features


def parse_example(buf):
    """-> dict name -> bytes (single bytes value) | list of bytes | list of ints | list of floats."""
    out = {}
    for num, _, feats in _parse(buf):
        if num != 1:
            continue
        for n2, _, entry in _parse(feats):
            if n2 != 1:
                continue
            key, val = None, None
            for n3, _, v in _parse(entry):
                if n3 == 1:
                    key = v.decode()
                elif n3 == 2:
                    val = v
            if key is None or val is None:
                continue
            for kind, _, lst in _parse(val):
                if kind == 1:                                        # BytesList
                    items = [v for n5, _, v in _parse(lst) if n5 == 1]
                    out[key] = items[0] if len(items) == 1 else items
                elif kind == 2:                                      # FloatList: packed or repeated fixed32
                    vals = []
                    for n5, wt, v in _parse(lst):
                        if n5 == 1 and wt == 2:
                            vals += list(struct.unpack("<%df" % (len(v) // 4), v))
                        elif n5 == 1:
                            vals.append(struct.unpack("<f", struct.pack("<I", v))[0])
                    out[key] = vals
                elif kind == 3:                                      # Int64List: packed or repeated varints
                    vals = []
                    for n5, wt, v in _parse(lst):
                        if n5 == 1 and wt == 2:
                            pos = 0
                            while pos < len(v):
                                x, pos = _read_varint(v, pos)
                                vals.append(x if x < (1 << 63) else x - (1 << 64))
                        elif n5 == 1:
                            vals.append(v if v < (1 << 63) else v - (1 << 64))
                    out[key] = vals
    return out


# ----------------------------------------------------------------------------------------------------------
# get_paired_input: the record fields on the host, the per-pixel work in one device pass per batch
# ----------------------------------------------------------------------------------------------------------
def distance_map_255(sketch_u8):
    """:90-100 (Config.pre_calculated_dist_map is False): binarise at 250, Euclidean distance transform, scale so the
    maximum is 255.  The reference runs exactly this on the host too (scipy inside tf.py_func); float32 [R,R,3]."""
    from scipy import ndimage
    binar = np.where(sketch_u8 < 250, 0.0, 255.0).astype(np.float32)
    dist = ndimage.distance_transform_edt(binar).astype(np.float32)
    return dist / dist.max() * np.float32(255.)


def raw_paired_example(ex, distance_map=False):
    """One parsed Example -> the raw sample: cartoon uint8 [384,384,3], sketch uint8 [384,384,3] (float32 0..255 distance
    map when `distance_map`), cls, category, name, color_text, text int32 [15] (:72-88, :117-124)."""
    cartoon = np.frombuffer(ex['cartoon_data'], dtype=np.uint8).reshape(RAW, RAW, 3)
    sketch = np.frombuffer(ex['sketch_data'], dtype=np.uint8).reshape(RAW, RAW, 3)
    if distance_map:
        sketch = distance_map_255(sketch)
    text = np.frombuffer(ex['Text_vocab_indices'], dtype=np.uint8).astype(np.int32).reshape(TEXT_LEN)
    cid = ex['Category_id']
    return dict(cartoon=cartoon, sketch=sketch, cls=int(cid[0] if isinstance(cid, list) else cid), category=ex['Category'].decode(),
                name=ex['ImageName'].decode(), color_text=ex['Color_text'].decode(), text=text)


def _record_files(data_base_dir, mode):
    d = os.path.join(data_base_dir, 'tfrecord', mode)
    files = sorted(os.path.join(d, f) for f in os.listdir(d) if os.path.isfile(os.path.join(d, f)))
    print("build_input_queue_paired from %s: paired file num: %d" % (d, len(files)))
    return files


def _device_batch(samples, ops, dim, seed, dequantize=True, want_d=False):
    """Raw samples -> the batch dict of the queues.  The uint8 payloads are stacked into (pinned) host memory, copied to the
    device of `ops` and resized / normalised / dequantised / transposed there by ONE call (ops.paired_input ->
    fgc_paired_input): 0.88 MB per sample cross the bus instead of the 0.88 MB of fp32 results plus the host arithmetic."""
    dev = torch.device(getattr(ops, 'device', 'cpu'))
    cartoon = torch.from_numpy(np.stack([s['cartoon'] for s in samples]))
    sketch = torch.from_numpy(np.stack([s['sketch'] for s in samples]))
    if dev.type == 'cuda':
        cartoon, sketch = cartoon.pin_memory().to(dev, non_blocking=True), sketch.pin_memory().to(dev, non_blocking=True)
    images, sketches = ops.paired_input(cartoon, sketch, dim, seed=seed, dequantize=dequantize)
    out = dict(sketch=sketches, images=images, cls=torch.tensor([s['cls'] for s in samples], dtype=torch.int32),
               text=torch.from_numpy(np.stack([s['text'] for s in samples])),
               categories=[s['category'] for s in samples], image_names=[s['name'] for s in samples],
               color_texts=[s['color_text'] for s in samples])
    if want_d:                                              # the discriminator's own queue: same record fields under *_d names
        out['images_d'], out['cls_d'] = out['images'], out['cls']
    return out


class PairedTrainInput:
    """build_input_queue_paired('train') (:131-157): an endless shuffled stream of batches.  Files are visited in a
    shuffled order every epoch (string_input_producer(shuffle=True)); samples pass through a shuffle buffer that holds
    at least `min_after_dequeue` records before one is drawn at random (tf.train.shuffle_batch).  Record parsing (and the
    distance transform, when asked for) runs in `num_threads` host threads, `prefetch` batches ahead of the consumer; the
    pixel work of a batch is one device call on the consumer's stream (`ops`: the model's operator set).  Under data
    parallelism give every rank its own `seed`.  main_procedure.train uses two of these (the second feeds images_d)."""

    def __init__(self, batch_size, ops, data_base_dir='data', small=False, distance_map=False, min_after_dequeue=512, seed=0,
                 num_threads=4, prefetch=4, mode='train'):
        from concurrent.futures import ThreadPoolExecutor
        self.n, self.ops, self.dim, self.dm = batch_size, ops, ((64, 64) if small else (192, 192)), distance_map
        self.files = _record_files(data_base_dir, mode)
        if not self.files:
            raise FileNotFoundError("no TFRecord files under %s" % os.path.join(data_base_dir, 'tfrecord', mode))
        self.rng = np.random.default_rng(seed)
        self.min_after = min_after_dequeue
        self.pool = ThreadPoolExecutor(max_workers=num_threads)
        self.buf = []
        self.raw = self._raw_stream()
        self.prefetch = prefetch
        self.pending = []

    def _raw_stream(self):
        while True:
            order = self.rng.permutation(len(self.files))
            for i in order:
                for rec in read_tfrecord(self.files[i]):
                    yield rec

    def _draw_raw(self):
        while len(self.buf) < self.min_after + 1:
            self.buf.append(next(self.raw))
        j = int(self.rng.integers(len(self.buf)))
        self.buf[j], self.buf[-1] = self.buf[-1], self.buf[j]
        return self.buf.pop()

    def _submit(self):
        raws = [self._draw_raw() for _ in range(self.n)]
        seed = int(self.rng.integers(1 << 62))              # the batch's dequantisation-noise stream
        return seed, [self.pool.submit(lambda r: raw_paired_example(parse_example(r), self.dm), r) for r in raws]

    def __iter__(self):
        return self

    def __next__(self):
        while len(self.pending) < self.prefetch:
            self.pending.append(self._submit())
        seed, futs = self.pending.pop(0)
        return _device_batch([f.result() for f in futs], self.ops, self.dim, seed, want_d=True)


class PairedEvalInput:
    """build_input_queue_paired_test('val' | 'test') (:160-181): one ordered epoch; the last partial batch is dropped as
    tf.train.batch does."""

    def __init__(self, mode, batch_size, ops, data_base_dir='data', small=False, distance_map=False):
        assert mode in ('test', 'val')
        self.n, self.ops, self.dim, self.dm = batch_size, ops, ((64, 64) if small else (192, 192)), distance_map
        self.files = _record_files(data_base_dir, mode)

    def __iter__(self):
        batch, k = [], 0
        for f in self.files:
            for rec in read_tfrecord(f):
                batch.append(raw_paired_example(parse_example(rec), self.dm))
                if len(batch) == self.n:
                    yield _device_batch(batch, self.ops, self.dim, seed=k)
                    batch, k = [], k + 1
