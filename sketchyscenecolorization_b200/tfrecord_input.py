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

from .tf_bundle import _field_bytes, _read_varint, _varint, crc32c, mask_crc

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


def _crc_view(view):
    """CRC-32C of a buffer without copying it (tf_bundle.crc32c takes the pointer of a numpy view)."""
    return crc32c(np.frombuffer(view, dtype=np.uint8))


def read_tfrecord_views(path, verify=True):
    """Like read_tfrecord, but the file is memory mapped and every payload is a memoryview into the mapping: the queues hold
    hundreds of 0.9 MB records in their shuffle buffer, and a view costs neither the allocation nor the copy (nor the
    memory) of a bytes object.  The mapping lives as long as any view of it."""
    import mmap
    size = os.path.getsize(path)
    if size == 0:
        return
    with open(path, "rb") as f:
        mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
    mv = memoryview(mm)
    pos = 0
    while pos < size:
        if pos + 12 > size:
            raise ValueError("%s: truncated record header" % path)
        length, lcrc = struct.unpack_from("<QI", mv, pos)
        if verify and mask_crc(_crc_view(mv[pos:pos + 8])) != lcrc:
            raise ValueError("%s: corrupt record length" % path)
        start, end = pos + 12, pos + 12 + length
        if end + 4 > size:
            raise ValueError("%s: truncated record" % path)
        data = mv[start:end]
        if verify and mask_crc(_crc_view(data)) != struct.unpack_from("<I", mv, end)[0]:
            raise ValueError("%s: corrupt record data" % path)
        yield data
        pos = end + 4


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


def _fields(mv):
    """(field number, wire type, value) of one message held in a memoryview; length-delimited values are VIEWS, not copies
    (a record carries two 442 KB payloads five messages deep: copying them at every level dominated the reader)."""
    pos, n = 0, len(mv)
    while pos < n:
        key, pos = _read_varint(mv, pos)
        num, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _read_varint(mv, pos)
        elif wt == 2:
            ln, pos = _read_varint(mv, pos)
            v = mv[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = struct.unpack_from("<I", mv, pos)[0]
            pos += 4
        elif wt == 1:
            v = struct.unpack_from("<Q", mv, pos)[0]
            pos += 8
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield num, wt, v


_SMALL = 4096       # byte values up to this size are returned as `bytes`, larger ones as memoryviews into the record


def parse_example(buf):
    """-> dict name -> bytes | memoryview (single bytes value; memoryview above 4 KB) | list of them | list of ints |
    list of floats."""
    out = {}
    for num, _, feats in _fields(memoryview(buf)):
        if num != 1:
            continue
        for n2, _, entry in _fields(feats):
            if n2 != 1:
                continue
            key, val = None, None
            for n3, _, v in _fields(entry):
                if n3 == 1:
                    key = bytes(v).decode()
                elif n3 == 2:
                    val = v
            if key is None or val is None:
                continue
            for kind, _, lst in _fields(val):
                if kind == 1:                                        # BytesList
                    items = [(v if len(v) > _SMALL else bytes(v)) for n5, _, v in _fields(lst) if n5 == 1]
                    out[key] = items[0] if len(items) == 1 else items
                elif kind == 2:                                      # FloatList: packed or repeated fixed32
                    vals = []
                    for n5, wt, v in _fields(lst):
                        if n5 == 1 and wt == 2:
                            vals += list(struct.unpack("<%df" % (len(v) // 4), v))
                        elif n5 == 1:
                            vals.append(struct.unpack("<f", struct.pack("<I", v))[0])
                    out[key] = vals
                elif kind == 3:                                      # Int64List: packed or repeated varints
                    vals = []
                    for n5, wt, v in _fields(lst):
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
    return dict(cartoon=cartoon, sketch=sketch, cls=int(cid[0] if isinstance(cid, list) else cid),
                category=bytes(ex['Category']).decode(), name=bytes(ex['ImageName']).decode(),
                color_text=bytes(ex['Color_text']).decode(), text=text)


def _record_files(data_base_dir, mode):
    d = os.path.join(data_base_dir, 'tfrecord', mode)
    files = sorted(os.path.join(d, f) for f in os.listdir(d) if os.path.isfile(os.path.join(d, f)))
    print("build_input_queue_paired from %s: paired file num: %d" % (d, len(files)))
    return files


def _host_buffers(n, sketch_dtype, pinned):
    shape = (n, RAW, RAW, 3)
    mk = lambda dt: torch.empty(shape, dtype=dt, pin_memory=pinned)          # noqa: E731
    return mk(torch.uint8), mk(torch.float32 if sketch_dtype == np.float32 else torch.uint8)


def _fill_slot(rec, distance_map, cartoon_np, sketch_np, i):
    """Worker-thread half of a batch: parse one record (zero copy) and copy its two payloads into slot i of the batch's host
    buffers (numpy's copy releases the GIL).  Returns the sample's metadata."""
    s = raw_paired_example(parse_example(rec), distance_map)
    np.copyto(cartoon_np[i], s.pop('cartoon'))
    np.copyto(sketch_np[i], s.pop('sketch'))
    return s


def _finish_batch(meta, cartoon, sketch, ops, dim, seed, dequantize=True, want_d=False, slot=None, copy_stream=None):
    """Consumer half: ONE copy of the raw uint8 batch to the device of `ops` (0.88 MB per sample instead of the fp32 results)
    and ONE device call for everything per pixel (ops.paired_input -> fgc_paired_input).  `slot`: the reusable host-buffer
    record of the batch; the event recorded after the copies tells the producer when the buffers may be overwritten."""
    dev = torch.device(getattr(ops, 'device', 'cpu'))
    if dev.type == 'cuda' and copy_stream is not None:
        # the copies go out on their own stream: the host runs ahead of the device, so they overlap the kernels of the
        # previous training step still executing on the compute stream, which then only waits for the event
        cur = torch.cuda.current_stream(dev)
        with torch.cuda.stream(copy_stream):
            cartoon, sketch = cartoon.to(dev, non_blocking=True), sketch.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        cur.wait_event(ev)
        cartoon.record_stream(cur)
        sketch.record_stream(cur)
        if slot is not None:
            slot['event'] = ev
    elif dev.type == 'cuda':
        cartoon, sketch = cartoon.to(dev, non_blocking=True), sketch.to(dev, non_blocking=True)
        if slot is not None:
            slot['event'] = torch.cuda.Event()
            slot['event'].record()
    images, sketches = ops.paired_input(cartoon, sketch, dim, seed=seed, dequantize=dequantize)
    out = dict(sketch=sketches, images=images, cls=torch.tensor([s['cls'] for s in meta], dtype=torch.int32),
               text=torch.from_numpy(np.stack([s['text'] for s in meta])),
               categories=[s['category'] for s in meta], image_names=[s['name'] for s in meta],
               color_texts=[s['color_text'] for s in meta])
    if want_d:                                              # the discriminator's own queue: same record fields under *_d names
        out['images_d'], out['cls_d'] = out['images'], out['cls']
    return out


class PairedTrainInput:
    """build_input_queue_paired('train') (:131-157): an endless shuffled stream of batches.  Files are visited in a
    shuffled order every epoch (string_input_producer(shuffle=True)); records pass through a shuffle buffer that holds
    at least `min_after_dequeue` of them before one is drawn at random (tf.train.shuffle_batch).  A reader thread streams
    the files (hardware CRC-32C of every frame) into the buffer; `num_threads` workers parse the drawn records without
    copying and write their payloads straight into the batch's pinned host buffers (plus the distance transform, when asked
    for); `prefetch` batches are in flight ahead of the consumer, whose share is one host-to-device copy and one device call
    per batch on its own stream (`ops`: the model's operator set).  Under data parallelism give every rank its own `seed`.
    main_procedure.train uses two of these (the second feeds images_d)."""

    def __init__(self, batch_size, ops, data_base_dir='data', small=False, distance_map=False, min_after_dequeue=512, seed=0,
                 num_threads=None, prefetch=4, mode='train', side_stream=None):
        import queue
        import threading
        from concurrent.futures import ThreadPoolExecutor
        self.n, self.ops, self.dim, self.dm = batch_size, ops, ((64, 64) if small else (192, 192)), distance_map
        self.files = _record_files(data_base_dir, mode)
        if not self.files:
            raise FileNotFoundError("no TFRecord files under %s" % os.path.join(data_base_dir, 'tfrecord', mode))
        self.rng = np.random.default_rng(seed)
        self.min_after = min_after_dequeue
        if num_threads is None:     # the reference uses 4; take more where the host has them (two queues per rank, N ranks per box)
            ranks = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
            num_threads = max(4, min(8, (os.cpu_count() or 8) // (2 * ranks)))
        self.pool = ThreadPoolExecutor(max_workers=num_threads)
        self.pinned = torch.device(getattr(ops, 'device', 'cpu')).type == 'cuda'
        if side_stream is None:                             # opt-in until it has been measured (FGC_INPUT_SIDE_STREAM=1)
            side_stream = os.environ.get("FGC_INPUT_SIDE_STREAM") == "1"
        self.copy_stream = torch.cuda.Stream(device=getattr(ops, 'device')) if (side_stream and self.pinned) else None
        self.buf = []
        self.prefetch = prefetch
        self.pending = []
        self._slots = []                                    # reusable (pinned) host buffers: allocated once, prefetch + 1 of them
        # the file order is drawn here (seeded) and handed to the reader, so that a seed fixes the whole stream
        self._orders = queue.Queue(maxsize=2)
        self._records = queue.Queue(maxsize=max(2 * batch_size, 64))
        self._closed = False
        self._orders.put(self.rng.permutation(len(self.files)))
        self._reader = threading.Thread(target=self._read_loop, daemon=True)
        self._reader.start()

    def _read_loop(self):
        try:
            while not self._closed:
                order = self._orders.get()
                for i in order:
                    for rec in read_tfrecord_views(self.files[i]):
                        self._records.put(rec)
                        if self._closed:
                            return
                self._records.put(None)                 # end of an epoch: the consumer draws the next file order
        except Exception as e:                          # noqa: BLE001 -- surfaced in the consumer thread
            self._records.put(e)

    def _next_record(self):
        while True:
            rec = self._records.get()
            if rec is None:
                self._orders.put(self.rng.permutation(len(self.files)))
                continue
            if isinstance(rec, Exception):
                raise rec
            return rec

    def _draw_raw(self):
        while len(self.buf) < self.min_after + 1:
            self.buf.append(self._next_record())
        j = int(self.rng.integers(len(self.buf)))
        self.buf[j], self.buf[-1] = self.buf[-1], self.buf[j]
        return self.buf.pop()

    def _submit(self):
        raws = [self._draw_raw() for _ in range(self.n)]
        seed = int(self.rng.integers(1 << 62))              # the batch's dequantisation-noise stream
        slot = self._free_slot()
        cn, sn = slot['cartoon'].numpy(), slot['sketch'].numpy()
        futs = [self.pool.submit(_fill_slot, r, self.dm, cn, sn, i) for i, r in enumerate(raws)]
        return seed, futs, slot

    def _free_slot(self):
        """A host-buffer pair no batch in flight is using (page-locked memory is expensive to allocate: cudaHostAlloc)."""
        busy = {id(p[2]) for p in self.pending}
        for sl in self._slots:
            if id(sl) not in busy:
                if sl.get('event') is not None:            # its last host-to-device copy must have left the buffers
                    sl['event'].synchronize()
                    sl['event'] = None
                return sl
        cartoon, sketch = _host_buffers(self.n, np.float32 if self.dm else np.uint8, self.pinned)
        self._slots.append(dict(cartoon=cartoon, sketch=sketch, event=None))
        return self._slots[-1]

    def __iter__(self):
        return self

    def __next__(self):
        while len(self.pending) < self.prefetch:
            self.pending.append(self._submit())
        seed, futs, slot = self.pending.pop(0)
        meta = [f.result() for f in futs]
        return _finish_batch(meta, slot['cartoon'], slot['sketch'], self.ops, self.dim, seed, want_d=True, slot=slot,
                             copy_stream=self.copy_stream)

    def close(self):
        """Stop the reader thread, the parser pool and drop the page-locked batch buffers (the NaN-restart loop of
        obj_colorization_main builds a fresh pair of queues every time)."""
        import queue
        self._closed = True
        try:                                    # the reader may be blocked in put() on a full queue, or in get() on the orders
            while True:
                self._records.get_nowait()
        except queue.Empty:
            pass
        try:
            self._orders.put_nowait(np.zeros(0, dtype=np.int64))
        except queue.Full:
            pass
        for _, futs, _ in self.pending:
            for f in futs:
                f.cancel()
        self.pool.shutdown(wait=False)
        self.pending, self.buf, self._slots = [], [], []


class PairedEvalInput:
    """build_input_queue_paired_test('val' | 'test') (:160-181): one ordered epoch; the last partial batch is dropped as
    tf.train.batch does."""

    def __init__(self, mode, batch_size, ops, data_base_dir='data', small=False, distance_map=False):
        assert mode in ('test', 'val')
        self.n, self.ops, self.dim, self.dm = batch_size, ops, ((64, 64) if small else (192, 192)), distance_map
        self.files = _record_files(data_base_dir, mode)

    def __iter__(self):
        pinned = torch.device(getattr(self.ops, 'device', 'cpu')).type == 'cuda'
        recs, k = [], 0
        for f in self.files:
            for rec in read_tfrecord_views(f):
                recs.append(rec)
                if len(recs) == self.n:
                    cartoon, sketch = _host_buffers(self.n, np.float32 if self.dm else np.uint8, pinned)
                    cn, sn = cartoon.numpy(), sketch.numpy()
                    meta = [_fill_slot(r, self.dm, cn, sn, i) for i, r in enumerate(recs)]
                    yield _finish_batch(meta, cartoon, sketch, self.ops, self.dim, seed=k)
                    recs, k = [], k + 1
