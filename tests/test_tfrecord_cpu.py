"""TFRecord / tf.train.Example input pipeline without TensorFlow (tfrecord_input.py): the record framing is checked against
tensorboard's independent RecordWriter, the Example encoding against the protobuf runtime (schema of feature.proto /
example.proto built with descriptor_pb2), the pre-processing against the TF-1 resize / normalisation rules of
input_pipeline.get_paired_input, and `--mode val` end to end on a small CPU model."""
import io
import os
import struct

import numpy as np
import pytest
import torch

from sketchyscenecolorization_b200 import tfrecord_input as TI


def _example(i, rng):
    img = rng.integers(0, 256, (384, 384, 3), dtype=np.uint8)
    sk = np.full((384, 384, 3), 255, np.uint8)
    sk[40 + i:44 + i, 20:300] = 0
    sk[100:300, 150 + 2 * i:153 + 2 * i] = 0
    text = np.array([0] * 9 + [24, 3, 6, 22, 5, 25], dtype=np.uint8)
    return dict(ImageName=("img%03d.png" % i).encode(), cartoon_data=img.tobytes(), sketch_data=sk.tobytes(), Category=b"bus",
                Category_id=2, Color_text=b"the bus is orange with gray windows", Text_vocab_indices=text.tobytes()), img, sk, text


def test_record_framing_matches_tensorboard(tmp_path):
    rw = pytest.importorskip("tensorboard.summary.writer.record_writer")
    payloads = [b"", b"x", os.urandom(1000), bytes(range(256)) * 3]
    path = str(tmp_path / "a.tfrecord")
    w = rw.RecordWriter(open(path, "wb"))
    for p in payloads:
        w.write(p)
    w.close()
    assert list(TI.read_tfrecord(path)) == payloads
    mine = str(tmp_path / "b.tfrecord")
    TI.write_tfrecord(mine, payloads)
    assert open(mine, "rb").read() == open(path, "rb").read()
    bad = bytearray(open(mine, "rb").read())
    bad[14] ^= 1                                            # first data byte of the second record
    open(mine, "wb").write(bad)
    with pytest.raises(ValueError):
        list(TI.read_tfrecord(mine))


def _feature_messages():
    """Example / Features / Feature / *List message classes from the protobuf runtime (schema of feature.proto)."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    F = descriptor_pb2.FieldDescriptorProto
    fd = descriptor_pb2.FileDescriptorProto(name="fgc_feature_test.proto", package="fgctest", syntax="proto3")

    def msg(name, fields, nested=None):
        m = fd.message_type.add(name=name)
        for fname, num, typ, label, tname in fields:
            f = m.field.add(name=fname, number=num, type=typ, label=label)
            if tname:
                f.type_name = tname
        return m

    msg("BytesList", [("value", 1, F.TYPE_BYTES, F.LABEL_REPEATED, None)])
    msg("FloatList", [("value", 1, F.TYPE_FLOAT, F.LABEL_REPEATED, None)])
    msg("Int64List", [("value", 1, F.TYPE_INT64, F.LABEL_REPEATED, None)])
    feat = msg("Feature", [("bytes_list", 1, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, ".fgctest.BytesList"),
                           ("float_list", 2, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, ".fgctest.FloatList"),
                           ("int64_list", 3, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, ".fgctest.Int64List")])
    feat.oneof_decl.add(name="kind")
    for f in feat.field:
        f.oneof_index = 0
    feats = msg("Features", [("feature", 1, F.TYPE_MESSAGE, F.LABEL_REPEATED, ".fgctest.Features.FeatureEntry")])
    entry = feats.nested_type.add(name="FeatureEntry")
    entry.field.add(name="key", number=1, type=F.TYPE_STRING, label=F.LABEL_OPTIONAL)
    entry.field.add(name="value", number=2, type=F.TYPE_MESSAGE, label=F.LABEL_OPTIONAL, type_name=".fgctest.Feature")
    entry.options.map_entry = True
    msg("Example", [("features", 1, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, ".fgctest.Features")])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("fgctest.Example"))


def test_example_encoding_matches_protobuf_runtime():
    Example = _feature_messages()
    feats = dict(ImageName=b"a.png", Category_id=7, Text_vocab_indices=bytes(range(15)), floats=[0.5, -2.0, 3.25],
                 ints=[1, -1, 1 << 40], blob=os.urandom(300))
    ex = Example()
    for k, v in feats.items():
        f = ex.features.feature[k]
        if isinstance(v, bytes):
            f.bytes_list.value.append(v)
        elif isinstance(v, list) and isinstance(v[0], float):
            f.float_list.value.extend(v)
        else:
            f.int64_list.value.extend(v if isinstance(v, list) else [v])
    # what the runtime wrote parses here ...
    got = TI.parse_example(ex.SerializeToString())
    assert got["ImageName"] == b"a.png" and got["Category_id"] == [7] and got["ints"] == [1, -1, 1 << 40]
    assert got["floats"] == [0.5, -2.0, 3.25] and got["blob"] == feats["blob"] and got["Text_vocab_indices"] == bytes(range(15))
    # ... and what is written here parses in the runtime to the same message
    back = Example()
    back.ParseFromString(TI.encode_example(feats))
    assert back == ex


def test_decode_follows_tf1_preprocessing():
    """The raw-sample view of a record, and the oracle restatement of get_paired_input's pixel work (the checker of
    fgc_paired_input; the product has no host implementation of it)."""
    from oracle import input_oracle as IO
    rng = np.random.default_rng(0)
    feats, img, sk, text = _example(3, rng)
    ex = TI.parse_example(TI.encode_example(feats))
    raw = TI.raw_paired_example(ex)
    assert raw["cartoon"].dtype == np.uint8 and np.array_equal(raw["cartoon"], img) and np.array_equal(raw["sketch"], sk)
    assert raw["cls"] == 2 and raw["category"] == "bus" and raw["name"] == "img003.png" and list(raw["text"]) == list(text)
    images, sketch = IO.paired_preprocess(raw["cartoon"], raw["sketch"], (192, 192))
    assert images.shape == (3, 192, 192) and sketch.shape == (3, 192, 192) and images.dtype == np.float32
    # BILINEAR 384 -> 192 in TF 1 (no half-pixel centres): the top-left pixel of every 2x2 block
    sub = img[::2, ::2].astype(np.float32)
    want = (sub - sub.min()) / (sub.max() - sub.min() + 1) * 2 - 1
    assert np.allclose(images, want.transpose(2, 0, 1), atol=1e-6)
    # AREA: 2x2 block mean; /255*2-1
    area = sk.astype(np.float32).reshape(192, 2, 192, 2, 3).mean(axis=(1, 3))
    assert np.allclose(sketch, (area / 255 * 2 - 1).transpose(2, 0, 1), atol=1e-6)
    assert sketch.min() >= -1 and sketch.max() <= 1
    # dequantisation noise: U[0, 1/256) before the [-1,1] map -> at most 2/256 above the noiseless value, never below
    noise = IO.splitmix_noise(1, 3 * 192 * 192)
    assert noise.min() >= 0 and noise.max() < 1 / 256 and abs(noise.mean() - 0.5 / 256) < 2e-5 and len(np.unique(noise)) > 100000
    assert np.array_equal(IO.splitmix_noise(1, 10, offset=5), noise[5:15])
    dn, _ = IO.paired_preprocess(raw["cartoon"], raw["sketch"], (192, 192), noise.reshape(3, 192, 192))
    delta = dn - images
    assert delta.min() >= -1e-6 and delta.max() <= 2.0 / 256 + 1e-6 and delta.std() > 1e-4
    # splitmix64 known answers (seed 1234567: the published reference outputs of the generator)
    with np.errstate(over="ignore"):
        i = np.arange(2, dtype=np.uint64)
        z = np.uint64(1234567) + (i + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    assert [int(v) for v in z] == [6457827717110365317, 3203168211198807973]
    assert np.array_equal(IO.splitmix_noise(1234567, 2), (z >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -32))
    # generic bilinear rule at a non-integer factor against a direct evaluation
    small = IO.resize_bilinear_tf1(img.astype(np.float32), (100, 100))
    y, x = 37, 81
    sy, sx = y * 3.84, x * 3.84
    y0, x0 = int(sy), int(sx)
    fy, fx = sy - y0, sx - x0
    ref = (img[y0, x0] * (1 - fx) + img[y0, x0 + 1] * fx) * (1 - fy) + (img[y0 + 1, x0] * (1 - fx) + img[y0 + 1, x0 + 1] * fx) * fy
    assert np.allclose(small[y, x], ref, atol=1e-3)
    # distance map branch: normalised EDT of the binarised sketch (host work in the reference as well), then the same pass
    dm = TI.raw_paired_example(ex, distance_map=True)["sketch"]
    assert dm.dtype == np.float32 and dm.max() == pytest.approx(255.0) and dm.min() == 0.0
    _, dms = IO.paired_preprocess(raw["cartoon"], dm, (192, 192))
    assert 0.95 < dms.max() <= 1.0 and dms.min() == pytest.approx(-1.0, abs=1e-5)   # block means of EDT / max


def _write_split(base, mode, n_per_file, files=("bus", "car")):
    d = os.path.join(base, "tfrecord", mode)
    os.makedirs(d, exist_ok=True)
    rng = np.random.default_rng(7)
    k = 0
    for name in files:
        recs = []
        for _ in range(n_per_file):
            recs.append(TI.encode_example(_example(k, rng)[0]))
            k += 1
        TI.write_tfrecord(os.path.join(d, name + ".tfrecord"), recs)
    return k


def test_train_and_eval_queues(tmp_path):
    base = str(tmp_path)
    total = _write_split(base, "train", 5)
    from torch_ops import TorchOps
    ops = TorchOps(torch.float32)
    q = TI.PairedTrainInput(4, ops, base, small=True, min_after_dequeue=6, seed=3, num_threads=2, prefetch=2)
    seen = set()
    for _ in range(12):
        b = next(q)
        assert b["sketch"].shape == (4, 3, 64, 64) and b["images"].shape == (4, 3, 64, 64) and b["text"].shape == (4, 15)
        assert b["cls"].dtype == torch.int32 and b["images_d"] is b["images"] and b["cls_d"] is b["cls"]
        assert float(b["images"].min()) >= -1.0 and float(b["images"].max()) <= 1.0
        seen.update(b["image_names"])
    assert len(seen) == total                               # the shuffle buffer lets every sample through
    a, b = next(TI.PairedTrainInput(4, ops, base, small=True, min_after_dequeue=6, seed=9)), next(TI.PairedTrainInput(4, ops, base, small=True, min_after_dequeue=6, seed=9))
    assert a["image_names"] == b["image_names"] and torch.equal(a["images"], b["images"])     # seeded: reproducible
    _write_split(base, "val", 3)
    batches = list(TI.PairedEvalInput("val", 4, ops, base, small=True))
    assert len(batches) == 1 and batches[0]["image_names"] == ["img000.png", "img001.png", "img002.png", "img003.png"]   # 6 // 4


def test_validation_mode_end_to_end(tmp_path, monkeypatch):
    """`--mode val`: ordered pass over data/tfrecord/val with a resident model -> <category>_<name>_{output,target,input}.png."""
    from sketchyscenecolorization_b200 import main_procedure
    from sketchyscenecolorization_b200.config import Config
    from sketchyscenecolorization_b200.trainer import FgColorModel
    from torch_ops import TorchOps
    base = str(tmp_path)
    _write_split(base, "val", 2, files=("bus",))
    Config.set_from_dict(dict(dataset_type="val", batch_size=2, ckpt_dir=os.path.join(base, "snapshot"),
                              results_dir=os.path.join(base, "validation_results"), data_format="NCHW", distance_map=0, small_img=1,
                              LSTM_hybrid=1, block_type="MRU", vocab_size=58))
    model = FgColorModel(TorchOps(torch.float64), "cpu", size=16, H=64, W=64, param_dtype=torch.float64, with_discriminator=False)
    model.initialize(seed=1, perturb_tables=0.1)
    n = main_procedure.validation(model=model, data_base_dir=base, noise=torch.zeros(2, 256))
    assert n == 1
    out = os.path.join(base, "validation_results", "with_text")
    import cv2
    for i in range(2):
        for kind in ("output", "target", "input"):
            p = os.path.join(out, "bus_img%03d_%s.png" % (i, kind))
            assert os.path.exists(p) and cv2.imread(p).shape == (64, 64, 3)
    # the input picture is the AREA-resized sketch: white background, dark strokes
    inp = cv2.imread(os.path.join(out, "bus_img000_input.png"))
    assert inp.max() >= 254 and inp.min() < 140


def test_test_mode_end_to_end(tmp_path):
    """`--mode test` (main_procedure.py:361-492): data/captions/<category>/test.json x data/images/<category>/sketch/<key> ->
    <category>_<stem>_{output,input}.png; class id = index of the category folder, strokes thickened for house / road."""
    import cv2
    import json
    from oracle import fgcolor_oracle as O
    from sketchyscenecolorization_b200 import main_procedure
    from sketchyscenecolorization_b200.config import Config
    from sketchyscenecolorization_b200.pipeline_fg import thicken_drawings
    from sketchyscenecolorization_b200.text_processing import default_vocab_dict, preprocess_sentence
    from sketchyscenecolorization_b200.trainer import FgColorModel
    from torch_ops import TorchOps
    base = tmp_path / "data"
    sketches = {}
    for cate, key, text in (("bus", "228_1.png", "A yellow bus with blue window"), ("road", "7_3.png", "the road is gray")):
        os.makedirs(base / "captions" / cate)
        os.makedirs(base / "images" / cate / "sketch")
        sk = np.full((64, 64, 3), 255, np.uint8)
        sk[20:22, 8:56] = 0
        sk[40:42, 8:56] = 0
        cv2.imwrite(str(base / "images" / cate / "sketch" / key), sk)
        json.dump([dict(key=key, color_text=text)], open(base / "captions" / cate / "test.json", "w"))
        sketches[cate] = (sk, key, text)
    res = str(tmp_path / "test_results")
    Config.set_from_dict(dict(dataset_type="test", batch_size=1, ckpt_dir=str(tmp_path / "snapshot"), results_dir=res,
                              data_format="NCHW", distance_map=0, small_img=1, LSTM_hybrid=1, block_type="MRU", vocab_size=58))
    model = FgColorModel(TorchOps(torch.float64), "cpu", size=16, H=64, W=64, param_dtype=torch.float64, with_discriminator=False)
    model.initialize(seed=1, perturb_tables=0.1)
    noise = torch.zeros(1, 256)
    assert main_procedure.test(model=model, noise=noise, data_base_dir=str(base)) == 2
    gp = {k: v.double() for k, v in model.gstore.state_dict().items()}
    for ci, cate in enumerate(("bus", "road")):                       # sorted folder order: bus -> 0, road -> 1
        sk, key, text = sketches[cate]
        out = cv2.imread(os.path.join(res, "%s_%s_output.png" % (cate, key[:-4])))
        inp = cv2.imread(os.path.join(res, "%s_%s_input.png" % (cate, key[:-4])))
        assert out.shape == (64, 64, 3) and inp.shape == (64, 64, 3)
        want_in = thicken_drawings(sk.astype(np.float32)) if cate == "road" else sk
        assert np.array_equal(inp, want_in)                           # the road's strokes are one pixel thicker
        x = torch.from_numpy(want_in.astype(np.float64) / 255 * 2 - 1).permute(2, 0, 1)[None]
        ids = torch.tensor([preprocess_sentence(text, default_vocab_dict(), 15)])
        ref = O.generator_forward(gp, x, ids, torch.tensor([ci]), noise.double(), 16)
        want = (((ref[0].permute(1, 2, 0).numpy() + 1) / 2) * 255)[:, :, ::-1].astype(np.uint8)
        assert np.abs(out.astype(int) - want.astype(int)).max() <= 1


def test_mapped_reader_and_hardware_crc(tmp_path):
    """read_tfrecord_views (memory-mapped, zero-copy payloads) yields what read_tfrecord yields and rejects the same corruption;
    the SSE4.2 CRC-32C path of fgc_crc32c equals the table walk (FGC_CRC_TABLE=1 forces the latter) on ragged lengths, unaligned
    starts and continued checksums."""
    import subprocess
    import sys
    payloads = [b"", b"x", os.urandom(100003), bytes(range(256)) * 5]
    path = str(tmp_path / "a.tfrecord")
    TI.write_tfrecord(path, payloads)
    assert [bytes(v) for v in TI.read_tfrecord_views(path)] == payloads
    empty = str(tmp_path / "empty.tfrecord")
    open(empty, "wb").close()
    assert list(TI.read_tfrecord_views(empty)) == []
    bad = bytearray(open(path, "rb").read())
    bad[12 + 4 + 13 + 40] ^= 0x10                                         # inside the third record's payload
    open(path, "wb").write(bad)
    with pytest.raises(ValueError):
        list(TI.read_tfrecord_views(path))
    open(path, "wb").write(bytes(bad[:-3]))
    with pytest.raises(ValueError):
        list(TI.read_tfrecord_views(path))
    code = ("import sys, random; sys.path.insert(0, %r)\n"
            "from sketchyscenecolorization_b200.tf_bundle import crc32c\n"
            "random.seed(1); buf = bytes(random.getrandbits(8) for _ in range(70001))\n"
            "print([crc32c(buf[a:b], c) for a, b, c in ((0, 70001, 0), (3, 77, 0), (1, 65536, 12345), (5, 5, 7), (7, 8, 0), (2, 41, 0xFFFFFFFF))])"
            % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    outs = []
    for env in ({}, {"FGC_CRC_TABLE": "1"}):
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, **env), timeout=300)
        assert r.returncode == 0, r.stderr[-1000:]
        outs.append(r.stdout.strip())
    assert outs[0] == outs[1] and outs[0].startswith("[")


def test_data_preparation_writes_what_the_queues_read(tmp_path):
    """data_preparation.py (images + captions -> per-category TFRecords) followed by the ordered queue: every field comes back,
    the class id is the index of the category folder, the caption ids are the reference's."""
    import cv2
    import importlib.util
    import json
    from oracle import input_oracle as IO
    from sketchyscenecolorization_b200.text_processing import default_vocab_dict, preprocess_sentence
    from torch_ops import TorchOps
    spec = importlib.util.spec_from_file_location("data_preparation", os.path.join(os.path.dirname(os.path.dirname(
        os.path.abspath(__file__))), "data_preparation", "data_preparation.py"))
    DP = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(DP)
    base = tmp_path / "data"
    rng = np.random.default_rng(3)
    pics = {}
    for cate, n in (("bus", 2), ("cat", 1)):
        for sub in ("cartoon", "edgemap"):
            os.makedirs(base / "images" / cate / sub)
        os.makedirs(base / "captions" / cate)
        entries = []
        for i in range(n):
            key = "%s_%d.png" % (cate, i)
            img = rng.integers(0, 256, (384, 384, 3), dtype=np.uint8)
            edge = np.full((384, 384, 3), 255, np.uint8)
            edge[100 + i:104 + i, 30:350] = 0
            cv2.imwrite(str(base / "images" / cate / "cartoon" / key), img[:, :, ::-1])
            cv2.imwrite(str(base / "images" / cate / "edgemap" / key), edge)
            entries.append(dict(key=key, color_text="the %s is yellow with blue windows" % cate))
            pics[key] = (img, edge)
        for split in ("train", "val"):
            json.dump(entries, open(base / "captions" / cate / (split + ".json"), "w"))
    written = DP.data_preparation(dataset="both", data_base_dir=str(base), text_len=15)
    assert written == {("train", "bus"): 2, ("train", "cat"): 1, ("val", "bus"): 2, ("val", "cat"): 1}
    assert sorted(os.listdir(base / "tfrecord" / "val")) == ["bus.tfrecord", "cat.tfrecord"]
    batches = list(TI.PairedEvalInput("val", 3, TorchOps(torch.float32), str(base)))
    assert len(batches) == 1
    b = batches[0]
    assert b["image_names"] == ["bus_0.png", "bus_1.png", "cat_0.png"] and b["categories"] == ["bus", "bus", "cat"]
    assert b["cls"].tolist() == [0, 0, 1]
    assert b["text"][2].tolist() == preprocess_sentence("the cat is yellow with blue windows", default_vocab_dict(), 15)
    for i, key in enumerate(b["image_names"]):
        img, edge = pics[key]
        _, want_s = IO.paired_preprocess(img, edge, (192, 192))
        assert np.array_equal(b["sketch"][i].numpy(), want_s)
        want_i, _ = IO.paired_preprocess(img, edge, (192, 192))
        d = b["images"][i].numpy() - want_i
        assert d.min() >= 0 and d.max() <= 2.0 / 256 + 1e-6
