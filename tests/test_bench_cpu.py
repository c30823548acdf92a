"""The driver's contract for bench.py, as far as it can be checked without a GPU: the reference arm prints ONE JSON line with the
agreed keys, and the product arm refuses to run without a CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["metric"].startswith("fg-colorization train images/sec") and d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "sample" in d["config"]


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_synthetic_record_files_feed_the_training_queue(tmp_path):
    """`bench.py --input tfrecord` writes its dataset in the reference's on-disk format; the queue reads it back."""
    import sys as _sys
    import torch
    _sys.path.insert(0, ROOT)
    import bench
    from sketchyscenecolorization_b200 import tfrecord_input as TI
    from torch_ops import TorchOps
    base = bench.write_synthetic_tfrecords(str(tmp_path), n_records=8, files=2)
    assert sorted(os.listdir(os.path.join(base, "tfrecord", "train"))) == ["part0.tfrecord", "part1.tfrecord"]
    q = TI.PairedTrainInput(4, TorchOps(torch.float32), base, min_after_dequeue=4, seed=1, num_threads=2, prefetch=2)
    b = next(q)
    q.close()
    assert b["images"].shape == (4, 3, 192, 192) and b["text"].shape == (4, 15) and int(b["text"].min()) >= 2
    assert float(b["sketch"].min()) == -1.0 and float(b["sketch"].max()) == 1.0 and 0 <= int(b["cls"].min()) and int(b["cls"].max()) < 25
    assert bench.block_type_flops("MRU") == (bench.GF_GFLOP, bench.DF_GFLOP) and bench.block_type_flops("Residual") is None
    gf, df = bench.block_type_flops("Pix2Pix")
    assert 6.0 < gf < 6.5 and 3.4 < df < 3.7
