"""GPU parity of the real-data input path: fgc_paired_input (csrc/input.cu, through the C-ABI via CudaOps.paired_input) against
oracle/input_oracle.py -- the numpy restatement of input_pipeline.get_paired_input (:72-126) -- bit for bit, dequantisation
noise included (both draw it from the same counter-based splitmix64 stream), and the TFRecord training queue end to end on
the device."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cu():
    from sketchyscenecolorization_b200.cuda_ops import CudaOps
    return CudaOps(torch.device("cuda:0"), torch.float32)


def _raw(N, R, seed, strokes=True):
    rng = np.random.default_rng(seed)
    cartoon = rng.integers(3, 251, (N, R, R, 3), dtype=np.uint8)          # min / max differ per sample below
    for n in range(N):
        cartoon[n] = np.clip(cartoon[n].astype(np.int32) // (n + 1) + 7 * n, 0, 255).astype(np.uint8)
    sketch = np.full((N, R, R, 3), 255, np.uint8)
    if strokes:
        for n in range(N):
            r0, c0 = R // 6 + n, R // 4 + 3 * n
            sketch[n, r0:r0 + 3, R // 10:R - R // 8] = 0
            sketch[n, R // 8:R - R // 8, c0:c0 + 3] = rng.integers(0, 120, (R - 2 * (R // 8), 3, 3), dtype=np.uint8)
    return cartoon, sketch


@pytest.mark.parametrize("dequantize", [False, True], ids=["plain", "noise"])
@pytest.mark.parametrize("case", [(3, 384, 192), (2, 384, 64), (2, 96, 48), (1, 40, 20), (2, 36, 18), (2, 36, 12)],
                         ids=["384to192", "384to64", "96to48", "40to20", "36to18-unaligned-rows", "36to12"])
def test_paired_input_matches_oracle(cu, case, dequantize):
    from oracle import input_oracle as IO
    N, R, O = case
    cartoon, sketch = _raw(N, R, seed=R + O)
    want_i, want_s = IO.paired_input(cartoon, sketch, (O, O), seed=987654321, dequantize=dequantize)
    got_i, got_s = cu.paired_input(torch.from_numpy(cartoon).cuda(), torch.from_numpy(sketch).cuda(), (O, O), seed=987654321,
                                   dequantize=dequantize)
    torch.cuda.synchronize()
    assert got_i.shape == (N, 3, O, O) and got_s.shape == (N, 3, O, O) and got_i.dtype == torch.float32
    gi, gs = got_i.cpu().numpy(), got_s.cpu().numpy()
    assert np.array_equal(gi, want_i), "image: max abs diff %.3e" % np.abs(gi - want_i).max()
    # block sums of uint8 are exact in fp32 in any order; the mean's division is one correctly rounded op on both sides
    assert np.array_equal(gs, want_s), "sketch: max abs diff %.3e" % np.abs(gs - want_s).max()


def test_paired_input_distance_map_and_errors(cu):
    """fp32 sketches (the 0..255 distance map the host computes with scipy, as the reference does through tf.py_func)."""
    from oracle import input_oracle as IO
    from sketchyscenecolorization_b200 import tfrecord_input as TI
    from sketchyscenecolorization_b200._lib import FgcError
    cartoon, sketch = _raw(2, 384, seed=5)
    dm = np.stack([TI.distance_map_255(s) for s in sketch])
    want_i, want_s = IO.paired_input(cartoon, dm, (192, 192), dequantize=False)
    got_i, got_s = cu.paired_input(torch.from_numpy(cartoon).cuda(), torch.from_numpy(dm).cuda(), (192, 192), dequantize=False)
    assert np.array_equal(got_i.cpu().numpy(), want_i)
    assert np.abs(got_s.cpu().numpy() - want_s).max() <= 1e-6            # fp32 block sums: summation order differs
    with pytest.raises(FgcError):                                        # AREA at a non-integer factor: refused, as in the oracle
        cu.paired_input(torch.from_numpy(cartoon).cuda(), torch.from_numpy(sketch).cuda(), (160, 160))
    with pytest.raises(NotImplementedError):
        IO.paired_preprocess(cartoon[0], sketch[0], (160, 160))


def test_paired_input_full_batch_properties(cu):
    """BASELINE size (bs 64, 384 -> 192): size-independent properties instead of the oracle."""
    N = 64
    g = torch.Generator(device="cuda").manual_seed(3)
    cartoon = torch.randint(0, 256, (N, 384, 384, 3), dtype=torch.uint8, device="cuda", generator=g)
    sketch = torch.randint(0, 256, (N, 384, 384, 3), dtype=torch.uint8, device="cuda", generator=g)
    plain_i, plain_s = cu.paired_input(cartoon, sketch, (192, 192), dequantize=False)
    noisy_i, noisy_s = cu.paired_input(cartoon, sketch, (192, 192), seed=11, dequantize=True)
    again_i, _ = cu.paired_input(cartoon, sketch, (192, 192), seed=11, dequantize=True)
    other_i, _ = cu.paired_input(cartoon, sketch, (192, 192), seed=12, dequantize=True)
    pick = cartoon[:, ::2, ::2].permute(0, 3, 1, 2).float()
    mn, mx = pick.amin(dim=(1, 2, 3), keepdim=True), pick.amax(dim=(1, 2, 3), keepdim=True)
    assert torch.equal(plain_i, ((pick - mn) / (mx - mn + 1)) * 2 - 1)                     # every op above is one rounding
    assert torch.equal(plain_i.amin(dim=(1, 2, 3)), torch.full((N,), -1.0, device="cuda"))  # the minimum maps to -1 exactly
    assert float(plain_i.max()) < 1.0
    area = sketch.float().reshape(N, 192, 2, 192, 2, 3).sum(dim=(2, 4)).permute(0, 3, 1, 2) * 0.25
    # (a tensor divisor: torch turns division by a Python scalar into a multiplication by its reciprocal on CUDA)
    assert torch.equal(plain_s, area / torch.full_like(area, 255.0) * 2 - 1) and torch.equal(noisy_s, plain_s)
    d = noisy_i - plain_i
    assert float(d.min()) >= 0.0 and float(d.max()) <= 2.0 / 256 + 1e-7 and abs(float(d.mean()) - 1.0 / 256) < 2e-5
    assert torch.equal(noisy_i, again_i) and not torch.equal(noisy_i, other_i)              # counter based: seed -> stream


def test_train_queue_on_device(cu, tmp_path):
    """TFRecord files -> PairedTrainInput with the CUDA operator set: batches arrive on the device and equal the oracle's
    treatment of the same raw samples.  (Green on a B200 with the first version of the queue, profiles/r1t; the rewritten host
    side is checked on the CPU in tests/test_tfrecord_cpu.py.)"""
    from oracle import input_oracle as IO
    from sketchyscenecolorization_b200 import tfrecord_input as TI
    d = os.path.join(str(tmp_path), "tfrecord", "train")
    os.makedirs(d)
    cartoon, sketch = _raw(6, 384, seed=9)
    recs = [TI.encode_example(dict(ImageName=("img%03d.png" % i).encode(), cartoon_data=cartoon[i].tobytes(),
                                   sketch_data=sketch[i].tobytes(), Category=b"car", Category_id=4, Color_text=b"the car is red",
                                   Text_vocab_indices=bytes([0] * 11 + [24, 5, 6, 30]))) for i in range(6)]
    TI.write_tfrecord(os.path.join(d, "car.tfrecord"), recs)
    q = TI.PairedTrainInput(4, cu, str(tmp_path), min_after_dequeue=2, seed=1, num_threads=2, prefetch=2)
    b = next(q)
    assert b["images"].is_cuda and b["sketch"].is_cuda and b["images"].shape == (4, 3, 192, 192) and b["images_d"] is b["images"]
    idx = [int(name[3:6]) for name in b["image_names"]]
    plain_i, plain_s = IO.paired_input(cartoon[idx], sketch[idx], (192, 192), dequantize=False)
    delta = b["images"].cpu().numpy() - plain_i
    assert delta.min() >= 0 and delta.max() <= 2.0 / 256 + 1e-7
    assert np.array_equal(b["sketch"].cpu().numpy(), plain_s)
    assert b["cls"].tolist() == [4] * 4 and b["text"].shape == (4, 15)
