"""CPU oracle for the real-data input path (raw TFRecord payloads -> the tensors the graph is fed).

TEST INFRASTRUCTURE ONLY.  Nothing under ``sketchyscenecolorization_b200/`` may import this module; ``tests/`` use it as
the checker of ``fgc_paired_input`` (csrc/input.cu) and ``tests/torch_ops.py`` uses it to give CPU-only host tests the
same operator.

PARITY UNPINNED: a numpy restatement of Foreground_Instance_Colorization/obj_lib/input_pipeline.get_paired_input
(:72-126), written from that file; the TensorFlow-1 kernels it calls (tf.image.resize_images BILINEAR / AREA, legacy
align_corners=False without half-pixel centres; tf.reduce_min/max; tf.random_uniform) cannot run here.  The reference
holds no test or golden vector for this function.
"""
from __future__ import annotations

import numpy as np

RAW = 384            # "cannot change" (input_pipeline.py:82,87)


def resize_bilinear_tf1(img, out_hw):
    """tf.image.resize_images(BILINEAR), TF-1 legacy kernel: src = dst * (in / out), no half-pixel offset (:104)."""
    H, W = img.shape[:2]
    oh, ow = out_hw
    ys, xs = np.arange(oh) * (H / oh), np.arange(ow) * (W / ow)
    y0, x0 = np.floor(ys).astype(int), np.floor(xs).astype(int)
    y1, x1 = np.minimum(y0 + 1, H - 1), np.minimum(x0 + 1, W - 1)
    fy, fx = (ys - y0)[:, None, None], (xs - x0)[None, :, None]
    top = img[y0][:, x0] * (1 - fx) + img[y0][:, x1] * fx
    bot = img[y1][:, x0] * (1 - fx) + img[y1][:, x1] * fx
    return (top * (1 - fy) + bot * fy).astype(np.float32)


def resize_area(img, out_hw):
    """tf.image.resize_images(AREA) at an integer factor: block mean (:105)."""
    H, W, C = img.shape
    oh, ow = out_hw
    if H % oh or W % ow:
        raise NotImplementedError("AREA resize at a non-integer factor (%dx%d -> %dx%d)" % (H, W, oh, ow))
    return img.astype(np.float32).reshape(oh, H // oh, ow, W // ow, C).mean(axis=(1, 3), dtype=np.float32)


def splitmix_noise(seed, count, offset=0):
    """Values offset .. offset+count-1 of splitmix64(seed), top 24 bits * 2^-32: U[0, 1/256) on a 2^-32 grid.  The CUDA
    kernel draws its dequantisation noise (tf.random_uniform(0, 1/256), :110) from this counter-based stream, indexed by
    the NCHW output position, so the noisy output can be compared bit for bit."""
    with np.errstate(over="ignore"):
        i = np.arange(offset, offset + count, dtype=np.uint64)
        z = np.uint64(seed) + (i + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -32)


def paired_preprocess(cartoon, sketch, out_hw=(192, 192), noise=None):
    """One sample.  cartoon uint8 [R,R,3]; sketch uint8 [R,R,3] or float32 0..255 (the normalised distance map of
    :90-100, which the reference computes on the host through tf.py_func) -> (image, sketch) float32 [3,H,W] in [-1,1]
    (:102-121).  noise: None, or float32 [3,H,W] added after the min-max normalisation (:110)."""
    image = cartoon.astype(np.float32)
    sk = sketch.astype(np.float32)
    if image.shape[0] != out_hw[0] and image.shape[1] != out_hw[1]:           # :103
        image = resize_bilinear_tf1(image, out_hw)
        sk = resize_area(sk, out_hw)
    image = (image - image.min()) / (image.max() - image.min() + np.float32(1))
    image = np.ascontiguousarray(image.transpose(2, 0, 1), dtype=np.float32)
    if noise is not None:
        image = image + noise.astype(np.float32)
    sk = np.ascontiguousarray((sk / np.float32(255)).transpose(2, 0, 1), dtype=np.float32)
    return image * np.float32(2) - np.float32(1), sk * np.float32(2) - np.float32(1)


def paired_input(cartoon, sketch, out_hw, seed=0, dequantize=True):
    """Batch form with the kernel's noise stream: cartoon uint8 [N,R,R,3], sketch uint8 | float32 [N,R,R,3]
    -> images, sketches float32 [N,3,H,W]."""
    N = cartoon.shape[0]
    per = 3 * out_hw[0] * out_hw[1]
    ims, sks = [], []
    for n in range(N):
        noise = splitmix_noise(seed, per, n * per).reshape(3, *out_hw) if dequantize else None
        a, b = paired_preprocess(cartoon[n], sketch[n], out_hw, noise)
        ims.append(a)
        sks.append(b)
    return np.stack(ims), np.stack(sks)
