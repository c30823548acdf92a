"""oracle/ primitives against known-answer vectors published in TensorFlow's own unit tests (tests/golden/tf_kats.json, each
entry citing the TF test it was transcribed from).  TensorFlow 1.x -- the reference's only arithmetic dependency -- cannot run
in this image, so these are the outputs of that dependency that exist outside this repository: BasicLSTMCell (gate order
i, j, f, o, forget_bias 1, [c, h] state), SAME padding with the odd pixel at the bottom / right, cross-correlation with HWIO
filters, the gradient conventions autograd must reproduce, legacy (no half-pixel) BILINEAR and AREA resizing, the channel order
of depth_to_space / space_to_depth, and the Adam update rule."""
import json
import math
import os

import numpy as np
import torch

from oracle import fgcolor_oracle as O
from oracle import input_oracle as IO

K = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "tf_kats.json")))


def _seq(shape):
    return torch.arange(1, int(np.prod(shape)) + 1, dtype=torch.float64).reshape(shape)


def test_basic_lstm_cell_matches_tf_rnn_cell_test():
    k = K["basic_lstm_cell"]
    x, m = torch.tensor(k["x"], dtype=torch.float64), torch.tensor(k["state"], dtype=torch.float64)
    kernel, bias = torch.full((4, 8), k["kernel_value"], dtype=torch.float64), torch.zeros(8, dtype=torch.float64)
    c0, h0 = O.basic_lstm_cell(x, m[:, 0:2], m[:, 2:4], kernel, bias)          # state_is_tuple=False: [c, h] per layer
    c1, h1 = O.basic_lstm_cell(h0, m[:, 4:6], m[:, 6:8], kernel, bias)
    assert np.allclose(h1.numpy(), k["expected_output"], rtol=0, atol=1e-7)
    assert np.allclose(torch.cat([c0, h0, c1, h1], 1).numpy(), k["expected_state"], rtol=0, atol=1e-7)


def test_conv2d_matches_tf_conv_ops_test():
    for case in K["conv2d"]:
        x, f = _seq(case["in"]), _seq(case["filter"])
        y = O.conv2d(x.permute(0, 3, 1, 2), f, None, stride=case["stride"]).permute(0, 2, 3, 1)      # the oracle is NCHW + SAME
        if case["padding"] == "VALID":      # SAME pads at the bottom / right only here (k <= 2): VALID is its top-left corner
            oh, ow = case["in"][1] - case["filter"][0] + 1, case["in"][2] - case["filter"][1] + 1
            y = y[:, :oh, :ow]
        assert y.flatten().tolist() == case["expected"], case["source"]


def test_conv2d_gradients_match_tf_conv_ops_test():
    k = K["conv2d_backprop"]
    x, f = _seq(k["in"]).requires_grad_(True), _seq(k["filter"]).requires_grad_(True)
    y = O.conv2d(x.permute(0, 3, 1, 2), f, None).permute(0, 2, 3, 1)[:, :k["out"][1], :k["out"][2]]
    y.backward(_seq(k["out"]))
    assert x.grad.flatten().tolist() == k["expected_input_grad"]
    assert f.grad.flatten().tolist() == k["expected_filter_grad"]


def test_resize_matches_tf_image_ops_test():
    for case in K["resize"]:
        img = np.array(case["data"], dtype=np.float64).reshape(case["shape"])
        if "area" in case["expected"]:
            assert IO.resize_area(img, tuple(case["target"])).flatten().tolist() == case["expected"]["area"], case["source"]
            fy = case["shape"][0] // case["target"][0]                 # the generator's sketch pyramid: AREA at an integer factor
            t = torch.tensor(img).permute(2, 0, 1)[None]
            assert fy == case["shape"][1] // case["target"][1]
            assert O.area_resize(t, fy)[0].permute(1, 2, 0).flatten().tolist() == case["expected"]["area"]
            assert O.mean_pool(t)[0].permute(1, 2, 0).flatten().tolist() == case["expected"]["area"]        # mru.mean_pool == AREA / 2
        assert IO.resize_bilinear_tf1(img, tuple(case["target"])).flatten().tolist() == case["expected"]["bilinear"], case["source"]


def _d2s(x, b):
    """tf.depth_to_space, NHWC, restated once here and pinned by TF's vectors below."""
    n, h, w, c = x.shape
    return x.reshape(n, h, w, b, b, c // (b * b)).transpose(0, 1, 3, 2, 4, 5).reshape(n, h * b, w * b, c // (b * b))


def test_depth_to_space_order_and_the_upsample_built_on_it():
    from torch_ops import TorchOps
    ops = TorchOps(torch.float64)
    for case in K["depth_to_space"]:
        x = np.array(case["x"], dtype=np.float64)
        assert _d2s(x, case["block"]).tolist() == np.array(case["expected"], dtype=np.float64).tolist(), case["source"]
        assert ops.depth_to_space(torch.tensor(x)).tolist() == np.array(case["expected"], dtype=np.float64).tolist()
    for case in K["space_to_depth"]:
        x = torch.tensor(np.array(case["x"], dtype=np.float64))
        assert ops.space_to_depth(x).tolist() == np.array(case["expected"], dtype=np.float64).tolist(), case["source"]
    # mru.upsample (mru.py:22-28): concat of four copies on the channel axis + depth_to_space(2) == nearest-neighbour x 2
    x = np.random.default_rng(0).normal(size=(2, 3, 5, 4))
    ref = _d2s(np.concatenate([x, x, x, x], axis=3), 2)
    got = O.upsample(torch.tensor(x).permute(0, 3, 1, 2)).permute(0, 2, 3, 1).numpy()
    assert np.array_equal(ref, got)


def test_adam_update_matches_tf_adam_test():
    k = K["adam"]

    def adam_update_numpy(param, g_t, t, m, v, alpha, beta1, beta2, epsilon):      # adam_test.py, verbatim rule
        alpha_t = alpha * np.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
        m_t = beta1 * m + (1 - beta1) * g_t
        v_t = beta2 * v + (1 - beta2) * g_t * g_t
        return param - alpha_t * m_t / (np.sqrt(v_t) + epsilon), m_t, v_t

    for var, grad in zip(k["var"], k["grad"]):
        p_np, g_np = np.array(var), np.array(grad)
        m, v = np.zeros(2), np.zeros(2)
        p_o, v_o = torch.tensor(var, dtype=torch.float64), torch.zeros(2, dtype=torch.float64)
        for t in range(1, k["steps"] + 1):      # the reference's optimiser: beta1 = 0, beta2 = 0.9 (graph_single.py:588)
            p_np, m, v = adam_update_numpy(p_np, g_np, t, m, v, 2e-4, 0.0, 0.9, 1e-8)
            p_o, v_o = O.adam_update(p_o, torch.tensor(g_np), v_o, 2e-4, t)
            assert np.allclose(p_o.numpy(), p_np, rtol=1e-14, atol=0) and np.allclose(v_o.numpy(), v, rtol=1e-14, atol=0)
    assert math.isclose(O.lr_decay(50000, 100000), 0.55) and O.lr_decay(10 ** 9, 100000) == 0.2       # graph_single.py:139
