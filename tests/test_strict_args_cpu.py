"""Every network variant under an operator set that enforces what the ctypes layer (cuda_ops.CudaOps) asserts about its
arguments -- contiguous buffers, fp32 parameters, one dtype per convolution, int32 labels -- in both storage modes (fp32 and
bf16 activations).  Catches host-code slips (a strided view handed to a kernel, a bf16 weight) on the CPU, before GPU time."""
import functools

import numpy as np
import pytest
import torch

from oracle import fgcolor_oracle as O
from sketchyscenecolorization_b200.bg import BgColorModel
from sketchyscenecolorization_b200.trainer import FgColorModel
from torch_ops import TorchOps


class StrictOps(TorchOps):
    def __getattribute__(self, name):
        a = object.__getattribute__(self, name)
        if name.startswith('_') or not callable(a):
            return a

        @functools.wraps(a)
        def wrapped(*args, **kw):
            def chk(t, path):
                if torch.is_tensor(t):
                    assert t.is_contiguous(), "%s: non-contiguous tensor at %s, shape %s" % (name, path, tuple(t.shape))
                elif isinstance(t, (list, tuple)):
                    for i, e in enumerate(t):
                        chk(e, "%s[%d]" % (path, i))
            chk(args, "args")
            chk(list(kw.values()), "kw")
            if name in ("conv_fwd", "conv_wgrad"):
                assert len({e[0].dtype for e in args[0]}) == 1, "%s: mixed source dtypes" % name
                assert all(e[0].dim() == 4 for e in args[0])
            if name == "conv_fwd":
                assert args[1].dtype == torch.float32 and (args[2] is None or args[2].dtype == torch.float32)
            if name == "conv_wgrad":
                assert args[2].dtype == torch.float32 and args[1].dtype == args[0][0][0].dtype
            if name == "conv_dgrad":
                assert args[1].dtype == torch.float32
            if name == "cbn_act_fwd":
                assert args[3].dtype == torch.float32 and args[4].dtype == torch.float32 and args[5].dtype == torch.int32
            if name == "cbn_act_bwd":
                assert args[0].dtype == args[1].dtype
            if name == "prelu_fwd":
                assert args[1].dtype == torch.float32
            return a(*args, **kw)
        return wrapped


@pytest.mark.parametrize("act", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("bt,size", [("MRU", 16), ("Pix2Pix", 16), ("Residual", 8)])
def test_fg_variants_hand_kernels_well_formed_arguments(bt, size, act):
    m = FgColorModel(StrictOps(act), "cpu", size=size, H=64, W=64, block_type=bt)
    m.initialize(seed=1)
    b = O.make_batch(2, 64, 64, 5, torch.float32, n_pad=3)
    bb = dict(b)
    bb["cls"], bb["cls_d"], bb["text"] = b["cls"].int(), b["cls_d"].int(), b["text"].numpy()
    rd, rg = m.d_step_grads(bb), m.g_step_grads(bb)
    out = m.generate(bb["sketch"], bb["text"], bb["cls"], bb["noise"])
    assert torch.isfinite(rd["loss"]) and torch.isfinite(rg["loss"]) and out.dtype == torch.float32 and out.shape == (2, 3, 64, 64)


@pytest.mark.parametrize("act", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_bg_generator_hands_kernels_well_formed_arguments(act):
    m = BgColorModel(StrictOps(act), "cpu", ngf=4, vocab_size=18)
    m.initialize(seed=1)
    out, reg = m.generate(torch.rand(1, 64, 64, 3) * 2 - 1, np.array([[0, 0, 2, 3, 4, 5, 8, 7]], dtype=np.int32))
    assert out.shape == (1, 64, 64, 3) and reg.shape == (1, 64, 64, 3) and torch.isfinite(out).all()
