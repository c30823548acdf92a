"""GPU parity of the background-colorization generator (BASELINE.json configs[3]) against oracle/bg_oracle.py.

Green on a B200 since round 2 (profiles/r2a_bg_gpu_tests.log).  The operators the
network is made of have run on a B200 (tests/test_ops_gpu.py, tests/test_pix2pix_gpu.py), the host code is checked against the
oracle on the CPU (tests/test_bg_cpu.py).  The published size -- 768 x 768, ngf 64, batch 1 -- is compared against the
oracle on the host in fp64, with the fp32 run beside it as the yardstick (seconds of CPU work; at this size the two differ
by 1.25e-3)."""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu]

INFER_TOL = 1e-3


@pytest.mark.parametrize("cfg", [(8, 96, 2), (64, 768, 1)], ids=["ngf8_96px_n2", "ngf64_768px_n1_cfg3"])
def test_generator_inference_parity(cfg):
    """~110 batch-normalised layers amplify rounding by three to four orders of magnitude at random initialisation: the oracle
    run in fp32 instead of fp64 on the CPU already moves the picture by about 1e-3 (measured at several sizes).  The 1e-3 bar of
    the MRU path is therefore at the noise floor of an fp32 reference here; bounds are stated against that yardstick -- fp32
    CUDA-core convolutions within 30 yardsticks, bf16x3 tensor-core convolutions within 100 (tests/test_residual_gpu.py)."""
    from oracle import bg_oracle as B
    from sketchyscenecolorization_b200.bg import BgColorModel
    from sketchyscenecolorization_b200.cuda_ops import CudaOps
    ngf, S, N = cfg
    m = BgColorModel(CudaOps("cuda:0", torch.float32), "cuda:0", ngf=ngf, vocab_size=18)
    m.initialize(seed=3, perturb_tables=0.1)
    gp = {k: v.detach().cpu().double() for k, v in m.gstore.state_dict().items()}
    g = torch.Generator().manual_seed(7)
    img = torch.rand(N, 3, S, S, generator=g, dtype=torch.float64) * 2 - 1
    ids = torch.randint(2, 18, (N, 8), generator=g)
    ids[0, :3] = 0
    with torch.no_grad():
        ref_out, ref_reg = B.generator_forward(gp, img, ids)
        ref32, _ = B.generator_forward({k: v.float() for k, v in gp.items()}, img.float(), ids)
    yard = (ref32.double() - ref_out).abs().max().item()
    x = img.float().permute(0, 2, 3, 1).contiguous()
    try:
        for impl, bound, tag in ((1, max(INFER_TOL, 30 * yard), "fp32 CUDA-core convolutions"),
                                 (0, max(INFER_TOL, 100 * yard), "bf16x3 tensor-core convolutions")):
            m.ops.lib.fgc_set_conv_impl(impl)
            out, reg = m.generate(x, ids.numpy())
            torch.cuda.synchronize()
            assert out.shape == (N, S, S, 3) and torch.isfinite(out).all() and torch.isfinite(reg).all()
            err = (out.cpu().double().permute(0, 3, 1, 2) - ref_out).abs().max().item()
            print("background generator, %s: max-abs err %.3e (fp32-oracle yardstick %.3e, bound %.3e)" % (tag, err, yard, bound))
            assert err <= bound, "background generator, %s: max-abs err %.3e > %.3e (yardstick %.3e)" % (tag, err, bound, yard)
            rscale = max(ref_reg.abs().max().item(), 1.0)
            assert (reg.cpu().double().permute(0, 3, 1, 2) - ref_reg).abs().max().item() <= bound * rscale
    finally:
        m.ops.lib.fgc_set_conv_impl(0)
    # CudaOps(conv_terms=3), the six-product split: measured on a B200 it does not move this network's error (1.1e-2 -> 1.3e-2,
    # 3.1e-2 -> 1.4e-2; profiles/r2g_sixproduct_gpu_tests.log) -- operand rounding is not the floor here, see
    # tests/test_ops_gpu.py::test_six_product_convolution.  Bound: the bf16x3 one.
    m6 = BgColorModel(CudaOps("cuda:0", torch.float32, conv_terms=3), "cuda:0", ngf=ngf, vocab_size=18)
    m6.gstore.load_state_dict(m.gstore.state_dict())
    out, reg = m6.generate(x, ids.numpy())
    torch.cuda.synchronize()
    err = (out.cpu().double().permute(0, 3, 1, 2) - ref_out).abs().max().item()
    bound = max(INFER_TOL, 100 * yard)
    print("background generator, six-product tensor-core convolutions: max-abs err %.3e (fp32-oracle yardstick %.3e, bound %.3e)"
          % (err, yard, bound))
    assert torch.isfinite(out).all() and err <= bound
