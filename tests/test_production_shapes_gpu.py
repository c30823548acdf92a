"""The benchmarked configuration, value-checked: bs 64 (and the 2N = 128 discriminator pass), 192 x 192, bf16, the tile
shapes the product routing picks there (`conv_halo_kernel<128,2,2>`, `<256,1,2>`, `<64,4,2>`, `conv_wgrad_halo_kernel<128,3>`,
...).  The small shapes of tests/test_ops_gpu.py never reach these instantiations' steady state (thousands of tiles per
persistent CTA, every mbarrier ring wrapping hundreds of times, 32-bit offsets beyond 2^31 bytes).

Checker: the CUDA-core direct convolution of the same library (`fgc_set_conv_impl(1)`, fp32 FMA on the same bf16 inputs;
itself pinned to fp64 torch in tests/test_ops_gpu.py) -- a CPU oracle would need minutes per case.  Both paths round the
same fp32-accumulated sums to bf16, so outputs agree to a bf16 ulp of the largest entry; weight gradients are fp32 sums of
~2.4 M products each and agree to 2e-3 of the largest entry (summation order).
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from sketchyscenecolorization_b200.ops_base import ACT_LRELU, ACT_NONE  # noqa: E402


@pytest.fixture(scope="module")
def cub():
    from sketchyscenecolorization_b200.cuda_ops import CudaOps
    return CudaOps("cuda:0", torch.bfloat16)


def _rnd(shape, seed, scale=1.0, dtype=torch.bfloat16):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale).to(dtype).contiguous()


def _counts(cu):
    import ctypes
    arr = (ctypes.c_longlong * 6)()
    cu.lib.fgc_debug_conv_counts(arr)
    return list(arr)


def _rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


# (N, H, W, [source channels], k, Cout, act)  -- layers of the bs-64 step
FWD_CASES = [
    (64, 192, 192, [128], 3, 128, ACT_NONE),          # G decoder unit 8 Conv_3 / the roofline kernel: halo<128,2,2>
    (128, 192, 192, [128], 3, 128, ACT_NONE),         # D unit 1 Conv_2 on the 2N real+fake pass
    (64, 192, 192, [128, 3], 3, 128, ACT_LRELU),      # decoder unit 8 reset gate: wide + sketch patch source
    (64, 192, 192, [128, 3], 3, 64, ACT_LRELU),       # decoder unit 8 update gate: 64-wide N tile, MT = 4
    (64, 192, 192, [64], 3, 64, ACT_NONE),            # decoder unit 8 Conv_3
    (128, 96, 96, [256], 3, 256, ACT_NONE),           # D unit 2 Conv_2: halo<256,1,2>
    (64, 96, 96, [128, 3, 8], 3, 128, ACT_LRELU),     # decoder unit 6 gates: three sources
]


@pytest.mark.parametrize("case", FWD_CASES, ids=[str(i) for i in range(len(FWD_CASES))])
def test_forward_and_input_gradient_at_production_shape(cub, case):
    N, H, W, cs, k, cout, act = case
    cin = sum(cs)
    xs = [_rnd((N, H, W, c), 10 + i) for i, c in enumerate(cs)]
    srcs = [(x, False, cub.small_patch(x, k) if x.shape[-1] < 64 else None) for x in xs]
    plain = [(x, False) for x in xs]
    w = _rnd((k, k, cin, cout), 3, 1.0 / math.sqrt(k * k * cin), torch.float32)
    b = _rnd((cout,), 4, 0.3, torch.float32)
    c0 = _counts(cub)
    got = cub.conv_fwd(srcs, w, b, act=act)
    c1 = _counts(cub)
    assert c1[0] - c0[0] == 1, "forward did not take the halo-reuse kernel: %r" % ([a - b_ for a, b_ in zip(c1, c0)],)
    cub.lib.fgc_set_conv_impl(1)
    try:
        want = cub.conv_fwd(plain, w, b, act=act)
    finally:
        cub.lib.fgc_set_conv_impl(0)
    torch.cuda.synchronize()
    assert torch.isfinite(got.float()).all()
    err = _rel(got, want)
    assert err <= 1.0 / 128, "forward: rel-to-max err %.3e" % err          # one bf16 ulp (2^-8) of the largest entry, doubled
    frac = ((got.float() - want.float()).abs() > 2e-3 * want.float().abs().max()).float().mean().item()
    assert frac <= 0.02, "forward: %.2f%% of the outputs differ by more than a rounding" % (100 * frac)
    # input gradient with respect to the first (wide) source
    gy = _rnd((N, H, W, cout), 5)
    wd = _rnd((k, k, cin, cout), 6, 1.0 / math.sqrt(k * k * cout), torch.float32)
    gotd = cub.conv_dgrad(gy, wd, 0, cs[0])
    cub.lib.fgc_set_conv_impl(1)
    try:
        wantd = cub.conv_dgrad(gy, wd, 0, cs[0])
    finally:
        cub.lib.fgc_set_conv_impl(0)
    torch.cuda.synchronize()
    err = _rel(gotd, wantd)
    assert err <= 1.0 / 128, "dgrad: rel-to-max err %.3e" % err


WGRAD_CASES = [
    (64, 192, 192, [128], 3, 128),                    # wgrad_halo<128,3>
    (128, 192, 192, [128], 3, 128),                   # the 2N discriminator pass
    (64, 192, 192, [128, 3], 3, 64),                  # wide + patch source, 64-wide gy
    (64, 192, 192, [64], 3, 64),
    (128, 96, 96, [256], 3, 256),                     # wgrad_halo<256,2>
]


@pytest.mark.parametrize("case", WGRAD_CASES, ids=[str(i) for i in range(len(WGRAD_CASES))])
def test_weight_gradient_at_production_shape(cub, case):
    N, H, W, cs, k, cout = case
    cin = sum(cs)
    xs = [_rnd((N, H, W, c), 20 + i) for i, c in enumerate(cs)]
    srcs = [(x, False, cub.small_patch(x, k) if x.shape[-1] < 64 else None) for x in xs]
    plain = [(x, False) for x in xs]
    gy = _rnd((N, H, W, cout), 7)
    dw, db = torch.zeros(k, k, cin, cout, device="cuda"), torch.zeros(cout, device="cuda")
    dw_ref, db_ref = torch.zeros_like(dw), torch.zeros_like(db)
    cub.conv_wgrad(srcs, gy, dw, db)
    cub.lib.fgc_set_conv_impl(1)
    try:
        cub.conv_wgrad(plain, gy, dw_ref, db_ref)
    finally:
        cub.lib.fgc_set_conv_impl(0)
    torch.cuda.synchronize()
    # fp64 spot check of the checker itself on one tap / one output channel (the small-shape suite pins it in general)
    c = 5
    col = torch.zeros(cs[0], dtype=torch.float64, device="cuda")                                  # tap (kh, kw) = (2, 2)
    for n in range(N):
        col += (xs[0][n, 1:H, 1:W, :].double() * gy[n, 0:H - 1, 0:W - 1, c:c + 1].double()).sum(dim=(0, 1))
    assert ((dw_ref[2, 2, :cs[0], c].double() - col).abs().max() / col.abs().max()).item() <= 2e-3
    assert _rel(dw, dw_ref) <= 2e-3, "dw: rel-to-max err %.3e" % _rel(dw, dw_ref)
    assert _rel(db, db_ref) <= 2e-3, "db: rel-to-max err %.3e" % _rel(db, db_ref)
