"""GPU parity of every libfgcolor op against the plain-torch op of the same name (tests/torch_ops.py, fp64).

All calls go through the C-ABI (ctypes) via CudaOps.  Convolutions are checked on both implementations: the
tcgen05 tensor-core path (the product) and the CUDA-core checker, on shapes that exercise every producer
path (big / small sources, concat, upsampled source, stride 2, 7x7, 1x1, FC rows, ragged M and N tiles).
"""
import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from sketchyscenecolorization_b200.ops_base import ACT_LRELU, ACT_MIU, ACT_NONE, ACT_TANH  # noqa: E402


@pytest.fixture(scope="module")
def env():
    from sketchyscenecolorization_b200.cuda_ops import CudaOps
    from torch_ops import TorchOps
    dev = torch.device("cuda:0")
    return dict(cu=CudaOps(dev, torch.float32), cub=CudaOps(dev, torch.bfloat16), ref=TorchOps(torch.float64, dev), dev=dev)


def rnd(shape, seed, dev, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g, dtype=torch.float64) * scale).to(dev)


def close(a, b, tol, what=""):
    a = a.double()
    b = b.double()
    scale = max(b.abs().max().item(), 1e-30)
    err = (a - b).abs().max().item() / scale
    assert err <= tol, "%s: rel-to-max err %.3e > %.1e" % (what, err, tol)


def _set_impl(cu, impl):
    """1: CUDA-core checker; 0: product routing (halo-reuse / direct narrow / gather kernels); 2: tensor-core gather
    kernels only (halo-reuse and direct narrow kernels switched off)."""
    cu.lib.fgc_set_conv_impl(1 if impl == 1 else 0)
    cu.lib.fgc_set_conv_flags(0 if impl == 2 else 1, 0 if impl == 2 else 1)


def conv_counts(cu):
    import ctypes
    arr = (ctypes.c_longlong * 6)()
    cu.lib.fgc_debug_conv_counts(arr)
    return list(arr)


CONV_CASES = [
    # (N, H, W, [(C, ups), ...], k, stride, Cout, act)
    (2, 16, 16, [(64, False)], 3, 1, 64, ACT_NONE),
    (3, 12, 12, [(128, False)], 3, 1, 96, ACT_LRELU),           # ragged M (432) and N (96 -> 128 tile)
    (2, 16, 16, [(64, True), (3, False), (8, False)], 3, 1, 128, ACT_LRELU),   # decoder-style concat + upsample
    (2, 24, 24, [(8, False), (3, False)], 3, 1, 8, ACT_LRELU),  # stem-level update gate 11 -> 8
    (2, 32, 32, [(3, False)], 7, 2, 8, ACT_NONE),               # generator stem 7x7 s2, asymmetric SAME pad
    (2, 16, 16, [(64, False)], 7, 1, 3, ACT_TANH),              # head 7x7 -> 3 + tanh
    (2, 12, 12, [(192, False)], 1, 1, 64, ACT_NONE),            # 1x1 projection
    (70, 1, 1, [(256, False)], 1, 1, 200, ACT_MIU),             # fully connected rows
    (5, 1, 1, [(512, False), (512, False)], 1, 1, 2048, ACT_NONE),   # LSTM gates
    (2, 6, 6, [(96, False)], 3, 1, 1, ACT_NONE),                # patch logits (Cout 1), partial K slab
    (64, 1, 1, [(100, False), (37, False)], 1, 1, 77, ACT_LRELU),   # skinny-product kernel: two sources, ragged K and N
    # halo-reuse kernel (bf16, stride 1, wide sources through tensor-map TMA)
    (8, 64, 80, [(64, False)], 3, 1, 128, ACT_LRELU),           # two sub-tiles per CTA (32 x 8 pixel tiles)
    (2, 28, 24, [(128, False), (3, False)], 3, 1, 64, ACT_NONE),    # wide + sketch source, ragged tile rows (28 = 16 + 12)
    (8, 64, 80, [(64, False), (8, False), (3, False)], 3, 1, 256, ACT_NONE),   # 256-wide N tile, two small sources
    (2, 32, 40, [(192, False)], 7, 1, 3, ACT_TANH),             # 7x7 over three channel groups, Cout 3
    (4, 24, 24, [(256, False)], 3, 1, 256, ACT_NONE),           # halo-reuse wgrad: 256-wide N tile, boxes shared across CTAs
    (2, 16, 24, [(128, False)], 3, 1, 96, ACT_NONE),            # halo-reuse wgrad: ragged N tile
    (2, 16, 16, [(64, False)], 5, 1, 64, ACT_NONE),             # 5x5: 12-row boxes
    (3, 30, 22, [(72, False)], 3, 1, 40, ACT_MIU),              # partial channel group (72 = 64 + 8), ragged rows and columns
]


def _conv_inputs(case, dev, seed=0):
    N, H, W, srcs, k, stride, cout, act = case
    xs = []
    for i, (c, ups) in enumerate(srcs):
        h, w = (H // 2, W // 2) if ups else (H, W)
        xs.append((rnd((N, h, w, c), seed + i, dev), ups))
    cin = sum(c for c, _ in srcs)
    w_ = rnd((k, k, cin, cout), seed + 10, dev, 1.0 / math.sqrt(k * k * cin))
    b = rnd((cout,), seed + 11, dev, 0.3)
    return xs, w_, b


@pytest.mark.parametrize("impl", [1, 0, 2], ids=["simple", "tcgen05", "tcgen05-gather"])
@pytest.mark.parametrize("case", CONV_CASES, ids=[str(i) for i in range(len(CONV_CASES))])
def test_conv_fwd(env, case, impl):
    cu, ref, dev = env["cu"], env["ref"], env["dev"]
    _set_impl(cu, impl)
    try:
        N, H, W, srcs, k, stride, cout, act = case
        xs, w, b = _conv_inputs(case, dev)
        want = ref.conv_fwd(xs, w, b, stride=stride, act=act)
        got = cu.conv_fwd([(x.float().contiguous(), u) for x, u in xs], w.float().contiguous(), b.float().contiguous(),
                          stride=stride, act=act)
        torch.cuda.synchronize()
        close(got, want, 1e-4 if impl != 1 else 2e-5, "conv_fwd fp32")
        gotb = env["cub"].conv_fwd([(x.bfloat16().contiguous(), u) for x, u in xs], w.float().contiguous(),
                                   b.float().contiguous(), stride=stride, act=act)
        torch.cuda.synchronize()
        assert gotb.dtype == torch.bfloat16
        close(gotb, want, 3e-2, "conv_fwd bf16")
    finally:
        _set_impl(cu, 0)


@pytest.mark.parametrize("case", [
    (4, 16, 16, 128, 128, 0, 128),      # the discriminator's Conv_2 form: 128 -> 128, exchanged-role epilogue
    (2, 32, 24, 64, 96, 0, 64),         # 64-wide tiles, channel slice at the start of a wider filter
    (2, 16, 16, 256, 256, 0, 256),      # 256-wide N tile
    (3, 16, 8, 72, 40, 8, 24),          # ragged channel group (72 = 64 + 8), ragged N, slice in the middle
    (2, 12, 12, 64, 64, 0, 64),         # width not a multiple of 8: falls back to the full-resolution form
], ids=["128x128", "64-slice", "256", "ragged", "fallback"])
def test_conv_dgrad_pooled_gy(env, case):
    """Input gradient of a 3x3 layer under the 2x2 mean pool from the LOW-resolution output gradient: four phase launches of
    2x2-tap convolutions (fgc_conv2d_fwd_phase) against conv_dgrad of the upsampled gradient."""
    cub, cu, ref, dev = env["cub"], env["cu"], env["ref"], env["dev"]
    N, h, w_, cout, cin_total, c_off, c_len = case
    g_low = rnd((N, h, w_, cout), 1, dev)
    w = rnd((3, 3, cin_total, cout), 2, dev, 1.0 / math.sqrt(9 * cout))
    want = ref.conv_dgrad(ref.unpool_bwd(g_low), w, c_off, c_len)
    before = cub.launch_count()
    cub.phase_dgrad = True
    got = cub.conv_dgrad_pooled_gy(g_low.bfloat16().contiguous(), w.float().contiguous(), c_off, c_len)
    torch.cuda.synchronize()
    assert got.shape == want.shape and got.dtype == torch.bfloat16
    close(got, want, 3e-2, "conv_dgrad_pooled_gy bf16")
    cub.phase_dgrad = False
    try:
        plain = cub.conv_dgrad_pooled_gy(g_low.bfloat16().contiguous(), w.float().contiguous(), c_off, c_len)
    finally:
        cub.phase_dgrad = os.environ.get("FGC_PHASE_DGRAD", "0") == "1"
    close(got, plain.double(), 2e-2, "phase form against the full-resolution form")
    got32 = cu.conv_dgrad_pooled_gy(g_low.float().contiguous(), w.float().contiguous(), c_off, c_len)      # fp32: full-resolution form
    close(got32, want, 1e-4, "conv_dgrad_pooled_gy fp32")
    assert cub.launch_count() > before


@pytest.mark.parametrize("acc", [False, True], ids=["write", "acc"])
@pytest.mark.parametrize("case", [CONV_CASES[i] for i in (6, 12)] + [(4, 12, 12, [(8, False)], 1, 1, 128, ACT_NONE)],
                         ids=["1x1", "halo", "skip-low-res"])
def test_conv_fwd_into(env, case, acc):
    """conv_fwd(out=, acc=): the pooled 1x1 skip of the encoder cell accumulates into mean_pool(h2) (blocks.enc_block_fwd)."""
    cu, ref, dev = env["cu"], env["ref"], env["dev"]
    N, H, W, srcs, k, stride, cout, act = case
    xs, w, b = _conv_inputs(case, dev)
    init = rnd((N, -(-H // stride), -(-W // stride), cout), 40, dev)
    want = ref.conv_fwd(xs, w, b, stride=stride, act=act, out=init.clone(), acc=acc)
    out = init.float().contiguous()
    got = cu.conv_fwd([(x.float().contiguous(), u) for x, u in xs], w.float().contiguous(), b.float().contiguous(),
                      stride=stride, act=act, out=out, acc=acc)
    torch.cuda.synchronize()
    assert got is out
    close(got, want, 1e-4, "conv_fwd(out=) fp32")
    outb = init.bfloat16().contiguous()
    env["cub"].conv_fwd([(x.bfloat16().contiguous(), u) for x, u in xs], w.float().contiguous(), b.float().contiguous(),
                        stride=stride, act=act, out=outb, acc=acc)
    torch.cuda.synchronize()
    close(outb, want, 3e-2, "conv_fwd(out=) bf16")


DGRAD_CASES = [
    # (N, H, W, Cin_total, c_off, c_len, k, Cout, ups, acc)
    (2, 16, 16, 64, 0, 64, 3, 64, False, False),
    (2, 12, 12, 131, 128, 3, 3, 128, False, True),      # gradient to the 3-channel image slice, accumulating
    (2, 16, 16, 75, 0, 64, 3, 96, True, False),         # upsampled hidden state: 2x2-summed low-res gradient
    (2, 16, 16, 75, 67, 8, 3, 96, False, False),
    (3, 8, 8, 64, 0, 64, 7, 3, False, False),           # head 7x7 (Cout 3)
    (2, 12, 12, 192, 0, 192, 1, 64, False, False),
    (70, 1, 1, 1024, 512, 512, 1, 2048, False, False),  # LSTM kernel slice
    (2, 6, 6, 96, 0, 96, 1, 1, False, True),            # from the 1-channel patch logits
    (8, 64, 80, 72, 64, 8, 3, 128, False, False),       # halo kernel, mirrored taps, narrow output (N tile 16)
    (2, 16, 24, 64, 0, 64, 7, 64, False, True),         # halo kernel, 7x7 mirrored taps, accumulating
    (2, 20, 20, 11, 8, 3, 3, 8, False, True),           # direct narrow kernel: 8-channel gy -> image slice, accumulating
    (2, 20, 20, 3, 0, 3, 7, 8, False, False),           # direct narrow kernel: stem 7x7, mirrored taps
    (2, 20, 20, 11, 0, 8, 3, 8, False, False),          # direct narrow kernel: 8 -> 8
    (8, 64, 80, 139, 0, 128, 3, 128, True, True),       # upsampled source, 2x2 sums folded into the halo epilogue, accumulating
    (48, 1, 1, 1024, 512, 512, 1, 2048, False, False),  # skinny-product kernel (<= 64 rows): LSTM kernel slice, float4 weights
    (33, 1, 1, 200, 7, 90, 1, 130, False, True),        # skinny-product kernel: ragged K / N, unaligned slice, accumulating
    (2, 16, 24, 139, 131, 8, 3, 128, False, True),      # column-folded narrow gradient (bf16), slice at the end, accumulating
    (2, 24, 16, 8, 0, 8, 3, 128, False, False),         # column-folded narrow gradient: the discriminator's unit-1 Conv_1
]


@pytest.mark.parametrize("impl", [1, 0, 2], ids=["simple", "tcgen05", "tcgen05-gather"])
@pytest.mark.parametrize("case", DGRAD_CASES, ids=[str(i) for i in range(len(DGRAD_CASES))])
def test_conv_dgrad(env, case, impl):
    cu, ref, dev = env["cu"], env["ref"], env["dev"]
    _set_impl(cu, impl)
    try:
        N, H, W, cin, c_off, c_len, k, cout, ups, acc = case
        gy = rnd((N, H, W, cout), 1, dev)
        w = rnd((k, k, cin, cout), 2, dev, 1.0 / math.sqrt(k * k * cout))
        oshape = (N, H // 2, W // 2, c_len) if ups else (N, H, W, c_len)
        init = rnd(oshape, 3, dev)
        want = ref.conv_dgrad(gy, w, c_off, c_len, ups=ups, out=init.clone() if acc else None, acc=acc)
        out = init.float().contiguous() if acc else None
        got = cu.conv_dgrad(gy.float().contiguous(), w.float().contiguous(), c_off, c_len, ups=ups, out=out, acc=acc)
        torch.cuda.synchronize()
        close(got, want, 2e-5 if impl == 1 else 1e-4, "conv_dgrad fp32")
        outb = init.bfloat16().contiguous() if acc else None
        gotb = env["cub"].conv_dgrad(gy.bfloat16().contiguous(), w.float().contiguous(), c_off, c_len, ups=ups, out=outb, acc=acc)
        torch.cuda.synchronize()
        close(gotb, want, 3e-2, "conv_dgrad bf16")
    finally:
        _set_impl(cu, 0)


@pytest.mark.parametrize("impl", [1, 0, 2], ids=["simple", "tcgen05", "tcgen05-gather"])
@pytest.mark.parametrize("case", CONV_CASES, ids=[str(i) for i in range(len(CONV_CASES))])
def test_conv_wgrad(env, case, impl):
    cu, ref, dev = env["cu"], env["ref"], env["dev"]
    _set_impl(cu, impl)
    try:
        N, H, W, srcs, k, stride, cout, act = case
        xs, w, b = _conv_inputs(case, dev, seed=20)
        OH, OW = -(-H // stride), -(-W // stride)
        gy = rnd((N, OH, OW, cout), 31, dev)
        dw0, db0 = rnd(w.shape, 32, dev), rnd(b.shape, 33, dev)
        dw_ref, db_ref = dw0.clone(), db0.clone()
        ref.conv_wgrad(xs, gy, dw_ref, db_ref, stride=stride)
        dw, db = dw0.float().contiguous(), db0.float().contiguous()
        cu.conv_wgrad([(x.float().contiguous(), u) for x, u in xs], gy.float().contiguous(), dw, db, stride=stride)
        torch.cuda.synchronize()
        close(dw, dw_ref, 2e-5 if impl == 1 else 1e-4, "conv_wgrad dw fp32")
        close(db, db_ref, 2e-5, "conv_wgrad db fp32")
        dwb, dbb = dw0.float().contiguous(), db0.float().contiguous()
        env["cub"].conv_wgrad([(x.bfloat16().contiguous(), u) for x, u in xs], gy.bfloat16().contiguous(), dwb, dbb, stride=stride)
        torch.cuda.synchronize()
        close(dwb, dw_ref, 3e-2, "conv_wgrad dw bf16")
    finally:
        _set_impl(cu, 0)


# ------------------------------------------------------------------------------------------------
# normalisation / activations / gating
# ------------------------------------------------------------------------------------------------
SHAPES = [(3, 10, 10, 64), (2, 6, 6, 8), (4, 5, 7, 3), (2, 4, 4, 768)]


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_cbn(env, shape):
    cu, ref, dev = env["cu"], env["ref"], env["dev"]
    N, H, W, Cc = shape
    x = rnd(shape, 1, dev) * 2 + 0.5
    gy = rnd(shape, 2, dev)
    scale, offset = rnd((25, Cc), 3, dev) * 0.2 + 1, rnd((25, Cc), 4, dev) * 0.2
    labels = torch.randint(0, 25, (N,), generator=torch.Generator().manual_seed(5)).to(dev).int()
    mean_r, rstd_r = ref.chan_stats(x)
    xf = x.float().contiguous()
    mean, rstd = cu.chan_stats(xf)
    close(mean, mean_r, 1e-5, "mean")
    close(rstd, rstd_r, 1e-5, "rstd")
    for act in (ACT_MIU, ACT_NONE):
        want = ref.cbn_act_fwd(x, mean_r, rstd_r, scale, offset, labels, act)
        got = cu.cbn_act_fwd(xf, mean, rstd, scale.float(), offset.float(), labels, act)
        close(got, want, 1e-5, "cbn fwd")
        ds_r, do_r = torch.zeros_like(scale), torch.zeros_like(offset)
        gx_r = ref.cbn_act_bwd(gy, x, mean_r, rstd_r, scale, offset, labels, ds_r, do_r, act)
        ds, do = torch.zeros_like(scale).float(), torch.zeros_like(offset).float()
        gx = cu.cbn_act_bwd(gy.float().contiguous(), xf, mean, rstd, scale.float(), offset.float(), labels, ds, do, act)
        close(gx, gx_r, 2e-4, "cbn bwd gx")
        close(ds, ds_r, 1e-4, "cbn dscale")
        close(do, do_r, 1e-4, "cbn doffset")
        # fused bias gradient (column sums of gx), accumulating
        db = torch.full((Cc,), 0.5, device=dev)
        ds2, do2 = torch.zeros_like(ds), torch.zeros_like(do)
        gx2 = cu.cbn_act_bwd(gy.float().contiguous(), xf, mean, rstd, scale.float(), offset.float(), labels, ds2, do2, act, dbias=db)
        close(gx2, gx_r, 2e-4, "cbn bwd gx (dbias variant)")
        want_db = gx_r.reshape(-1, Cc).sum(0) + 0.5
        assert (db.double() - want_db).abs().max().item() <= 1e-4 * max(1.0, gx_r.abs().sum(dim=(0, 1, 2)).max().item()), "cbn dbias"
    # bf16 storage
    xb = x.bfloat16().contiguous()
    mb, rb = env["cub"].chan_stats(xb)
    got = env["cub"].cbn_act_fwd(xb, mb, rb, scale.float(), offset.float(), labels, ACT_MIU)
    m2, r2 = ref.chan_stats(xb.double())
    close(got, ref.cbn_act_fwd(xb.double(), m2, r2, scale, offset, labels, ACT_MIU), 1e-2, "cbn fwd bf16")


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_prelu_minmax_actbwd(env, shape):
    cu, ref, dev = env["cu"], env["ref"], env["dev"]
    x, gy = rnd(shape, 1, dev), rnd(shape, 2, dev)
    a = torch.tensor(0.23, device=dev, dtype=torch.float64)
    xf, gf, af = x.float().contiguous(), gy.float().contiguous(), a.float()
    close(cu.prelu_fwd(xf, af), ref.prelu_fwd(x, a), 1e-6, "prelu fwd")
    da_r, da = torch.zeros((), device=dev, dtype=torch.float64), torch.zeros((), device=dev)
    close(cu.prelu_bwd(gf, xf, af, da), ref.prelu_bwd(gy, x, a, da_r), 1e-6, "prelu bwd")
    close(da, da_r, 1e-4, "prelu da")
    Cc = shape[-1]
    da2, db = torch.zeros((), device=dev), torch.full((Cc,), -0.25, device=dev)
    gx_r = ref.prelu_bwd(gy, x, a, None)
    close(cu.prelu_bwd(gf, xf, af, da2, dbias=db), gx_r, 1e-6, "prelu bwd (dbias variant)")
    close(da2, da_r, 1e-4, "prelu da (dbias variant)")
    close(db, gx_r.reshape(-1, Cc).sum(0) - 0.25, 1e-5, "prelu dbias")
    acc0 = rnd(shape, 9, dev)
    da3 = torch.zeros((), device=dev)
    got = cu.prelu_bwd(gf, xf, af, da3, acc_into=acc0.float().contiguous())
    close(got, acc0 + gx_r, 1e-6, "prelu bwd (accumulating form)")
    close(da3, da_r, 1e-4, "prelu da (accumulating form)")
    cs = torch.full((Cc,), 1.5, device=dev)
    cu.colsum_(gf, cs)
    close(cs, gy.reshape(-1, Cc).sum(0) + 1.5, 1e-5, "colsum")
    # min-max (with an exact tie for the maximum in one map)
    xl = torch.where(x > 0, x, 0.2 * x)
    xl[0, 0, 0, 0] = xl[0, 1, 1, 0] = xl[0, :, :, 0].max() + 0.5
    xlf = xl.float().contiguous()
    g_r, mn_r, mx_r = ref.minmax_fwd(xlf.double())
    g, mn, mx = cu.minmax_fwd(xlf)
    close(g, g_r, 1e-5, "minmax fwd")
    close(mn, mn_r, 1e-7, "mn")
    close(mx, mx_r, 1e-7, "mx")
    gmm_r = ref.minmax_bwd(gy, xlf.double(), mn_r, mx_r)
    close(cu.minmax_bwd(gf, xlf, mn, mx), gmm_r, 2e-4, "minmax bwd")
    dbm = torch.zeros((Cc,), device=dev)
    close(cu.minmax_bwd(gf, xlf, mn, mx, dbias=dbm), gmm_r, 2e-4, "minmax bwd (dbias variant)")
    assert (dbm.double() - gmm_r.reshape(-1, Cc).sum(0)).abs().max().item() <= 2e-4 * gmm_r.abs().sum(dim=(0, 1, 2)).max().item(), "minmax dbias"
    # activation backward from the output
    y_t, y_m = torch.tanh(x), (x + torch.sqrt(0.09 + x * x)) / 2
    close(cu.act_bwd(gf, y_t.float().contiguous(), ACT_TANH), ref.act_bwd(gy, y_t, ACT_TANH), 1e-5, "tanh bwd")
    close(cu.act_bwd(gf, y_m.float().contiguous(), ACT_MIU), ref.act_bwd(gy, y_m, ACT_MIU), 1e-4, "miu bwd")


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_gate_prelu(env, shape, dt):
    """prelu(ht + rg*im) as one pass and its backward from the operands (the discriminator's cell, blocks.enc_block_*)."""
    cu, ref, dev = (env["cu"] if dt == torch.float32 else env["cub"]), env["ref"], env["dev"]
    tol = 1e-5 if dt == torch.float32 else 2e-2
    ht, rg, im, gp, g0 = (rnd(shape, s_, dev).to(dt).contiguous() for s_ in (1, 2, 3, 4, 5))
    D = lambda t: t.double()  # noqa: E731
    for aval in (0.2, 1.7, -0.3):
        a = torch.tensor(aval, dtype=torch.float64, device=dev)
        af = a.float()
        close(cu.gate_prelu_fwd(ht, rg, im, af), ref.gate_prelu_fwd(D(ht), D(rg), D(im), a), tol, "gate_prelu fwd a=%g" % aval)
        da_r, da = torch.zeros((), dtype=torch.float64, device=dev), torch.full((), 0.5, device=dev)
        gh_r = D(g0).clone()
        grg_r, gim_r = ref.gate_prelu_bwd(D(gp), D(ht), D(rg), D(im), a, da_r, g_ht=gh_r, acc=True)
        gh = g0.clone()
        grg, gim = cu.gate_prelu_bwd(gp, ht, rg, im, af, da, g_ht=gh, acc=True)
        torch.cuda.synchronize()
        close(grg, grg_r, tol, "g_rg")
        close(gim, gim_r, tol, "g_im")
        close(gh, gh_r, tol, "g_ht (accumulated)")
        assert abs(da.item() - 0.5 - da_r.item()) <= (1e-4 if dt == torch.float32 else 2e-2) * max(1.0, (D(gp) * D(ht)).abs().sum().item())
        gh2 = torch.empty_like(g0)
        cu.gate_prelu_bwd(gp, ht, rg, im, af, None, g_ht=gh2, acc=False)
        close(gh2, gh_r - D(g0), tol, "g_ht (written)")


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_minmax_in_sample_chunks(env, dt):
    """Large gate tensors run the reduce / apply pair a few samples at a time (L2 reuse, CudaOps._mm_chunks): statistics are
    per (sample, channel), so the result -- bias-gradient column sums included -- is the unchunked one."""
    from sketchyscenecolorization_b200.cuda_ops import CudaOps
    dev = env["dev"]
    x = rnd((7, 12, 10, 24), 3, dev).to(dt).contiguous()
    gy = rnd((7, 12, 10, 24), 4, dev).to(dt).contiguous()
    whole, parts = CudaOps(dev, dt), CudaOps(dev, dt)
    whole._MM_CHUNK_BYTES = 0
    parts._MM_CHUNK_BYTES = 2 * x[0].numel() * x.element_size()         # 2 samples per launch pair: 4 chunks, the last ragged
    assert len(parts._mm_chunks(x)) == 4 and len(whole._mm_chunks(x)) == 1
    g0, mn0, mx0 = whole.minmax_fwd(x)
    g1, mn1, mx1 = parts.minmax_fwd(x)
    assert torch.equal(g0, g1) and torch.equal(mn0, mn1) and torch.equal(mx0, mx1)
    db0, db1 = torch.zeros(24, device=dev), torch.zeros(24, device=dev)
    b0, b1 = whole.minmax_bwd(gy, x, mn0, mx0, dbias=db0), parts.minmax_bwd(gy, x, mn1, mx1, dbias=db1)
    torch.cuda.synchronize()
    assert torch.equal(b0, b1)
    close(db1, db0, 1e-5, "dbias across chunks")


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_gating(env, shape, dt):
    cu, ref, dev = (env["cu"] if dt == torch.float32 else env["cub"]), env["ref"], env["dev"]
    tol = 1e-5 if dt == torch.float32 else 2e-2
    N, h, w, Cc = shape
    full = (N, 2 * h, 2 * w, Cc)
    lo, a, b, c, g = (rnd(shape, 1, dev).to(dt), rnd(full, 2, dev).to(dt), rnd(full, 3, dev).to(dt), rnd(full, 4, dev).to(dt),
                      rnd(full, 5, dev).to(dt))
    D = lambda t: t.double()  # noqa: E731
    close(cu.gate_fma_fwd(a, b, c), ref.gate_fma_fwd(D(a), D(b), D(c)), tol, "gate_fma")
    for got, want in zip(cu.gate_fma_bwd(g, b, c), ref.gate_fma_bwd(D(g), D(b), D(c))):
        close(got, want, tol, "gate_fma_bwd")
    close(cu.mul_up_fwd(a, lo), ref.mul_up_fwd(D(a), D(lo)), tol, "mul_up")
    for got, want in zip(cu.mul_up_bwd(g, a, lo), ref.mul_up_bwd(D(g), D(a), D(lo))):
        close(got, want, tol, "mul_up_bwd")
    close(cu.blend_fwd(lo, a, b), ref.blend_fwd(D(lo), D(a), D(b)), tol, "blend")
    for got, want in zip(cu.blend_bwd(g, lo, a, b), ref.blend_bwd(D(g), D(lo), D(a), D(b))):
        close(got, want, tol, "blend_bwd")
    close(cu.addpool_fwd(a, b), ref.addpool_fwd(D(a), D(b)), tol, "addpool")
    close(cu.meanpool_fwd(a), ref.meanpool_fwd(D(a)), tol, "meanpool")
    close(cu.unpool_bwd(lo), ref.unpool_bwd(D(lo)), tol, "unpool")
    assert torch.equal(cu.upsample_fwd(lo), ref.upsample_fwd(lo))
    close(cu.spatial_mean_fwd(a), ref.spatial_mean_fwd(D(a)), tol, "spatial_mean")
    sm = cu.spatial_mean_fwd(a)
    close(cu.spatial_mean_bwd(sm, 2 * h, 2 * w), ref.spatial_mean_bwd(D(sm), 2 * h, 2 * w), tol, "spatial_mean_bwd")
    d = a.clone()
    close(cu.add_(d, b), D(a) + D(b), tol, "add_")
    nchw = rnd((N, Cc, 2 * h, 2 * w), 7, dev).float()
    assert torch.equal(cu.nchw_to_nhwc(nchw, out_dtype=torch.float32), nchw.permute(0, 2, 3, 1).contiguous())
    assert torch.equal(cu.nhwc_to_nchw(a, out_dtype=dt), a.permute(0, 3, 1, 2).contiguous())
    assert torch.equal(cu.cast(a, torch.bfloat16), a.to(torch.bfloat16))


# ------------------------------------------------------------------------------------------------
# caption encoder pieces, spectral norm, losses, Adam
# ------------------------------------------------------------------------------------------------
def test_text_ops(env):
    cu, ref, dev = env["cu"], env["ref"], env["dev"]
    N, P, D, T = 3, 4, 64, 5
    R = N * P
    x, gy = rnd((R, D), 1, dev), rnd((R, D), 2, dev)
    y_r, inv_r = ref.l2norm_rows_fwd(x)
    y, inv = cu.l2norm_rows_fwd(x.float())
    close(y, y_r, 1e-6, "l2norm")
    close(inv, inv_r, 1e-6, "l2norm inv")
    close(cu.l2norm_rows_bwd(gy.float(), y, inv), ref.l2norm_rows_bwd(gy, y_r, inv_r), 1e-5, "l2norm bwd")
    table = rnd((58, D), 3, dev)
    ids = torch.tensor([[0, 0, 5, 7, 9], [0, 3, 3, 3, 3], [2, 4, 6, 8, 10]], device=dev, dtype=torch.int32)
    for t in (0, 1, 4):
        close(cu.embedding_fwd(table.float(), ids, t), ref.embedding_fwd(table, ids, t), 1e-7, "embedding")
        dt_r, dt_ = torch.zeros_like(table), torch.zeros_like(table).float()
        g = rnd((N, D), 4, dev)
        ref.embedding_bwd(g, ids, t, dt_r)
        cu.embedding_bwd(g.float(), ids, t, dt_)
        close(dt_, dt_r, 1e-6, "embedding bwd")
        gates, gates2, grow = rnd((R, 4 * D), 5, dev), rnd((R, 4 * D), 6, dev), rnd((N, 4 * D), 7, dev)
        c0, h0 = rnd((R, D), 8, dev), rnd((R, D), 9, dev)
        want = ref.lstm_cell_fwd(gates, gates2, grow, c0, h0, ids, t, P)
        got = cu.lstm_cell_fwd(gates.float(), gates2.float(), grow.float(), c0.float(), h0.float(), ids, t, P)
        for a, b in zip(got, want):
            close(a, b, 1e-5, "lstm fwd")
        gc, gh = rnd((R, D), 10, dev), rnd((R, D), 11, dev)
        wantb = ref.lstm_cell_bwd(gc, gh, want[2], c0, want[0], ids, t, P)
        gotb = cu.lstm_cell_bwd(gc.float(), gh.float(), got[2], c0.float(), got[0], ids, t, P)
        for a, b in zip(gotb, wantb):
            close(a, b, 1e-5, "lstm bwd")
        # word-LSTM form: no gates2 / grow, P = 1
        want1 = ref.lstm_cell_fwd(grow, None, None, c0[:N], h0[:N], ids, t, 1)
        got1 = cu.lstm_cell_fwd(grow.float(), None, None, c0[:N].float().contiguous(), h0[:N].float().contiguous(), ids, t, 1)
        for a, b in zip(got1, want1):
            close(a, b, 1e-5, "lstm fwd P=1")
    close(cu.rows_group_sum(x.float(), P), ref.rows_group_sum(x, P), 1e-6, "rows_group_sum")
    h = torch.tanh(rnd((R, D), 12, dev))
    close(cu.atanh_relu_fwd(h.float()), ref.atanh_relu_fwd(h), 1e-5, "atanh_relu")
    close(cu.atanh_relu_bwd(gy.float(), h.float()), ref.atanh_relu_bwd(gy, h), 1e-4, "atanh_relu bwd")


@pytest.mark.parametrize("case", [(2, 24, 24, 128, 3, 128), (2, 16, 16, 512, 3, 64), (64, 1, 1, 1024, 1, 512)], ids=str)
def test_six_product_convolution(env, case):
    """CudaOps(conv_terms=3): x = x1 + x2 + x3, w = w1 + w2 + w3 in bf16 terms, the six products of order <= 2^-16 kept
    (one bf16x3 pass + three accumulating bf16 passes, fgc_conv2d_fwd_acc / fgc_split_term) -- against fp64, next to bf16x3
    and the fp32 CUDA-core convolution.  What it shows (printed; B200, K = 4608: fp32 FMA 2.4e-6, bf16x3 1.60e-5, six products
    1.58e-5 of the largest entry): adding the missing operand terms changes nothing -- the error of the tensor path is the
    tensor core's fp32 ACCUMULATION (it aligns and truncates the addends, where an FMA chain rounds to nearest), about 6x an
    fp32 FMA chain's at this depth.  That, amplified by ~110 batch-normalised layers, is the floor of the Residual and
    background generators on the tensor path; the MRU / Pix2Pix networks sit far above it."""
    from sketchyscenecolorization_b200.cuda_ops import CudaOps
    cu, ref, dev = env["cu"], env["ref"], env["dev"]
    N, H, W, cin, k, cout = case
    x, w = rnd((N, H, W, cin), 1, dev), rnd((k, k, cin, cout), 2, dev, 1.0 / math.sqrt(k * k * cin))
    b = rnd((cout,), 3, dev, 0.3)
    want = ref.conv_fwd([(x, False)], w, b)
    xf, wf, bf = x.float().contiguous(), w.float().contiguous(), b.float().contiguous()
    want32 = ref.conv_fwd([(xf.double(), False)], wf.double(), bf.double())        # fp64 arithmetic on the fp32-rounded operands
    cu6 = CudaOps(dev, torch.float32, conv_terms=3)
    errs = {}
    _set_impl(cu, 1)
    try:
        errs["fp32 CUDA cores"] = cu.conv_fwd([(xf, False)], wf, bf)
    finally:
        _set_impl(cu, 0)
    errs["bf16x3"] = cu.conv_fwd([(xf, False)], wf, bf)
    errs["six products"] = cu6.conv_fwd([(xf, False)], wf, bf)
    torch.cuda.synchronize()
    scale = want32.abs().max().item()
    e = {kk: (v.double() - want32).abs().max().item() / scale for kk, v in errs.items()}
    print("conv %s, max-abs error / largest entry vs fp64 on the same fp32 operands: %s" % (case, {kk: "%.2e" % v for kk, v in e.items()}))
    assert e["six products"] <= 1.05 * e["bf16x3"] + 1e-7 and e["bf16x3"] <= 1e-4 and e["fp32 CUDA cores"] <= 2e-5
    close(errs["six products"], want, 1e-4, "six-product conv vs fp64 of the fp64 operands")


@pytest.mark.parametrize("shape", [(15, 64, 512), (5, 3, 128), (4, 70, 64), (2, 1, 16), (15, 130, 512)], ids=str)
def test_word_lstm_sequence_kernels(env, shape):
    """fgc_lstm_seq_fwd / _bwd (the word LSTM's recurrence and its BPTT, one persistent launch each, grid barrier per step)
    and the all-steps embedding gather / scatter against the plain-torch operators: the production size (T 15, N 64,
    D 512: 128 co-resident CTAs), ragged batches, more than one 64-sample chunk, captions with leading <pad>s."""
    cu, ref, dev = env["cu"], env["ref"], env["dev"]
    T, N, D = shape
    g = torch.Generator().manual_seed(T * 1000 + N)
    ids = torch.randint(1, 58, (N, T), generator=g)
    for n in range(N):                                      # pads are a prefix (text_processing.py:50-52); one all-pad row
        ids[n, :int(torch.randint(0, T, (1,), generator=g))] = 0
    if N > 2:
        ids[1, :] = 0
    ids = ids.to(dev, torch.int32).contiguous()
    table = rnd((58, D), 1, dev)
    e_r, e = ref.embedding_all_fwd(table, ids), cu.embedding_all_fwd(table.float(), ids)
    assert e.shape == (T, N, D)
    close(e, e_r, 1e-7, "embedding_all")
    gx, kh = rnd((T, N, 4 * D), 2, dev), rnd((D, 4 * D), 3, dev, 1.0 / math.sqrt(D))
    want = ref.lstm_seq_fwd(gx, kh, ids)
    got = cu.lstm_seq_fwd(gx.float().contiguous(), kh.float().contiguous(), ids)
    torch.cuda.synchronize()
    for a, b, what in zip(got, want, ("h_all", "c_all", "pre_all")):
        assert a.shape == b.shape
        close(a, b, 2e-5, "lstm_seq_fwd " + what)
    assert float(got[0][0].abs().max()) == 0.0 and float(got[1][0].abs().max()) == 0.0      # slot 0: the zero initial state
    g_hext = rnd((T, N, D), 4, dev)
    want_b = ref.lstm_seq_bwd(g_hext, want[2], want[1], kh, ids)
    got_b = cu.lstm_seq_bwd(g_hext.float().contiguous(), got[2], got[1], kh.float().contiguous(), ids)
    torch.cuda.synchronize()
    close(got_b, want_b, 5e-5, "lstm_seq_bwd")
    # against autograd through the plain recurrence (pins the hand-written BPTT of both implementations)
    gxa, kha = gx.clone().requires_grad_(True), kh.clone().requires_grad_(True)
    h, c = torch.zeros(N, D, dtype=torch.float64, device=dev), torch.zeros(N, D, dtype=torch.float64, device=dev)
    loss = 0.0
    for t in range(T):
        pre = gxa[t] + h @ kha
        i, j, f, o = pre.chunk(4, dim=1)
        c2 = c * torch.sigmoid(f + 1.0) + torch.sigmoid(i) * torch.tanh(j)
        h2 = torch.tanh(c2) * torch.sigmoid(o)
        m = (ids[:, t] != 0)[:, None]
        c, h = torch.where(m, c2, c), torch.where(m, h2, h)
        loss = loss + (h * g_hext[t]).sum()
    loss.backward()
    close(got_b, gxa.grad, 5e-5, "lstm_seq_bwd vs autograd")
    dt_r, dt_ = torch.zeros_like(table), torch.zeros_like(table).float()
    ref.embedding_all_bwd(g_hext, ids, dt_r)
    cu.embedding_all_bwd(g_hext.float().contiguous(), ids, dt_)
    close(dt_, dt_r, 1e-5, "embedding_all bwd")


@pytest.mark.parametrize("kc", [(27 * 8, 8), (1152, 128), (768, 1), (768, 25), (6912, 768)], ids=str)
def test_spectral_norm(env, kc):
    cu, ref, dev = env["cu"], env["ref"], env["dev"]
    K, Cc = kc
    w, u, gw = rnd((K, Cc), 1, dev, 0.02), rnd((1, Cc), 2, dev), rnd((K, Cc), 3, dev)
    wbar_r, ctx_r = ref.sn_fwd(w, u)
    wbar, ctx = cu.sn_fwd(w.float().contiguous(), u.float().contiguous())
    close(wbar, wbar_r, 2e-5, "wbar")
    close(ctx["u_new"], ctx_r["u_new"], 2e-5, "u_new")
    dw_r, dw = rnd((K, Cc), 4, dev), None
    dw = dw_r.float().contiguous()
    ref.sn_bwd(gw, w, ctx_r, dw_r)
    cu.sn_bwd(gw.float().contiguous(), w.float().contiguous(), ctx, dw)
    close(dw, dw_r, 1e-4, "sn dw")


def test_losses_and_adam(env):
    cu, ref, dev = env["cu"], env["ref"], env["dev"]
    d = rnd((4, 12, 12, 1), 1, dev) * 3
    for sign in (1.0, -1.0):
        l_r, g_r = ref.softplus_mean(d, sign)
        l, g = cu.softplus_mean(d.float(), sign)
        close(l, l_r, 1e-5, "softplus")
        close(g, g_r, 1e-5, "softplus grad")
    logits = rnd((6, 1, 1, 25), 2, dev) * 2
    labels = torch.tensor([0, 24, 3, 3, 7, 11], device=dev, dtype=torch.int32)
    for focal, wgt in ((True, 1.0), (False, 0.5)):
        l_r, g_r = ref.ce_loss(logits, labels, focal, wgt)
        l, g = cu.ce_loss(logits.float(), labels, focal, wgt)
        close(l, l_r, 1e-5, "ce")
        close(g, g_r, 1e-5, "ce grad")
    t, gen = rnd((2, 16, 16, 3), 3, dev) * 1.5, rnd((2, 16, 16, 3), 4, dev)
    l_r, g_r = ref.smooth_l1(t, gen, 100.0)
    l, g = cu.smooth_l1(t.float(), gen.float(), 100.0)
    close(l, l_r, 1e-5, "smooth_l1")
    close(g, g_r, 1e-6, "smooth_l1 grad")
    # reg loss + Adam on a real parameter store
    from sketchyscenecolorization_b200.params import ParamStore, discriminator_vars
    st = ParamStore(discriminator_vars(8), dev)
    st.initialize(3)
    st_r = ParamStore(discriminator_vars(8), dev, torch.float64)
    st_r.load_state_dict(st.state_dict())
    close(cu.reg_loss(st), ref.reg_loss(st_r), 1e-5, "reg")
    gr = rnd((st.n_flat,), 5, dev, 0.01)
    for it in range(3):
        st.grad.copy_(gr.float())
        st_r.grad.copy_(gr)
        cu.adam_step(st, 1e-3)
        ref.adam_step(st_r, 1e-3)
    for k in st.p:       # (the flat buffers also hold alignment padding that only the torch reference touches)
        close(st.p[k], st_r.p[k], 1e-5, "adam params " + k)
    o = st.offsets["discriminator/Conv_1/weights"]
    close(st.adam_v[o:o + 64], st_r.adam_v[o:o + 64], 1e-5, "adam v")


def test_conv_routing(env):
    """The product routing really uses the kernels it claims: halo-reuse for wide stride-1 bf16 layers (also when the TMA box
    is taller than the image), the direct kernels for the narrow stem layers, the gather kernel for fp32 / upsampled sources."""
    cub, cu, dev = env["cub"], env["cu"], env["dev"]
    _set_impl(cu, 0)

    def delta(fn):
        c0 = conv_counts(cu)
        fn()
        torch.cuda.synchronize()
        return [b - a for a, b in zip(c0, conv_counts(cu))]

    x128 = rnd((2, 16, 24, 128), 1, dev).bfloat16()
    w = rnd((3, 3, 128, 64), 2, dev, 0.05).float()
    b = torch.zeros(64, device=dev)
    assert delta(lambda: cub.conv_fwd([(x128, False)], w, b))[:2] == [1, 0]               # 18-row box over a 16-row image
    assert delta(lambda: cu.conv_fwd([(x128.float(), False)], w, b))[:2] == [0, 1]        # fp32 (bf16x3): gather kernel
    xlow = rnd((2, 8, 12, 128), 3, dev).bfloat16()
    assert delta(lambda: cub.conv_fwd([(xlow, True)], w, b))[:2] == [0, 1]                # upsampled source: gather kernel
    gy = rnd((2, 16, 24, 64), 4, dev).bfloat16()
    assert delta(lambda: cub.conv_dgrad(gy, w, 0, 128))[:2] == [1, 0]
    x3 = rnd((2, 24, 24, 3), 5, dev).bfloat16()
    w3 = rnd((7, 7, 3, 8), 6, dev, 0.05).float()
    b8 = torch.zeros(8, device=dev)
    assert delta(lambda: cub.conv_fwd([(x3, False)], w3, b8))[2] == 1
    gy8 = rnd((2, 24, 24, 8), 7, dev).bfloat16()
    dw, db = torch.zeros_like(w3), torch.zeros_like(b8)
    assert delta(lambda: cub.conv_wgrad([(x3, False)], gy8, dw, db))[3] == 1
    dw2, db2 = torch.zeros_like(w), torch.zeros_like(b)
    assert delta(lambda: cub.conv_wgrad([(x128, False)], gy, dw2, db2))[4:6] == [0, 1]    # halo-reuse wgrad
    assert delta(lambda: cub.conv_wgrad([(xlow, True)], gy, dw2, db2))[4:6] == [1, 0]     # upsampled source: per-tap wgrad


@pytest.mark.parametrize("case", [(8, 64, 80, [(64, False), (8, False), (3, False)], 3, 64),
                                  (2, 28, 24, [(3, False)], 3, 128),
                                  (3, 32, 64, [(128, False), (3, True)], 3, 32)], ids=str)
def test_conv_with_patch_sources(env, case):
    """Narrow sources handed over as pre-flattened patch tensors (fgc_im2col_small): forward and weight gradient fetch them by
    TMA (halo-reuse kernel / tiled wgrad); values equal the plain-source result."""
    cub, ref, dev = env["cub"], env["ref"], env["dev"]
    N, H, W, srcs, k, cout = case
    xs, w, b = _conv_inputs((N, H, W, srcs, k, 1, cout, ACT_NONE), dev, seed=40)
    want = ref.conv_fwd(xs, w, b)
    xb = []
    for x, u in xs:
        t = x.bfloat16().contiguous()
        p = cub.small_patch(t, k, ups=u) if t.shape[-1] < 64 else None
        if t.shape[-1] < 64:
            assert p is not None and p.shape[-1] % 8 == 0
        xb.append((t, u, p))
    c0 = conv_counts(cub)
    got = cub.conv_fwd(xb, w.float().contiguous(), b.float().contiguous())
    torch.cuda.synchronize()
    assert conv_counts(cub)[0] == c0[0] + 1          # halo-reuse kernel, no gathered slabs needed
    close(got, want, 3e-2, "conv_fwd with patch sources")
    gy = rnd((N, H, W, cout), 51, dev)
    dw0, db0 = rnd(w.shape, 52, dev), rnd(b.shape, 53, dev)
    dw_ref, db_ref = dw0.clone(), db0.clone()
    ref.conv_wgrad(xs, gy, dw_ref, db_ref)
    dw, db = dw0.float().contiguous(), db0.float().contiguous()
    cub.conv_wgrad(xb, gy.bfloat16().contiguous(), dw, db)
    torch.cuda.synchronize()
    close(dw, dw_ref, 3e-2, "conv_wgrad with patch sources")


@pytest.mark.parametrize("case", [(5000, 3, "bf16"), (77, 3, "f32"), (64, 2048, "f32"), (40, 300, "bf16"), (9000, 11, "f32"),
                                  (300, 256, "bf16")], ids=str)
def test_colsum_generic_forms(env, case):
    """fgc_colsum outside the vectorised form: narrow rows (the bias gradient of a 3-channel layer: row lanes of a block combined in
    shared memory, one atomic per channel and block) and wide rows over few samples (columns spread over grid.y)."""
    cu, dev = env["cu"], env["dev"]
    M, Cc, dt = case
    x = rnd((M, Cc), 31, dev)
    xd = (x.bfloat16() if dt == "bf16" else x.float()).contiguous()
    out = torch.full((Cc,), 0.5, device=dev)
    cu.colsum_(xd, out)
    torch.cuda.synchronize()
    close(out, xd.double().sum(0) + 0.5, 2e-5, "colsum %s" % (case,))


def test_keep_packed_measurement_switch(env):
    """fgc_debug_keep_packed (bench.py's kernel-only timing of the dominant layer): a repeated identical call that skips the
    weight-packing launch returns the same bytes."""
    cub, dev = env["cub"], env["dev"]
    x = rnd((2, 32, 32, 128), 5, dev).bfloat16().contiguous()
    w = (rnd((3, 3, 128, 128), 6, dev) * 0.05).float().contiguous()
    b = rnd((128,), 7, dev).float().contiguous()
    arr, N, H, W, dt = cub._srcs([(x, False)])
    ws = cub._ws([128], 3, 128, dt)
    ys = [torch.empty((2, 32, 32, 128), dtype=torch.bfloat16, device=dev) for _ in range(2)]
    st = torch.cuda.current_stream().cuda_stream
    try:
        for i, y in enumerate(ys):
            cub.lib.fgc_debug_keep_packed(i)
            rc = cub.lib.fgc_conv2d_fwd(arr, 1, dt, N, H, W, w.data_ptr(), 3, 128, 128, b.data_ptr(), 1, 1, 1, H, W, ACT_NONE,
                                        y.data_ptr(), 1, ws.data_ptr(), st)
            assert rc == 0
    finally:
        cub.lib.fgc_debug_keep_packed(0)
    torch.cuda.synchronize()
    assert torch.equal(ys[0].view(torch.int16), ys[1].view(torch.int16))
    assert ys[0].float().abs().max().item() > 0


def _patch_reference(x, k, ups, mirror):
    """out[n,h,w,(kh*k+kw)*C + c] = x[n, h + s*(kh-pad), w + s*(kw-pad), c] (s = -1: mirrored taps), zeros outside the image and in
    the tail that pads k*k*C to a multiple of 8 -- plain indexing on the device, no arithmetic."""
    if ups:
        x = x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
    N, H, W, Cc = x.shape
    pad = (k - 1) // 2
    cp = 8 * ((k * k * Cc + 7) // 8)
    xp = torch.nn.functional.pad(x, (0, 0, pad, pad, pad, pad))
    out = torch.zeros((N, H, W, cp), dtype=x.dtype, device=x.device)
    s = -1 if mirror else 1
    for kh in range(k):
        for kw in range(k):
            dh, dw = s * (kh - pad), s * (kw - pad)
            out[..., (kh * k + kw) * Cc:(kh * k + kw + 1) * Cc] = xp[:, pad + dh:pad + dh + H, pad + dw:pad + dw + W, :]
    return out


@pytest.mark.parametrize("case", [(3, 40, 24, 3, 7), (2, 37, 16, 3, 3), (2, 21, 32, 8, 3), (1, 9, 8, 11, 3), (2, 5, 24, 5, 5),
                                  (2, 192, 192, 3, 7), (2, 192, 192, 8, 3), (1, 16, 8, 40, 3)], ids=str)
@pytest.mark.parametrize("mirror", [False, True])
def test_patch_tensor_bit_exact(env, case, mirror):
    """fgc_im2col_small is byte movement: the row-tiled kernel (sources that are not upsampled; rows staged in shared memory with
    the SAME padding as zeros) and the per-element kernel (upsampled sources) reproduce the indexed reference bit for bit, from
    bf16 and from fp32 sources (fp32: round-to-nearest-even to bf16)."""
    cub, dev = env["cub"], env["dev"]
    N, H, W, Cc, k = case
    x32 = rnd((N, H, W, Cc), 77, dev).float().contiguous()
    xb = x32.bfloat16().contiguous()
    want = _patch_reference(xb, k, False, mirror)
    got = cub.small_patch(xb, k, mirror=mirror)
    torch.cuda.synchronize()
    assert got is not None and got.shape == want.shape
    assert torch.equal(got.view(torch.int16), want.view(torch.int16)), "patch tensor (bf16 source) differs"
    # fp32 source through the C entry point (CudaOps.small_patch hands bf16 tensors over only)
    out = torch.empty_like(want)
    rc = cub.lib.fgc_im2col_small(x32.data_ptr(), 0, N, H, W, Cc, 0, k, 1 if mirror else 0, out.data_ptr(),
                                  torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert rc == 0 and torch.equal(out.view(torch.int16), want.view(torch.int16)), "patch tensor (fp32 source) differs"
    if H % 2 == 0 and W % 2 == 0 and H <= 64:       # the upsampled form keeps the per-element kernel
        lo = xb[:, :H // 2, :W // 2].contiguous()
        want_u = _patch_reference(lo, k, True, mirror)
        got_u = cub.small_patch(lo, k, ups=True, mirror=mirror)
        torch.cuda.synchronize()
        assert torch.equal(got_u.view(torch.int16), want_u.view(torch.int16)), "patch tensor (upsampled source) differs"


@pytest.mark.parametrize("case", [(2, 64, 24, [(128, False)], 3, 128, ACT_LRELU), (3, 64, 16, [(64, False), (3, False)], 3, 64, ACT_NONE),
                                  (2, 128, 8, [(64, False)], 7, 3, ACT_TANH)], ids=str)
def test_conv_halo_tall_tiles(env, case):
    """64 x 8 pixel tiles (four 128-pixel sub-tiles per CTA, single or double buffered accumulators), forced on small inputs."""
    cub, ref, dev = env["cub"], env["ref"], env["dev"]
    N, H, W, srcs, k, cout, act = case
    xs, w, b = _conv_inputs((N, H, W, srcs, k, 1, cout, act), dev, seed=60)
    want = ref.conv_fwd(xs, w, b, act=act)
    xb = [(x.bfloat16().contiguous(), u, cub.small_patch(x.bfloat16().contiguous(), k) if x.shape[-1] < 64 else None) for x, u in xs]
    cub.lib.fgc_set_conv_flags(3, 1)
    try:
        c0 = conv_counts(cub)
        got = cub.conv_fwd(xb, w.float().contiguous(), b.float().contiguous(), act=act)
        torch.cuda.synchronize()
        assert conv_counts(cub)[0] == c0[0] + 1
        close(got, want, 3e-2, "conv_fwd, 64x8 tiles")
        gy = rnd((N, H, W, cout), 61, dev)
        wd = rnd((k, k, sum(c for c, _ in srcs), cout), 62, dev, 1.0 / math.sqrt(k * k * cout))
        want_d = ref.conv_dgrad(gy, wd, 0, srcs[0][0])
        if cout >= 64:
            got_d = cub.conv_dgrad(gy.bfloat16().contiguous(), wd.float().contiguous(), 0, srcs[0][0])
            torch.cuda.synchronize()
            close(got_d, want_d, 3e-2, "conv_dgrad, 64x8 tiles")
    finally:
        cub.lib.fgc_set_conv_flags(1, 1)


def test_conv_head_backward_with_gy_patches(env):
    """7x7, 64 -> 3 head: input gradient from the mirrored patch tensor of the 3-channel gy (halo-reuse kernel) and weight
    gradient from its plain patch tensor (operand-swapped halo-reuse wgrad) equal the reference."""
    cub, ref, dev = env["cub"], env["ref"], env["dev"]
    N, H, W, cin, cout, k = 3, 32, 40, 64, 3, 7
    x = rnd((N, H, W, cin), 70, dev)
    gy = rnd((N, H, W, cout), 71, dev)
    w = rnd((k, k, cin, cout), 72, dev, 1.0 / math.sqrt(k * k * cout))
    xb, gb = x.bfloat16().contiguous(), gy.bfloat16().contiguous()
    want = ref.conv_dgrad(gy, w, 0, cin)
    c0 = conv_counts(cub)
    got = cub.conv_dgrad(gb, w.float().contiguous(), 0, cin, gy_patch=cub.small_patch(gb, k, mirror=True))
    torch.cuda.synchronize()
    assert conv_counts(cub)[0] == c0[0] + 1
    close(got, want, 3e-2, "head dgrad via gy patch")
    dw0, db0 = rnd(w.shape, 73, dev), rnd((cout,), 74, dev)
    dw_ref, db_ref = dw0.clone(), db0.clone()
    ref.conv_wgrad([(x, False)], gy, dw_ref, db_ref)
    dw, db = dw0.float().contiguous(), db0.float().contiguous()
    c0 = conv_counts(cub)
    cub.conv_wgrad([(xb, False)], gb, dw, db, gy_patch=cub.small_patch(gb, k))
    torch.cuda.synchronize()
    assert conv_counts(cub)[5] == c0[5] + 1
    close(dw, dw_ref, 3e-2, "head wgrad via gy patch")
    close(db, db_ref, 1e-2, "head db")
