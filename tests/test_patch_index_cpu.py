"""Index arithmetic of the row-tiled patch-tensor kernel (csrc/conv_small.cu: im2col_rows_kernel), restated in numpy and checked
against the definition of the patch tensor -- out[n,h,w,(kh*k+kw)*C + c] = x[n, h + s*(kh-pad), w + s*(kw-pad), c], zeros outside
the image and in the tail that pads k*k*C to a multiple of 8 (s = -1: mirrored taps, the input-gradient form).  The kernel's
shared-memory layout (left / right margins of >= pad*C zeros, rows beyond the image zero, per-thread constant offsets) is what
this walks through; the GPU test `test_patch_tensor_bit_exact` checks the kernel itself bit for bit."""
import numpy as np
import pytest


def patch_reference(x, k, sign):
    N, H, W, C = x.shape
    pad = (k - 1) // 2
    cp = 8 * ((k * k * C + 7) // 8)
    out = np.zeros((N, H, W, cp), x.dtype)
    xp = np.pad(x, ((0, 0), (pad, pad), (pad, pad), (0, 0)))
    for kh in range(k):
        for kw in range(k):
            dh, dw = sign * (kh - pad), sign * (kw - pad)
            out[..., (kh * k + kw) * C:(kh * k + kw + 1) * C] = xp[:, pad + dh:pad + dh + H, pad + dw:pad + dw + W, :]
    return out


def patch_rows_emulation(x, k, sign, TH=8, block=256):
    """Thread-for-thread restatement of im2col_rows_kernel (scalar staging path)."""
    N, H, W, C = x.shape
    pad = (k - 1) // 2
    KK = k * k * C
    CP = 8 * ((KK + 7) // 8)
    CPV = CP // 8
    WC = W * C
    LP = 8 * ((pad * C + 7) // 8)
    RS = 8 * ((LP + WC + pad * C + 7) // 8)
    TH = min(TH, H)
    tiles_h = (H + TH - 1) // TH
    out = np.full((N, H, W, CP), -7, x.dtype)
    xf = x.reshape(-1)
    for b in range(N * tiles_h):
        n = b // tiles_h
        h0 = (b - n * tiles_h) * TH
        th = min(TH, H - h0)
        R = th + k - 1
        sh = np.full(((TH + k - 1) * RS,), -99, x.dtype)          # unwritten shared memory must never be read
        for r in range(R):
            ih = h0 - pad + r
            for e in range(RS):
                j = e - LP
                sh[r * RS + e] = xf[(n * H + ih) * WC + j] if (0 <= ih < H and 0 <= j < WC) else 0
        lanes = block // CPV
        for t in range(block):
            ch, lane = t % CPV, t // CPV
            if lane >= lanes:
                continue
            off, valid = [0] * 8, 0
            for e in range(8):
                q = ch * 8 + e
                if q < KK:
                    tap, c = divmod(q, C)
                    kh, kw = divmod(tap, k)
                    off[e] = (pad + sign * (kh - pad)) * RS + LP + sign * (kw - pad) * C + c
                    valid |= 1 << e
            hl, w = divmod(lane, W)
            dh, dw = divmod(lanes, W)
            while hl < th:
                pix = hl * RS + w * C
                for e in range(8):
                    out[n, h0 + hl, w, ch * 8 + e] = sh[pix + off[e]] if (valid >> e) & 1 else 0
                hl += dh
                w += dw
                if w >= W:
                    w -= W
                    hl += 1
    return out


@pytest.mark.parametrize("case", [(2, 5, 8, 3, 7), (1, 11, 16, 3, 3), (2, 9, 8, 8, 3), (1, 3, 8, 11, 3), (1, 20, 24, 5, 5)], ids=str)
@pytest.mark.parametrize("sign", [1, -1])
def test_row_tiled_patch_indexing(case, sign):
    N, H, W, C, k = case
    x = np.random.default_rng(0).integers(1, 1000, (N, H, W, C)).astype(np.int64)
    want = patch_reference(x, k, sign)
    got = patch_rows_emulation(x, k, sign)
    assert (got == want).all()
