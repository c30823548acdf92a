"""Which convolution arithmetic does each network need for 1e-3 inference parity?  (CPU only)

Emulates the tensor-core operand split inside the oracles: every convolution's input and filter are split into bf16 terms
(x = x0 + x1 + ..., each term the bf16 rounding of what is left), the products kept by a scheme are evaluated in fp64 and summed
exactly; everything else stays fp64.  Prints the max-abs deviation of the generator output from the fp64 oracle."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch                                    # noqa: E402
import torch.nn.functional as F                 # noqa: E402
from oracle import bg_oracle as B               # noqa: E402
from oracle import fgcolor_oracle as O          # noqa: E402
from oracle import pix2pix_oracle as P          # noqa: E402
from oracle import residual_oracle as R         # noqa: E402

SCHEMES = [("bf16, one product", 1, 1), ("bf16x3 (hi*hi + hi*lo + lo*hi)", 2, 2), ("bf16x4 (+ lo*lo)", 2, 3),
           ("three bf16 terms, six products", 3, 3), ("three bf16 terms, nine products", 3, 5)]
MODE = {}


def _terms(x, n):
    out, rest = [], x
    for _ in range(n):
        t = rest.to(torch.bfloat16).to(x.dtype)
        out.append(t)
        rest = rest - t
    return out


def _emulated(fn):
    def run(x, w, *a, **k):
        acc = None
        for i, xt in enumerate(_terms(x, MODE["terms"])):
            for j, wt in enumerate(_terms(w, MODE["terms"])):
                if i + j < MODE["order"]:               # keep the products down to this combined order of smallness
                    y = fn(xt, wt, *a, **k)
                    acc = y if acc is None else acc + y
        return acc
    return run


def study(specs, fwd, args):
    p = O.init_params(specs, 3, torch.float64, perturb_tables=0.1)
    first = lambda o: o[0] if isinstance(o, tuple) else o      # noqa: E731
    res = {}
    with torch.no_grad():
        ref = first(fwd(p, *args))
        conv, convt = F.conv2d, F.conv_transpose2d
        for name, terms, order in SCHEMES:
            MODE.update(terms=terms, order=order)
            F.conv2d, F.conv_transpose2d = _emulated(conv), _emulated(convt)
            try:
                res[name] = (first(fwd(p, *args)) - ref).abs().max().item()
            finally:
                F.conv2d, F.conv_transpose2d = conv, convt
        p32 = {k: v.float() for k, v in p.items()}
        a32 = [a.float() if torch.is_tensor(a) and a.is_floating_point() else a for a in args]
        res["fp32 throughout"] = (first(fwd(p32, *a32)).double() - ref).abs().max().item()
    return res


if __name__ == "__main__":
    b = O.make_batch(2, 64, 64, 11, torch.float64, n_pad=4)
    args = (b["sketch"], b["text"], b["cls"], b["noise"])
    nets = [("MRU, size 16", O.generator_specs(16, 58, 64, 64), lambda p, *a: O.generator_forward(p, *a, 16), args),
            ("Pix2Pix, size 16", P.generator_specs(16, 58, 64, 64), lambda p, *a: P.generator_forward(p, *a, 16), args),
            ("Residual, size 16", R.generator_specs(16, 58, 64, 64), lambda p, *a: R.generator_forward(p, *a, 16), args)]
    g = torch.Generator().manual_seed(7)
    img = torch.rand(2, 3, 96, 96, generator=g, dtype=torch.float64) * 2 - 1
    ids = torch.randint(2, 18, (2, 8), generator=g)
    nets.append(("background, ngf 8", B.generator_specs(8, 18), lambda p, *a: B.generator_forward(p, *a), (img, ids)))
    for name, specs, fwd, a in nets:
        print(name)
        for k, v in study(specs, fwd, a).items():
            print("    %-34s %.1e" % (k, v))
