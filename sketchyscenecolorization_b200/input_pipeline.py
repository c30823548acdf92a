"""Host-side image preparation of the fg-colorization path.

Mirrors the two numpy/PIL helpers of the reference's obj_lib/input_pipeline.py that inference needs:
`resize_and_padding_mask_image` (:199-239).  The TFRecord queue (:45-196) is replaced by an iterator
protocol: anything yielding dicts with `sketch`, `images`, `images_d` [N,3,H,W] float32 in [-1,1], `cls`,
`cls_d` int32 [N], `text` int32 [N,15] (SyntheticInput below follows SURVEY 8(d))."""
from __future__ import annotations

import numpy as np
import torch

CATEGORIES = ("bench bird bus butterfly car cat chair chicken cloud cow dog duck grass horse house moon person pig "
              "rabbit road sheep star sun tree truck").split()      # sorted(os.listdir('data/captions')), 25 classes


def resize_and_padding_mask_image(image, new_size, resample_method=None, margin_size=10):
    """PIL image -> uint8 [new_size, new_size, 3]: scale the longer side (+2*margin) to new_size, centre, pad white.
    (reference default resample Image.ANTIALIAS == LANCZOS, removed from Pillow >= 10)"""
    from PIL import Image
    if resample_method is None:
        resample_method = Image.LANCZOS
    scale = new_size / max(image.height + 2 * margin_size, image.width + 2 * margin_size)
    new_h, new_w = int(round(image.height * scale)), int(round(image.width * scale))
    assert new_h <= new_size and new_w <= new_size
    if scale != 1:
        image = image.resize((new_w, new_h), resample=resample_method)
    plane = np.array(image, dtype=np.uint8)
    if plane.ndim == 3:
        plane = plane[:, :, 0]
    top, left = (new_size - new_h) // 2, (new_size - new_w) // 2
    canvas = np.full((new_size, new_size), 255, dtype=np.uint8)
    canvas[top:top + new_h, left:left + new_w] = plane
    return np.repeat(canvas[:, :, None], 3, axis=2)


class SyntheticInput:
    """Seeded stand-in for the two TFRecord shuffle queues of main_procedure.train (:109-122)."""

    def __init__(self, batch_size, H=192, W=192, vocab_size=58, seed=1234):
        self.n, self.H, self.W, self.vocab = batch_size, H, W, vocab_size
        self.g = torch.Generator().manual_seed(seed)

    def _sketch(self):
        n, H, W, g = self.n, self.H, self.W, self.g
        sk = torch.ones(n, 1, H, W)
        for i in range(n):
            for _ in range(6):
                p = torch.rand(5, 2, generator=g) * torch.tensor([H - 1.0, W - 1.0])
                for a, b in zip(p[:-1], p[1:]):
                    L = int(max(abs(b[0] - a[0]), abs(b[1] - a[1]))) + 1
                    ys = torch.linspace(a[0].item(), b[0].item(), L).round().long()
                    xs = torch.linspace(a[1].item(), b[1].item(), L).round().long()
                    sk[i, 0, ys, xs] = -1.0
        return sk.expand(n, 3, H, W).contiguous()

    def __iter__(self):
        return self

    def __next__(self):
        n, g = self.n, self.g
        return dict(sketch=self._sketch(), images=torch.rand(n, 3, self.H, self.W, generator=g) * 2 - 1,
                    images_d=torch.rand(n, 3, self.H, self.W, generator=g) * 2 - 1,
                    cls=torch.randint(0, 25, (n,), generator=g).int(), cls_d=torch.randint(0, 25, (n,), generator=g).int(),
                    text=torch.randint(2, self.vocab, (n, 15), generator=g).int())
