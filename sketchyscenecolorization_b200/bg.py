"""Background-colorization generator, inference (BASELINE.json configs[3]: 768x768, one picture per call).

Reference: Background_Colorization/bg_colorization_main.py -- create_residual_generator (:302-420) with the defaults of
bg_colorization (:733-736: residual encoder, [3, 4, 6, 3] units per level), encode_feat_with_text (:117-214), the test loop's
pre- / post-processing (:752-765, 838-882).  It is the fg `Residual` generator (residual.py) with a 1024-channel fifth level,
the caption fused at 24 x 24 (576 mLSTM positions instead of 36), no noise input and a second output: three region logits per
pixel from a 1x1 projection of the bottleneck upsampled by five transposed convolutions.  Built from residual.Layers, i.e. the
stride-1 SAME tensor-core convolutions (4x4 layers in phase form), plain batch norm with batch statistics -- at batch 1, as the
reference runs it, also at test time -- and the fused LSTM kernels.  Precision: with the bf16x3 tensor-core convolutions the
picture is within ~1e-2 of exact arithmetic (about two grey levels; an fp32 run of the reference is itself 1e-3 away: this
network amplifies rounding ~1000x); FGC_CONV_IMPL=simple runs the fp32 CUDA-core convolutions instead.
Forward only: training this network (its own
discriminator, the segmentation loss, Adam with beta1 = 0.5) is outside SURVEY 8.
"""
from __future__ import annotations

import numpy as np
import torch

from . import text_fusion
from .params import BG_TEXT_SCOPE, ParamStore, bg_generator_vars
from .residual import UNITS, Layers


class BgGenerator:
    def __init__(self, ops, store, ngf=64):
        self.ops, self.store, self.ngf = ops, store, ngf

    def forward(self, image_nhwc, text_ids_host):
        """image NHWC [N,H,W,3] in [-1,1] (the reference graph is NHWC), H and W multiples of 32; ids [N,T] host ints.
        Returns (picture NHWC [N,H,W,3] in (-1,1), region logits NHWC [N,H,W,3])."""
        ops, st, p = self.ops, self.store, "generator"
        L = Layers(ops, st, need_wgrad=False)
        L.begin(image_nhwc.shape[0], image_nhwc.device)
        h, _ = L.unit_fwd("c7s2", p + "/encoder_1", [image_nhwc], "lrelu")                               # :320-325
        z = [h]
        for lvl in range(4):                                                                              # :334-344
            h, _ = L.block_fwd("en", "%s/encoder_%d_0" % (p, lvl + 2), [z[-1]], "lrelu")
            for u in range(1, UNITS[lvl]):
                h, _ = L.block_fwd("pu", "%s/encoder_%d_%d" % (p, lvl + 2, u), [h], "lrelu")
            z.append(h)
        feat, _ = text_fusion.text_fusion_fwd(ops, st, z[4], text_ids_host, save=False, prefix=BG_TEXT_SCOPE)   # :346-352
        r, _ = L.unit_fwd("c1", p + "/region_br_projection", [z[4]], "relu")                             # :358-363
        d = None
        for i in range(4):                                                                                # :374-402
            skip = 4 - i
            srcs = [feat] if i == 0 else [d, z[skip]]
            d, _ = L.block_fwd("de", "%s/decoder_%d_0" % (p, skip + 1), srcs, "relu")
            for u in range(1, UNITS[skip - 1]):
                d, _ = L.block_fwd("pu", "%s/decoder_%d_%d" % (p, skip + 1, u), [d], "relu")
            r, _ = L.unit_fwd("deconv", "%s/region_br_%d" % (p, skip + 1), [r], "relu")
        y, _ = L.unit_fwd("deconv", p + "/decoder_1", [d, z[0]], None)                                   # :405-412
        r, _ = L.unit_fwd("deconv", p + "/region_br_1", [r], "relu")                                     # :414-419
        return ops.tanh_fwd(y), r


class BgColorModel:
    """The generator and its parameter store on one device, with the test loop's uint8 boundary."""

    def __init__(self, ops, device, *, ngf=64, vocab_size=18, seg_classes=3, param_dtype=torch.float32):
        self.ops, self.device = ops, device
        self.gstore = ParamStore(bg_generator_vars(ngf, vocab_size, seg_classes), device, param_dtype)
        self.dstore = None
        self.G = BgGenerator(ops, self.gstore, ngf)

    def initialize(self, seed=0, perturb_tables=0.0):
        self.gstore.initialize(seed, perturb_tables)

    def generate(self, image_nhwc, text_ids_host):
        """float boundary: image [N,H,W,3] in [-1,1] -> (picture, region logits), both fp32 NHWC."""
        x = image_nhwc.to(self.device)
        x = x if x.dtype == self.ops.act_dtype else self.ops.cast(x.float().contiguous(), self.ops.act_dtype)
        out, reg = self.G.forward(x.contiguous(), text_ids_host)
        return self.ops.cast(out, torch.float32), self.ops.cast(reg, torch.float32)

    def colorize_u8(self, input_u8, text_ids_host):
        """The test loop's feed / fetch (:752-774, 838-853): uint8 [1,H,W,3] -> (output picture uint8 [H,W,3], region class map
        int64 [H,W]).  tf.image.convert_image_dtype: uint8 -> float is x / 255; float -> uint8 with saturate is
        trunc(clip(x * 255.5, 0, 255))."""
        x = torch.from_numpy(np.ascontiguousarray(input_u8)).float() / 255.0 * 2.0 - 1.0
        out, reg = self.generate(x, text_ids_host)
        pic = ((out[0].float().cpu().numpy() + 1.0) / 2.0) * 255.5
        return np.clip(pic, 0, 255).astype(np.uint8), reg[0].float().cpu().numpy().argmax(axis=2)
