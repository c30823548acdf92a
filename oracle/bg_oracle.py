"""CPU oracle for the background-colorization generator (BASELINE.json configs[3]: 768x768 inference).

TEST INFRASTRUCTURE ONLY.  Nothing under ``sketchyscenecolorization_b200/`` may import this module.

PARITY UNPINNED, like the other oracles: a torch-CPU functional restatement written from
Background_Colorization/bg_colorization_main.py -- conv / conv_ex / deconv / lrelu / batchnorm (:41-98), encode_feat_with_text
(:117-214, the fg function under scope `mLSTM_G`), bottleneck_residual_en / _de / _pu (:217-299, the same blocks as
obj_lib/residual_util.py), create_residual_generator (:302-420) with residual_enc_g = multi_residual = True (:733-736).  The
reference graph is NHWC; this restatement runs NCHW and takes / returns NCHW tensors.
"""
from __future__ import annotations

import torch

from .fgcolor_oracle import PSpec, conv2d, encode_feat_with_text
from .pix2pix_oracle import lrelu, nchw_deconv
from .residual_oracle import (UNITS, _block_specs, _bn, _bnp, _filter, bottleneck_residual_de, bottleneck_residual_en,
                              bottleneck_residual_pu)

TEXT_SCOPE = "generator/mLSTM_G"


def enc_channels(ngf=64):             # :328-333
    return [ngf * 2, ngf * 4, ngf * 8, ngf * 16]


def dec_channels(ngf=64):             # :367-372
    return [ngf * 8, ngf * 4, ngf * 2, ngf]


def generator_specs(ngf=64, vocab_size=18, seg_classes=3):
    """Variables of `generator/` (:585).  Every batch norm is bg_colorization_main.batchnorm, i.e. scoped (:86-98)."""
    p = "generator"
    sp = _filter(p + "/encoder_1", "conv_ex", 7, 3, ngf) + _bn(p + "/encoder_1", ngf)
    cin = ngf
    for lvl, co in enumerate(enc_channels(ngf)):
        sp += _block_specs("%s/encoder_%d_0" % (p, lvl + 2), "en", cin, co)
        for u in range(1, UNITS[lvl]):
            sp += _block_specs("%s/encoder_%d_%d" % (p, lvl + 2, u), "pu", co, co)
        cin = co
    d = cin
    sp.append(PSpec(TEXT_SCOPE + "/embedding", (vocab_size, d), ("uniform", 0.08)))
    for cell, kin in (("WLSTM", 2 * d), ("ALSTM", 4 * d)):
        base = TEXT_SCOPE + "/RNN/%s/multi_rnn_cell/cell_0/basic_lstm_cell" % cell
        sp.append(PSpec(base + "/kernel", (kin, 4 * d), ("glorot_uniform", None)))
        sp.append(PSpec(base + "/bias", (4 * d,), ("const", 0.0)))
    sp += _filter(p + "/region_br_projection", "conv_ex", 1, d, seg_classes) + _bn(p + "/region_br_projection", seg_classes)
    skip_ch = [ngf] + enc_channels(ngf)
    for i, co in enumerate(dec_channels(ngf)):
        skip = 4 - i
        sp += _block_specs("%s/decoder_%d_0" % (p, skip + 1), "de", cin, co)
        for u in range(1, UNITS[skip - 1]):
            sp += _block_specs("%s/decoder_%d_%d" % (p, skip + 1, u), "pu", co, co)
        sp += _filter("%s/region_br_%d" % (p, skip + 1), "deconv", 4, seg_classes, seg_classes)
        sp += _bn("%s/region_br_%d" % (p, skip + 1), seg_classes)
        cin = co + skip_ch[skip - 1]
    sp += _filter(p + "/decoder_1", "deconv", 4, cin, 3) + _bn(p + "/decoder_1", 3)
    sp += _filter(p + "/region_br_1", "deconv", 4, seg_classes, seg_classes) + _bn(p + "/region_br_1", seg_classes)
    return sp


def generator_forward(params, image, ids):
    """create_residual_generator (:302-420).  image [N,3,H,W] in [-1,1] (the foreground picture), ids [N,T] ->
    (background picture [N,3,H,W] in (-1,1), region logits [N,3,H,W] >= 0)."""
    p = "generator"
    h = lrelu(_bnp(params, p + "/encoder_1", conv2d(image, params[p + "/encoder_1/conv_ex/filter"], None, 2)), 0.2)   # :320-325
    z = [h]
    for lvl in range(4):                                                                                            # :334-344
        h = bottleneck_residual_en(params, "%s/encoder_%d_0" % (p, lvl + 2), z[-1])
        for u in range(1, UNITS[lvl]):
            h = bottleneck_residual_pu(params, "%s/encoder_%d_%d" % (p, lvl + 2, u), h, True)
        z.append(h)
    feat = encode_feat_with_text(params, z[-1], ids, TEXT_SCOPE)                                                    # :346-352
    r = conv2d(z[-1], params[p + "/region_br_projection/conv_ex/filter"], None)                                     # :358-363
    r = torch.relu(_bnp(params, p + "/region_br_projection", r))
    for i in range(4):                                                                                              # :374-402
        skip = 4 - i
        inp = feat if i == 0 else torch.cat([z[-1], z[skip]], 1)
        h = bottleneck_residual_de(params, "%s/decoder_%d_0" % (p, skip + 1), inp)
        for u in range(1, UNITS[skip - 1]):
            h = bottleneck_residual_pu(params, "%s/decoder_%d_%d" % (p, skip + 1, u), h, False)
        z.append(h)
        s = "%s/region_br_%d" % (p, skip + 1)
        r = torch.relu(_bnp(params, s, nchw_deconv(r, params[s + "/deconv/filter"])))
    inp = torch.cat([z[-1], z[0]], 1)                                                                               # :405-412
    out = torch.tanh(_bnp(params, p + "/decoder_1", nchw_deconv(inp, params[p + "/decoder_1/deconv/filter"])))
    s = p + "/region_br_1"                                                                                          # :414-419
    r = torch.relu(_bnp(params, s, nchw_deconv(r, params[s + "/deconv/filter"])))
    return out, r
