"""CPU oracle for `--block_type Residual`: the bottleneck-residual U-Net generator and discriminator.

TEST INFRASTRUCTURE ONLY.  Nothing under ``sketchyscenecolorization_b200/`` may import this module.

PARITY UNPINNED, like the other oracles: a torch-CPU functional restatement (NCHW, TF filter layouts, TF variable names)
written from the reference files cited on each function; TensorFlow cannot run here and the reference holds no test or golden
vector for these networks.

Reference (Foreground_Instance_Colorization/obj_lib/): residual_util.py -- conv (:16-25, 4x4 with explicit pad 1), conv_ex
(:28-34, SAME, default 4x4), batchnorm (:56-68), deconv (:71-80), bottleneck_residual_en / _de / _pu (:83-175);
models_collection.py -- image_encoder_residual (:541-576), generate_residual (:579-672), discriminate_residual (:844-893).
"""
from __future__ import annotations

import torch

from .fgcolor_oracle import (NOISE_DIM, NUM_CLASSES, SIZE, PSpec, conv2d, encode_feat_with_text, losses, miu_relu, reg_loss,
                             spectral_normed_weight)
from .pix2pix_oracle import batchnorm, lrelu, nchw_conv, nchw_deconv

UNITS = [3, 4, 6, 3]                  # models_collection.py:609


def enc_channels(size=SIZE):          # :562-567 (after the 7x7 stem of `size` channels)
    return [size * 2, size * 4, size * 8, size * 8]


def dec_channels(size=SIZE):          # :637-642
    return [size * 8, size * 4, size * 2, size]


# --------------------------------------------------------------------------------------
# variables
# --------------------------------------------------------------------------------------
def _filter(scope, kind, k, cin, cout):
    """kind 'conv' | 'conv_ex': [k,k,cin,cout]; 'deconv': [4,4,cout,cin] (residual_util.py:19,31,74); N(0, 0.02), no bias."""
    shape = (4, 4, cout, cin) if kind == "deconv" else (k, k, cin, cout)
    return [PSpec("%s/%s/filter" % (scope, kind), shape, ("normal", 0.02))]


def _bn(scope, c, scoped=True):
    """residual_util.batchnorm lives in variable_scope('batchnorm') (:57); models_collection.batchnorm creates `offset` /
    `scale` directly in the caller's scope (models_collection.py:40-42)."""
    s = scope + "/batchnorm" if scoped else scope
    return [PSpec(s + "/offset", (c,), ("const", 0.0)), PSpec(s + "/scale", (c,), ("normal1", 0.02))]


def _block_specs(scope, kind, cin, cout):
    q = int(round(cout / 4))
    sp = []
    if kind == "en":
        sp += _filter(scope + "/block_1", "conv", 4, cin, q)
    elif kind == "de":
        sp += _filter(scope + "/block_1", "deconv", 4, cin, q)
    else:
        sp += _filter(scope + "/block_1", "conv_ex", 4, cin, q)
    sp += _bn(scope + "/block_1", q)
    sp += _filter(scope + "/block_2", "conv_ex", 3, q, q) + _bn(scope + "/block_2", q)
    sp += _filter(scope + "/block_3", "conv_ex", 1, q, cout) + _bn(scope + "/block_3", cout)
    if kind == "en":
        sp += _filter(scope + "/block_add", "conv", 4, cin, cout) + _bn(scope + "/block_add", cout)
    elif kind == "de":
        sp += _filter(scope + "/block_add", "deconv", 4, cin, cout) + _bn(scope + "/block_add", cout)
    return sp


def generator_specs(size=SIZE, vocab_size=58, H=192, W=192):
    assert H % 32 == 0 and W % 32 == 0
    p = "generator"
    sp = _filter(p + "/encoder_1", "conv_ex", 7, 3, size) + _bn(p + "/encoder_1", size, scoped=False)
    cin = size
    for lvl, co in enumerate(enc_channels(size)):
        sp += _block_specs("%s/encoder_%d_0" % (p, lvl + 2), "en", cin, co)
        for u in range(1, UNITS[lvl]):
            sp += _block_specs("%s/encoder_%d_%d" % (p, lvl + 2, u), "pu", co, co)
        cin = co
    d = cin
    sp.append(PSpec(p + "/TextLSTM/embedding", (vocab_size, d), ("uniform", 0.08)))
    for cell, kin in (("WLSTM", 2 * d), ("ALSTM", 4 * d)):
        base = p + "/TextLSTM/RNN/%s/multi_rnn_cell/cell_0/basic_lstm_cell" % cell
        sp.append(PSpec(base + "/kernel", (kin, 4 * d), ("glorot_uniform", None)))
        sp.append(PSpec(base + "/bias", (4 * d,), ("const", 0.0)))
    nfc = (d // 8) * (H // 32) * (W // 32)
    sp.append(PSpec(p + "/fully_connected/weights", (NOISE_DIM, nfc), ("xavier", None), reg=1e-6))
    sp.append(PSpec(p + "/fully_connected/biases", (nfc,), ("const", 0.0)))
    skip_ch = [size] + enc_channels(size)            # channels of z_encoded[0..4]
    cin = d + d // 8
    for i, co in enumerate(dec_channels(size)):
        skip = 4 - i                                 # :645-646
        sp += _block_specs("%s/decoder_%d_0" % (p, skip + 1), "de", cin, co)
        for u in range(1, UNITS[skip - 1]):          # :653
            sp += _block_specs("%s/decoder_%d_%d" % (p, skip + 1, u), "pu", co, co)
        cin = co + skip_ch[skip - 1]
    sp += _filter(p + "/decoder_1", "deconv", 4, cin, 3) + _bn(p + "/decoder_1", 3, scoped=False)
    return sp


def discriminator_specs(size=SIZE):
    p = "discriminator"
    chans = [6, size, size * 2, size * 4, size * 8, 512]
    sp = []
    for k in range(1, 6):
        sp += _block_specs("%s/layer_%d" % (p, k), "en", chans[k - 1], chans[k])
    sp += _filter(p + "/layer_5", "conv_ex", 4, 512, 1)
    fc = p + "/fully_connected"
    sp.append(PSpec(fc + "/weights", (chans[4], NUM_CLASSES), ("xavier", None), reg=1e-6, sn=True))
    sp.append(PSpec(fc + "/" + fc + "/u", (1, NUM_CLASSES), ("trunc_normal", 1.0), trainable=False))
    sp.append(PSpec(fc + "/biases", (NUM_CLASSES,), ("const", 0.0)))
    return sp


# --------------------------------------------------------------------------------------
# blocks (residual_util.py:83-175)
# --------------------------------------------------------------------------------------
def _bnp(params, scope, x, scoped=True):
    s = scope + "/batchnorm" if scoped else scope
    return batchnorm(x, params[s + "/offset"], params[s + "/scale"])


def _conv_ex(params, scope, x, stride=1):
    return conv2d(x, params[scope + "/conv_ex/filter"], None, stride)          # SAME; the odd pad pixel goes bottom / right


def bottleneck_residual_en(params, scope, x, stride=2):
    """:83-111"""
    assert stride == 2
    h = lrelu(_bnp(params, scope + "/block_1", nchw_conv(x, params[scope + "/block_1/conv/filter"], 2)), 0.2)
    h = lrelu(_bnp(params, scope + "/block_2", _conv_ex(params, scope + "/block_2", h)), 0.2)
    h = _bnp(params, scope + "/block_3", _conv_ex(params, scope + "/block_3", h))
    o = _bnp(params, scope + "/block_add", nchw_conv(x, params[scope + "/block_add/conv/filter"], 2))
    return lrelu(h + o, 0.2)


def bottleneck_residual_de(params, scope, x):
    """:114-144 (need_relu=True at every call site)"""
    h = torch.relu(_bnp(params, scope + "/block_1", nchw_deconv(x, params[scope + "/block_1/deconv/filter"])))
    h = torch.relu(_bnp(params, scope + "/block_2", _conv_ex(params, scope + "/block_2", h)))
    h = _bnp(params, scope + "/block_3", _conv_ex(params, scope + "/block_3", h))
    o = _bnp(params, scope + "/block_add", nchw_deconv(x, params[scope + "/block_add/deconv/filter"]))
    return torch.relu(h + o)


def bottleneck_residual_pu(params, scope, x, is_encoder):
    """:147-175 -- block_1 is conv_ex with its DEFAULT 4x4 filter, SAME."""
    act = (lambda t: lrelu(t, 0.2)) if is_encoder else torch.relu
    h = act(_bnp(params, scope + "/block_1", _conv_ex(params, scope + "/block_1", x)))
    h = act(_bnp(params, scope + "/block_2", _conv_ex(params, scope + "/block_2", h)))
    h = _bnp(params, scope + "/block_3", _conv_ex(params, scope + "/block_3", h))
    return act(h + x)


# --------------------------------------------------------------------------------------
# networks
# --------------------------------------------------------------------------------------
def generator_forward(params, sketch, ids, labels, noise, size=SIZE, lstm_hybrid=True):
    """generate_residual (:579-672); `labels` unused (plain batch norm)."""
    p = "generator"
    h = lrelu(_bnp(params, p + "/encoder_1", _conv_ex(params, p + "/encoder_1", sketch, 2), scoped=False), 0.2)     # :555-559
    z = [h]
    for lvl in range(4):                                                                                          # :568-574
        h = bottleneck_residual_en(params, "%s/encoder_%d_0" % (p, lvl + 2), z[-1])
        for u in range(1, UNITS[lvl]):
            h = bottleneck_residual_pu(params, "%s/encoder_%d_%d" % (p, lvl + 2, u), h, True)
        z.append(h)
    feat = encode_feat_with_text(params, z[-1], ids) if lstm_hybrid else z[-1]                                    # :620-625
    N, d, hh, ww = z[-1].shape
    nz = miu_relu(noise @ params[p + "/fully_connected/weights"] + params[p + "/fully_connected/biases"])
    nz = nz.reshape(N, d // 8, hh, ww)                                                                            # :627-634
    for i in range(4):                                                                                            # :644-658
        skip = 4 - i
        inp = torch.cat([feat, nz], 1) if i == 0 else torch.cat([z[-1], z[skip]], 1)
        h = bottleneck_residual_de(params, "%s/decoder_%d_0" % (p, skip + 1), inp)
        for u in range(1, UNITS[skip - 1]):
            h = bottleneck_residual_pu(params, "%s/decoder_%d_%d" % (p, skip + 1, u), h, False)
        z.append(h)
    inp = torch.cat([z[-1], z[0]], 1)                                                                             # :661-666
    out = _bnp(params, p + "/decoder_1", nchw_deconv(inp, params[p + "/decoder_1/deconv/filter"]), scoped=False)
    return torch.tanh(out)


def discriminator_forward(params, sketch, image, size=SIZE, return_u=False):
    """discriminate_residual (:844-893): five stride-2 bottleneck blocks over the (sketch, image) pair; patch logits from
    the fifth ([N,1,H/32,W/32]), class logits from the mean of the FOURTH block's output (`rectified`, :876-886)."""
    p = "discriminator"
    h = torch.cat([sketch, image], 1)
    for k in range(1, 5):
        h = bottleneck_residual_en(params, "%s/layer_%d" % (p, k), h)
    c = bottleneck_residual_en(params, p + "/layer_5", h)
    disc = _conv_ex(params, p + "/layer_5", c)
    fc = p + "/fully_connected"
    w, u_new = spectral_normed_weight(params[fc + "/weights"], params[fc + "/" + fc + "/u"])
    logits = h.mean(dim=(2, 3)) @ w + params[fc + "/biases"]
    if return_u:
        return disc, logits, {fc + "/" + fc + "/u": u_new.detach()}
    return disc, logits


def d_step_loss(gp, dp, gspecs, dspecs, batch, size=SIZE):
    with torch.no_grad():
        fake = generator_forward(gp, batch["sketch"], batch["text"], batch["cls"], batch["noise"], size)
    rd, rl, u_new = discriminator_forward(dp, batch["sketch"], batch["images_d"], size, return_u=True)
    fd, fl = discriminator_forward(dp, batch["sketch"], fake, size)
    _, loss_d, terms = losses(rd, rl, fd, fl, batch["cls_d"], batch["cls"], batch["images"], fake,
                              reg_loss(gp, gspecs), reg_loss(dp, dspecs))
    return loss_d, terms, u_new


def g_step_loss(gp, dp, gspecs, dspecs, batch, size=SIZE):
    fake = generator_forward(gp, batch["sketch"], batch["text"], batch["cls"], batch["noise"], size)
    rd, rl, u_new = discriminator_forward(dp, batch["sketch"], batch["images_d"], size, return_u=True)
    fd, fl = discriminator_forward(dp, batch["sketch"], fake, size)
    loss_g, _, terms = losses(rd, rl, fd, fl, batch["cls_d"], batch["cls"], batch["images"], fake,
                              reg_loss(gp, gspecs), reg_loss(dp, dspecs))
    return loss_g, terms, u_new, fake
