"""CPU oracle for `--block_type Pix2Pix`: the U-Net generator and PatchGAN discriminator variants.

TEST INFRASTRUCTURE ONLY.  Nothing under ``sketchyscenecolorization_b200/`` may import this module.

PARITY UNPINNED, like oracle/fgcolor_oracle.py: a torch-CPU functional restatement (NCHW, TF filter layouts, TF variable
names) written from the reference files cited on each function; TensorFlow cannot run here and the reference holds no test
or golden vector for these networks.

Reference (Foreground_Instance_Colorization/obj_lib/models_collection.py): nchw_conv (:380-391), nchw_deconv (:394-405),
batchnorm non-conditional branch (:36-46), lrelu (:51-53), image_encoder_pix2pix (:408-441), generate_pix2pix (:444-538),
discriminate_pix2pix (:789-841); text fusion, noise FC, spectral norm and the losses are shared with the MRU variant
(fgcolor_oracle).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .fgcolor_oracle import (NOISE_DIM, NUM_CLASSES, SIZE, PSpec, encode_feat_with_text, losses, miu_relu, reg_loss,
                             spectral_normed_weight)

N_ENC = 5


def enc_channels(size=SIZE):          # :421-432
    return [size, size * 2, size * 4, size * 8, size * 8]


def dec_channels(size=SIZE):          # :498-503, decoder_5 .. decoder_2 (decoder_1 -> output_channel)
    return [size * 8, size * 4, size * 2, size]


def _bn_specs(scope, c):
    # :40-42 -- `offset` zeros, `scale` N(1, 0.02), created directly in the layer's variable scope
    return [PSpec(scope + "/offset", (c,), ("const", 0.0)), PSpec(scope + "/scale", (c,), ("normal1", 0.02))]


def generator_specs(size=SIZE, vocab_size=58, H=192, W=192):
    """Variables of `generator/` in creation order.  Filters have no bias and no regulariser (:384-385,398-399)."""
    assert H % 32 == 0 and W % 32 == 0
    p = "generator"
    ch = enc_channels(size)
    sp = [PSpec(p + "/encoder_1/conv/filter", (4, 4, 3, ch[0]), ("normal", 0.02))]
    for k in range(2, N_ENC + 1):
        sp.append(PSpec(p + "/encoder_%d/conv/filter" % k, (4, 4, ch[k - 2], ch[k - 1]), ("normal", 0.02)))
        sp += _bn_specs(p + "/encoder_%d" % k, ch[k - 1])
    d = ch[4]
    sp.append(PSpec(p + "/TextLSTM/embedding", (vocab_size, d), ("uniform", 0.08)))
    for cell, kin in (("WLSTM", 2 * d), ("ALSTM", 4 * d)):
        base = p + "/TextLSTM/RNN/%s/multi_rnn_cell/cell_0/basic_lstm_cell" % cell
        sp.append(PSpec(base + "/kernel", (kin, 4 * d), ("glorot_uniform", None)))
        sp.append(PSpec(base + "/bias", (4 * d,), ("const", 0.0)))
    nfc = (d // 8) * (H // 32) * (W // 32)                               # :479-485: noise at the bottleneck resolution
    sp.append(PSpec(p + "/fully_connected/weights", (NOISE_DIM, nfc), ("xavier", None), reg=1e-6))
    sp.append(PSpec(p + "/fully_connected/biases", (nfc,), ("const", 0.0)))
    cin = d + d // 8
    for i, co in enumerate(dec_channels(size)):
        k = N_ENC - i                                                    # decoder_5 .. decoder_2
        sp.append(PSpec(p + "/decoder_%d/deconv/filter" % k, (4, 4, co, cin), ("normal", 0.02)))
        sp += _bn_specs(p + "/decoder_%d" % k, co)
        cin = co + ch[k - 2]                                             # next input: [this output, encoder_(k-1) output]
    sp.append(PSpec(p + "/decoder_1/deconv/filter", (4, 4, 3, cin), ("normal", 0.02)))
    return sp


def discriminator_specs(size=SIZE):
    p = "discriminator"
    chans = [6, size, size * 2, size * 4, size * 8, 1]
    sp = []
    for k in range(1, 6):
        sp.append(PSpec(p + "/layer_%d/conv/filter" % k, (4, 4, chans[k - 1], chans[k]), ("normal", 0.02)))
        if 2 <= k <= 4:
            sp += _bn_specs(p + "/layer_%d" % k, chans[k])
    fc = p + "/fully_connected"
    sp.append(PSpec(fc + "/weights", (chans[4], NUM_CLASSES), ("xavier", None), reg=1e-6, sn=True))
    sp.append(PSpec(fc + "/" + fc + "/u", (1, NUM_CLASSES), ("trunc_normal", 1.0), trainable=False))
    sp.append(PSpec(fc + "/biases", (NUM_CLASSES,), ("const", 0.0)))
    return sp


def nchw_conv(x, filt, stride):
    """:380-391 -- tf.pad by 1 on every side, then 4x4 VALID cross-correlation; filter HWIO."""
    return F.conv2d(F.pad(x, (1, 1, 1, 1)), filt.permute(3, 2, 0, 1), stride=stride)


def nchw_deconv(x, filt):
    """:394-405 -- tf.nn.conv2d_transpose, 4x4, stride 2, SAME, output 2H x 2W; filter [kh, kw, out, in].  It is the input
    gradient of the SAME stride-2 convolution (pad 1 top/left, 1 bottom/right) that maps the output back to the input."""
    return F.conv_transpose2d(x, filt.permute(3, 2, 0, 1), stride=2, padding=1)


def batchnorm(x, offset, scale):
    """:36-46 -- batch statistics over N,H,W (biased variance), eps 1e-5."""
    mean = x.mean(dim=(0, 2, 3), keepdim=True)
    var = ((x - mean) ** 2).mean(dim=(0, 2, 3), keepdim=True)
    return (x - mean) * torch.rsqrt(var + 1e-5) * scale.reshape(1, -1, 1, 1) + offset.reshape(1, -1, 1, 1)


def lrelu(x, leak):                   # :51-53
    return torch.maximum(leak * x, x)


def generator_forward(params, sketch, ids, labels, noise, size=SIZE, lstm_hybrid=True, return_taps=False):
    """generate_pix2pix (:444-538).  `labels` is accepted and unused, as in the reference (plain batch norm)."""
    p = "generator"
    enc = [nchw_conv(sketch, params[p + "/encoder_1/conv/filter"], 2)]                       # :423-425
    for k in range(2, N_ENC + 1):                                                             # :434-439
        c = nchw_conv(lrelu(enc[-1], 0.2), params[p + "/encoder_%d/conv/filter" % k], 2)
        enc.append(batchnorm(c, params[p + "/encoder_%d/offset" % k], params[p + "/encoder_%d/scale" % k]))
    feat = encode_feat_with_text(params, enc[-1], ids) if lstm_hybrid else enc[-1]            # :474-477
    N, d, hh, ww = enc[-1].shape
    nz = miu_relu(noise @ params[p + "/fully_connected/weights"] + params[p + "/fully_connected/biases"])
    nz = nz.reshape(N, d // 8, hh, ww)                                                        # :479-492
    taps = {"enc%d" % (i + 1): e for i, e in enumerate(enc)}
    taps["text"] = feat
    z = list(enc)
    for i in range(len(dec_channels(size))):                                                  # :505-521
        k = N_ENC - i
        inp = torch.cat([feat, nz], 1) if i == 0 else torch.cat([z[-1], z[k - 1]], 1)
        out = nchw_deconv(torch.relu(inp), params[p + "/decoder_%d/deconv/filter" % k])
        out = batchnorm(out, params[p + "/decoder_%d/offset" % k], params[p + "/decoder_%d/scale" % k])
        z.append(out)
        taps["dec%d" % k] = out
    inp = torch.cat([z[-1], z[0]], 1)                                                         # :524-529
    out = torch.tanh(nchw_deconv(torch.relu(inp), params[p + "/decoder_1/deconv/filter"]))
    if return_taps:
        return out, taps
    return out


def discriminator_forward(params, sketch, image, size=SIZE, return_u=False):
    """discriminate_pix2pix (:789-841): looks at (sketch, image) pairs; plain convs, BN on layers 2-4, SN only on the class
    head (Config.sn).  Returns (patch logits [N,1,h-2,w-2] at h = H/8, class logits [N,25])."""
    p = "discriminator"
    x = torch.cat([sketch, image], 1)                                                         # :811
    h = lrelu(nchw_conv(x, params[p + "/layer_1/conv/filter"], 2), 0.2)                      # :814-817
    for k in (2, 3, 4):                                                                       # :822-829
        c = nchw_conv(h, params[p + "/layer_%d/conv/filter" % k], 1 if k == 4 else 2)
        h = lrelu(batchnorm(c, params[p + "/layer_%d/offset" % k], params[p + "/layer_%d/scale" % k]), 0.2)
    disc = nchw_conv(h, params[p + "/layer_5/conv/filter"], 1)                                # :832-833
    fc = p + "/fully_connected"
    w, u_new = spectral_normed_weight(params[fc + "/weights"], params[fc + "/" + fc + "/u"])
    logits = h.mean(dim=(2, 3)) @ w + params[fc + "/biases"]                                  # :836-837
    if return_u:
        return disc, logits, {fc + "/" + fc + "/u": u_new.detach()}
    return disc, logits


def d_step_loss(gp, dp, gspecs, dspecs, batch, size=SIZE):
    """sess.run([opt_d, loss_d]) with block_type Pix2Pix: graph_single.py:269-272 (two discriminator instantiations on
    (sketches, images_d) and (sketches, image_gens)), losses as in the MRU variant."""
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
