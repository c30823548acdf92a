"""Variable registry of the fg-colorization networks, with the reference's TF variable names.

Every trainable tensor lives in ONE flat fp32 buffer per network (`flat`), with a matching flat
gradient buffer and Adam second-moment buffer, so that the NCCL all-reduce and the fused Adam
kernel touch a single contiguous range (reference: graph_single.py:33-68 `average_gradients`,
:584-593 `get_optimizer`).  Per-tensor views keep the reference layout (conv `weights` HWIO
[k,k,Cin,Cout], `biases` [1,C,1,1], cBN tables [25,C]) so a state dict keyed by the reference
names round-trips without transposition.

Naming follows tf.variable_scope uniquification in the reference:
  generator: models_collection.py:68-147 (encoder), :150-248 (TextLSTM), :310-377 (decoder)
  discriminator: models_collection.py:676-786; SN `u` vectors: sn.py:17-18
  blocks: mru.py:353-461 (v3, scopes update_gate/Conv/Conv_1..3), mru.py:527-591 (v2, Conv..Conv_4)
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass

import torch

NUM_CLASSES = 25      # input_pipeline.py:11
NOISE_DIM = 256       # models_collection.py:310
ALIGN = 4             # every tensor starts on a 16-byte boundary inside the flat buffer


@dataclass
class VarSpec:
    name: str
    shape: tuple
    init: tuple            # (kind, arg)
    reg: float = 0.0       # l2_regularizer scale (loss = reg * sum(w^2)/2)
    sn: bool = False       # spectrally normalised (discriminator weights)
    trainable: bool = True

    @property
    def numel(self):
        return int(math.prod(self.shape)) if self.shape else 1


def _conv(scope, k, cin, cout, reg, sn, bias=0.0):
    v = [VarSpec(scope + "/weights", (k, k, cin, cout), ("normal", 0.02), reg, sn)]
    if sn:
        v.append(VarSpec(scope + "/" + scope + "/u", (1, cout), ("trunc_normal", 1.0), trainable=False))
    v.append(VarSpec(scope + "/biases", (1, cout, 1, 1), ("const", bias)))
    return v


def _cbn(scope, c):
    return [VarSpec(scope + "/offset", (NUM_CLASSES, c), ("const", 0.0)),
            VarSpec(scope + "/scale", (NUM_CLASSES, c), ("const", 1.0))]


def _prelu(scope):
    return [VarSpec(scope + "/prelu/param", (), ("const", 0.2))]


def _norm(scope, c, kind):
    return _cbn(scope, c) if kind == "cbn" else _prelu(scope)


def _enc_unit(prefix, unit, cin, cout, kind, sn):
    s = "%s/mru_conv_unit_t_%d_layer_0" % (prefix, unit)
    v = _norm(s + "/norm_activation_in", cin, kind)
    v += _conv(s + "/update_gate", 3, cin + 3, cin, 1e-5, sn, bias=0.5)
    v += _conv(s + "/Conv", 3, 3, cin, 1e-5, sn)
    v += _norm(s + "/norm_activation_merge_1", cin, kind)
    v += _conv(s + "/Conv_1", 3, cin, cout, 1e-5, sn)
    v += _norm(s + "/Conv_1", cout, kind)
    v += _conv(s + "/Conv_2", 3, cout, cout, 1e-5, sn)
    if cin != cout:
        v += _conv(s + "/Conv_3", 1, cin, cout, 1e-5, sn)
    return v


def encoder_channels(size):
    return [8, size, size * 2, size * 4, size * 8]


def disc_channels(size):
    return [8, size * 2, size * 4, size * 8, size * 12]


def decoder_plan(size):
    """(unit, [extra source channel counts after the 3-ch sketch], hidden C, out C)."""
    return [(0, size, size * 8, size * 6), (2, size * 2, size * 6, size * 4), (4, size, size * 4, size * 2),
            (6, 8, size * 2, size * 2), (8, 0, size * 2, size)]


def generator_vars(size=64, vocab_size=58, H=192, W=192):
    assert H % 32 == 0 and W % 32 == 0, "image size must be a multiple of 32"
    p = "generator"
    ch = encoder_channels(size)
    v = _conv(p + "/Conv", 7, 3, 8, 0.0, False)
    for u in range(1, 5):
        v += _enc_unit(p, u, ch[u - 1], ch[u], "cbn", False)
    v += _cbn(p + "/mru_conv_unit_last_norm", ch[4])
    d = ch[4]
    v.append(VarSpec(p + "/TextLSTM/embedding", (vocab_size, d), ("uniform", 0.08)))
    for cell, kin in (("WLSTM", 2 * d), ("ALSTM", 4 * d)):
        b = p + "/TextLSTM/RNN/%s/multi_rnn_cell/cell_0/basic_lstm_cell" % cell
        v.append(VarSpec(b + "/kernel", (kin, 4 * d), ("glorot_uniform", None)))
        v.append(VarSpec(b + "/bias", (4 * d,), ("const", 0.0)))
    nfc = (d // 8) * (H // 16) * (W // 16)
    v.append(VarSpec(p + "/fully_connected/weights", (NOISE_DIM, nfc), ("xavier", None), reg=1e-6))
    v.append(VarSpec(p + "/fully_connected/biases", (nfc,), ("const", 0.0)))
    for (u, cx, chid, cout) in decoder_plan(size):
        s = "%s/mru_deconv_unit_t_%d_layer_0" % (p, u)
        cin = chid + 3 + cx
        v += _conv(s + "/Conv", 3, cin, chid, 1e-5, False)
        v += _conv(s + "/Conv_1", 3, cin, cout, 1e-5, False)
        v += _conv(s + "/Conv_2", 3, cin, cout, 1e-5, False) + _cbn(s + "/Conv_2", cout)
        v += _conv(s + "/Conv_3", 3, cout, cout, 1e-5, False) + _cbn(s + "/Conv_3", cout)
        if chid != cout:
            v += _conv(s + "/Conv_4", 1, chid, cout, 0.0, False) + _cbn(s + "/Conv_4", cout)
    v += _conv(p + "/Conv_1", 7, size, 3, 0.0, False)
    return v


def discriminator_vars(size=64):
    p = "discriminator"
    ch = disc_channels(size)
    v = _conv(p + "/Conv", 7, 3, 8, 0.0, True) + _prelu(p + "/Conv")
    for u in range(1, 5):
        v += _enc_unit(p, u, ch[u - 1], ch[u], "prelu", True)
    v += _prelu(p + "/mru_conv_unit_last_norm")
    v += _conv(p + "/Conv_1", 1, ch[4], 1, 0.0, True)
    fc = p + "/fully_connected"
    v.append(VarSpec(fc + "/weights", (ch[4], NUM_CLASSES), ("xavier", None), reg=1e-6, sn=True))
    v.append(VarSpec(fc + "/" + fc + "/u", (1, NUM_CLASSES), ("trunc_normal", 1.0), trainable=False))
    v.append(VarSpec(fc + "/biases", (NUM_CLASSES,), ("const", 0.0)))
    return v


# ------------------------------------------------------------------------------------------------
# --block_type Pix2Pix (models_collection.py:380-538, 789-841): 4x4 filters without bias or regulariser, plain batch norm
# (`offset` zeros, `scale` N(1, 0.02), created directly in the layer scope), the text LSTM / noise FC / SN class head of
# the MRU variant under the same names
# ------------------------------------------------------------------------------------------------
def pix2pix_enc_channels(size):
    return [size, size * 2, size * 4, size * 8, size * 8]


def pix2pix_dec_channels(size):
    return [size * 8, size * 4, size * 2, size]


def _bn(scope, c):
    return [VarSpec(scope + "/offset", (c,), ("const", 0.0)), VarSpec(scope + "/scale", (c,), ("normal1", 0.02))]


def pix2pix_generator_vars(size=64, vocab_size=58, H=192, W=192):
    assert H % 32 == 0 and W % 32 == 0, "image size must be a multiple of 32"
    p = "generator"
    ch = pix2pix_enc_channels(size)
    v = [VarSpec(p + "/encoder_1/conv/filter", (4, 4, 3, ch[0]), ("normal", 0.02))]
    for k in range(2, 6):
        v.append(VarSpec(p + "/encoder_%d/conv/filter" % k, (4, 4, ch[k - 2], ch[k - 1]), ("normal", 0.02)))
        v += _bn(p + "/encoder_%d" % k, ch[k - 1])
    d = ch[4]
    v.append(VarSpec(p + "/TextLSTM/embedding", (vocab_size, d), ("uniform", 0.08)))
    for cell, kin in (("WLSTM", 2 * d), ("ALSTM", 4 * d)):
        b = p + "/TextLSTM/RNN/%s/multi_rnn_cell/cell_0/basic_lstm_cell" % cell
        v.append(VarSpec(b + "/kernel", (kin, 4 * d), ("glorot_uniform", None)))
        v.append(VarSpec(b + "/bias", (4 * d,), ("const", 0.0)))
    nfc = (d // 8) * (H // 32) * (W // 32)
    v.append(VarSpec(p + "/fully_connected/weights", (NOISE_DIM, nfc), ("xavier", None), reg=1e-6))
    v.append(VarSpec(p + "/fully_connected/biases", (nfc,), ("const", 0.0)))
    cin = d + d // 8
    for i, co in enumerate(pix2pix_dec_channels(size)):
        k = 5 - i
        v.append(VarSpec(p + "/decoder_%d/deconv/filter" % k, (4, 4, co, cin), ("normal", 0.02)))
        v += _bn(p + "/decoder_%d" % k, co)
        cin = co + ch[k - 2]
    v.append(VarSpec(p + "/decoder_1/deconv/filter", (4, 4, 3, cin), ("normal", 0.02)))
    return v


def pix2pix_discriminator_vars(size=64):
    p = "discriminator"
    chans = [6, size, size * 2, size * 4, size * 8, 1]
    v = []
    for k in range(1, 6):
        v.append(VarSpec(p + "/layer_%d/conv/filter" % k, (4, 4, chans[k - 1], chans[k]), ("normal", 0.02)))
        if 2 <= k <= 4:
            v += _bn(p + "/layer_%d" % k, chans[k])
    fc = p + "/fully_connected"
    v.append(VarSpec(fc + "/weights", (chans[4], NUM_CLASSES), ("xavier", None), reg=1e-6, sn=True))
    v.append(VarSpec(fc + "/" + fc + "/u", (1, NUM_CLASSES), ("trunc_normal", 1.0), trainable=False))
    v.append(VarSpec(fc + "/biases", (NUM_CLASSES,), ("const", 0.0)))
    return v


# ------------------------------------------------------------------------------------------------
# --block_type Residual (residual_util.py:16-175; models_collection.py:541-672, 844-893).  residual_util.batchnorm keeps its
# tables in a `batchnorm` sub-scope (:57); the generator's first and last layer use models_collection.batchnorm (no sub-scope)
# ------------------------------------------------------------------------------------------------
RESIDUAL_UNITS = [3, 4, 6, 3]          # models_collection.py:609


def _res_filter(scope, kind, k, cin, cout):
    shape = (4, 4, cout, cin) if kind == "deconv" else (k, k, cin, cout)
    return [VarSpec("%s/%s/filter" % (scope, kind), shape, ("normal", 0.02))]


def _res_block(scope, kind, cin, cout):
    q = int(round(cout / 4))
    first = {"en": "conv", "de": "deconv", "pu": "conv_ex"}[kind]
    v = _res_filter(scope + "/block_1", first, 4, cin, q) + _bn(scope + "/block_1/batchnorm", q)
    v += _res_filter(scope + "/block_2", "conv_ex", 3, q, q) + _bn(scope + "/block_2/batchnorm", q)
    v += _res_filter(scope + "/block_3", "conv_ex", 1, q, cout) + _bn(scope + "/block_3/batchnorm", cout)
    if kind != "pu":
        v += _res_filter(scope + "/block_add", first, 4, cin, cout) + _bn(scope + "/block_add/batchnorm", cout)
    return v


def residual_generator_vars(size=64, vocab_size=58, H=192, W=192):
    assert H % 32 == 0 and W % 32 == 0, "image size must be a multiple of 32"
    p = "generator"
    ench = [size * 2, size * 4, size * 8, size * 8]
    v = _res_filter(p + "/encoder_1", "conv_ex", 7, 3, size) + _bn(p + "/encoder_1", size)
    cin = size
    for lvl, co in enumerate(ench):
        v += _res_block("%s/encoder_%d_0" % (p, lvl + 2), "en", cin, co)
        for u in range(1, RESIDUAL_UNITS[lvl]):
            v += _res_block("%s/encoder_%d_%d" % (p, lvl + 2, u), "pu", co, co)
        cin = co
    d = cin
    v.append(VarSpec(p + "/TextLSTM/embedding", (vocab_size, d), ("uniform", 0.08)))
    for cell, kin in (("WLSTM", 2 * d), ("ALSTM", 4 * d)):
        b = p + "/TextLSTM/RNN/%s/multi_rnn_cell/cell_0/basic_lstm_cell" % cell
        v.append(VarSpec(b + "/kernel", (kin, 4 * d), ("glorot_uniform", None)))
        v.append(VarSpec(b + "/bias", (4 * d,), ("const", 0.0)))
    nfc = (d // 8) * (H // 32) * (W // 32)
    v.append(VarSpec(p + "/fully_connected/weights", (NOISE_DIM, nfc), ("xavier", None), reg=1e-6))
    v.append(VarSpec(p + "/fully_connected/biases", (nfc,), ("const", 0.0)))
    skip_ch = [size] + ench
    cin = d + d // 8
    for i, co in enumerate([size * 8, size * 4, size * 2, size]):
        skip = 4 - i
        v += _res_block("%s/decoder_%d_0" % (p, skip + 1), "de", cin, co)
        for u in range(1, RESIDUAL_UNITS[skip - 1]):
            v += _res_block("%s/decoder_%d_%d" % (p, skip + 1, u), "pu", co, co)
        cin = co + skip_ch[skip - 1]
    v += _res_filter(p + "/decoder_1", "deconv", 4, cin, 3) + _bn(p + "/decoder_1", 3)
    return v


def residual_discriminator_vars(size=64):
    p = "discriminator"
    chans = [6, size, size * 2, size * 4, size * 8, 512]
    v = []
    for k in range(1, 6):
        v += _res_block("%s/layer_%d" % (p, k), "en", chans[k - 1], chans[k])
    v += _res_filter(p + "/layer_5", "conv_ex", 4, 512, 1)
    fc = p + "/fully_connected"
    v.append(VarSpec(fc + "/weights", (chans[4], NUM_CLASSES), ("xavier", None), reg=1e-6, sn=True))
    v.append(VarSpec(fc + "/" + fc + "/u", (1, NUM_CLASSES), ("trunc_normal", 1.0), trainable=False))
    v.append(VarSpec(fc + "/biases", (NUM_CLASSES,), ("const", 0.0)))
    return v


# ------------------------------------------------------------------------------------------------
# background-colorization generator (Background_Colorization/bg_colorization_main.py:302-420, scope `generator`, :585):
# the residual generator with a 1024-channel fifth level, the caption cells under generator/mLSTM_G, no noise input, a
# region-segmentation branch, and EVERY batch norm in its `batchnorm` sub-scope (:86-98)
# ------------------------------------------------------------------------------------------------
BG_TEXT_SCOPE = "generator/mLSTM_G"


def bg_generator_vars(ngf=64, vocab_size=18, seg_classes=3):
    p = "generator"
    ench = [ngf * 2, ngf * 4, ngf * 8, ngf * 16]
    v = _res_filter(p + "/encoder_1", "conv_ex", 7, 3, ngf) + _bn(p + "/encoder_1/batchnorm", ngf)
    cin = ngf
    for lvl, co in enumerate(ench):
        v += _res_block("%s/encoder_%d_0" % (p, lvl + 2), "en", cin, co)
        for u in range(1, RESIDUAL_UNITS[lvl]):
            v += _res_block("%s/encoder_%d_%d" % (p, lvl + 2, u), "pu", co, co)
        cin = co
    d = cin
    v.append(VarSpec(BG_TEXT_SCOPE + "/embedding", (vocab_size, d), ("uniform", 0.08)))
    for cell, kin in (("WLSTM", 2 * d), ("ALSTM", 4 * d)):
        b = BG_TEXT_SCOPE + "/RNN/%s/multi_rnn_cell/cell_0/basic_lstm_cell" % cell
        v.append(VarSpec(b + "/kernel", (kin, 4 * d), ("glorot_uniform", None)))
        v.append(VarSpec(b + "/bias", (4 * d,), ("const", 0.0)))
    v += _res_filter(p + "/region_br_projection", "conv_ex", 1, d, seg_classes) + _bn(p + "/region_br_projection/batchnorm", seg_classes)
    skip_ch = [ngf] + ench
    for i, co in enumerate([ngf * 8, ngf * 4, ngf * 2, ngf]):
        skip = 4 - i
        v += _res_block("%s/decoder_%d_0" % (p, skip + 1), "de", cin, co)
        for u in range(1, RESIDUAL_UNITS[skip - 1]):
            v += _res_block("%s/decoder_%d_%d" % (p, skip + 1, u), "pu", co, co)
        s = "%s/region_br_%d" % (p, skip + 1)
        v += _res_filter(s, "deconv", 4, seg_classes, seg_classes) + _bn(s + "/batchnorm", seg_classes)
        cin = co + skip_ch[skip - 1]
    v += _res_filter(p + "/decoder_1", "deconv", 4, cin, 3) + _bn(p + "/decoder_1/batchnorm", 3)
    s = p + "/region_br_1"
    v += _res_filter(s, "deconv", 4, seg_classes, seg_classes) + _bn(s + "/batchnorm", seg_classes)
    return v


class ParamStore:
    """Flat fp32 parameter / gradient / Adam-v buffers of one network plus named views."""

    def __init__(self, specs, device, dtype=torch.float32):
        self.specs = list(specs)
        self.device = torch.device(device)
        self.dtype = dtype
        self.offsets = OrderedDict()
        off = 0
        for s in self.specs:
            if s.trainable:
                self.offsets[s.name] = off
                off += (s.numel + ALIGN - 1) // ALIGN * ALIGN
        self.n_flat = off
        self.flat = torch.zeros(off, dtype=dtype, device=self.device)
        self.grad = torch.zeros(off, dtype=dtype, device=self.device)
        self.adam_v = torch.zeros(off, dtype=dtype, device=self.device)     # Adam v / RMSProp rms / Adadelta accum / Adagrad accum
        self.opt_s2 = None                                                  # Adadelta accum_update (allocated on demand)
        self.optimizer = 'adam'
        self.adam_t = 0
        self.p, self.g = OrderedDict(), OrderedDict()
        self.state = OrderedDict()      # non-trainable (SN `u`)
        for s in self.specs:
            if s.trainable:
                o = self.offsets[s.name]
                self.p[s.name] = self.flat[o:o + s.numel].view(s.shape)
                self.g[s.name] = self.grad[o:o + s.numel].view(s.shape)
            else:
                self.state[s.name] = torch.zeros(s.shape, dtype=dtype, device=self.device)
        # chunk table for the fused Adam / weight-decay kernels: (start, length, reg) rows
        rows = []
        CH = 1 << 15
        for s in self.specs:
            if not s.trainable:
                continue
            o = self.offsets[s.name]
            for c0 in range(0, s.numel, CH):
                rows.append((o + c0, min(CH, s.numel - c0), s.reg))
        self.chunk_start = torch.tensor([r[0] for r in rows], dtype=torch.int64, device=self.device)
        self.chunk_len = torch.tensor([r[1] for r in rows], dtype=torch.int32, device=self.device)
        self.chunk_reg = torch.tensor([r[2] for r in rows], dtype=torch.float32, device=self.device)

    OPT_SLOT_INIT = {'adam': 0.0, 'rmsprop': 1.0, 'adadelta': 0.0, 'adagrad': 0.1}   # TF-1 slot initialisers

    def set_optimizer(self, kind):
        """Choose the optimiser whose slots this store carries (graph_single.get_optimizer, :584-593) and reset them."""
        kind = kind.lower()
        if kind not in self.OPT_SLOT_INIT:
            raise ValueError("optimizer %r (Adam, RMSprop, AdaDelta, AdaGrad)" % kind)
        self.optimizer = kind
        self.adam_v.fill_(self.OPT_SLOT_INIT[kind])
        self.opt_s2 = torch.zeros_like(self.adam_v) if kind == 'adadelta' else None
        self.adam_t = 0

    # --- counts (reference: main_procedure.print_parameter_count, :28-59) ---
    def num_trainable_tensors(self):
        return len(self.p)

    def num_params(self):
        return sum(s.numel for s in self.specs if s.trainable)

    # --- state dict keyed by the reference variable names -------------------
    def state_dict(self):
        d = OrderedDict((k, v.detach().clone()) for k, v in self.p.items())
        d.update((k, v.detach().clone()) for k, v in self.state.items())
        return d

    def load_state_dict(self, d, strict=True):
        for k, v in list(self.p.items()) + list(self.state.items()):
            if k in d:
                v.copy_(torch.as_tensor(d[k]).reshape(v.shape).to(v.dtype))
            elif strict:
                raise KeyError("missing variable %s" % k)

    def initialize(self, seed, perturb_tables=0.0):
        """Reference initialisers: N(0,0.02) convs (models_collection.py:896), xavier FCs (mru.py:54),
        U(+-0.08) embedding (:181), glorot-uniform LSTM kernels (TF default), zeros/ones cBN tables (:30-31),
        0.5 update-gate biases (mru.py:359), 0.2 PReLU (:58), truncated-normal `u` (sn.py:18)."""
        g = torch.Generator().manual_seed(seed)
        f64 = torch.float64
        for s in self.specs:
            kind, arg = s.init
            shape = s.shape
            if kind == "normal":
                t = torch.randn(shape, generator=g, dtype=f64) * arg
            elif kind == "normal1":
                t = 1.0 + torch.randn(shape, generator=g, dtype=f64) * arg
            elif kind == "trunc_normal":
                t = torch.randn(shape, generator=g, dtype=f64)
                for _ in range(8):
                    bad = t.abs() > 2
                    if not bad.any():
                        break
                    t = torch.where(bad, torch.randn(shape, generator=g, dtype=f64), t)
                t = t.clamp(-2, 2) * arg
            elif kind == "const":
                t = torch.full(shape, float(arg), dtype=f64)
                if perturb_tables and (s.name.endswith("/offset") or s.name.endswith("/scale")
                                       or s.name.endswith("prelu/param")):
                    t = t + torch.randn(shape, generator=g, dtype=f64) * perturb_tables
            elif kind == "uniform":
                t = (torch.rand(shape, generator=g, dtype=f64) * 2 - 1) * arg
            else:  # glorot_uniform / xavier
                lim = math.sqrt(6.0 / (shape[0] + shape[1]))
                t = (torch.rand(shape, generator=g, dtype=f64) * 2 - 1) * lim
            dst = self.p[s.name] if s.trainable else self.state[s.name]
            dst.copy_(t.to(self.dtype))
        self.adam_v.fill_(self.OPT_SLOT_INIT[self.optimizer])
        if self.opt_s2 is not None:
            self.opt_s2.zero_()
        self.adam_t = 0


# ------------------------------------------------------------------------------------------------------------------
# Instance-matching model (BASELINE.json configs[4]): Instance_Matching/deeplab_model.py (scope `ResNet`, :51-52) and
# RMI_model.py (scope `text_sketchyscene`, :113-151).  Inference only; the stored batch-norm moments are kept in the store
# like every other variable so that a state dict keyed by the reference names loads as it is.
# ------------------------------------------------------------------------------------------------------------------
RMI_UNITS = (3, 4, 23, 3)                      # deeplab_model.py:13
RMI_FILTERS = (64, 256, 512, 1024, 2048)       # :20


def _rmi_bn(scope, c):
    return [VarSpec(scope + "/beta", (c,), ("const", 0.0)), VarSpec(scope + "/gamma", (c,), ("const", 1.0)),
            VarSpec(scope + "/factor", (1,), ("const", 1.0)), VarSpec(scope + "/mean", (c,), ("const", 0.0)),
            VarSpec(scope + "/variance", (c,), ("const", 1.0))]


def _rmi_dw(scope, k, cin, cout):
    return [VarSpec(scope + "/DW", (k, k, cin, cout), ("normal", math.sqrt(2.0 / (k * k * cout))))]      # deeplab_model.py:281-284


def rmi_unit_plan(units=RMI_UNITS, filters=RMI_FILTERS):
    """[(scope, cin, cout, stride, rate)] of the bottleneck units in graph order (deeplab_model.py:77-103)."""
    plan = []
    for g, (stride, rate) in enumerate(((1, 1), (2, 1), (1, 2), (1, 4))):
        for i in range(units[g]):
            plan.append(("ResNet/group_%d_%d" % (g + 2, i), filters[g] if i == 0 else filters[g + 1], filters[g + 1],
                         stride if i == 0 else 1, rate))
    return plan


def rmi_vars(units=RMI_UNITS, filters=RMI_FILTERS, vocab_size=59, w_emb=1000, v_emb=1000, m_rnn=500, w_rnn=1000):
    v = _rmi_dw("ResNet/group_1/conv1", 7, 3, filters[0]) + _rmi_bn("ResNet/group_1/bn_conv1", filters[0])
    for scope, cin, cout, _, _ in rmi_unit_plan(units, filters):
        for blk, k, a, b in (("block_1", 1, cin, cout // 4), ("block_2", 3, cout // 4, cout // 4), ("block_3", 1, cout // 4, cout)):
            v += _rmi_dw("%s/%s/conv" % (scope, blk), k, a, b) + _rmi_bn("%s/%s/bn" % (scope, blk), b)
        if cin != cout:
            v += _rmi_dw(scope + "/block_add/conv", 1, cin, cout) + _rmi_bn(scope + "/block_add/bn", cout)
    p = "text_sketchyscene"
    xav = lambda cin, cout: ("uniform", math.sqrt(6.0 / (cin + cout)))            # xavier_initializer_conv2d, 1x1 (RMI_model.py:290-291)
    v += [VarSpec(p + "/visual_feat_projection/DW", (1, 1, filters[4], v_emb), xav(filters[4], v_emb)),
          VarSpec(p + "/visual_feat_projection/biases", (v_emb,), ("const", 0.0)),
          VarSpec(p + "/embedding", (vocab_size, w_emb), ("uniform", 0.08)),                             # :128-129
          VarSpec(p + "/wLSTM/lstm_cell/kernel", (w_emb + w_rnn, 4 * w_rnn), ("glorot_uniform", None)),
          VarSpec(p + "/wLSTM/lstm_cell/bias", (4 * w_rnn,), ("const", 0.0)),
          VarSpec(p + "/mLSTM/lstm_cell/kernel", (v_emb + w_emb + w_rnn + 8 + m_rnn, 4 * m_rnn), ("glorot_uniform", None)),
          VarSpec(p + "/mLSTM/lstm_cell/bias", (4 * m_rnn,), ("const", 0.0)),
          VarSpec(p + "/m_lstm_output_projection/DW", (1, 1, m_rnn, 1), xav(m_rnn, 1)),
          VarSpec(p + "/m_lstm_output_projection/biases", (1,), ("const", 0.0))]
    return v
