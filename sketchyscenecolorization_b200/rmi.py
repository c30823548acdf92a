"""Instance-matching model, inference (BASELINE.json configs[4]: referring segmentation of a 768 x 768 scene sketch).

Reference: Instance_Matching/RMI_model.py (`RMI_model`, :11-151; fusion module :169-222 with use_attn False; output
processing :275-286) over the trunk of Instance_Matching/deeplab_model.py (`DeepLab(is_intermediate=True)`, :65-107: ResNet-101
with output stride 8 -- groups 4 and 5 dilated by 2 and 4), in the default configuration of matching_main.py (weights
'deeplab', fusion_type 'RMI', mode 'eval').  Forward only: training it is outside SURVEY 8.

How it sits on the kernels of the colorization path:
  * every contraction is `ops.conv_fwd` (1x1 / 3x3 / 7x7, stride 1 / 2; the LSTM and projection matrices as k = 1 rows);
  * tf.nn.atrous_conv2d(rate r) is a plain SAME convolution in the space-to-batch form -- as inside TensorFlow -- and since
    everything else in a residual unit is per pixel, groups 4 and 5 run WHOLLY in that form: one permutation in, one out;
  * the stored-moment batch norm is an affine map per channel.  Where a relu follows directly (stem, block_1, block_2) its
    scale is folded into the filter once per set of weights and shift + relu ride in the convolution's epilogue; at the end
    of a unit the last batch norm, the shortcut's own batch norm, the residual sum and the relu are one pass
    (`ops.affine_act`);
  * the multimodal LSTM runs over N*96*96 rows x up to 15 steps.  As in text_fusion.py its input product is split by blocks of
    the kernel: [visual, spatial] rows once, [word embedding, l2n(word-LSTM output)] once per (sample, step), and only the
    [rows, 500] x [500, 2000] recurrent product per step -- 9.2 instead of 64.7 GMAC per picture and step.  Steps past a
    sample's `sequence_length` keep the state (dynamic_rnn), via the per-sample mask of the cell kernel.
"""
from __future__ import annotations

import numpy as np
import torch

from .ops_base import ACT_RELU
from .params import RMI_FILTERS, RMI_UNITS, ParamStore, rmi_unit_plan, rmi_vars
from .text_fusion import _mat, _op, _word_lstm_fwd

BN_EPS = 0.001                                              # deeplab_model.py:231-233
MU = np.array((104.00698793, 116.66876762, 122.67891434))   # matching_main.py:78


def spatial_rows(N, fh, fw):
    """utils/processing_tools.generate_spatial_batch (:5-17) as [N*fh*fw, 8] float32 rows."""
    w, h = np.arange(fw, dtype=np.float64), np.arange(fh, dtype=np.float64)
    xmin, xmax = w / fw * 2 - 1, (w + 1) / fw * 2 - 1
    ymin, ymax = h / fh * 2 - 1, (h + 1) / fh * 2 - 1
    v = np.zeros((fh, fw, 8), dtype=np.float32)
    v[..., 0], v[..., 2], v[..., 4] = xmin[None, :], xmax[None, :], ((xmin + xmax) / 2)[None, :]
    v[..., 1], v[..., 3], v[..., 5] = ymin[:, None], ymax[:, None], ((ymin + ymax) / 2)[:, None]
    v[..., 6], v[..., 7] = 1 / fw, 1 / fh
    return np.broadcast_to(v, (N, fh, fw, 8)).reshape(N * fh * fw, 8).copy()


def preprocess_sentence(sentence, vocab_dict, T=15):
    """data_processing/text_processing.preprocess_sentence (:89-102): (ids padded at the END with <pad> to T, length)."""
    import re
    words = [w.lower() for w in re.split(r'(\W+)', sentence.strip()) if len(w.strip()) > 0 and w != '-']
    if words and words[-1] == '.':
        words = words[:-1]
    ids = [vocab_dict.get(w, vocab_dict['<unk>']) for w in words][:T]
    return ids + [vocab_dict['<pad>']] * (T - len(ids)), len(ids)


class RMIModel:
    def __init__(self, ops, device, *, units=RMI_UNITS, filters=RMI_FILTERS, vocab_size=59, w_emb=1000, v_emb=1000, m_rnn=500,
                 w_rnn=1000, param_dtype=torch.float32):
        self.ops, self.device, self.units, self.filters = ops, torch.device(device), tuple(units), tuple(filters)
        self.dims = dict(w_emb=w_emb, v_emb=v_emb, m_rnn=m_rnn, w_rnn=w_rnn)
        self.store = ParamStore(rmi_vars(units, filters, vocab_size, w_emb, v_emb, m_rnn, w_rnn), device, param_dtype)
        self._folded = None

    def initialize(self, seed=0):
        self.store.initialize(seed)
        self._folded = None

    def load_state_dict(self, d, strict=True):
        self.store.load_state_dict(d, strict)
        self._folded = None

    # ---- parameter preparation (once per set of weights; the arithmetic of the forward pass is all in libfgcolor)
    def _prepare(self):
        """Stored-moment batch norm -> (scale, shift) per channel (deeplab_model.py:213-233), and the [visual, spatial] rows of
        the mLSTM kernel side by side."""
        if self._folded is not None:
            return self._folded
        P = self.store.p
        bn = {}
        for name in P:
            if name.endswith("/gamma"):
                s = name[:-len("/gamma")]
                inv = 1.0 / P[s + "/factor"].double()
                scale = P[s + "/gamma"].double() / torch.sqrt(P[s + "/variance"].double() * inv + BN_EPS)
                shift = P[s + "/beta"].double() - P[s + "/mean"].double() * inv * scale
                bn[s] = (scale.to(self.store.dtype).contiguous(), shift.to(self.store.dtype).contiguous())
        # convolution -> batch norm -> relu: the scale goes into the filter (HWIO: per output channel), the shift is the bias
        wf = {}
        for s in bn:
            conv = (s[:-len("/bn")] + "/conv") if s.endswith("/bn") else s.replace("/bn_conv1", "/conv1")
            if s.endswith("/block_1/bn") or s.endswith("/block_2/bn") or s.endswith("/bn_conv1"):
                wf[conv] = (P[conv + "/DW"].double() * bn[s][0].double()).to(self.store.dtype).contiguous()
        d = self.dims
        km = P["text_sketchyscene/mLSTM/lstm_cell/kernel"]
        o = d["v_emb"] + d["w_emb"] + d["w_rnn"]
        k_pos = torch.cat([km[0:d["v_emb"]], km[o:o + 8]], 0).contiguous()
        # recurrent rows of the mLSTM kernel, zero rows appended up to a multiple of 8: a bf16 operand row of 500 elements is
        # not 16-byte aligned, and the tensor path's operand fetches fall back to 2-byte loads on it (14.6 ms instead of
        # 1.4 ms per step at 32 x 96 x 96 rows, profiles/r2p_prof_rmi.log)
        Dm = d["m_rnn"]
        Dp = (Dm + 7) // 8 * 8
        kh = km[o + 8:]
        kh_p = torch.cat([kh, kh.new_zeros(Dp - Dm, kh.shape[1])], 0).contiguous() if Dp != Dm else kh
        wo = P["text_sketchyscene/m_lstm_output_projection/DW"]
        wo_p = torch.cat([wo, wo.new_zeros(1, 1, Dp - Dm, wo.shape[3])], 2).contiguous() if Dp != Dm else wo
        self._folded = dict(bn=bn, wf=wf, k_pos=k_pos, kh_p=kh_p, wo_p=wo_p, Dp=Dp, spatial={})
        return self._folded

    # ---- trunk: DeepLab._build_model (:65-107)
    def _unit(self, x, scope, cin, cout, stride, bn):
        """_bottleneck_residual (:237-264); x is already in the batch form of the group's dilation."""
        ops, P = self.ops, self.store.p
        wf = self._folded["wf"]
        h = ops.conv_fwd([(x, False)], wf[scope + "/block_1/conv"], bn[scope + "/block_1/bn"][1], stride=stride, act=ACT_RELU)
        h = ops.conv_fwd([(h, False)], wf[scope + "/block_2/conv"], bn[scope + "/block_2/bn"][1], act=ACT_RELU)
        h = ops.conv_fwd([(h, False)], P[scope + "/block_3/conv/DW"], None)
        if cin != cout:
            sc = ops.conv_fwd([(x, False)], P[scope + "/block_add/conv/DW"], None, stride=stride)
            rs, rt = bn[scope + "/block_add/bn"]
            return ops.affine_act(h, *bn[scope + "/block_3/bn"], res=sc, rscale=rs, rshift=rt, relu=True)
        return ops.affine_act(h, *bn[scope + "/block_3/bn"], res=x, relu=True)

    def trunk(self, im_nhwc):
        """[N,H,W,3] (activation dtype) -> `intermediate_feat` [N,H/8,W/8,filters[4]]."""
        ops, P = self.ops, self.store.p
        pre = self._prepare()
        bn = pre["bn"]
        x = ops.conv_fwd([(im_nhwc, False)], pre["wf"]["ResNet/group_1/conv1"], bn["ResNet/group_1/bn_conv1"][1], stride=2, act=ACT_RELU)
        x = ops.maxpool3x3s2(x)
        cur = 1                                               # dilation whose batch form x is in
        for scope, cin, cout, stride, rate in rmi_unit_plan(self.units, self.filters):
            if rate != cur:
                if cur > 1:
                    x = ops.batch_to_space(x, cur)
                x = ops.space_to_batch(x, rate)
                cur = rate
            x = self._unit(x, scope, cin, cout, stride, bn)
        if cur > 1:
            x = ops.batch_to_space(x, cur)
        return x                                              # group_last's relu (:105-106) is idempotent here

    # ---- fusion: RMI_model.build_graph (:119-151)
    def fuse(self, feat, words, lengths, H, W):
        """feat [N,h,w,C]; words int [N,T], lengths int [N] (host arrays or tensors).  Returns (pred [N,h,w,1], up [N,H,W,1],
        sigm [N,H,W,1]), fp32."""
        ops, P, d = self.ops, self.store.p, self.dims
        pre = self._prepare()
        f32 = torch.float32
        p = "text_sketchyscene"
        N, fh, fw, _ = feat.shape
        Pn, R = fh * fw, N * fh * fw
        V, E, L, Dm, Dw = d["v_emb"], d["w_emb"], d["w_rnn"], d["m_rnn"], d["w_rnn"]
        dev = feat.device
        words_np = np.asarray(words.cpu() if torch.is_tensor(words) else words, dtype=np.int64)
        len_np = np.asarray(lengths.cpu() if torch.is_tensor(lengths) else lengths, dtype=np.int64).reshape(-1)
        T = words_np.shape[1]
        ids = torch.as_tensor(words_np.astype(np.int32), device=dev).contiguous()
        # dynamic_rnn's sequence_length as the cell kernels' per-sample step mask (non-zero = the step runs)
        live = torch.as_tensor((np.arange(T)[None, :] < len_np[:, None]).astype(np.int32), device=dev).contiguous()
        steps = int(min(T, len_np.max())) if len_np.size else 0
        # visual rows: 1x1 projection + bias, l2-normalised over channels (:120-123)
        vis = ops.conv_fwd([(feat, False)], P[p + "/visual_feat_projection/DW"], P[p + "/visual_feat_projection/biases"], out_dtype=f32)
        vis, _ = ops.l2norm_rows_fwd(vis.view(R, V))
        key = (N, fh, fw)
        if key not in pre["spatial"]:
            pre["spatial"][key] = torch.from_numpy(spatial_rows(N, fh, fw)).to(dev)
        sp = pre["spatial"][key]
        km, bm = P[p + "/mLSTM/lstm_cell/kernel"], P[p + "/mLSTM/lstm_cell/bias"]
        kw, bw = P[p + "/wLSTM/lstm_cell/kernel"], P[p + "/wLSTM/lstm_cell/bias"]
        g_pos = ops.conv_fwd([(_op(ops, vis), False), (_op(ops, sp), False)], _mat(pre["k_pos"]), bm, out_dtype=f32).view(R, 4 * Dm)
        # word LSTM over all steps (:153-167): embeddings, their gate products in one go, then the recurrence
        e_all = ops.embedding_all_fwd(P[p + "/embedding"], ids)                                   # [T,N,E]
        e_rows = _op(ops, e_all.view(T * N, E))
        gx = ops.conv_fwd([(e_rows, False)], _mat(kw[0:E]), bw, out_dtype=f32).view(T, N, 4 * Dw)
        hw_all, _, _ = _word_lstm_fwd(ops, gx, kw[E:E + Dw], live, T, N, Dw)
        lang, _ = ops.l2norm_rows_fwd(hw_all[1:].view(T * N, Dw))                                 # :173 (rows past the length are never read)
        g_row = ops.conv_fwd([(e_rows, False), (_op(ops, lang), False)], _mat(km[V:V + E + L]), None, out_dtype=f32).view(T, N, 4 * Dm)
        # mLSTM recurrence (:188-200): the only per-step product is h @ kernel[-Dm:]
        Dp, odt = pre["Dp"], ops.act_dtype

        def rows_op(x):          # [R, Dm] fp32 -> matrix-product operand, padded to an aligned width
            return ops.pad_cast_rows(x, Dp, odt).view(R, 1, 1, Dp) if Dp != Dm else _op(ops, x)
        c, h = ops.zeros_f32((R, Dm)), ops.zeros_f32((R, Dm))
        for t in range(steps):
            ga = ops.conv_fwd([(rows_op(h), False)], _mat(pre["kh_p"]), None, out_dtype=f32).view(R, 4 * Dm)
            c, h, _ = ops.lstm_cell_fwd(ga, g_pos, g_row[t], c, h, live, t, Pn, save_pre=False)
        m = ops.atanh_relu_fwd(h)                                                                 # :277-279
        pred = ops.conv_fwd([(rows_op(m).view(N, fh, fw, Dp), False)], pre["wo_p"],
                            P[p + "/m_lstm_output_projection/biases"], out_dtype=f32)             # :284-285
        up, sigm = ops.resize_bilinear_sigmoid(pred, H, W)                                        # :150-151
        return pred, up, sigm

    def forward(self, im_nhwc, words, lengths):
        """im [N,H,W,3] float (BGR minus MU, matching_main.py:447), H and W multiples of 32 -> (up, sigm) fp32 [N,H,W,1]."""
        x = im_nhwc.to(self.device)
        x = x.contiguous() if x.dtype == self.ops.act_dtype else self.ops.cast(x.float().contiguous(), self.ops.act_dtype)
        feat = self.trunk(x)
        _, up, sigm = self.fuse(feat, words, lengths, x.shape[1], x.shape[2])
        return up, sigm

    def predict_mask(self, sketch_u8, caption, vocab_dict, T=15, score_thresh=1e-9):
        """matching_main.inference's feed / fetch (:443-466): uint8 sketch [H,W,3] + caption -> binary mask [H,W] restricted to
        the strokes (pixel value 0 in the first channel)."""
        img = np.asarray(sketch_u8, dtype=np.float32) - MU.astype(np.float32)
        ids, n = preprocess_sentence(caption, vocab_dict, T)
        up, _ = self.forward(torch.from_numpy(img[None]), np.asarray([ids]), np.asarray([n]))
        pred = (up[0, :, :, 0].float().cpu().numpy() >= score_thresh).astype(np.float32)
        strokes = np.asarray(sketch_u8)[:, :, 0].copy()        # :449-451: strokes (0) -> 1, paper (255) -> 0
        strokes[strokes == 0] = 1
        strokes[strokes == 255] = 0
        return pred * strokes
