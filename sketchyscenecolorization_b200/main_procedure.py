"""Session loops of the fg-colorization path on the B200 kernels: train / validation / test / inference.

Same entry points, `Config` keys, status codes and on-disk layout as the reference's obj_lib/main_procedure.py
(`train` :62-242 -> 0 ok / -1 NaN, `inference` :495-621 writes <stem>_output.png and <stem>_input.png);
the TF1 session, queue runners and Saver are replaced by FgColorTrainer, an input iterator and checkpoint.py."""
from __future__ import annotations

import json
import math
import os
from time import time

import numpy as np
import torch

from . import checkpoint, graph_single
from .config import Config
from .input_pipeline import CATEGORIES, SyntheticInput, resize_and_padding_mask_image
from .text_processing import default_vocab_dict, load_vocab_dict_from_file, preprocess_sentence

SIZE = {True: (64, 64), False: (192, 192)}
T = 15      # main_procedure.py:503


def _dtype(name):
    return torch.bfloat16 if name == 'bf16' else torch.float32


def _build_model(precision, img_dim, vocab_size, lstm_hybrid, device=None, with_discriminator=True):
    block_type = getattr(Config, 'block_type', 'MRU') or 'MRU'
    if block_type not in ('MRU', 'Pix2Pix', 'Residual'):
        raise ValueError("block_type %r (MRU, Pix2Pix, Residual)" % (block_type,))
    from .cuda_ops import CudaOps
    from .trainer import FgColorModel
    dev = device or "cuda:%d" % int(os.environ.get("LOCAL_RANK", "0"))
    ops = CudaOps(dev, _dtype(precision), conv_terms=int(getattr(Config, 'conv_terms', 2)))
    return FgColorModel(ops, dev, H=img_dim[0], W=img_dim[1], vocab_size=vocab_size, lstm_hybrid=lstm_hybrid,
                        with_discriminator=with_discriminator, block_type=block_type)


def _categories():
    base = os.path.join('data', 'captions')
    if os.path.isdir(base):
        return sorted(os.listdir(base))            # main_procedure.py:506-508
    return list(CATEGORIES)


def _vocab():
    f = os.path.join('data', 'vocab.txt')
    return load_vocab_dict_from_file(f) if os.path.exists(f) else default_vocab_dict()


def print_parameter_count(model, verbose=False):
    """main_procedure.print_parameter_count (:28-59)."""
    g, d = model.gstore.num_params(), model.dstore.num_params() if model.dstore is not None else 0
    print('generator: %d trainable parameters in %d tensors' % (g, model.gstore.num_trainable_tensors()))
    if model.dstore is not None:
        print('discriminator: %d trainable parameters in %d tensors' % (d, model.dstore.num_trainable_tensors()))
    return g, d


def any_rank_true(flag, process_group=None, world_size=1, device="cpu"):
    """Data-parallel agreement on a per-rank condition (a NaN loss): every rank must leave the session loop in the same
    iteration, or the others would wait for it in the next all-reduce.  One tiny MAX all-reduce; a no-op for one process."""
    if world_size <= 1:
        return bool(flag)
    import torch.distributed as dist
    t = torch.tensor([1.0 if flag else 0.0], device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=process_group)
    return bool(t.item() > 0)


def shared_string(value, process_group=None, world_size=1, src=0):
    """Rank `src`'s string on every rank (the run directory's time stamp must not differ between ranks)."""
    if world_size <= 1:
        return value
    import torch.distributed as dist
    box = [value]
    dist.broadcast_object_list(box, src=src, group=process_group)
    return box[0]


class TrainSession:
    """One process' share of the training session (main_procedure.py:124-237): the trainer, the two input queues and
    `iteration(i)` = `disc_iterations` D steps + one G step, each on a freshly dequeued batch, the loss scalars read back
    and the NaN verdict agreed across ranks.  `train` below is a loop over it; `bench.py` times the same object.

    On the CUDA operator set the two steps replay CUDA graphs (`use_cuda_graphs`, default on; `--cuda_graphs 0` /
    kwargs `use_cuda_graphs=False` launches every kernel from Python): caption ids then stay on the device.  Host traffic
    per iteration: the batches' host-to-device copies and ONE device-to-host read of the loss scalars -- with more than one
    rank the NaN flags are MAX-reduced on the device first, so the verdict rides in the same read."""

    D_KEYS = ('sketch', 'images_d', 'cls', 'cls_d', 'text', 'noise')
    G_KEYS = ('sketch', 'images', 'cls', 'text', 'noise')

    def __init__(self, model, *, batch_size, max_iter, lr_g, lr_d, optimizer='Adam', disc_iterations=1, iter_from=0,
                 input_iter=None, input_iter_d=None, process_group=None, world_size=1, use_cuda_graphs=None,
                 small=False, vocab_size=58, distance_map=False, data_base_dir='data', synthetic_input=None):
        from .trainer import FgColorTrainer
        self.model, self.n, self.diters = model, batch_size, disc_iterations
        self.pg, self.world = process_group, int(world_size)
        self.dev = model.device
        on_cuda = bool(getattr(model.ops, 'supports_cuda_graphs', False))
        if use_cuda_graphs is None:
            use_cuda_graphs = os.environ.get("FGC_CUDA_GRAPHS", "1") != "0"
        self.graphs = bool(use_cuda_graphs) and on_cuda
        self.tr = FgColorTrainer(model, lr_g=lr_g, lr_d=lr_d, max_iter=max_iter, process_group=process_group,
                                 world_size=self.world, optimizer=optimizer, use_cuda_graphs=self.graphs)
        self.tr.counter = iter_from                                  # sess.run(counter.assign(iter_from)), :176
        rank = int(os.environ.get("RANK", "0"))
        self._own = []                                               # queues created here (closed by close())
        q1, q2 = input_iter, input_iter_d
        if q1 is None or q2 is None:
            rec_dir = os.path.join(data_base_dir, 'tfrecord', 'train')
            synthetic_input = bool(synthetic_input) or os.environ.get("FGC_SYNTHETIC_INPUT") == "1"
            if os.path.isdir(rec_dir) and os.listdir(rec_dir):
                # the reference's two independent TFRecord shuffle queues (main_procedure.py:109-122)
                from .tfrecord_input import PairedTrainInput
                for seed, have in ((1234 + rank, q1), (4321 + rank, q2)):
                    if have is None:
                        self._own.append(PairedTrainInput(batch_size, model.ops, data_base_dir, small=small,
                                                          distance_map=distance_map, seed=seed))
                it = iter(self._own)
                q1, q2 = q1 or next(it), q2 or next(it)
            elif synthetic_input:      # asked for (bench, tests): seeded synthetic batches of the SURVEY 8(d) shape
                q1 = q1 or SyntheticInput(batch_size, *SIZE[small], vocab_size=vocab_size, seed=1234 + rank)
                q2 = q2 or SyntheticInput(batch_size, *SIZE[small], vocab_size=vocab_size, seed=4321 + rank)
            else:                      # the reference fails in os.listdir of build_input_queue_paired (input_pipeline.py:133)
                raise FileNotFoundError(
                    "%s holds no TFRecord files (the FOREGROUND dataset is not part of the repository).  Pass "
                    "synthetic_input=True / set FGC_SYNTHETIC_INPUT=1 to train on seeded synthetic batches." % rec_dir)
        self.q1, self.q2 = q1, q2
        self.last_d = self.last_g = None
        self.prefetch = os.environ.get("FGC_INPUT_PREFETCH", "1") != "0"
        self._copy_stream, self._staged, self._stage_bufs, self._stage_free = None, {}, {}, {}

    # ---- batches: opt_d dequeues both queues, opt_g only the first (graph_single.py:257-289; main_procedure.py:202-227)
    def _text(self, t):
        if self.graphs:
            return torch.as_tensor(t).to(self.dev, non_blocking=True)
        return t.numpy() if torch.is_tensor(t) else np.asarray(t)

    def _noise(self):
        return torch.randn(self.n, 256, device=self.dev)

    def _host_d(self):
        a, b = next(self.q1), next(self.q2)
        out = {k: a[k] for k in ('sketch', 'cls')}
        out.update({k: b[k] for k in ('images_d', 'cls_d')})
        out['text'] = torch.as_tensor(a['text'])
        return out

    def _host_g(self):
        a = next(self.q1)
        out = {k: a[k] for k in ('sketch', 'images', 'cls')}
        out['text'] = torch.as_tensor(a['text'])
        return out

    def fetch_d(self):
        a = self._host_d()
        out = {k: (v if self.graphs else v.to(self.dev, non_blocking=True)) for k, v in a.items() if k != 'text'}
        out['text'], out['noise'] = self._text(a['text']), self._noise()
        return out

    def fetch_g(self):
        a = self._host_g()
        out = {k: (v if self.graphs else v.to(self.dev, non_blocking=True)) for k, v in a.items() if k != 'text'}
        out['text'], out['noise'] = self._text(a['text']), self._noise()
        return out

    # ---- input prefetch (graph-replay mode): the NEXT batch crosses PCIe on a copy stream while the current step computes;
    # at its turn the step copies it device-to-device into the captured graph's input buffers.  One staging set per kind: the
    # D batch of iteration i+1 is staged under the G step of iteration i, the G batch under the D step, and a set is rewritten
    # only after the step that read it has finished (event).  Batches are drawn from the queues in the reference's order.
    def _prefetch(self, kind):
        try:
            host = self._host_d() if kind == 'd' else self._host_g()
        except StopIteration:
            self._staged[kind] = None
            return
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.dev)
        side, bufs = self._copy_stream, self._stage_bufs.setdefault(kind, {})
        free = self._stage_free.get(kind)
        if free is not None:
            side.wait_event(free)
        dev = {}
        with torch.cuda.stream(side):
            for k, v in host.items():
                v = torch.as_tensor(v)
                if v.is_cuda:                       # the TFRecord queue hands over device tensors already
                    dev[k] = v
                    continue
                b = bufs.get(k)
                if b is None or b.shape != v.shape or b.dtype != v.dtype:
                    b = bufs[k] = torch.empty(v.shape, dtype=v.dtype, device=self.dev)
                b.copy_(v, non_blocking=True)
                dev[k] = b
            ready = torch.cuda.Event()
            ready.record(side)
        self._staged[kind] = (dev, ready, host)     # the host tensors stay referenced until the copy has been consumed

    def _take(self, kind):
        if kind not in self._staged:
            self._prefetch(kind)
        item = self._staged.pop(kind)
        if item is None:
            raise StopIteration
        dev, ready, _ = item
        torch.cuda.current_stream().wait_event(ready)
        out = dict(dev)
        out['noise'] = self._noise()
        return out

    def _step_done(self, kind):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self._stage_free[kind] = ev

    # ---- one iteration of the session loop
    def iteration(self):
        """Returns (loss_d, loss_g, nan_d, nan_g) as Python values: one device-to-host read."""
        if self.graphs and self.prefetch:
            for j in range(self.diters):
                od = self.tr.d_step(self._take('d'))
                self._step_done('d')
                self._prefetch('d' if j + 1 < self.diters else 'g')
            og = self.tr.g_step(self._take('g'))
            self._step_done('g')
            self._prefetch('d')
        else:
            for _ in range(self.diters):
                od = self.tr.d_step(self.fetch_d())
            og = self.tr.g_step(self.fetch_g())
        self.last_d, self.last_g = od, og
        lo = torch.stack([od['loss'].float().reshape(()), og['loss'].float().reshape(())])
        bad = torch.isnan(lo).float()
        if self.world > 1:            # every rank must leave the loop in the same iteration
            import torch.distributed as dist
            dist.all_reduce(bad, op=dist.ReduceOp.MAX, group=self.pg)
        v = torch.cat([lo, bad]).tolist()
        return v[0], v[1], v[2] > 0, v[3] > 0

    def close(self):
        self._staged.clear()
        self._stage_bufs.clear()
        for q in self._own:
            q.close()
        self._own = []


def train(**kwargs):
    """Alternating D / G optimisation (main_procedure.py:62-242).  kwargs: iter_from, and optionally `input_iter` /
    `input_iter_d` (two-queue stand-ins yielding batch dicts), `process_group` / `world_size` for data parallelism,
    `use_cuda_graphs`, `synthetic_input`, a resident `model`."""
    status = 0
    batch_size, max_iter_step, diters = Config.batch_size, Config.max_iter_step, Config.disc_iterations
    log_dir, ckpt_dir = Config.log_dir, Config.ckpt_dir
    small, lstm_hybrid = Config.small_img != 0, Config.LSTM_hybrid != 0
    if Config.block_type not in ('MRU', 'Pix2Pix', 'Residual'):
        raise ValueError("block_type %r (MRU, Pix2Pix, Residual)" % (Config.block_type,))
    if Config.optimizer.lower() not in ('adam', 'rmsprop', 'adadelta', 'adagrad'):
        raise ValueError("optimizer %r (RMSprop, Adam, AdaDelta, AdaGrad)" % (Config.optimizer,))
    iter_from = kwargs['iter_from']
    world = int(kwargs.get('world_size', 1))
    print('Iteration starts from: %d' % iter_from)

    model = kwargs.get('model')                                     # a resident model (tests); else built on the CUDA operator set
    if model is None:
        precision = Config.train_precision_residual if Config.block_type == 'Residual' else Config.train_precision
        model = _build_model(precision, SIZE[small], Config.vocab_size, lstm_hybrid)
        model.initialize(seed=int(kwargs.get('seed', 0)))
    if iter_from > 0:
        prefix = checkpoint.latest_checkpoint(ckpt_dir)
        print('Restore:', prefix)
        checkpoint.restore(model, prefix)
    graphs = kwargs.get('use_cuda_graphs')
    if graphs is None and getattr(Config, 'cuda_graphs', None) is not None:
        graphs = bool(Config.cuda_graphs)
    sess = TrainSession(model, batch_size=batch_size, max_iter=max_iter_step, lr_g=Config.lr_G, lr_d=Config.lr_D,
                        optimizer=Config.optimizer, disc_iterations=diters, iter_from=iter_from,
                        input_iter=kwargs.get('input_iter'), input_iter_d=kwargs.get('input_iter_d'),
                        process_group=kwargs.get('process_group'), world_size=world, use_cuda_graphs=graphs, small=small,
                        vocab_size=Config.vocab_size, distance_map=Config.distance_map != 0,
                        data_base_dir=kwargs.get('data_base_dir', 'data'),
                        synthetic_input=kwargs.get('synthetic_input') or getattr(Config, 'synthetic_input', 0))
    tr = sess.tr
    print_parameter_count(model)
    rank = int(os.environ.get("RANK", "0"))
    summ = open(os.path.join(log_dir, 'summaries.jsonl'), 'a') if rank == 0 else None
    nvtx = torch.cuda.nvtx if torch.cuda.is_available() else None
    try:
        prev_time = float("-inf")
        for i in range(iter_from, max_iter_step):
            if i % Config.count_left_time_freq == 0:
                curr_time = time()
                elapsed = curr_time - prev_time
                print("Now at iteration %d. Elapsed time: %.5fs. Average time: %.5fs/iter"
                      % (i, elapsed, elapsed / float(Config.count_left_time_freq)))
                if elapsed != float("inf"):                               # main_procedure.py:182-188
                    left_sec = (max_iter_step - i) * (elapsed / float(Config.count_left_time_freq))
                    left_day = int(left_sec / 24 / 60 / 60)
                    left_hour = int((left_sec - (24 * 60 * 60) * left_day) / 60 / 60)
                    left_min = int((left_sec - (24 * 60 * 60) * left_day - (60 * 60) * left_hour) / 60)
                    print("Left time:%dd %dh %dm" % (left_day, left_hour, left_min))
                prev_time = curr_time
            want_summary = i % Config.summary_write_freq == 0
            if nvtx is not None:
                nvtx.range_push("train_iteration_%d" % i)
            loss_d_out, loss_g_out, nan_d, nan_g = sess.iteration()
            if nvtx is not None:
                nvtx.range_pop()
            if nan_d:                                                     # main_procedure.py:213-216
                print("NaN occurred during training D")
                return -1
            if nan_g:                                                     # :229-232
                print("NaN occurred during training G")
                return -1
            if want_summary and summ is not None:                         # scalar names of graph_single.py:71-98
                od, og = sess.last_d, sess.last_g
                rec = {"step": i, "GAN_loss/G": float(og['gan']), "GAN_loss/D": float(od['gan']), "ACGAN_loss/G": float(og['ac']),
                       "ACGAN_loss/D": float(od['ac']), "l1_perceptual_loss": float(og['l1']), "total_loss/g": loss_g_out,
                       "total_loss/d": loss_d_out, "learning_rate_g": Config.lr_G * max(0.2, 1 - 0.9 * i / max_iter_step)}
                summ.write(json.dumps(rec) + "\n")
                summ.flush()
            if i % Config.save_model_freq == Config.save_model_freq - 1 and rank == 0:
                checkpoint.save(model, ckpt_dir, i, tr.counter)
                print('Save model_{}.ckpt'.format(i))
    finally:
        sess.close()
        if summ is not None:
            summ.close()
    return status


def _load_sketch(path, img_dim, category, thicken=None):
    from PIL import Image
    im = Image.open(path).convert("RGB")
    if im.width != img_dim[0] or im.height != img_dim[1]:
        arr = resize_and_padding_mask_image(im, img_dim[0], margin_size=0 if category in ['road'] else 10).astype(np.float32)
    else:
        arr = np.array(im, dtype=np.float32)
    if thicken is not None:                                        # main_procedure.py:443-444 ('house', 'road' in test())
        arr = thicken(arr).astype(np.float32)
    arr = arr / 255. * 2. - 1
    return np.transpose(arr[None], [0, 3, 1, 2])                    # [1, 3, H, W]


def _write_pair(folder, stem, generated, sketch):
    """NCHW -> NHWC, (x+1)/2*255 truncated to uint8, generated image RGB->BGR for cv2 (main_procedure.py:601-619)."""
    import cv2
    gen = np.transpose(generated, (0, 2, 3, 1))
    sk = np.transpose(sketch, (0, 2, 3, 1))
    gen = (((gen + 1) / 2.) * 255)[:, :, :, ::-1].astype(np.uint8)
    sk = (((sk + 1) / 2.) * 255).astype(np.uint8)
    cv2.imwrite(os.path.join(folder, stem + '_output.png'), gen[0])
    cv2.imwrite(os.path.join(folder, stem + '_input.png'), sk[0])


def inference(img_name, instruction, model=None, noise=None):
    """One sketch + caption -> colourised PNG (main_procedure.py:495-621)."""
    wild_cate = img_name[:img_name.find('.png')]
    categories = _categories()
    if wild_cate not in categories:
        wild_cate = categories[2]                                   # 'bus', :510-511
    small, lstm_hybrid = Config.small_img != 0, Config.LSTM_hybrid != 0
    img_dim = SIZE[small]
    os.makedirs(Config.results_dir, exist_ok=True)
    print('output_folder:', Config.results_dir)
    if model is None:
        model = _build_model(Config.infer_precision, img_dim, Config.vocab_size, lstm_hybrid, with_discriminator=False)
        prefix = checkpoint.latest_checkpoint(Config.ckpt_dir)
        print('Restore trained model:', prefix)
        if prefix is None:
            raise RuntimeError("no snapshot in %s" % Config.ckpt_dir)
        checkpoint.restore(model, prefix, strict=True)
    sketch = _load_sketch(os.path.join('examples', img_name), img_dim, wild_cate)
    class_id = np.array([categories.index(wild_cate)])
    ids = np.array(preprocess_sentence(instruction, _vocab(), T), dtype=np.int32)[None]
    ret = graph_single.build_single_graph(sketch, sketch, None, class_id, None, ids, batch_size=1, training=False,
                                          LSTM_hybrid=lstm_hybrid, vocab_size=Config.vocab_size, data_format=Config.data_format,
                                          distance_map=Config.distance_map != 0, block_type=Config.block_type, model=model,
                                          noise=noise)
    generated, input_sketch = ret[0].cpu().numpy(), ret[2].cpu().numpy()
    _write_pair(Config.results_dir, img_name[:-4], generated, input_sketch)
    print('Saved file %s' % (img_name[:-4] + '_output.png'))
    return generated


def test(model=None, noise=None, data_base_dir='data'):
    """Batch-1 inference over every entry of <data>/captions/<category>/test.json (main_procedure.py:361-492): the sketch is
    <data>/images/<category>/sketch/<key>, resized / padded like `inference` (no margin for 'road'), strokes thickened for
    'house' and 'road' (:443-444), class id = index of the category folder, caption = the entry's `color_text`; writes
    <results_dir>/<category>_<stem>_{output,input}.png.  A failing sample prints the error and ends its category, as in the
    reference (:463-465).  Returns the number of pictures written.  Extra arguments: a resident `model`, fixed `noise`."""
    from .pipeline_fg import thicken_drawings
    small, lstm_hybrid = Config.small_img != 0, Config.LSTM_hybrid != 0
    img_dim = SIZE[small]
    captions_base_dir, images_base_dir = os.path.join(data_base_dir, 'captions'), os.path.join(data_base_dir, 'images')
    if not os.path.isdir(captions_base_dir):
        raise FileNotFoundError("%s is missing: the FOREGROUND dataset is not part of the repository" % captions_base_dir)
    categories = sorted(os.listdir(captions_base_dir))
    print(categories)
    os.makedirs(Config.results_dir, exist_ok=True)
    vocab_file = os.path.join(data_base_dir, 'vocab.txt')
    vocab = load_vocab_dict_from_file(vocab_file) if os.path.exists(vocab_file) else _vocab()
    if model is None:
        model = _build_model(Config.infer_precision, img_dim, Config.vocab_size, lstm_hybrid, with_discriminator=False)
        prefix = checkpoint.latest_checkpoint(Config.ckpt_dir)
        print('Restore trained model:', prefix)
        if prefix is None:
            raise RuntimeError("no snapshot in %s" % Config.ckpt_dir)
        checkpoint.restore(model, prefix, strict=True)
    written = 0
    for cate in categories:
        with open(os.path.join(captions_base_dir, cate, 'test.json')) as f:
            json_data = json.load(f)
        print(len(json_data), 'inference datas')
        for entry in json_data:
            input_name, input_text = entry['key'], entry['color_text']
            sketch = _load_sketch(os.path.join(images_base_dir, cate, 'sketch', input_name), img_dim, cate,
                                  thicken=thicken_drawings if cate in ['house', 'road'] else None)
            class_id = np.array([categories.index(cate)])
            ids = np.array(preprocess_sentence(input_text, vocab, T), dtype=np.int32)[None]
            try:
                ret = graph_single.build_single_graph(sketch, sketch, None, class_id, None, ids, batch_size=1, training=False,
                                                      LSTM_hybrid=lstm_hybrid, vocab_size=Config.vocab_size,
                                                      data_format=Config.data_format, distance_map=Config.distance_map != 0,
                                                      block_type=Config.block_type, model=model, noise=noise)
            except Exception as e:          # noqa: BLE001 -- printed and swallowed, the reference's convention (:463-465)
                print(e.args)
                break
            stem = cate + '_' + input_name[:-4]
            _write_pair(Config.results_dir, stem, ret[0].cpu().numpy(), ret[2].cpu().numpy())
            print('Saved file %s' % (stem + '_output.png'))
            written += 1
    return written


def validation(**kwargs):
    """One ordered pass over data/tfrecord/<Config.dataset_type> (main_procedure.py:245-358): for every sample writes
    <category>_<name>_{output,target,input}.png into <results_dir>/with_text (or without_text).  Batches of
    Config.batch_size, as in the reference (conditional BN sees the batch).  kwargs: optional `model`, `data_base_dir`."""
    import cv2
    from .tfrecord_input import PairedEvalInput
    small, lstm_hybrid = Config.small_img != 0, Config.LSTM_hybrid != 0
    batch_size = Config.batch_size
    output_folder = os.path.join(Config.results_dir, 'with_text' if lstm_hybrid else 'without_text')
    print(output_folder)
    os.makedirs(output_folder, exist_ok=True)
    model = kwargs.get('model')
    if model is None:
        model = _build_model(Config.infer_precision, SIZE[small], Config.vocab_size, lstm_hybrid, with_discriminator=False)
        prefix = checkpoint.latest_checkpoint(Config.ckpt_dir)
        print('Restore trained model:', prefix)
        if prefix is None:
            raise RuntimeError("no snapshot in %s" % Config.ckpt_dir)
        checkpoint.restore(model, prefix, strict=True)
    queue = PairedEvalInput(Config.dataset_type, batch_size, model.ops, data_base_dir=kwargs.get('data_base_dir', 'data'), small=small,
                            distance_map=Config.distance_map != 0)
    counter, prev_time = 0, float("-inf")
    for b in queue:
        ret = graph_single.build_single_graph(b['images'], b['sketch'], None, b['cls'], None, b['text'], batch_size=batch_size,
                                              training=False, LSTM_hybrid=lstm_hybrid, vocab_size=Config.vocab_size,
                                              data_format=Config.data_format, distance_map=Config.distance_map != 0,
                                              block_type=Config.block_type, model=model, noise=kwargs.get('noise'))
        if counter % 100 == 0:
            curr_time = time()
            print("Now at iteration %d. Elapsed time: %.5fs." % (counter, curr_time - prev_time))
            prev_time = curr_time
        to_img = lambda t: ((np.transpose(t.detach().float().cpu().numpy(), (0, 2, 3, 1)) + 1) / 2.) * 255   # noqa: E731
        generated, target, sketch = to_img(ret[0])[:, :, :, ::-1].astype(np.uint8), to_img(ret[1])[:, :, :, ::-1].astype(np.uint8), \
            to_img(ret[2]).astype(np.uint8)
        for i in range(batch_size):
            stem = '%s_%s' % (b['categories'][i], b['image_names'][i][:-4])
            cv2.imwrite(os.path.join(output_folder, stem + '_output.png'), generated[i])
            cv2.imwrite(os.path.join(output_folder, stem + '_target.png'), target[i])
            cv2.imwrite(os.path.join(output_folder, stem + '_input.png'), sketch[i])
            print('Saved file %s' % b['categories'][i])
        counter += 1
    return counter
