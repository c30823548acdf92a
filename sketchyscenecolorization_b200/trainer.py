"""The two graphs the reference's session loop evaluates every iteration, as explicit step functions.

Reference: graph_single.build_single_graph (:221-314) wires G, D(real), D(fake) and get_losses (:317-581);
main_procedure.train (:202-227) alternates `sess.run([opt_d, loss_d])` and `sess.run([opt_g, loss_g, ...])`,
each on a freshly dequeued batch.  Here:

  d_step = G forward (no tape) + D(real) + D(fake) + loss_d + D backward            (Gf + 6 Df conv FLOPs)
  g_step = G forward + D(fake) + loss_g + D input-gradient + G backward + SN u<-u'  (3 Gf + 2 Df)

followed by `average_gradients` (NCCL all-reduce when world_size > 1, graph_single.py:33-68) and Adam
(beta1=0, beta2=0.9, lr*decay; graph_single.py:139-142,588).
"""
from __future__ import annotations

import math
import os

import torch

from .discriminator import Discriminator
from .generator import Generator
from .params import (ParamStore, discriminator_vars, generator_vars, pix2pix_discriminator_vars, pix2pix_generator_vars,
                     residual_discriminator_vars, residual_generator_vars)
from .pix2pix import Pix2PixDiscriminator, Pix2PixGenerator
from .residual import ResidualDiscriminator, ResidualGenerator

_NETS = {'MRU': (generator_vars, Generator, discriminator_vars, Discriminator),
         'Pix2Pix': (pix2pix_generator_vars, Pix2PixGenerator, pix2pix_discriminator_vars, Pix2PixDiscriminator),
         'Residual': (residual_generator_vars, ResidualGenerator, residual_discriminator_vars, ResidualDiscriminator)}


class _Range:
    """NVTX range (tracing) around a phase of the step; a no-op without CUDA."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        self.on = torch.cuda.is_available()
        if self.on:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if self.on:
            torch.cuda.nvtx.range_pop()
        return False


def lr_decay(counter, max_iter):
    """graph_single.py:139 -- max(0.2, 1 - 0.9*counter/max_iter)."""
    return max(0.2, 1.0 - (float(counter) / max_iter * 0.9))


class FgColorModel:
    """Generator + discriminator + parameter stores on one device."""

    def __init__(self, ops, device, *, size=64, H=192, W=192, vocab_size=58, lstm_hybrid=True,
                 param_dtype=torch.float32, with_discriminator=True, block_type='MRU'):
        if block_type not in _NETS:
            raise NotImplementedError("block_type %r (the reference has MRU, Pix2Pix, Residual)" % (block_type,))
        self.ops, self.device, self.size, self.H, self.W, self.block_type = ops, device, size, H, W, block_type
        gvars, gcls, dvars, dcls = _NETS[block_type]
        self.gstore = ParamStore(gvars(size, vocab_size, H, W), device, param_dtype)
        self.G = gcls(ops, self.gstore, size, lstm_hybrid)
        self.dstore = None
        self.D = None
        if with_discriminator:
            self.dstore = ParamStore(dvars(size), device, param_dtype)
            self.D = dcls(ops, self.dstore, size)

    def initialize(self, seed=0, perturb_tables=0.0):
        self.gstore.initialize(seed, perturb_tables)
        if self.dstore is not None:
            self.dstore.initialize(seed + 1, perturb_tables)

    # ---- inference graph: build_single_graph(training=False) -> image_gens (graph_single.py:257-266)
    def generate(self, sketch_nchw, text_ids_host, labels, noise):
        out, _ = self.G.forward(sketch_nchw, text_ids_host, labels, noise, save=False)
        return self.ops.nhwc_to_nchw(out, out_dtype=torch.float32)

    def generate_replay(self, sketch_nchw, text_ids, labels, noise):
        """`generate` as a CUDA-graph replay (CUDA operator set): batch-1 inference is ~830 launches of microsecond kernels,
        i.e. bound by the host's launch rate; the first call at a batch shape runs eagerly, the second is captured, later ones
        copy the four inputs into the graph's buffers and replay it.  Caption ids go to the device (the <pad> steps are masked
        inside the cells instead of skipped).  The returned tensor is the graph's output buffer: consume it before the next call."""
        if not getattr(self.ops, 'supports_cuda_graphs', False):
            return self.generate(sketch_nchw, text_ids, labels, noise)
        dev = self.device
        key = tuple(sketch_nchw.shape)
        st = self.__dict__.setdefault('_gen_graphs', {}).setdefault(key, dict(calls=0))
        st['calls'] += 1
        ins = dict(sketch=torch.as_tensor(sketch_nchw).to(dev, torch.float32), text=torch.as_tensor(text_ids).to(dev, torch.int32),
                   cls=torch.as_tensor(labels).to(dev, torch.int32), noise=torch.as_tensor(noise).to(dev, torch.float32))
        if st['calls'] == 1:
            return self.generate(ins['sketch'], ins['text'], ins['cls'], ins['noise'])
        if 'graph' not in st:
            st['in'] = {k: torch.empty_like(v) for k, v in ins.items()}
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                st['out'] = self.generate(st['in']['sketch'], st['in']['text'], st['in']['cls'], st['in']['noise'])
            st['graph'] = graph
        for k, v in ins.items():
            st['in'][k].copy_(v, non_blocking=True)
        st['graph'].replay()
        return st['out']

    def _weight_view_ahead(self, need_wgrad):
        """The discriminator's weight view with its spectral normalisation already evaluated -- aside (ops.run_aside: a side
        stream on the CUDA operator set), concurrently with the generator's forward pass that every step starts with; join()
        makes the current stream wait for it.  ~90 launch-latency-bound kernels per step leave the critical path this way
        (99.55 -> 99.0 ms per iteration, profiles/r2t_*)."""
        wv = self.D.new_weight_view(need_wgrad=need_wgrad)
        if self.block_type != 'MRU' or os.environ.get("FGC_SN_SIDE_STREAM", "1") == "0":
            return wv, (lambda: None)
        _, join = self.ops.run_aside(lambda: self.D.prefetch_weights(wv))
        return wv, join

    # ---- loss_d and dL/dtheta_D
    def d_step_grads(self, batch, grads_ready=None):
        """batch: dict(sketch, images, images_d [N,3,H,W] fp32; cls, cls_d int32 [N]; text host [N,15]; noise [N,256]).
        Leaves dL_d/dtheta_D in dstore.grad; returns dict of fp32 0-d loss tensors (total under 'loss').
        grads_ready(lo, hi): see Discriminator.backward (MRU networks; the others finish their gradients at the end)."""
        ops = self.ops
        if self.block_type != 'MRU':
            return self._d_step_grads_pairs(batch)
        self.dstore.grad.zero_()
        wv, join = self._weight_view_ahead(True)
        fake, _ = self.G.forward(batch["sketch"], batch["text"], batch["cls"], batch["noise"], save=False)
        join()
        real = ops.nchw_to_nhwc(batch["images_d"])
        # D(real) and D(fake) as ONE pass over the 2N images: the discriminator has no batch statistics (PReLU, per-sample
        # min-max gates, spectral norm), so this equals the reference's two instantiations (graph_single.py:269,271) while
        # every kernel sees twice the work per launch.  The loss means stay per half (:401-402).
        N = real.shape[0]
        both = torch.empty((2 * N,) + tuple(real.shape[1:]), dtype=real.dtype, device=real.device)
        both[:N].copy_(real)
        both[N:].copy_(fake)
        del real, fake
        d, lg, dctx = self.D.forward(both, wv)
        l_real, g_rd = ops.softplus_mean(d[:N], -1.0)                    # graph_single.py:402
        l_fake, g_fd = ops.softplus_mean(d[N:], 1.0)                     # :402
        l_ac, g_rl = ops.ce_loss(lg[:N], batch["cls_d"], True, 1.0)     # :343-348 (focal, ld1 = 1), real images only
        g_d = torch.empty_like(d)
        g_d[:N].copy_(g_rd)
        g_d[N:].copy_(g_fd)
        g_l = torch.zeros_like(lg)
        g_l[:N].copy_(g_rl)
        self.D.backward(g_d, g_l, dctx, need_x_grad=False, grads_ready=grads_ready)
        del dctx
        wv.finish_backward()
        reg = ops.reg_loss(self.dstore)                                  # :571
        return dict(loss=l_real + l_fake + l_ac + reg, gan=l_real + l_fake, ac=l_ac, reg=reg)

    def _d_step_grads_pairs(self, batch):
        """Pix2Pix / Residual discriminators: they look at (sketch, image) pairs and batch-normalise, so the real and the fake
        pass stay two instantiations with their own batch statistics (graph_single.py:269-272)."""
        ops = self.ops
        self.dstore.grad.zero_()
        fake, _ = self.G.forward(batch["sketch"], batch["text"], batch["cls"], batch["noise"], save=False)
        wv = self.D.new_weight_view(need_wgrad=True)
        sk = ops.nchw_to_nhwc(batch["sketch"])
        d, lg, dctx = self.D.forward(sk, ops.nchw_to_nhwc(batch["images_d"]), wv)
        l_real, g_rd = ops.softplus_mean(d, -1.0)
        l_ac, g_rl = ops.ce_loss(lg, batch["cls_d"], True, 1.0)
        self.D.backward(g_rd, g_rl, dctx, need_x_grad=False)
        d, lg, dctx = self.D.forward(sk, fake, wv)
        l_fake, g_fd = ops.softplus_mean(d, 1.0)
        self.D.backward(g_fd, None, dctx, need_x_grad=False)
        del dctx
        wv.finish_backward()
        reg = ops.reg_loss(self.dstore)
        return dict(loss=l_real + l_fake + l_ac + reg, gan=l_real + l_fake, ac=l_ac, reg=reg)

    # ---- loss_g and dL/dtheta_G (+ SN u update)
    def g_step_grads(self, batch, grads_ready=None):
        ops = self.ops
        self.gstore.grad.zero_()
        wv, join = self._weight_view_ahead(False)
        fake, gctx = self.G.forward(batch["sketch"], batch["text"], batch["cls"], batch["noise"], save=True)
        join()
        if self.block_type != 'MRU':
            fd, fl, fctx = self.D.forward(ops.nchw_to_nhwc(batch["sketch"]), fake, wv)
        else:
            fd, fl, fctx = self.D.forward(fake, wv)
        l_gan, g_fd = ops.softplus_mean(fd, -1.0)                        # graph_single.py:401
        l_ac, g_fl = ops.ce_loss(fl, batch["cls"], False, 0.5)          # :350-352 (ld2 = 0.5)
        g_fake = self.D.backward(g_fd, g_fl, fctx, need_x_grad=True)
        del fctx
        target = ops.nchw_to_nhwc(batch["images"])
        l_l1, g_l1 = ops.smooth_l1(target, fake, 100.0)                  # :552-555,575
        ops.add_(g_fake, g_l1)
        if grads_ready is not None and self.block_type == 'MRU':
            self.G.backward(g_fake, gctx, grads_ready=grads_ready)
        else:
            self.G.backward(g_fake, gctx)
        wv.commit_u()                                                    # :178-180,208-210
        reg = ops.reg_loss(self.gstore)
        return dict(loss=l_gan + l_ac + l_l1 + reg, gan=l_gan, ac=l_ac, l1=l_l1, reg=reg)


class FgColorTrainer:
    """Alternating D / G optimisation with optional data-parallel gradient averaging.

    use_cuda_graphs=True (CUDA operator set only): the first call of each step runs eagerly, the second one is captured
    into a CUDA graph (kernels, memsets, the NCCL all-reduce and the fused Adam), and every later call copies the batch
    into the graph's static input buffers and replays it -- the ~3500 per-step kernel launches cost one graph launch
    on the host.  Nothing inside a captured step depends on host data: caption ids stay on the device (pad steps are
    masked in the LSTM kernel instead of skipped) and the Adam step size is read from a device scalar."""

    D_KEYS = ("sketch", "images_d", "cls", "cls_d", "text", "noise")
    G_KEYS = ("sketch", "images", "cls", "text", "noise")

    def __init__(self, model, *, lr_g=2e-4, lr_d=1e-4, max_iter=100000, process_group=None, world_size=1,
                 use_cuda_graphs=False, optimizer='Adam', overlap_allreduce=None):
        self.m, self.lr_g, self.lr_d, self.max_iter = model, lr_g, lr_d, max_iter
        # gradient averaging: one all-reduce at the end of the backward pass (default), or started bucket by bucket under it
        # (FGC_OVERLAP_ALLREDUCE=1 / overlap_allreduce=True).  Measured at 2 GPUs (profiles/r2s_*): 99.6 against 99.1 ms per
        # iteration -- the collective's CTAs share the SMs with persistent one-CTA-per-SM convolution kernels, and at NVLink
        # speed the exposed time they save (~0.5 ms) is less than what they cost; it is an option for slower fabrics.
        self.overlap = (os.environ.get("FGC_OVERLAP_ALLREDUCE", "0") == "1") if overlap_allreduce is None else bool(overlap_allreduce)
        self.optimizer = optimizer.lower()                    # graph_single.get_optimizer (:584-593)
        for store in (model.gstore, model.dstore):
            if store is not None and store.optimizer != self.optimizer:
                store.set_optimizer(self.optimizer)
        self.pg, self.world = process_group, world_size
        self.counter = 0
        self.use_graphs = use_cuda_graphs
        self._g = {}          # kind -> dict(graph, inputs, outputs, lr, calls)
        self.launches_per_step = {}

    def _allreduce(self, store):
        """average_gradients (graph_single.py:33-68): one all-reduce of the flat fp32 gradient bucket."""
        if self.world > 1:
            import torch.distributed as dist
            if dist.get_backend(self.pg) == "nccl":       # ncclAvg: the division rides in the collective
                dist.all_reduce(store.grad, op=dist.ReduceOp.AVG, group=self.pg)
            else:                                         # gloo (CPU tests) has no AVG
                dist.all_reduce(store.grad, op=dist.ReduceOp.SUM, group=self.pg)
                store.grad.mul_(1.0 / self.world)

    class _Buckets:
        """The same average, started piecewise: ready(lo, hi) launches the all-reduce of grad[lo:hi] asynchronously (NCCL: on
        its own stream, ordered after everything enqueued so far -- inside a captured step a parallel branch of the graph) as
        soon as the backward pass has finished that range (deep layers first: they hold most parameters and finish first);
        finish() sends whatever was not announced and joins.  Every rank issues the same collectives in the same order."""

        def __init__(self, trainer, store):
            import torch.distributed as dist
            self.dist, self.tr, self.store = dist, trainer, store
            self.avg = dist.get_backend(trainer.pg) == "nccl"
            self.works, self.done = [], []

        def ready(self, lo, hi):
            if hi <= lo:
                return
            t = self.store.grad[lo:hi]
            op = self.dist.ReduceOp.AVG if self.avg else self.dist.ReduceOp.SUM
            self.works.append((self.dist.all_reduce(t, op=op, group=self.tr.pg, async_op=True), t))
            self.done.append((lo, hi))

        def finish(self):
            pos = 0
            for lo, hi in sorted(self.done):
                assert lo >= pos, "overlapping gradient buckets"
                self.ready(pos, lo)
                pos = hi
            self.ready(pos, self.store.n_flat)
            for w, t in self.works:
                w.wait()
                if not self.avg:
                    t.mul_(1.0 / self.tr.world)

    # ---- eager steps
    def _apply(self, store, lr, lr_dev):
        if self.optimizer == 'adam':
            self.m.ops.adam_step(store, lr, lr_dev=lr_dev)
        else:
            self.m.ops.optimizer_step(store, self.optimizer, lr, lr_dev=lr_dev)

    def _d_eager(self, batch, lr_dev=None):
        bk = self._Buckets(self, self.m.dstore) if self.world > 1 and self.overlap else None
        with _Range("fgc.d_step.forward_backward"):
            out = self.m.d_step_grads(batch, grads_ready=bk.ready) if bk else self.m.d_step_grads(batch)
        with _Range("fgc.d_step.allreduce"):
            bk.finish() if bk else self._allreduce(self.m.dstore)
        with _Range("fgc.d_step.optimizer"):
            self._apply(self.m.dstore, self.lr_d * lr_decay(self.counter, self.max_iter), lr_dev)
        return out

    def _g_eager(self, batch, lr_dev=None):
        bk = self._Buckets(self, self.m.gstore) if self.world > 1 and self.overlap else None
        with _Range("fgc.g_step.forward_backward"):
            out = self.m.g_step_grads(batch, grads_ready=bk.ready) if bk else self.m.g_step_grads(batch)
        with _Range("fgc.g_step.allreduce"):
            bk.finish() if bk else self._allreduce(self.m.gstore)
        with _Range("fgc.g_step.optimizer"):
            self._apply(self.m.gstore, self.lr_g * lr_decay(self.counter, self.max_iter), lr_dev)
        return out

    # ---- CUDA-graph steps
    def _graph_step(self, kind, batch):
        store = self.m.dstore if kind == "d" else self.m.gstore
        keys = self.D_KEYS if kind == "d" else self.G_KEYS
        eager = self._d_eager if kind == "d" else self._g_eager
        base_lr = self.lr_d if kind == "d" else self.lr_g
        st = self._g.setdefault(kind, dict(calls=0))
        st["calls"] += 1
        dev = self.m.device
        if st["calls"] == 1:                    # lazy one-time initialisation inside the library happens here
            onto = {}
            for k in keys:
                v = batch[k] if torch.is_tensor(batch[k]) else torch.as_tensor(batch[k])
                onto[k] = v.to(dev, dtype=torch.float32 if v.is_floating_point() else torch.int32, non_blocking=True)
            return eager(onto)
        if "graph" not in st:
            st["inputs"] = {}
            for k in keys:
                v = batch[k]
                v = torch.as_tensor(v) if not torch.is_tensor(v) else v
                st["inputs"][k] = torch.empty(v.shape, dtype=torch.int32 if not v.is_floating_point() else torch.float32,
                                              device=dev)
            st["lr"] = torch.zeros((), dtype=torch.float32, device=dev)
            n0 = self.m.ops.launch_count() if hasattr(self.m.ops, "launch_count") else 0
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                st["outputs"] = eager(dict(st["inputs"]), lr_dev=st["lr"])
            st["graph"] = graph
            if hasattr(self.m.ops, "launch_count"):
                self.launches_per_step[kind] = self.m.ops.launch_count() - n0
        for k in keys:
            v = batch[k]
            v = torch.as_tensor(v) if not torch.is_tensor(v) else v
            st["inputs"][k].copy_(v, non_blocking=True)
        store.adam_t += 1
        bias = math.sqrt(1.0 - 0.9 ** store.adam_t) if self.optimizer == 'adam' else 1.0      # Adam's step-size correction
        st["lr"].fill_(base_lr * lr_decay(self.counter, self.max_iter) * bias)
        with _Range("fgc.%s_step.graph_replay" % kind):
            st["graph"].replay()
        return st["outputs"]

    def d_step(self, batch):
        if self.use_graphs:
            return self._graph_step("d", batch)
        return self._d_eager(batch)

    def g_step(self, batch):
        out = self._graph_step("g", batch) if self.use_graphs else self._g_eager(batch)
        self.counter += 1                                                # counter_addition_op, main_procedure.py:106,221
        return out
