"""Data-parallel host logic on CPU: world_size-2 gloo run of FgColorTrainer (one all-reduce(avg) of the flat gradient
bucket per optimiser step, per-rank BN statistics -- the reference's tower semantics, graph_single.py:33-68,146-166),
and the snapshot round trip in the reference's file layout."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import fgcolor_oracle as O

SIZE, H, W, N = 16, 64, 64, 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _batch(seed):
    b = O.make_batch(N, H, W, seed, torch.float64)
    bb = dict(b)
    bb["cls"], bb["cls_d"], bb["text"] = b["cls"].int(), b["cls_d"].int(), b["text"].numpy()
    return bb


def _make_model(dtype=torch.float64):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from sketchyscenecolorization_b200.trainer import FgColorModel
    from torch_ops import TorchOps
    m = FgColorModel(TorchOps(dtype), "cpu", size=SIZE, H=H, W=W, param_dtype=dtype)
    m.initialize(seed=3, perturb_tables=0.1)
    return m


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sketchyscenecolorization_b200.trainer import FgColorTrainer
    m = _make_model()
    tr = FgColorTrainer(m, max_iter=100, process_group=dist.group.WORLD, world_size=world)
    tr.d_step(_batch(10 + rank))
    tr.g_step(_batch(20 + rank))
    # session-loop agreement helpers (main_procedure.train under data parallelism)
    from sketchyscenecolorization_b200 import main_procedure as MP
    agree = [MP.any_rank_true(rank == 1, dist.group.WORLD, world), MP.any_rank_true(False, dist.group.WORLD, world)]
    stamp = MP.shared_string("stamp-from-rank-%d" % rank, dist.group.WORLD, world)
    torch.save(dict(d=m.dstore.flat.clone(), g=m.gstore.flat.clone(), dg=m.dstore.grad.clone(), gg=m.gstore.grad.clone(),
                    agree=agree, stamp=stamp),
               os.path.join(out_dir, "rank%d.pt" % rank))
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_matches_mean_of_shard_gradients(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = torch.load(tmp_path / "rank0.pt"), torch.load(tmp_path / "rank1.pt")
    # replicas stay bit-identical: same averaged gradient, same Adam update
    for k in ("d", "g", "dg", "gg"):
        assert torch.equal(r0[k], r1[k]), k
    # a NaN seen by ONE rank ends the loop on every rank; the run directory's time stamp is rank 0's everywhere
    assert r0["agree"] == [True, False] and r1["agree"] == [True, False]
    assert r0["stamp"] == r1["stamp"] == "stamp-from-rank-0"
    # and the averaged D gradient is the mean of the two shard gradients computed independently
    from torch_ops import TorchOps
    ops = TorchOps(torch.float64)
    shard = []
    for rank in range(world):
        m = _make_model()
        m.d_step_grads(_batch(10 + rank))
        shard.append(m.dstore.grad.clone())
    m = _make_model()
    mean = (shard[0] + shard[1]) / 2
    m.dstore.grad.copy_(mean)
    ops.add_reg_grad(m.dstore)            # the step adds the decay gradient before Adam
    assert (m.dstore.grad - r0["dg"]).abs().max().item() < 1e-12


def _worker_modes(rank, world, port, out_dir):
    """The bucketed all-reduce (started under the backward pass) against the single one at its end."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sketchyscenecolorization_b200.trainer import FgColorTrainer
    res = {}
    for overlap in (True, False):
        m = _make_model()
        tr = FgColorTrainer(m, max_iter=100, process_group=dist.group.WORLD, world_size=world, overlap_allreduce=overlap)
        seen = []
        if overlap:
            orig = tr._Buckets.ready

            def spy(self, lo, hi, _o=orig):
                seen.append((self.store.n_flat, lo, hi))
                return _o(self, lo, hi)
            tr._Buckets.ready = spy
        tr.d_step(_batch(10 + rank))
        tr.g_step(_batch(20 + rank))
        if overlap:
            tr._Buckets.ready = orig
        res[overlap] = dict(dg=m.dstore.grad.clone(), gg=m.gstore.grad.clone(), d=m.dstore.flat.clone(), g=m.gstore.flat.clone(),
                            seen=seen, nd=m.dstore.n_flat, ng=m.gstore.n_flat)
    torch.save(res, os.path.join(out_dir, "modes%d.pt" % rank))
    dist.destroy_process_group()


def test_bucketed_allreduce_equals_single_allreduce(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker_modes, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for rank in range(world):
        r = torch.load(tmp_path / ("modes%d.pt" % rank))
        for k in ("dg", "gg", "d", "g"):
            assert torch.equal(r[True][k], r[False][k]), k          # elementwise sums: the split does not change a bit
        seen = r[True]["seen"]
        # several buckets per network were announced before the end of the backward pass, and together they tile each buffer
        assert r[True]["nd"] != r[True]["ng"]
        for n in (r[True]["nd"], r[True]["ng"]):
            edges = sorted((lo, hi) for nf, lo, hi in seen if nf == n and hi > lo)
            assert len(edges) >= 3
            assert edges[0][0] == 0 and edges[-1][1] == n and all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
