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


def test_snapshot_round_trip_reference_layout(tmp_path):
    from sketchyscenecolorization_b200 import checkpoint
    m = _make_model(torch.float32)        # snapshots hold fp32 (the product's master-weight type)
    m.dstore.adam_v.uniform_(0, 1)
    m.gstore.adam_t, m.dstore.adam_t = 7, 9
    ck = str(tmp_path / "snapshot")
    prefix = checkpoint.save(m, ck, 99, counter=100)
    assert os.path.basename(prefix) == "model_99.ckpt-99"
    for suffix in (".index", ".data-00000-of-00001"):
        assert os.path.exists(prefix + suffix)
    assert 'model_checkpoint_path: "model_99.ckpt-99"' in open(os.path.join(ck, "checkpoint")).read()
    assert checkpoint.latest_checkpoint(ck) == prefix
    assert int(os.path.split(prefix)[1].split('-')[1]) + 1 == 100        # iter_from rule, obj_colorization_main.py:62
    m2 = _make_model(torch.float32)
    m2.initialize(seed=11)
    assert checkpoint.restore(m2, prefix) == 100
    assert torch.equal(m2.gstore.flat, m.gstore.flat) and torch.equal(m2.dstore.flat, m.dstore.flat)
    for k, v in m.dstore.p.items():       # (alignment padding between variables is not part of a snapshot)
        o = m.dstore.offsets[k]
        assert torch.equal(m2.dstore.adam_v[o:o + v.numel()], m.dstore.adam_v[o:o + v.numel()]), k
    assert (m2.gstore.adam_t, m2.dstore.adam_t) == (7, 9)
    for k in m.dstore.state:
        assert torch.equal(m2.dstore.state[k], m.dstore.state[k])
    checkpoint.save(m, ck, 199, counter=200)
    assert checkpoint.latest_checkpoint(ck).endswith("model_199.ckpt-199")
    assert open(os.path.join(ck, "checkpoint")).read().count("all_model_checkpoint_paths") == 2


def test_num_gpu_relaunches_under_torch_distributed_run(monkeypatch):
    """`--num_gpu N` (N in-graph towers in the reference) re-launches the same command with one process per GPU."""
    import sys
    import obj_colorization_main as M
    seen = {}

    def fake_execv(exe, cmd):
        seen["cmd"] = cmd
        raise SystemExit(0)
    monkeypatch.setattr(os, "execv", fake_execv)
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    try:
        M.main(["--mode", "train", "--num_gpu", "4", "--batch_size", "64"])
    except SystemExit:
        pass
    cmd = seen["cmd"]
    assert cmd[0] == sys.executable and cmd[1:3] == ["-m", "torch.distributed.run"]
    assert cmd[cmd.index("--nproc-per-node") + 1] == "4" and cmd[cmd.index("--master-addr") + 1] == "127.0.0.1"
    assert cmd[-6:] == ["--mode", "train", "--num_gpu", "4", "--batch_size", "64"] and cmd[-7].endswith("obj_colorization_main.py")
    assert M.main_procedure.any_rank_true(True) is True and M.main_procedure.shared_string("x") == "x"
