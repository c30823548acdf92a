#!/usr/bin/env python
"""Headline benchmark: fg-colorization training images/sec @192x192, bs 64 per GPU (BASELINE.json configs[1]).

One step = one training iteration of the reference's session loop (main_procedure.py:202-227): a D step
(G forward, D(real), D(fake), D backward, Adam) and a G step (G forward, D(fake), D input-gradient, G backward,
Adam, SN u update) on two different synthetic batches, through hand-written sm_100a kernels (libfgcolor.so).

  python bench.py [--gpus N] [--steps K] [--warmup W]             # product arm (one process per GPU under torchrun)
  python bench.py --impl reference ...                            # CPU restatement of the reference graph (oracle)

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H = W = 192
BS = 64
SIZE = 64
GF_GFLOP = 57.989          # generator conv forward, GFLOP / image   
DF_GFLOP = 64.633          # discriminator conv forward
STEP_GFLOP = 4 * GF_GFLOP + 8 * DF_GFLOP      # 749.02 GFLOP / image / iteration
METRIC = "fg-colorization train images/sec @192x192 bs64 per GPU"
WORKLOAD = ("fg-colorization MRU G+D training iteration (D step + G step), 192x192, bs 64/GPU, 15-token captions "
            "(BASELINE.json configs[1])")


def block_type_flops(block_type):
    """(generator, discriminator) conv forward GFLOP / image at 192x192 -- algorithmic (true 4x4 taps, not the zero-padded
    phase form the Pix2Pix layers run in).  None where not tabulated."""
    if block_type == "MRU":
        return GF_GFLOP, DF_GFLOP
    if block_type == "Pix2Pix":       # models_collection.py:408-538, 789-841 (size 64)
        enc = [(96, 3, 64), (48, 64, 128), (24, 128, 256), (12, 256, 512), (6, 512, 512)]            # (output side, Cin, Cout)
        dec = [(6, 576, 512), (12, 1024, 256), (24, 512, 128), (48, 256, 64), (96, 128, 3)]          # (INPUT side, Cin, Cout)
        dis = [(96, 6, 64), (48, 64, 128), (24, 128, 256), (23, 256, 512), (22, 512, 1)]
        mac = lambda rows: sum(s * s * 16 * ci * co for s, ci, co in rows)                           # noqa: E731
        return 2 * (mac(enc) + mac(dec)) / 1e9, 2 * mac(dis) / 1e9
    return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], burst=d["bf16_tflops"], sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured")
    return dict(hbm=6650.0, burst=1590.0, sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def synth_batch(n, seed):
    """+-1 stroke sketches (~5% black), U(-1,1) images, labels U{0..24}, 15 non-pad ids, N(0,1) noise."""
    import torch
    g = torch.Generator().manual_seed(seed)
    sk = torch.ones(n, 1, H, W)
    for i in range(n):
        for _ in range(6):
            p = torch.rand(5, 2, generator=g) * torch.tensor([H - 1.0, W - 1.0])
            for a, b in zip(p[:-1], p[1:]):
                L = int(max(abs(b[0] - a[0]), abs(b[1] - a[1]))) + 1
                ys = torch.linspace(a[0].item(), b[0].item(), L).round().long()
                xs = torch.linspace(a[1].item(), b[1].item(), L).round().long()
                sk[i, 0, ys, xs] = -1.0
    return dict(sketch=sk.expand(n, 3, H, W).contiguous(),
                images=torch.rand(n, 3, H, W, generator=g) * 2 - 1,
                images_d=torch.rand(n, 3, H, W, generator=g) * 2 - 1,
                cls=torch.randint(0, 25, (n,), generator=g).int(), cls_d=torch.randint(0, 25, (n,), generator=g).int(),
                text=torch.randint(2, 58, (n, 15), generator=g).int(), noise=torch.randn(n, 256, generator=g))


def write_synthetic_tfrecords(base, n_records=256, files=4, seed=7):
    """A synthetic dataset in the reference's own on-disk format (data_preparation.py:21-32: raw 384x384x3 uint8 picture and
    sketch, class id, 15 caption ids) for `--input tfrecord`: the e2e leg then starts from RECORD FILES -- framing, CRC-32C,
    proto parsing, shuffle queue, raw uint8 host-to-device copy, fgc_paired_input -- instead of from ready fp32 tensors."""
    import numpy as np
    from sketchyscenecolorization_b200 import tfrecord_input as TI
    d = os.path.join(base, "tfrecord", "train")
    os.makedirs(d, exist_ok=True)
    rng = np.random.default_rng(seed)
    per = n_records // files
    for f in range(files):
        recs = []
        for i in range(per):
            sk = np.full((384, 384, 3), 255, np.uint8)
            for _ in range(6):
                r, c = rng.integers(8, 370, 2)
                sk[r:r + 2, c // 2:c // 2 + 180] = 0
                sk[r // 2:r // 2 + 180, c:c + 2] = 0
            recs.append(TI.encode_example(dict(
                ImageName=("img%d_%d.png" % (f, i)).encode(), cartoon_data=rng.integers(0, 256, (384, 384, 3), dtype=np.uint8).tobytes(),
                sketch_data=sk.tobytes(), Category=b"bus", Category_id=int(rng.integers(0, 25)), Color_text=b"synthetic",
                Text_vocab_indices=rng.integers(2, 58, 15, dtype=np.uint8).tobytes())))
        TI.write_tfrecord(os.path.join(d, "part%d.tfrecord" % f), recs)
    return base


# ----------------------------------------------------------------------------------------------------------
# reference arm: the CPU restatement of the reference graph (TensorFlow 1.x cannot be installed here)
# ----------------------------------------------------------------------------------------------------------
def cpu_reference_images_per_sec(steps, warmup, bs=2, seed=0):
    import torch
    from oracle import fgcolor_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    gspecs, dspecs = O.generator_specs(SIZE, 58, H, W), O.discriminator_specs(SIZE)
    gp = {k: v.requires_grad_(v.is_floating_point()) for k, v in O.init_params(gspecs, seed).items()}
    dp = {k: v.requires_grad_(v.is_floating_point()) for k, v in O.init_params(dspecs, seed + 1).items()}
    vg = {k: torch.zeros_like(v) for k, v in gp.items()}
    vd = {k: torch.zeros_like(v) for k, v in dp.items()}
    bA, bB = O.make_batch(bs, H, W, 1), O.make_batch(bs, H, W, 2)

    def one_iter(t):
        ld, _, _ = O.d_step_loss(gp, dp, gspecs, dspecs, bA, SIZE)
        gd = O.grads_of(ld, dp, dspecs)
        with torch.no_grad():
            for k, g in gd.items():
                new, vd[k] = O.adam_update(dp[k], g, vd[k], 1e-4, t)
                dp[k].copy_(new)
        lg, _, u_new, _ = O.g_step_loss(gp, dp, gspecs, dspecs, bB, SIZE)
        gg = O.grads_of(lg, gp, gspecs)
        with torch.no_grad():
            for k, g in gg.items():
                new, vg[k] = O.adam_update(gp[k], g, vg[k], 2e-4, t)
                gp[k].copy_(new)
            for k, u in u_new.items():
                dp[k].copy_(u)
        return float(ld.detach()), float(lg.detach())

    for i in range(warmup):
        one_iter(i + 1)
    t0 = time.perf_counter()
    for i in range(steps):
        one_iter(warmup + i + 1)
    dt = time.perf_counter() - t0
    return bs * steps / dt, dt / steps, cores, bs


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 3)), min(args.warmup, 1)
    ips, spi, cores, bs = cpu_reference_images_per_sec(steps, warmup)
    sample = "bs %d (the reference default batch), %d timed iteration(s) after %d warm-up, same 192x192 graph" % (bs, steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": spi * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "sample": sample,
                       "note": "CPU restatement of the reference TF1 graph (oracle/fgcolor_oracle.py, torch-CPU autograd); "
                               "TensorFlow 1.x is not installable in this image"},
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------
# product arm
# ----------------------------------------------------------------------------------------------------------
def _profile_json(name):
    """A committed ncu-derived summary under profiles/ (written by scripts/ncu_traffic.py / scripts/tensor_pipe_summary.py from
    captures of THIS kernel set); None when absent -- nothing is pasted into this file."""
    p = os.path.join(ROOT, "profiles", name)
    try:
        return json.load(open(p))
    except Exception:
        return None


def dominant_kernel_roofline(ops, torch, pk, reps=7):
    """The dominant kernel of the step: the 128->128 3x3 conv at 192x192 (D unit-1 Conv_2 / G decoder unit 8), bs 64,
    single-pass bf16.  Algorithmic FLOPs per launch = 2 * 64 * 192^2 * 9 * 128 * 128.  Timed twice with CUDA events on the
    launching stream: the whole fgc_conv2d_fwd call (weight-packing launch + convolution kernel: what a layer costs inside the
    step) and the convolution kernel ALONE (fgc_debug_keep_packed: the workspace keeps the tiles of the previous identical call)
    -- `achieved` is the latter, per launch of that kernel, like the ncu figures beside it."""
    from sketchyscenecolorization_b200._lib import check
    from sketchyscenecolorization_b200.ops_base import ACT_NONE
    dev = ops.device
    x = torch.randn(BS, H, W, 128, device=dev).to(torch.bfloat16)
    wgt = (torch.randn(3, 3, 128, 128, device=dev) * 0.02).contiguous()
    b = torch.zeros(128, device=dev)
    y = torch.empty((BS, H, W, 128), dtype=torch.bfloat16, device=dev)
    flop = 2.0 * BS * H * W * 9 * 128 * 128
    arr, n_, h_, w_, dt = ops._srcs([(x, False)])
    ws = ops._ws([128], 3, 128, dt)
    stream = torch.cuda.current_stream().cuda_stream

    def call():
        check(ops.lib.fgc_conv2d_fwd(arr, 1, dt, n_, h_, w_, wgt.data_ptr(), 3, 128, 128, b.data_ptr(), 1, 1, 1, h_, w_, ACT_NONE,
                                     y.data_ptr(), ops._dt(y), ws.data_ptr(), stream), "conv2d_fwd (roofline)")

    def timed():
        for _ in range(2):
            call()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for e0, e1 in evs:        # input (604 MB) + output (604 MB) per launch exceed the 126 MB L2
            e0.record()
            call()
            e1.record()
        torch.cuda.synchronize()
        return sorted(e0.elapsed_time(e1) for e0, e1 in evs)[len(evs) // 2]

    ms_call = timed()
    ops.lib.fgc_debug_keep_packed(1)
    try:
        ms = timed()
    finally:
        ops.lib.fgc_debug_keep_packed(0)
    ach = flop / (ms * 1e-3) / 1e12
    tr = _profile_json("dominant_kernel_traffic.json") or {}
    return {"bound": "tensor", "kernel": tr.get("kernel", "conv_halo_kernel<128,2,2,swap> 3x3 128->128 @192x192 bs64"),
            "achieved": round(ach, 2), "peak": pk["burst"], "unit": "TFLOP/s", "frac": round(ach / pk["burst"], 4),
            "peak_source": pk["src"] + " bf16 burst", "ms_per_launch": round(ms, 4), "flop_per_launch": flop,
            "ms_per_call_incl_weight_pack": round(ms_call, 4),
            "achieved_incl_weight_pack": round(flop / (ms_call * 1e-3) / 1e12, 2),
            # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed `ncu --set full` capture
            "traffic": tr.get("dram_bytes_per_launch"), "traffic_unit": "bytes/launch (DRAM)",
            "traffic_source": tr.get("source"), "l2_to_sm_bytes_per_launch": tr.get("l2_to_sm_bytes_per_launch"),
            "tensor_pipe_pct": tr.get("tensor_pipe_pct")}


def run_product(args):
    import torch
    import torch.distributed as dist
    from sketchyscenecolorization_b200.cuda_ops import CudaOps
    from sketchyscenecolorization_b200.main_procedure import TrainSession
    from sketchyscenecolorization_b200.trainer import FgColorModel
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    pg = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(dev))
        pg = dist.group.WORLD
    ops = CudaOps(dev, torch.bfloat16)          # training mode: bf16 NHWC activations, fp32 master weights / stats / Adam
    model = FgColorModel(ops, dev, size=SIZE, H=H, W=W, block_type=args.block_type)
    model.initialize(seed=0)                    # identical on every rank (reference initialisers)
    fl = block_type_flops(args.block_type)
    # D step = Gf + 6 Df for MRU; the pair discriminators run real and fake separately, same conv work
    step_gflop = (4 * fl[0] + 8 * fl[1]) if fl else None
    workload = WORKLOAD if args.block_type == "MRU" else WORKLOAD.replace("MRU G+D", args.block_type + " G+D").replace(
        " (BASELINE.json configs[1])", " (--block_type %s; BASELINE.json configs[1] is the MRU default)" % args.block_type)
    graphs = not args.no_graphs

    hostA, hostB = synth_batch(BS, 1234 + rank), synth_batch(BS, 4321 + rank)
    pin = lambda b: {k: v.pin_memory() for k, v in b.items()}  # noqa: E731
    hostA, hostB = pin(hostA), pin(hostB)

    class Cycle:            # the two queues of main_procedure.train: here pinned host batches, handed out again and again
        def __init__(self, *batches):
            self.b, self.i = batches, 0

        def __iter__(self):
            return self

        def __next__(self):
            self.i += 1
            return self.b[(self.i - 1) % len(self.b)]

    # The session object of the reference-facing entry point (main_procedure.train is a loop over TrainSession.iteration):
    # the e2e leg below times exactly what `obj_colorization_main.py --mode train` runs per iteration.
    queues = None
    if args.input == "tfrecord":        # start from record files in the reference's dataset format (rank 0 writes them)
        import tempfile
        from sketchyscenecolorization_b200.tfrecord_input import PairedTrainInput
        base = os.path.join(tempfile.gettempdir(), "fgc_bench_records_%d" % os.getuid())
        if local == 0 and not os.path.isdir(os.path.join(base, "tfrecord", "train")):
            write_synthetic_tfrecords(base)
        if world > 1:
            dist.barrier()
        queues = [PairedTrainInput(BS, ops, base, min_after_dequeue=128, seed=s0 + rank, num_threads=6) for s0 in (1234, 4321)]
    sess = TrainSession(model, batch_size=BS, max_iter=100000, lr_g=2e-4, lr_d=1e-4, process_group=pg, world_size=world,
                        use_cuda_graphs=graphs,
                        input_iter=queues[0] if queues else Cycle(hostA, hostB),
                        input_iter_d=queues[1] if queues else Cycle(hostA))
    tr = sess.tr

    def to_dev(hb):
        d = {k: v.to(dev, non_blocking=True) for k, v in hb.items() if k != "text"}
        # caption ids: eager mode keeps a host copy (it drives the pad-skip control flow); graph mode keeps them on the device
        d["text"] = hb["text"].to(dev, non_blocking=True) if graphs else hb["text"].numpy()
        return d

    def h2d_bytes(hb, keys):
        return sum(hb[k].numel() * hb[k].element_size() for k in keys)

    devA, devB = to_dev(hostA), to_dev(hostB)
    torch.cuda.synchronize()

    def iteration(bA, bB):
        od = tr.d_step(bA)
        og = tr.g_step(bB)
        return od["loss"], og["loss"]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup + (2 if graphs else 0)):      # +2 untimed set-up steps: eager run, then graph capture
        iteration(devA, devB)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = ops.launch_count()
    ncu_range = bool(os.environ.get("FGC_NCU_RANGE"))      # `ncu --profile-from-start off`: list only the timed region
    if ncu_range:
        torch.cuda.profiler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t_host0 = time.perf_counter()
    for _ in range(args.steps):
        ld, lg = iteration(devA, devB)
    host_ms = (time.perf_counter() - t_host0) * 1e3 / args.steps      # time to ENQUEUE one step (no sync inside)
    e1.record()
    barrier()
    if ncu_range:
        torch.cuda.profiler.stop()
    launches = ops.launch_count() - n0
    if graphs:          # launches are recorded once at capture; every replay re-issues all of them
        launches = args.steps * sum(tr.launches_per_step.values())
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    ld_v, lg_v = float(ld), float(lg)

    # ---- end to end, through the session loop's own iteration: pinned host buffers (or record files) -> device every step,
    # the loss scalars and the NaN verdict read back every step
    d_keys = ("sketch", "images_d", "cls", "cls_d", "text")
    g_keys = ("sketch", "images", "cls", "text")
    for _ in range(2):                  # fill the shuffle buffers / prefetch pipeline outside the timed region
        sess.iteration()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        e2e_ld, e2e_lg, nan_d, nan_g = sess.iteration()
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)

    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()
    if rank == 0:
        pk = peaks()
        ips = BS * world * args.steps / (ms * 1e-3)
        ips_e2e = BS * world * args.steps / (ms_e2e * 1e-3)
        roof = dominant_kernel_roofline(ops, torch, pk)
        step_tflops = (step_gflop * 1e9 * ips / world / 1e12) if step_gflop else None
        cores = os.cpu_count() or 1
        cpu = None
        if world == 1 and not args.no_cpu_baseline and args.block_type == "MRU":     # the CPU arm restates the MRU graph
            c_ips, c_spi, cores, c_bs = cpu_reference_images_per_sec(1, 1)
            cpu = {"value": c_ips, "unit": "images/s", "cores": cores, "kind": "port",
                   "sample": "oracle (torch-CPU restatement of the TF1 graph), bs %d, 1 timed training iteration after 1 warm-up" % c_bs}
        line = {"metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": {"workload": workload,
                           "global_batch": BS * world, "parallelism": "dp%d" % world,
                           "precision": "bf16 activations + single-pass bf16 tcgen05 convs, fp32 accumulate/master/BN/LSTM/Adam",
                           "l2_policy": "per-step working set (tens of GB of activations) far exceeds the 126 MB L2",
                           "conv_flop_per_image": step_gflop * 1e9 if step_gflop else None,
                           "step_conv_tflops_per_gpu": round(step_tflops, 2) if step_tflops else None,
                           "step_conv_frac_of_sustained_peak": round(step_tflops / pk["sustained"], 4) if step_tflops else None,
                           "host_enqueue_ms_per_step": round(host_ms, 2),
                           "cuda_graphs": graphs,
                           "g_conv_tensor_pipe_pct": (_profile_json("g_conv_tensor_pipe.json") or {}).get("g_conv_tensor_pipe_pct"),
                           "g_conv_tensor_pipe_source": (_profile_json("g_conv_tensor_pipe.json") or {}).get("source"),
                           "loss_d": ld_v, "loss_g": lg_v},
                "clocks": clocks,
                "e2e": {"value": ips_e2e, "unit": "images/s",
                        "h2d_bytes_per_step": (h2d_bytes(hostA, d_keys) + h2d_bytes(hostB, g_keys)) if queues is None
                        else 2 * BS * (2 * 384 * 384 * 3 + 15 * 4 + 4),
                        "d2h_bytes_per_step": 16,
                        "api": "main_procedure.TrainSession.iteration (the body of main_procedure.train's loop)",
                        "input": "pinned fp32 tensors" if queues is None else
                        "TFRecord files -> mapped reader + CRC-32C + proto parse + shuffle queue (host threads) -> raw uint8 H2D -> "
                        "fgc_paired_input"},
                "gpu_launches": launches,
                "roofline": roof}
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        # Leave without tearing NCCL down: destroying the process group while CUDA graphs that captured its all-reduce
        # are still alive hung the 2-GPU run after the result line was printed (r1j).  Everything is flushed; exit now.
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


# ----------------------------------------------------------------------------------------------------------
# the other configurations of BASELINE.json, one line each (not the driver's headline): --mode infer | bg
# ----------------------------------------------------------------------------------------------------------
def _time_calls(torch, fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for e0, e1 in ev:
        e0.record()
        fn()
        e1.record()
    torch.cuda.synchronize()
    ms = sorted(e0.elapsed_time(e1) for e0, e1 in ev)
    return ms[len(ms) // 2], ms[0]


def run_infer(args):
    """BASELINE.json configs[0]: one 192x192 sketch + 'the bus is orange' through the generator (main_procedure.py:593-597), in
    the parity mode inference runs in (fp32 storage, bf16x3 tensor-core convolutions), checked against the fp64 oracle in the
    same run.  value = latency of one picture resident on the device; e2e = host sketch in, host picture out."""
    import torch
    from oracle import fgcolor_oracle as O
    from sketchyscenecolorization_b200.cuda_ops import CudaOps
    from sketchyscenecolorization_b200.trainer import FgColorModel
    dev = "cuda:0"
    ops = CudaOps(dev, torch.float32)
    m = FgColorModel(ops, dev, size=SIZE, H=H, W=W, with_discriminator=False)
    m.initialize(seed=0, perturb_tables=0.1)
    b = O.make_batch(1, H, W, 11, torch.float64, n_pad=12)
    b["text"][0, 12:] = torch.tensor([24, 3, 6])            # 'the bus is orange' (tests/golden/text_ids.json), class 2 = bus
    b["cls"][0] = 2
    sk, noise = b["sketch"].float().to(dev), b["noise"].float().to(dev)
    ids_host, ids_dev, cls = b["text"].numpy(), b["text"].int().to(dev), b["cls"].int().to(dev)
    eager_ms, _ = _time_calls(torch, lambda: m.generate(sk, ids_host, cls, noise), args.steps, args.warmup)
    n0 = ops.launch_count()
    out = m.generate(sk, ids_host, cls, noise)
    launches = ops.launch_count() - n0
    replay_ms, replay_best = _time_calls(torch, lambda: m.generate_replay(sk, ids_dev, cls, noise), args.steps, args.warmup + 2)
    sk_host, noise_host = b["sketch"].float().pin_memory(), b["noise"].float().pin_memory()

    def e2e():
        return m.generate_replay(sk_host, ids_dev, cls, noise_host).cpu()
    e2e_ms, _ = _time_calls(torch, e2e, args.steps, args.warmup)
    gp = {k: v.detach().cpu().double() for k, v in m.gstore.state_dict().items()}
    with torch.no_grad():
        t0 = time.perf_counter()
        ref = O.generator_forward(gp, b["sketch"], b["text"], b["cls"], b["noise"], SIZE)
        cpu_s = time.perf_counter() - t0
    err = (m.generate_replay(sk, ids_dev, cls, noise).cpu().double() - ref).abs().max().item()
    err_eager = (out.cpu().double() - ref).abs().max().item()
    infer_gflop = GF_GFLOP + 3 * 0.151 * 2
    print(json.dumps({
        "metric": "fg-colorization inference latency @192x192 bs1", "value": replay_ms, "unit": "ms/image", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": replay_ms, "higher_is_better": False, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 (bf16x3 split tensor-core products)", "data": "synthetic",
        "config": {"workload": "single 192x192 sketch + 'the bus is orange' fg-colorization inference (BASELINE.json configs[0])",
                   "max_abs_err_vs_fp64_oracle": err, "max_abs_err_eager_vs_fp64_oracle": err_eager, "tolerance": 1e-3,
                   "cuda_graph_replay_ms": replay_ms, "cuda_graph_replay_best_ms": replay_best, "eager_ms": eager_ms,
                   "gflop_per_image": round(infer_gflop, 2), "tflops": round(infer_gflop / replay_ms, 2)},
        "e2e": {"value": e2e_ms, "unit": "ms/image", "h2d_bytes_per_step": sk_host.numel() * 4 + noise_host.numel() * 4,
                "d2h_bytes_per_step": 3 * H * W * 4, "api": "FgColorModel.generate_replay (host sketch in, host picture out)"},
        "gpu_launches": launches * args.steps,
        "cpu_baseline": {"value": cpu_s * 1e3, "unit": "ms/image", "cores": os.cpu_count() or 1, "kind": "port",
                         "sample": "oracle generator_forward, fp64, one picture"}}), flush=True)


def run_bg(args):
    """BASELINE.json configs[3]: the background generator, 768x768, batch 1, 8-token caption."""
    import numpy as np
    import torch
    from sketchyscenecolorization_b200.bg import BgColorModel
    from sketchyscenecolorization_b200.cuda_ops import CudaOps
    ids = np.array([[0, 2, 3, 4, 5, 8, 3, 7]], dtype=np.int32)
    res = {}
    for name, dt in (("parity", torch.float32), ("bf16", torch.bfloat16)):
        ops = CudaOps("cuda:0", dt)
        m = BgColorModel(ops, "cuda:0", ngf=64, vocab_size=18)
        m.initialize(seed=0)
        img = torch.rand(1, 768, 768, 3, device="cuda") * 2 - 1
        n0 = ops.launch_count()
        m.generate(img, ids)
        launches = ops.launch_count() - n0
        med, best = _time_calls(torch, lambda: m.generate(img, ids), args.steps, args.warmup)
        res[name] = dict(ms=med, best_ms=best, launches=launches)
        del m, ops
        torch.cuda.empty_cache()
    print(json.dumps({
        "metric": "background-colorization generator inference latency @768x768 bs1", "value": res["parity"]["ms"], "unit": "ms/image",
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["parity"]["ms"], "higher_is_better": False,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 (bf16x3 split tensor-core products)", "data": "synthetic",
        "config": {"workload": "background-colorization generator 768x768 inference, bs 1 (BASELINE.json configs[3])",
                   "parity_mode": res["parity"], "single_pass_bf16": res["bf16"],
                   "parity": "tests/test_bg_gpu.py: max-abs vs the fp64 oracle against the fp32-oracle yardstick (DESIGN.md section 7)"},
        "gpu_launches": res["parity"]["launches"] * args.steps}), flush=True)


def run_rmi(args):
    """BASELINE.json configs[4]: the instance-matching model (ResNet-101 trunk at output stride 8 + word LSTM + multimodal LSTM
    over 96 x 96 positions), 768x768, bs 32, 15-token captions of mixed length.  value = pictures per second with the batch
    resident on the device, in the throughput mode (bf16 storage, single-pass tensor-core products); the parity mode (fp32
    storage, bf16x3 products; tests/test_rmi_gpu.py) is timed beside it at the batch that fits its fp32 intermediates."""
    import numpy as np
    import torch
    from sketchyscenecolorization_b200.cuda_ops import CudaOps
    from sketchyscenecolorization_b200.rmi import RMIModel
    rs = np.random.RandomState(0)
    res = {}
    for name, dt, N in (("bf16", torch.bfloat16, args.rmi_batch), ("parity", torch.float32, max(1, args.rmi_batch // 4))):
        ops = CudaOps("cuda:0", dt)
        m = RMIModel(ops, "cuda:0")
        m.initialize(seed=0)
        im = (torch.rand(N, 768, 768, 3, device="cuda") * 255.0 - 115.0).contiguous()
        im_dev = im if dt == torch.float32 else ops.cast(im, dt)
        words = rs.randint(2, 59, size=(N, 15))
        lengths = rs.randint(4, 16, size=(N,))
        n0 = ops.launch_count()
        up, sg = m.forward(im_dev, words, lengths)
        launches = ops.launch_count() - n0
        assert torch.isfinite(up).all().item() and up.shape == (N, 768, 768, 1)
        med, best = _time_calls(torch, lambda: m.forward(im_dev, words, lengths), args.steps, args.warmup)
        t_med, _ = _time_calls(torch, lambda: m.trunk(im_dev), args.steps, 1)
        res[name] = dict(batch=N, ms=med, best_ms=best, trunk_ms=t_med, launches=launches, images_per_s=N / med * 1e3)
        if name == "bf16":
            im_host = im.cpu().pin_memory()

            def e2e():
                u, s_ = m.forward(ops.cast(im_host.to("cuda", non_blocking=True), dt), words, lengths)
                return (u >= 1e-9).to(torch.uint8).cpu()
            e_med, _ = _time_calls(torch, e2e, max(2, args.steps // 2), 1)
            res["e2e"] = dict(ms=e_med, images_per_s=N / e_med * 1e3, h2d=im_host.numel() * 4, d2h=N * 768 * 768)
        del m, ops, im, im_dev, up, sg
        torch.cuda.empty_cache()
    trunk_gflop, fusion_gflop = 2 * 398.1, 2 * (18.9 + 18.4 + 9.2 * 9.5)     # per picture; mean caption length 9.5 steps
    b = res["bf16"]
    print(json.dumps({
        "metric": "instance-matching (RMI) inference images/sec @768x768 bs%d" % b["batch"], "value": b["images_per_s"], "unit": "images/s",
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": b["ms"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "Instance_Matching RMI_model (DeepLab ResNet-101 trunk + LSTM) referring-segmentation inference, "
                               "768x768, bs %d (BASELINE.json configs[4])" % b["batch"],
                   "single_pass_bf16": b, "parity_mode": res["parity"],
                   "gflop_per_image": round(trunk_gflop + fusion_gflop, 1),
                   "tflops": round((trunk_gflop + fusion_gflop) * b["images_per_s"] / 1e3, 1),
                   "parity": "tests/test_rmi_gpu.py: max-abs of the 768x768 score map vs the fp64 oracle"},
        "e2e": {"value": res["e2e"]["images_per_s"], "unit": "images/s", "h2d_bytes_per_step": res["e2e"]["h2d"],
                "d2h_bytes_per_step": res["e2e"]["d2h"], "api": "RMIModel.forward (pinned host pictures in, host masks out)"},
        "gpu_launches": b["launches"] * args.steps}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--block-type", dest="block_type", default="MRU", choices=["MRU", "Pix2Pix", "Residual"],
                    help="network family (obj_colorization_main.py --block_type); the BASELINE metric is the MRU default")
    ap.add_argument("--input", default="tensors", choices=["tensors", "tfrecord"],
                    help="source of the e2e leg: ready pinned fp32 tensors (default) or synthetic TFRecord files through the "
                         "real input pipeline")
    ap.add_argument("--no-graphs", action="store_true", help="launch every kernel from Python instead of replaying CUDA graphs")
    ap.add_argument("--mode", default="train", choices=["train", "infer", "bg", "rmi"],
                    help="train: the headline (BASELINE.json configs[1]); infer: configs[0] latency + parity; bg: configs[3] latency; "
                         "rmi: configs[4] throughput")
    ap.add_argument("--rmi-batch", dest="rmi_batch", type=int, default=32)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.mode == "infer":
        run_infer(args)
    elif args.mode == "bg":
        run_bg(args)
    elif args.mode == "rmi":
        run_rmi(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
