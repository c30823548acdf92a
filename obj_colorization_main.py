#!/usr/bin/env python
"""Command line of the foreground-instance-colorization path on B200 -- same flags, defaults, run-directory
layout (outputs/<UTC ts>/{log/param_<iter>.json, snapshot/, *_results/}), resume and NaN-restart behaviour as the
reference's Foreground_Instance_Colorization/obj_colorization_main.py (:17-246).

    python obj_colorization_main.py --mode train --batch_size 64 --max_iter 100000
    python obj_colorization_main.py --mode inference --resume_from <ts> --infer_name car.png --instruction 'the car is red'
Under `torchrun --nproc-per-node N` training is data parallel (one process per GPU, one NCCL all-reduce of the
flat gradient bucket per optimiser step)."""
import argparse
import json
import os
from time import gmtime, strftime

from sketchyscenecolorization_b200 import checkpoint, main_procedure
from sketchyscenecolorization_b200.config import Config

OUTPUTS = 'outputs'


def _valid(appendix):
    return bool(appendix) and len(appendix.split('-')) == 6


def _process_group():
    """(process group, world size) under torchrun, else (None, 1)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return None, 1
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        backend = os.environ.get("FGC_DIST_BACKEND", "nccl")
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group(backend)
    return dist.group.WORLD, world


def launch_training(**kwargs):
    # objects a caller may hand through (a resident model, input iterators): not parameters, not written to param_<iter>.json
    runtime = {k: kwargs.pop(k) for k in ('model', 'input_iter', 'input_iter_d') if k in kwargs}
    appendix = kwargs["resume_from"]
    pg, world = _process_group()
    if appendix is None or appendix == '':
        # one time stamp for all ranks: each would otherwise read its own clock and may land in a different second / directory
        cur_time = main_procedure.shared_string(strftime("%Y-%m-%d-%H-%M-%S", gmtime()), pg, world)
        log_dir, ckpt_dir = os.path.join(OUTPUTS, cur_time, 'log'), os.path.join(OUTPUTS, cur_time, 'snapshot')
        os.makedirs(log_dir, exist_ok=True)
        os.makedirs(ckpt_dir, exist_ok=True)
        kwargs.update(log_dir=log_dir, ckpt_dir=ckpt_dir, resume_from=appendix, iter_from=0)
        appendix = cur_time
        with open(os.path.join(log_dir, 'param_0.json'), 'w') as fp:
            json.dump(kwargs, fp, indent=4)
        print("Launching new train: %s" % cur_time)
    else:
        if not _valid(appendix):
            print("Invalid resume folder")
            return
        log_dir, ckpt_dir = os.path.join(OUTPUTS, appendix, 'log'), os.path.join(OUTPUTS, appendix, 'snapshot')
        ckpt_file = checkpoint.latest_checkpoint(ckpt_dir)
        if ckpt_file is None:
            raise RuntimeError("no snapshot to resume from in %s" % ckpt_dir)
        iter_from = int(os.path.split(ckpt_file)[1].split('-')[1]) + 1
        kwargs.update(log_dir=log_dir, ckpt_dir=ckpt_dir, iter_from=iter_from)
        with open(os.path.join(log_dir, 'param_%d.json' % iter_from), 'w') as fp:
            json.dump(kwargs, fp, indent=4)
        print("Launching training from checkpoint: %s" % appendix)
    Config.set_from_dict(kwargs)
    extra = dict(process_group=pg, world_size=world) if world > 1 else {}
    status = main_procedure.train(**kwargs, **extra, **runtime)
    return status, appendix


def _launch_eval(kind, **kwargs):
    appendix = kwargs["resume_from"]
    if not _valid(appendix):
        print("Invalid resume folder")
        return False
    kwargs['log_dir'] = os.path.join(OUTPUTS, appendix, 'log')
    kwargs['ckpt_dir'] = os.path.join(OUTPUTS, appendix, 'snapshot')
    kwargs['results_dir'] = os.path.join(OUTPUTS, appendix, kind + '_results')
    Config.set_from_dict(kwargs)
    print("Launching %s from checkpoint: %s" % (kind, appendix))
    return True


def launch_val(**kwargs):
    if _launch_eval('validation', **kwargs):
        main_procedure.validation(**kwargs)


def launch_test(**kwargs):
    if _launch_eval('test', **kwargs):
        main_procedure.test()


def launch_inference(**kwargs):
    if _launch_eval('inference', **kwargs):
        main_procedure.inference(kwargs["infer_name"], kwargs["instruction"])


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument('--mode', '-md', type=str, choices=['train', 'val', 'test', 'inference'], default='train')
    p.add_argument('--resume_from', '-rf', type=str, default='')
    p.add_argument('--batch_size', '-bs', type=int, default=2, help="Batch size per gpu")
    p.add_argument('--max_iter', '-mi', type=int, default=100000)
    p.add_argument('--optimizer', '-opt', type=str, choices=["RMSprop", "Adam", "AdaDelta", "AdaGrad"], default='Adam')
    p.add_argument('--lr_G', '-lrg', type=float, default=2e-4)
    p.add_argument('--lr_D', '-lrd', type=float, default=1e-4)
    p.add_argument('--small_img', '-si', type=int, choices=[0, 1], default=0)
    p.add_argument('--lstm_hybrid', '-lh', type=int, choices=[0, 1], default=1)
    p.add_argument('--distance_map', '-dm', type=int, choices=[0, 1], default=0)
    p.add_argument('--block_type', '-bt', type=str, choices=['MRU', 'Pix2Pix', 'Residual'], default='MRU')
    p.add_argument('--vocab_size', '-vs', type=int, default=58)
    p.add_argument('--disc_iterations', '-di', type=int, default=1)
    p.add_argument('--ld', '-ld', type=int, default=10)
    p.add_argument('--num_gpu', '-gpu', type=int, default=1,
                   help="GPUs to train on; N > 1 re-launches this command under torch.distributed.run, one process per GPU")
    p.add_argument('--extra_info', '-ei', type=str, default='')
    p.add_argument('--summary_write_freq', '-swf', type=int, default=100)
    p.add_argument('--save_model_freq', '-smf', type=int, default=10000)
    p.add_argument('--count_left_time_freq', '-clt', type=int, default=100)
    p.add_argument('--count_inception_score_freq', '-cis', type=int, default=-1)
    p.add_argument('--infer_name', '-in', type=str, default='')
    p.add_argument('--instruction', '-ins', type=str, default='')
    return p


def _relaunch_data_parallel(n, argv):
    """`--num_gpu N` in the reference builds N in-graph towers (graph_single.py:107-218); here it means N processes, one per
    GPU: replace this process by `python -m torch.distributed.run --nproc-per-node N <this command>`."""
    import socket
    import sys
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.abspath(__file__)] + list(argv)
    print("Launching %d data-parallel processes: %s" % (n, " ".join(cmd)))
    os.execv(sys.executable, cmd)


def main(argv=None):
    import sys
    argv = list(sys.argv[1:] if argv is None else argv)
    args = build_parser().parse_args(argv)
    if args.mode == 'train' and args.num_gpu > 1 and "WORLD_SIZE" not in os.environ:
        _relaunch_data_parallel(args.num_gpu, argv)
    if args.mode == 'inference':
        assert args.infer_name != '' and args.instruction != ''
    d_params = {
        "dataset_type": args.mode, "resume_from": args.resume_from, "batch_size": args.batch_size,
        "max_iter_step": args.max_iter, "disc_iterations": args.disc_iterations, "optimizer": args.optimizer,
        "lr_G": args.lr_G, "lr_D": args.lr_D, "num_gpu": args.num_gpu, "small_img": args.small_img,
        "distance_map": args.distance_map, "LSTM_hybrid": args.lstm_hybrid, "block_type": args.block_type,
        "vocab_size": args.vocab_size, "ld": args.ld, "extra_info": args.extra_info,
        "summary_write_freq": args.summary_write_freq, "save_model_freq": args.save_model_freq,
        "count_left_time_freq": args.count_left_time_freq, "count_inception_score_freq": args.count_inception_score_freq,
        "infer_name": args.infer_name, "instruction": args.instruction,
    }
    if args.mode == 'train':
        status, appendix = launch_training(**d_params)
        while status == -1:                     # NaN during training: restart from the latest snapshot
            print("Training ended with status -1. Restarting..")
            d_params["resume_from"] = appendix
            status, appendix = launch_training(**d_params)
    elif args.mode == 'val':
        launch_val(**d_params)
    elif args.mode == 'test':
        launch_test(**d_params)
    elif args.mode == 'inference':
        launch_inference(**d_params)


if __name__ == "__main__":
    main()
