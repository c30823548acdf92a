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
    runtime = {k: kwargs.pop(k) for k in ('model', 'input_iter', 'input_iter_d', 'synthetic_input', 'use_cuda_graphs', 'data_base_dir')
               if k in kwargs}
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


# (flag, short flag, type, default, choices) -- obj_colorization_main.py:160-206 of the reference
_FLAGS = [
    ('mode', 'md', str, 'train', ['train', 'val', 'test', 'inference']),
    ('resume_from', 'rf', str, '', None),
    ('batch_size', 'bs', int, 2, None),                      # per GPU
    ('max_iter', 'mi', int, 100000, None),
    ('optimizer', 'opt', str, 'Adam', ["RMSprop", "Adam", "AdaDelta", "AdaGrad"]),
    ('lr_G', 'lrg', float, 2e-4, None),
    ('lr_D', 'lrd', float, 1e-4, None),
    ('small_img', 'si', int, 0, [0, 1]),
    ('lstm_hybrid', 'lh', int, 1, [0, 1]),
    ('distance_map', 'dm', int, 0, [0, 1]),
    ('block_type', 'bt', str, 'MRU', ['MRU', 'Pix2Pix', 'Residual']),
    ('vocab_size', 'vs', int, 58, None),
    ('disc_iterations', 'di', int, 1, None),
    ('ld', 'ld', int, 10, None),
    ('num_gpu', 'gpu', int, 1, None),                        # N > 1: re-launched under torch.distributed.run, one process per GPU
    ('extra_info', 'ei', str, '', None),
    ('summary_write_freq', 'swf', int, 100, None),
    ('save_model_freq', 'smf', int, 10000, None),
    ('count_left_time_freq', 'clt', int, 100, None),
    ('count_inception_score_freq', 'cis', int, -1, None),
    ('infer_name', 'in', str, '', None),
    ('instruction', 'ins', str, '', None),
    # B200 additions (not in the reference)
    ('cuda_graphs', 'cg', int, 1, [0, 1]),                   # training steps replay CUDA graphs (0: every kernel launched from Python)
    ('synthetic_input', 'syn', int, 0, [0, 1]),              # train on seeded synthetic batches when data/tfrecord/train is absent
]
# Config key <- flag, where the two names differ (:208-232)
_RENAMED = {'dataset_type': 'mode', 'max_iter_step': 'max_iter', 'LSTM_hybrid': 'lstm_hybrid'}


def build_parser():
    p = argparse.ArgumentParser()
    for name, short, typ, default, choices in _FLAGS:
        p.add_argument('--' + name, '-' + short, type=typ, default=default, choices=choices)
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
    d_params = {name: getattr(args, name) for name, *_ in _FLAGS if name not in _RENAMED.values()}
    d_params.update({key: getattr(args, flag) for key, flag in _RENAMED.items()})
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
