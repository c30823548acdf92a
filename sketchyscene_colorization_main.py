"""Whole-pipeline entry point: one instruction colours either an object instance (FG) or the background (BG) of scene `image_id`
and is appended to the scene's editing records.  Flags, defaults, directory contract and the signature of `colorization_main`
are those of the reference's sketchyscene_colorization_main.py (:19-60, :63-112), so existing invocations keep working.

FG instructions need the indices of the instances they refer to.  As in the reference they come from the instance-matching model
(`pipeline_match.build_instance_matching`: rmi.RMIModel restored from --match_snapshot_root + the occupied-share rule over the
segmentation file); --matched_inst_indices, `matched_inst_indices=` or a `matcher=` callable override it.
"""
import argparse
import os

from sketchyscenecolorization_b200 import customization_util as records
from sketchyscenecolorization_b200.pipeline_bg import build_background_colorization
from sketchyscenecolorization_b200.pipeline_fg import build_instance_colorization

# (long flag, short flag, type, default) -- the reference's fixed parameters, :76-106
_FLAGS = [
    ("data_base_dir", "dbd", str, "examples"),
    ("results_base_dir", "rbd", str, "outputs"),
    ("match_snapshot_root", "msr", str, "Instance_Matching/outputs/snapshots"),
    ("match_vocab_path", "mvp", str, "Instance_Matching/data/vocab.txt"),
    ("match_vocab_size", "mvs", int, 76),
    ("match_max_len", "ml", int, 15),
    ("fgcolor_snapshot_root", "fgsr", str, "Foreground_Instance_Colorization/outputs/2019-00-00-00-00-00/snapshot"),
    ("fgcolor_vocab_path", "fgvp", str, "Foreground_Instance_Colorization/data/vocab.txt"),
    ("fgcolor_vocab_size", "fgvs", int, 58),
    ("fgcolor_max_len", "fgl", int, 15),
    ("bg_snapshot_root", "bgsr", str, "Background_Colorization/outputs/2019-00-00-00-00-00/snapshot"),
    ("bg_vocab_path", "bgvp", str, "Background_Colorization/data/bg_vocab.txt"),
    ("bg_vocab_size", "bgvs", int, 18),
    ("bg_max_len", "bgl", int, 8),
]


def withdraw_last_record(image_id, results_base_dir):
    records.withdraw_records(image_id, results_base_dir)


def _scene_files(data_base_dir, image_id):
    """(sketch png, Mask-RCNN segmentation npz, inner-mask mat) of a scene under the examples directory."""
    at = lambda sub, name: os.path.join(data_base_dir, sub, name)         # noqa: E731
    return at('sketches', '%s.png' % image_id), at('seg_data', '%s_datas.npz' % image_id), at('inner_masks', '%s.mat' % image_id)


def colorization_main(image_id, input_text, data_base_dir, results_base_dir,
                      match_vocab_path, match_vocab_size, match_snapshot_root, match_max_len,
                      fgcolor_vocab_path, fgcolor_vocab_size, fgcolor_snapshot_root, fgcolor_max_len,
                      bg_vocab_path, bg_vocab_size, bg_snapshot_root, bg_max_len, *,
                      matched_inst_indices=None, matcher=None, fg_model=None, bg_model=None, ops=None):
    """-> (colorization type, name of the result picture)."""
    kind = records.judge_colorize_type(input_text)
    print('colorization_type:', kind)
    sketch_path, segm_npz, inner_mat = _scene_files(data_base_dir, image_id)
    new_name, last_name, last_bg_text, history = records.fetch_records(image_id, results_base_dir)
    if kind == 'BG':
        bg_text = build_background_colorization(image_id, input_text, sketch_path, inner_mat, segm_npz, results_base_dir,
                                                bg_vocab_size, bg_max_len, bg_vocab_path, bg_snapshot_root, new_name, last_name,
                                                last_bg_text, model=bg_model, ops=ops)
    else:
        assert input_text
        if matched_inst_indices is None and matcher is not None:
            matched_inst_indices = matcher(data_base_dir, sketch_path, input_text, segm_npz, match_vocab_path, match_vocab_size,
                                           match_snapshot_root, match_max_len)
        if matched_inst_indices is None:          # reference: Pipeline_utils.fg_matching_utils.build_instance_matching (:33-37)
            from sketchyscenecolorization_b200.pipeline_match import build_instance_matching
            matched_inst_indices = build_instance_matching(data_base_dir, sketch_path, input_text, segm_npz, match_vocab_path,
                                                           match_vocab_size, match_snapshot_root, match_max_len, ops=ops)
        assert type(matched_inst_indices) is list
        print('matched_inst_indices', matched_inst_indices)
        build_instance_colorization(data_base_dir, image_id, input_text, matched_inst_indices, sketch_path, inner_mat, segm_npz,
                                    results_base_dir, fgcolor_vocab_size, fgcolor_max_len, fgcolor_vocab_path,
                                    fgcolor_snapshot_root, new_name, last_name, model=fg_model, ops=ops)
        bg_text = last_bg_text                       # an FG edit leaves the background caption as it was
    records.update_records(image_id, input_text, results_base_dir, kind, new_name, bg_text, history)
    return kind, new_name


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument('--command', '-c', type=str, choices=['color', 'withdraw'], default='color')
    p.add_argument('--image_id', '-id', type=int, default=-1)
    p.add_argument('--instruction', '-it', type=str, default='')
    for name, short, typ, default in _FLAGS:
        p.add_argument('--' + name, '-' + short, type=typ, default=default)
    p.add_argument('--matched_inst_indices', type=str, default='',
                   help="comma-separated instance indices for an FG instruction (stands in for the matching model)")
    return p


def main(argv=None):
    a = build_parser().parse_args(argv)
    if a.image_id == -1:
        raise SystemExit("--image_id is required")
    if a.command == 'withdraw':
        return withdraw_last_record(a.image_id, a.results_base_dir)
    if not a.instruction:
        raise SystemExit("--instruction is required for --command color")
    picked = [int(t) for t in a.matched_inst_indices.split(',') if t.strip()] or None
    return colorization_main(a.image_id, a.instruction, a.data_base_dir, a.results_base_dir,
                             a.match_vocab_path, a.match_vocab_size, a.match_snapshot_root, a.match_max_len,
                             a.fgcolor_vocab_path, a.fgcolor_vocab_size, a.fgcolor_snapshot_root, a.fgcolor_max_len,
                             a.bg_vocab_path, a.bg_vocab_size, a.bg_snapshot_root, a.bg_max_len, matched_inst_indices=picked)


if __name__ == '__main__':
    main()
