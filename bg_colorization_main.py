"""Drop-in for Background_Colorization/bg_colorization_main.py, TEST mode (the 768x768 inference of BASELINE.json configs[3]).

Same flags (:979-1005), same directory contract: data/{foreground,background,segment}/test/, data/captions/test.json
(`fg_name`, `bg_name`, `color_text`), data/bg_vocab.txt; the snapshot is the newest TensorFlow V2 bundle under
outputs/<resume_from>/snapshot (`snapshot-<step>`, :816-823, read without TensorFlow by tf_bundle.py); results go to
outputs/<resume_from>/results/<bg stem>_{inputs,outputs,targets}.png, and the foreground is pasted back over the output through
the segment mask (:872-882).  `--mode train` is not built: training this network (its own residual discriminator, the
segmentation loss, Adam with beta1 = 0.5) lies outside SURVEY 8.
"""
import argparse
import json
import os

import numpy as np

from sketchyscenecolorization_b200.text_processing import (bg_vocab_dict, load_vocab_dict_from_file,  # noqa: F401
                                                           preprocess_sentence)


def load_image(imname, image_size):
    """data_processing/image_processing.py:5-11 -> uint8 [1,H,W,3] RGB."""
    from PIL import Image
    im = Image.open(imname).convert("RGB")
    if im.width != image_size or im.height != image_size:
        im = im.resize((image_size, image_size), resample=Image.BILINEAR)
    return np.expand_dims(np.array(im, dtype=np.uint8), axis=0)


def _build_model(ngf, vocab_size, seg_classes, snapshot_dir):
    import torch
    from sketchyscenecolorization_b200 import checkpoint, tf_bundle
    from sketchyscenecolorization_b200.bg import BgColorModel
    from sketchyscenecolorization_b200.cuda_ops import CudaOps
    dev = "cuda:%d" % int(os.environ.get("LOCAL_RANK", "0"))
    model = BgColorModel(CudaOps(dev, torch.float32), dev, ngf=ngf, vocab_size=vocab_size, seg_classes=seg_classes)
    prefix = checkpoint.latest_checkpoint(snapshot_dir)
    print("loading model from checkpoint", prefix)
    if prefix is None:
        raise RuntimeError("no snapshot in %s" % snapshot_dir)
    model.gstore.load_state_dict(tf_bundle.read_bundle(prefix), strict=True)
    return model


def bg_colorization(**kwargs):
    """bg_colorization (:703-882), test branch.  kwargs as the reference's; optional `model` (a resident BgColorModel)."""
    from PIL import Image
    mode = kwargs['mode']
    resume_from = kwargs.get('resume_from', '')
    data_base_dir = kwargs.get('data_base_dir', 'data')
    image_size = kwargs.get('image_size', 768)
    T, vocab_size = kwargs.get('text_len', 8), kwargs.get('vocab_size', 18)
    vocab_file = kwargs.get('vocab_file', 'data/bg_vocab.txt')
    if mode != "test":
        raise NotImplementedError("bg_colorization: only --mode test is built (training the background model is outside the "
                                  "scope of this package, DESIGN.md)")
    if resume_from == '':
        raise Exception("checkpoint required for test mode")
    output_dir = os.path.join("outputs", resume_from)
    inputs_base_dir = os.path.join(data_base_dir, 'foreground', mode)
    targets_base_dir = os.path.join(data_base_dir, 'background', mode)
    segment_base_dir = os.path.join(data_base_dir, 'segment', mode)
    with open(os.path.join(data_base_dir, 'captions', mode + '.json')) as fp:
        json_data = json.load(fp)
    nImgs = len(json_data)
    print('## nImgs =', nImgs, '\n')
    vocab_dict = load_vocab_dict_from_file(vocab_file) if os.path.exists(vocab_file) else bg_vocab_dict()
    model = kwargs.get('model') or _build_model(kwargs.get('ngf', 64), vocab_size, kwargs.get('seg_classes', 3),
                                                os.path.join(output_dir, "snapshot"))
    print("parameter_count =", model.gstore.num_params())
    image_dir = os.path.join(output_dir, "results")
    os.makedirs(image_dir, exist_ok=True)
    for image_idx in range(nImgs):
        input_name, target_name = json_data[image_idx]['fg_name'], json_data[image_idx]['bg_name']
        print('Processing', image_idx, '/', nImgs)
        input_data = load_image(os.path.join(inputs_base_dir, input_name), image_size)
        target_data = load_image(os.path.join(targets_base_dir, target_name), image_size)
        vocab_indices = np.array(preprocess_sentence(json_data[image_idx]['color_text'], vocab_dict, T), dtype=np.int32)[None]
        output, _region = model.colorize_u8(input_data, vocab_indices)
        stem = target_name[:-4]
        Image.fromarray(input_data[0], 'RGB').save(os.path.join(image_dir, stem + "_inputs.png"), 'PNG')
        Image.fromarray(target_data[0], 'RGB').save(os.path.join(image_dir, stem + "_targets.png"), 'PNG')
        # post-processing: cover the FG to generation (:872-882); in the segment picture 0 is foreground
        inner_mask = np.array(Image.open(os.path.join(segment_base_dir, input_name)).convert('RGB'), dtype=np.uint8)[:, :, 0]
        if inner_mask.shape != output.shape[:2]:
            raise ValueError("segment mask %s is %s, the picture %s" % (input_name, inner_mask.shape, output.shape[:2]))
        output[inner_mask == 0] = input_data[0][inner_mask == 0]
        Image.fromarray(output, 'RGB').save(os.path.join(image_dir, stem + "_outputs.png"), 'PNG')
    return nImgs


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("--mode", type=str, default='train', choices=["train", "test"])
    parser.add_argument("--resume_from", type=str, default='', help="where to put output files")
    parser.add_argument("--data_base_dir", type=str, default='data', help="where to put data")
    parser.add_argument("--image_size", type=int, default=768, help="image size")
    parser.add_argument("--batch_size", type=int, default=1, help="number of images in batch")
    parser.add_argument("--max_steps", type=int, default=100000)
    parser.add_argument("--lr", type=float, default=0.0002)
    parser.add_argument("--l1_weight", type=float, default=100.0)
    parser.add_argument("--gan_weight", type=float, default=1.0)
    parser.add_argument("--seg_weight", type=float, default=100.0)
    parser.add_argument("--seg_classes", type=int, default=3, help="number of categories of seg")
    parser.add_argument("--ngf", type=int, default=64, help="number of generator filters in first conv layer")
    parser.add_argument("--ndf", type=int, default=64)
    parser.add_argument("--text_len", type=int, default=8, help="the longest length of text")
    parser.add_argument("--vocab_size", type=int, default=18, help="vocab size")
    parser.add_argument("--vocab_file", type=str, default='data/bg_vocab.txt', help="path of vocab")
    parser.add_argument("--summary_freq", type=int, default=200)
    parser.add_argument("--progress_freq", type=int, default=50)
    parser.add_argument("--save_freq", type=int, default=20000)
    a = parser.parse_args()
    if a.mode != "test":        # the reference's default is 'train'; say what is there instead of a traceback
        parser.error("--mode train is not built here: training the background model is outside the scope of this package "
                     "(DESIGN.md section 7).  Use --mode test --resume_from <run> for the 768x768 inference path.")
    bg_colorization(**vars(a))
