"""Drop-in for Foreground_Instance_Colorization/data_preparation/data_preparation.py: images + captions -> the per-category
TFRecord files the training / validation queues read, written without TensorFlow (tfrecord_input.encode_example /
write_tfrecord; framing and encoding are pinned to tensorboard's writer and the protobuf runtime in tests/test_tfrecord_cpu.py).

Same flags (:99-116) and directory contract (:37-96):
    <data>/captions/<category>/{train,val}.json   list of {key, color_text}
    <data>/images/<category>/{cartoon,edgemap}/<key>
    <data>/vocab.txt
 -> <data>/tfrecord/{train,val}/<category>.tfrecord, one Example per picture with the features of :21-32
    (ImageName, cartoon_data, sketch_data, Category, Category_id, Color_text, Text_vocab_indices).
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from sketchyscenecolorization_b200.text_processing import (default_vocab_dict, load_vocab_dict_from_file,  # noqa: E402
                                                           preprocess_sentence)
from sketchyscenecolorization_b200.tfrecord_input import encode_example, write_tfrecord  # noqa: E402


def _raw_rgb(path):
    from PIL import Image
    return np.array(Image.open(path).convert("RGB"), dtype=np.uint8).tobytes()


def data_preparation(**kwargs):
    dataset, data_base_dir, text_len = kwargs['dataset'], kwargs['data_base_dir'], kwargs['text_len']
    dataset_types = ['train', 'val'] if dataset == 'both' else [dataset]
    caption_data_base_dir = os.path.join(data_base_dir, 'captions')
    image_data_base_dir = os.path.join(data_base_dir, 'images')
    categories = sorted(os.listdir(caption_data_base_dir))
    vocab_file = os.path.join(data_base_dir, 'vocab.txt')
    vocab_dict = load_vocab_dict_from_file(vocab_file) if os.path.exists(vocab_file) else default_vocab_dict()
    written = {}
    for dataset_type in dataset_types:
        split_dir = os.path.join(data_base_dir, 'tfrecord', dataset_type)
        os.makedirs(split_dir, exist_ok=True)
        for category_id, category_name in enumerate(categories):
            with open(os.path.join(caption_data_base_dir, category_name, dataset_type + '.json')) as fp:
                json_data = json.load(fp)
            print(dataset_type, category_name, len(json_data))

            def examples():
                for entry in json_data:
                    image_name, color_text = entry['key'], entry['color_text']
                    ids = np.array(preprocess_sentence(color_text, vocab_dict, text_len), dtype=np.uint8).tobytes()
                    yield encode_example(dict(
                        ImageName=image_name.encode(),
                        cartoon_data=_raw_rgb(os.path.join(image_data_base_dir, category_name, 'cartoon', image_name)),
                        sketch_data=_raw_rgb(os.path.join(image_data_base_dir, category_name, 'edgemap', image_name)),
                        Category=category_name.encode(), Category_id=category_id, Color_text=color_text.encode(),
                        Text_vocab_indices=ids))
            write_tfrecord(os.path.join(split_dir, category_name + '.tfrecord'), examples())
            written[(dataset_type, category_name)] = len(json_data)
    return written


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument('--dataset', '-ds', type=str, choices=['train', 'val', 'both'], default='both', help="choose a dataset")
    parser.add_argument('--data_base_dir', '-db', type=str, default='../data', help="set the data base dir")
    parser.add_argument('--text_len', '-tl', type=int, default=15, help="the longest length of text")
    args = parser.parse_args()
    data_preparation(dataset=args.dataset, data_base_dir=args.data_base_dir, text_len=args.text_len)
