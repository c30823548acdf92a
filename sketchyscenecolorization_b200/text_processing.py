"""Caption -> fixed-length vocabulary ids, the host-side text front end of the fg-colorization path.

Same public names and results as the reference's data_processing/text_processing.py
(`load_vocab_dict_from_file` :36-40, `sentence2vocab_indices` :10-31, `preprocess_sentence` :43-53):
split on runs of non-word characters, lower-case, drop a final '.', drop a leading 'a', drop every 'the',
map ',' to 'and', unknown words -> <unk>, keep the first T ids, left-pad with <pad>.  T is 15 on this path
(main_procedure.py:503,538).  The result is pinned against ids produced by the reference's own module in
tests/golden/text_ids.json.
"""
from __future__ import annotations

import re

UNK_IDENTIFIER = "<unk>"
PAD_IDENTIFIER = "<pad>"
T_DEFAULT = 15

# the reference's 58-entry vocabulary (Foreground_Instance_Colorization/data/vocab.txt), index = line number
DEFAULT_VOCAB = (
    "<pad> <unk> bench is light gray orange red purple brown dark green black cyan pink blue yellow bird has body and "
    "wing with white bus windows butterfly edge car cat chair chicken tail head cloud cow dog duck horse house roof moon "
    "person hair in shirt pants skirt pig rabbit road sheep star sun tree truck carriage grass").split()

# the background model's 18-entry vocabulary (Background_Colorization/data/bg_vocab.txt)
BG_VOCAB = "<pad> <unk> sky is blue and grass green ground gray purple black yellow brown cyan pink orange red".split()

_TOKEN_BREAK = re.compile(r"(\W+)")


def default_vocab_dict():
    return {w: i for i, w in enumerate(DEFAULT_VOCAB)}


def bg_vocab_dict():
    return {w: i for i, w in enumerate(BG_VOCAB)}


def load_vocab_dict_from_file(dict_file):
    with open(dict_file) as f:
        return {line.strip(): i for i, line in enumerate(f.readlines())}


def sentence2vocab_indices(sentence, vocab_dict):
    tokens = [t.lower() for t in _TOKEN_BREAK.split(sentence.strip()) if t.strip()]
    if tokens[-1] == ".":
        tokens.pop()
    if tokens[0] == "a":
        tokens.pop(0)
    unk = vocab_dict[UNK_IDENTIFIER]
    ids = []
    for t in tokens:
        if t == "the":
            continue
        if t in (",", ", "):      # separators keep their blanks: only these two spell 'and' (reference :23-26)
            t = "and"
        ids.append(vocab_dict.get(t, unk))
    return ids


def preprocess_sentence(sentence, vocab_dict, T=T_DEFAULT):
    ids = sentence2vocab_indices(sentence, vocab_dict)[:T]
    return [vocab_dict[PAD_IDENTIFIER]] * (T - len(ids)) + ids
