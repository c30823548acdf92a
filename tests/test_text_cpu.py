"""Host-side caption processing against golden ids produced by the reference's own text_processing.py
(tests/golden/make_text_golden.py)."""
import json
import os

from sketchyscenecolorization_b200 import text_processing as tp

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "text_ids.json")))


def test_vocab_matches_reference_file():
    assert list(tp.DEFAULT_VOCAB) == GOLD["vocab"]
    assert len(tp.DEFAULT_VOCAB) == 58 and tp.default_vocab_dict()["<pad>"] == 0 and tp.default_vocab_dict()["<unk>"] == 1


def test_preprocess_sentence_golden():
    vd = tp.default_vocab_dict()
    for case in GOLD["cases"]:
        assert tp.preprocess_sentence(case["sentence"], vd, GOLD["T"]) == case["ids"], case["sentence"]


def test_survey_known_answers():
    vd = tp.default_vocab_dict()
    assert tp.preprocess_sentence("the bus is orange", vd, 15) == [0] * 12 + [24, 3, 6]
    assert tp.preprocess_sentence("the car is yellow with blue window", vd, 15) == [0] * 9 + [28, 3, 16, 22, 15, 1]


def test_load_vocab_file(tmp_path):
    p = tmp_path / "vocab.txt"
    p.write_text("\n".join(tp.DEFAULT_VOCAB) + "\n")
    assert tp.load_vocab_dict_from_file(str(p)) == tp.default_vocab_dict()
