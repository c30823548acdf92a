"""Generates tests/golden/text_ids.json by importing the REFERENCE's own text_processing.py
(/root/reference/Foreground_Instance_Colorization/data_processing/text_processing.py) and its vocab.txt.
Run in the build container only (the reference tree does not exist on the GPU box); the JSON is committed."""
import json
import os
import sys

REF = "/root/reference/Foreground_Instance_Colorization"
sys.path.insert(0, os.path.join(REF, "data_processing"))
import text_processing as ref  # noqa: E402

SENTENCES = [
    "the bus is orange",
    "the bus is orange with gray windows",
    "the car is yellow with blue window",
    "the person has black hair, in red shirt and blue pants.",
    "a dog is brown",
    "The Cat Is White .",
    "the tree is dark green",
    "the house is red with a blue roof and the windows is cyan",
    "the chicken has red head , yellow body and brown tail and the wing is light gray with pink edge and purple",
    "sun",
    "the butterfly has purple wing with black edge.",
    "the truck is green with dark gray carriage",
    "zebra unicorn",
    "the  road   is gray",
]

vocab = ref.load_vocab_dict_from_file(os.path.join(REF, "data", "vocab.txt"))
out = {"T": 15, "vocab": [w for w, _ in sorted(vocab.items(), key=lambda kv: kv[1])],
       "cases": [{"sentence": s, "ids": [int(i) for i in ref.preprocess_sentence(s, vocab, 15)]} for s in SENTENCES]}
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "text_ids.json"), "w") as f:
    json.dump(out, f, indent=1)
print("wrote", len(out["cases"]), "cases")
