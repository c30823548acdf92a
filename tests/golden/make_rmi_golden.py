"""Generates tests/golden/rmi_helpers.json by importing the REFERENCE's own Instance_Matching helpers
(/root/reference/Instance_Matching/utils/processing_tools.py, data_processing/text_processing.py) and its vocab.txt.
Run in the build container only (the reference tree does not exist on the GPU box); the JSON is committed."""
import json
import os
import sys

REF = "/root/reference/Instance_Matching"
sys.path.insert(0, os.path.join(REF, "utils"))
sys.path.insert(0, os.path.join(REF, "data_processing"))
import processing_tools as pt  # noqa: E402
import text_processing as tp  # noqa: E402

SENTENCES = [
    "the dog on the right",
    "the person in front of the house",
    "all the trees on the left.",
    "The two clouds in the middle of the sky",
    "the left - most bird",
    "zebra on the road",
    "the sun",
    "the second tree on the right of the house near the road and the person in the middle front of it and more",
]

vocab = tp.load_vocab_dict_from_file(os.path.join(REF, "data", "vocab.txt"))
out = {"T": 15, "vocab": [w for w, _ in sorted(vocab.items(), key=lambda kv: kv[1])], "sentences": [], "spatial": []}
for s in SENTENCES:
    ids, n = tp.preprocess_sentence(s, vocab, 15)
    out["sentences"].append({"sentence": s, "ids": [int(i) for i in ids], "len": int(n)})
for (N, fh, fw) in ((1, 3, 4), (2, 5, 5), (1, 8, 6)):
    v = pt.generate_spatial_batch(N, fh, fw)
    out["spatial"].append({"N": N, "h": fh, "w": fw, "values": [float(x) for x in v.reshape(-1)]})
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "rmi_helpers.json"), "w") as f:
    json.dump(out, f)
print("wrote", len(out["sentences"]), "sentences,", len(out["spatial"]), "spatial grids")
