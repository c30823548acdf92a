"""Generates tests/golden/pipeline_fg.json from the REFERENCE's own helper functions (run in the build container only:
it reads /root/reference).  Pipeline_utils/fg_color_utils.py imports tensorflow at module level, so the pure helpers are
lifted out of its source with `ast` and executed against the importable Instance_Matching text module."""
import ast
import importlib.util
import json
import os
import sys

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from pipeline_cases import SENTENCES, road_sketch          # noqa: E402  (shared with the test)

spec = importlib.util.spec_from_file_location("match_text_processing",
                                              os.path.join(REF, "Instance_Matching/data_processing/text_processing.py"))
match_text_processing = importlib.util.module_from_spec(spec)
spec.loader.exec_module(match_text_processing)

src = open(os.path.join(REF, "Pipeline_utils/fg_color_utils.py")).read()
tree = ast.parse(src)
wanted = {"judging_preposition", "segment_user_input_text", "is_road_not_single_line"}
mod = ast.Module(body=[n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in wanted], type_ignores=[])
import re  # noqa: E402
ns = {"re": re, "np": np, "match_text_processing": match_text_processing}
exec(compile(mod, "fg_color_utils_helpers", "exec"), ns)

out = {"segment_user_input_text": {s: ns["segment_user_input_text"](s) for s in SENTENCES},
       "is_road_not_single_line": {name: bool(ns["is_road_not_single_line"](road_sketch(name))) for name in
                                   ("two_edges_vertical", "two_edges_horizontal", "single_line", "blank", "grey_edges", "diagonal_pair")}}
json.dump(out, open(os.path.join(HERE, "pipeline_fg.json"), "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1)[:1500])
