"""Generates tests/golden/pipeline_bg.json from the REFERENCE's own helper functions (run in the build container only: it reads
/root/reference).  Pipeline_utils/bg_utils.py imports tensorflow and skimage at module level, so its pure helpers are lifted out of
the source with `ast`; `skimage.color` is replaced by the standard library's colorsys (the same HSV definition, float64).
customization_util.judge_colorize_type is imported as it is (its only dependency is the importable matching text module)."""
import ast
import colorsys
import json
import os
import re
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, REF)
from pipeline_bg_cases import COMBINE_CASES, TYPE_SENTENCES, gradient_case      # noqa: E402


def _map(fn, arr):
    arr = np.asarray(arr, dtype=np.float64)
    flat = arr.reshape(-1, 3)
    return np.array([fn(*px) for px in flat], dtype=np.float64).reshape(arr.shape)


color = types.SimpleNamespace(rgb2hsv=lambda a: _map(colorsys.rgb_to_hsv, a), hsv2rgb=lambda a: _map(colorsys.hsv_to_rgb, a))
skimage = types.SimpleNamespace(color=color)

src = open(os.path.join(REF, "Pipeline_utils/bg_utils.py")).read()
tree = ast.parse(src)
wanted = {"get_text_type", "check_duplicated_color", "combine_bg_input_text", "add_color_gradient"}
body = [n for n in tree.body if (isinstance(n, ast.FunctionDef) and n.name in wanted) or
        (isinstance(n, ast.Assign) and getattr(n.targets[0], "id", "") in ("input_text_types", "ALL_COLOR"))]
ns = {"re": re, "np": np, "skimage": skimage}
exec(compile(ast.Module(body=body, type_ignores=[]), "bg_utils_helpers", "exec"), ns)

from Pipeline_utils.customization_util import judge_colorize_type           # noqa: E402

out = {"get_text_type": {s: ns["get_text_type"](s) for s in TYPE_SENTENCES}, "combine_bg_input_text": [],
       "judge_colorize_type": {s: judge_colorize_type(s) for s in TYPE_SENTENCES}, "add_color_gradient": {}}
for new, prev in COMBINE_CASES:
    try:
        out["combine_bg_input_text"].append([new, prev, ns["combine_bg_input_text"](new, prev)])
    except Exception as e:            # the reference signals unusable instructions with exceptions
        out["combine_bg_input_text"].append([new, prev, "EXC:" + str(e.args[0] if e.args else type(e).__name__)])
for name in ("blue_sky", "two_tone", "low_horizon"):
    img, mask = gradient_case(name)
    res = ns["add_color_gradient"](img, mask)
    out["add_color_gradient"][name] = {"sha_rows": [int(r.astype(np.int64).sum()) for r in res], "top_left": res[0, 0].tolist(),
                                       "mid": res[res.shape[0] // 4, res.shape[1] // 2].tolist(), "full": res.tolist()}
json.dump(out, open(os.path.join(HERE, "pipeline_bg.json"), "w"))
print({k: (len(v) if hasattr(v, "__len__") else v) for k, v in out.items()})
