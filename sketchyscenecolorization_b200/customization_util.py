"""Editing records of the interactive pipeline and the FG / BG dispatch of an instruction.

Same names, arguments and files as the reference's Pipeline_utils/customization_util.py: `judge_colorize_type` (:8-17),
`fetch_records` (:20-52), `update_records` (:55-70), `withdraw_records` (:73-106).  The record file is
<results_base_dir>/update_records/<image_id>_records.json, a list of {colorization_type, result_name, input_text,
proc_bg_text}; results are <results_base_dir>/results/<image_id>/<image_id>_<k>.png.
"""
from __future__ import annotations

import collections
import json
import os

from .pipeline_fg import _self_category

_KEYS = ("colorization_type", "result_name", "input_text", "proc_bg_text")


def judge_colorize_type(text):
    """'FG' when the instruction names an object category, else 'BG' (search_for_self_category, :15-17)."""
    return 'BG' if _self_category(text) is None else 'FG'


def _records_path(image_id, results_base_dir):
    records_dir = os.path.join(results_base_dir, 'update_records')
    os.makedirs(records_dir, exist_ok=True)
    return os.path.join(records_dir, str(image_id) + '_records.json')


def _copy(rec):
    return collections.OrderedDict((k, rec[k]) for k in _KEYS)


def fetch_records(image_id, results_base_dir):
    """-> (new_result_image_name, last_result_image_name, last_bg_text, summary_data)."""
    path = _records_path(image_id, results_base_dir)
    if not os.path.isfile(path):
        return str(image_id) + '_1.png', '', "", []
    with open(path) as fp:
        records = json.load(fp)
    print(len(records), 'editing records')
    summary = [_copy(r) for r in records]
    last_bg_text = records[-1]["proc_bg_text"] if records else ""
    return str(image_id) + '_' + str(len(records) + 1) + '.png', records[-1]['result_name'], last_bg_text, summary


def update_records(image_id, input_text, results_base_dir, colorization_type, new_result_image_name, proc_bg_text, summary_data):
    path = _records_path(image_id, results_base_dir)
    summary_data.append(collections.OrderedDict(zip(_KEYS, (colorization_type, new_result_image_name, input_text, proc_bg_text))))
    with open(path, "w") as f:
        f.write(json.dumps(summary_data, indent=4))


def withdraw_records(image_id, results_base_dir):
    results_dir = os.path.join(results_base_dir, 'results', str(image_id))
    path = os.path.join(results_base_dir, 'update_records', str(image_id) + '_records.json')
    if not os.path.isfile(path):
        raise Exception('No record to withdraw.')
    with open(path) as fp:
        records = json.load(fp)
    print('Original: ', len(records), 'editing records')
    os.remove(os.path.join(results_dir, str(image_id) + '_' + str(len(records)) + '.png'))
    if len(records) == 1:
        os.remove(path)
    else:
        with open(path, "w") as f:
            f.write(json.dumps([_copy(r) for r in records[:-1]], indent=4))
