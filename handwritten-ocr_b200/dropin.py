"""Install this package behind the reference's `ocr_agent.tools` name.

`ocr_agent/nodes.py:8-14` binds `compare_versions, merge_versions, preprocess_image, run_ocr,
unload_ocr_model` by name at import time, `transcribe.py:33` / `eval_final.py:22` bind `evaluate`
and `parse_ground_truth`.  `install()` (a) registers our tools module as `ocr_agent.tools` so later
imports bind to it and (b) re-binds the names on reference modules that were imported already.
"""
from __future__ import annotations

import sys

HOT = ("preprocess_image", "run_ocr", "unload_ocr_model", "compare_versions", "merge_versions", "evaluate",
       "tier1_metrics", "cer", "wer", "levenshtein", "normalize_text")


def install():
    from . import tools
    sys.modules["ocr_agent.tools"] = tools
    pkg = sys.modules.get("ocr_agent")
    if pkg is not None:
        setattr(pkg, "tools", tools)
    for modname in ("ocr_agent.nodes", "ocr_agent.transcribe", "ocr_agent.eval_final"):
        m = sys.modules.get(modname)
        if m is None:
            continue
        for name in HOT:
            if hasattr(m, name):
                setattr(m, name, getattr(tools, name))
    return tools
