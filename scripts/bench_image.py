#!/usr/bin/env python
"""Time the preprocessing transforms on device-resident pages (CUDA events, warm, median of reps)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import handwritten_ocr_b200  # noqa: E402,F401
from handwritten_ocr_b200 import preprocess as pp, synth  # noqa: E402


def timeit(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=9)
    ap.add_argument("--pages", type=int, default=1)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    rgb = pp.to_device([synth.page(100 + i) for i in range(a.pages)])
    gray = pp.to_gray(rgb)
    ruled = pp.to_device([synth.rule_lines(synth.page(100 + i)) for i in range(a.pages)])
    out = {}
    cases = {"remove_lines_ruled": lambda: pp.remove_lines(ruled),"denoise_rgb": lambda: pp.denoise(rgb), "denoise_gray": lambda: pp.denoise(gray),
             "high_contrast": lambda: pp.high_contrast(rgb), "binarize": lambda: pp.binarize(rgb),
             "sharpen": lambda: pp.sharpen(rgb), "deskew": lambda: pp.deskew(rgb),
             "remove_lines_mask": lambda: pp.remove_lines_mask(rgb)}
    for k, fn in cases.items():
        if a.only and k not in a.only.split(","):
            continue
        out[k] = round(timeit(fn, a.reps), 4)
    print(json.dumps({"pages": a.pages, "shape": list(rgb.shape), "ms": out}))


if __name__ == "__main__":
    main()
