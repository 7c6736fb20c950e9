#!/usr/bin/env python
"""Generate the committed golden fixtures from the UNMODIFIED reference.

Run in the build container only (needs /root/reference, cv2 4.13.0.92, transformers):
    python tests/golden/make_golden.py
The reference has no tests or fixtures of its own, so the pins are
outputs of its own functions (`ocr_agent.tools.*`, imported with only `ollama`
stubbed) and of the HF image processor it calls, on the seeded synthetic inputs
of handwritten-ocr_b200/synth.py.  /root/reference is never read by tests at run time.

Writes:
  image_small.npz     small pages + reference outputs of every transform / strategy chain
  image_hashes.json   sha256 of reference outputs on full-size (1024x768 / 768x1024) pages,
                      cv2 minAreaRect angles, HF pixel_values hashes
  text_golden.json    levenshtein / compare_versions / merge_versions / tier1_metrics cases
"""
import hashlib
import importlib.util
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")
sys.modules.setdefault("ollama", types.ModuleType("ollama"))
from ocr_agent import tools as ref  # noqa: E402
from ocr_agent import config as ref_config  # noqa: E402
from PIL import Image  # noqa: E402

spec = importlib.util.spec_from_file_location("synth", os.path.join(ROOT, "handwritten-ocr_b200", "synth.py"))
synth = importlib.util.module_from_spec(spec)
spec.loader.exec_module(synth)

GPU_TRANSFORMS = ["high_contrast", "binarize", "sharpen", "deskew"]
CHAINS = [["deskew", "high_contrast", "binarize"], ["high_contrast", "binarize"],
          ["deskew", "high_contrast", "sharpen"]]


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def apply_chain(arr: np.ndarray, chain) -> np.ndarray:
    img = Image.fromarray(arr)
    for step in chain:
        img = ref._TRANSFORMS[step](img)
    return np.array(img)


def cv_angle(arr: np.ndarray):
    import cv2
    gray = cv2.cvtColor(arr, cv2.COLOR_RGB2GRAY) if arr.ndim == 3 else arr
    coords = np.column_stack(np.where(gray < 128))
    if len(coords) <= 100:
        return None
    a = cv2.minAreaRect(coords)[-1]
    return float(-(90 + a) if a < -45 else -a)


def main():
    # ---- small images, full arrays ----
    small = {}
    cases = [("rgb_256x192", synth.page(11, 256, 192)), ("rgb_203x157", synth.page(12, 203, 157)),
             ("gray_256x192", synth.page(13, 256, 192, gray=True)),
             ("rgb_blank_128x96", np.full((96, 128, 3), 230, np.uint8)),
             ("rgb_ruled_320x240", synth.page(14, 320, 240, ruled=True))]
    for name, arr in cases:
        small[f"{name}/input"] = arr
        for t in GPU_TRANSFORMS:
            small[f"{name}/{t}"] = apply_chain(arr, [t])
        for ch in CHAINS:
            small[f"{name}/{'+'.join(ch)}"] = apply_chain(arr, ch)
        ang = cv_angle(arr)
        small[f"{name}/angle"] = np.array([np.nan if ang is None else ang], np.float64)
    np.savez_compressed(os.path.join(HERE, "image_small.npz"), **small)

    # ---- full-size pages, hashes only ----
    from transformers.models.qwen2_vl.image_processing_qwen2_vl import Qwen2VLImageProcessor
    ip = Qwen2VLImageProcessor(min_pixels=ref_config.OCR_MIN_PIXELS, max_pixels=ref_config.OCR_MAX_PIXELS)
    hashes = {}
    for seed, (w, h) in [(0, (1024, 768)), (1, (768, 1024)), (2, (1024, 768)), (3, (1024, 768)),
                         (5, (640, 480)), (6, (1600, 1200))]:
        arr = synth.page(seed, w, h)
        ent = {"w": w, "h": h, "input": sha(arr), "angle": cv_angle(arr)}
        for t in GPU_TRANSFORMS:
            ent[t] = sha(apply_chain(arr, [t]))
        for ch in CHAINS:
            out = apply_chain(arr, ch)
            ent["+".join(ch)] = sha(out)
            rgb = np.array(Image.fromarray(out).convert("RGB"))
            o = ip(images=[Image.fromarray(rgb)], return_tensors="pt")
            ent["pv:" + "+".join(ch)] = sha(o["pixel_values"].numpy().astype(np.float32))
            ent["grid:" + "+".join(ch)] = o["image_grid_thw"].numpy().tolist()
        o = ip(images=[Image.fromarray(arr)], return_tensors="pt")
        ent["pv:original"] = sha(o["pixel_values"].numpy().astype(np.float32))
        ent["grid:original"] = o["image_grid_thw"].numpy().tolist()
        hashes[f"seed{seed}"] = ent
    with open(os.path.join(HERE, "image_hashes.json"), "w") as f:
        json.dump(hashes, f, indent=1, sort_keys=True)

    # ---- text ----
    tg = {"levenshtein": [], "levenshtein_words": [], "compare_versions": [], "merge_versions": [],
          "tier1_metrics": [], "normalize_text": []}
    edge_pairs = [("", ""), ("", "abc"), ("abc", ""), ("kitten", "sitting"), ("flaw", "lawn"),
                  ("a", "a"), ("İstanbul “x” — y", "istanbul \"x\" - y"), ("a b\tc\n", "a b c"),
                  ("abc" * 40, "abd" * 37)]
    for a, b in edge_pairs:
        tg["levenshtein"].append({"a": a, "b": b, "d": ref.levenshtein(a, b)})
    for seed in range(6):
        base = synth.text(seed, 60 + 30 * seed)
        v1 = synth.corrupt(base, 10 + seed, 0.04)
        v2 = synth.corrupt(base, 20 + seed, 0.08)
        v3 = synth.corrupt(base, 30 + seed, 0.02 + 0.03 * seed)
        tg["levenshtein"].append({"a": v1, "b": v2, "d": ref.levenshtein(v1, v2)})
        tg["levenshtein_words"].append({"a": v1.split(), "b": v2.split(),
                                        "d": ref._levenshtein_words(v1.split(), v2.split())})
        tg["compare_versions"].append({"v1": v1, "v2": v2, "out": ref.compare_versions(v1, v2)})
        tg["merge_versions"].append({"versions": [v1, v2, v3], "out": ref.merge_versions([v1, v2, v3])})
        tg["merge_versions"].append({"versions": [v2, v1], "out": ref.merge_versions([v2, v1])})
        for lower in (False, True):
            tg["tier1_metrics"].append({"gt": base, "ocr": v3, "lower": lower,
                                        "out": ref.tier1_metrics(base, v3, lower)})
        tg["normalize_text"].append({"in": base, "out": ref.normalize_text(base),
                                     "out_lower": ref.normalize_text(base, True)})
    for vs in ([], ["only One  version\n"], ["a b c d", "a x c d", "a y c d"], ["a b c", "a B c"],
               ["the cat sat", "the cat sat on", "The cat"], ["x", "y"], ["", "a b"]):
        tg["merge_versions"].append({"versions": vs, "out": ref.merge_versions(vs)})
    for a, b in [("", ""), ("same text here", "same text here"), ("a b c d e f", "a c d x f g h"),
                 ("one two three four five six seven eight nine ten eleven twelve", "zero")]:
        tg["compare_versions"].append({"v1": a, "v2": b, "out": ref.compare_versions(a, b)})
    tg["tier1_metrics"].append({"gt": "", "ocr": "abc", "lower": False, "out": ref.tier1_metrics("", "abc")})
    with open(os.path.join(HERE, "text_golden.json"), "w") as f:
        json.dump(tg, f, indent=1, ensure_ascii=True)
    print("golden fixtures written")


if __name__ == "__main__":
    main()


# remove_lines.json (ruled-line mask of tools._apply_remove_lines, tools.py:598-614) was generated in the same
# container by the snippet recorded in tests/test_remove_lines.py::GOLDEN_RECIPE (reference function for the
# unchanged-page cases, the reference's own cv2 call sequence for the masks of pages with drawn ruled lines).
