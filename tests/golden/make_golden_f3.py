#!/usr/bin/env python
"""Golden fixtures for SURVEY §8 row f3 (denoise; remove_lines incl. the Telea inpaint), from the UNMODIFIED reference.

Run in the build container only (needs /root/reference and cv2 4.13.0.92):  python tests/golden/make_golden_f3.py
Writes denoise.json (sha256 of `ocr_agent.tools._apply_denoise` outputs and of the configured strategy chain
["deskew", "denoise", "high_contrast"], config.py:33) and denoise_small.npz (two small pages with full outputs);
inpaint.json (sha256 of `ocr_agent.tools._apply_remove_lines` outputs on pages ruled by synth.rule_lines, and of the
configured chain ["deskew", "remove_lines", "high_contrast"], config.py:34) and inpaint_small.npz (one small ruled page
with its full output).
"""
import json
import os
import sys

import numpy as np

from make_golden import HERE, apply_chain, sha, synth

CASES = [("rgb_1024x768", 41, 1024, 768, False), ("rgb_768x1024", 42, 768, 1024, False),
         ("gray_640x480", 43, 640, 480, True), ("rgb_259x197", 44, 259, 197, False), ("gray_131x97", 45, 131, 97, True)]
CHAIN = ["deskew", "denoise", "high_contrast"]


RULED = [("ruled_rgb_1024x768", 51, 1024, 768, False), ("ruled_rgb_517x389", 52, 517, 389, False),
         ("ruled_gray_640x480", 53, 640, 480, True), ("ruled_rgb_259x197", 54, 259, 197, False),
         ("ruled_rgb_768x1024", 55, 768, 1024, False)]
CHAIN_RL = ["deskew", "remove_lines", "high_contrast"]


def main_inpaint():
    out, small = {}, {}
    for name, seed, w, h, gray in RULED:
        page = synth.rule_lines(synth.page(seed, w, h, gray=gray))
        res = apply_chain(page, ["remove_lines"])
        out[name] = {"seed": seed, "w": w, "h": h, "gray": gray, "input": sha(page), "remove_lines": sha(res),
                     "changed_px": int((res != page).reshape(h, w, -1).any(-1).sum()),
                     "+".join(CHAIN_RL): sha(apply_chain(page, CHAIN_RL))}
        if w * h < 60000:
            small[f"{name}/input"] = page
            small[f"{name}/remove_lines"] = res
    with open(os.path.join(HERE, "inpaint.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "inpaint_small.npz"), **small)
    print("inpaint fixtures written")


def main():
    out, small = {}, {}
    for name, seed, w, h, gray in CASES:
        page = synth.page(seed, w, h, gray=gray)
        den = apply_chain(page, ["denoise"])
        out[name] = {"seed": seed, "w": w, "h": h, "gray": gray, "input": sha(page), "denoise": sha(den),
                     "+".join(CHAIN): sha(apply_chain(page, CHAIN))}
        if w * h < 60000:
            small[f"{name}/input"] = page
            small[f"{name}/denoise"] = den
    with open(os.path.join(HERE, "denoise.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "denoise_small.npz"), **small)
    print("denoise fixtures written")


if __name__ == "__main__":
    if "--inpaint-only" not in sys.argv:
        main()
    main_inpaint()
