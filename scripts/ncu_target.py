"""Profiling target for ncu: ONE read of the configs[1] / configs[2] workload (P pages x 3 strategies, 7B-class VLM, a few
greedy tokens, no CUDA graph so every kernel is its own launch) inside the NVTX range "capture", after one untimed warm
read.  Run plainly first (must exit 0), then under `ncu --nvtx --nvtx-include "capture/"` (scripts/ncu_r02.sh)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import handwritten_ocr_b200  # noqa
from handwritten_ocr_b200 import engine, preprocess, synth, textops, vlm
from handwritten_ocr_b200.vlm_config import VLMConfig

ap = argparse.ArgumentParser()
ap.add_argument("--pages", type=int, default=1)
ap.add_argument("--new-tokens", type=int, default=4)
a = ap.parse_args()
STRATEGIES = [["deskew", "high_contrast", "binarize"], ["high_contrast", "binarize"], ["deskew", "high_contrast", "sharpen"]]
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
cfg = VLMConfig.olmocr_7b()
w = vlm.VLMWeights.random(cfg, dev, seed=0)
eng = engine.OcrEngine(w, max_batch=3 * a.pages, max_new_tokens=512, max_prompt=1600)
x = preprocess.to_device([synth.page(i) for i in range(a.pages)])


def step():
    cands = [preprocess.apply_strategy(x, st) for st in STRATEGIES]
    batch = torch.stack(cands, 1).reshape((a.pages * 3,) + tuple(cands[0].shape[1:]))
    toks = eng.read_batch(batch, max_new_tokens=a.new_tokens, use_graph=False)
    texts = [eng.detokenize(t) for t in toks]
    for p in range(a.pages):
        tp = texts[3 * p: 3 * p + 3]
        textops.compare_versions(tp[0], tp[1]); textops.merge_versions(tp)
    return toks


step()
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("capture")
toks = step()
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("ok", len(toks), eng.timings)
