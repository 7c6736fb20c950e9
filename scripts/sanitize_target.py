"""Small end-to-end pass over every kernel family for compute-sanitizer (scripts/sanitize.sh): tiny VLM dimensions, small
pages, every preprocessing transform, a batched read (vision tower incl. the tcgen05 flash attention, prefill, paged decode
with and without the CUDA graph, PDL on), Levenshtein / LCS.  Prints `sanitize target ok` when the results also match
the oracle, so a sanitizer-clean run is a correct run."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import handwritten_ocr_b200  # noqa: F401
from handwritten_ocr_b200 import _lib, engine, preprocess, synth, textops, vlm
from handwritten_ocr_b200.vlm_config import VLMConfig
from oracle import image_ref, text_ref

torch.cuda.set_device(0)
page = synth.rule_lines(synth.page(7, 504, 392))
x = preprocess.to_device(page)
S = [["deskew", "high_contrast", "binarize"], ["high_contrast", "binarize"], ["deskew", "high_contrast", "sharpen"]]
outs = [preprocess.apply_strategy(x, s) for s in S]
for s, o in zip(S, outs):
    assert np.array_equal(o[0].cpu().numpy(), image_ref.apply_strategy(page, s)), s
crop = np.ascontiguousarray(page[40:140, 100:260])
xc = preprocess.to_device(crop)
for name in ("denoise", "remove_lines"):
    assert np.array_equal(preprocess.apply_transform(xc, name)[0].cpu().numpy(), image_ref.TRANSFORMS[name](crop)), name
cfg = VLMConfig.tiny()
w = vlm.VLMWeights.random(cfg, torch.device("cuda", 0), seed=0)
eng = engine.OcrEngine(w, max_batch=4, max_new_tokens=12, max_prompt=400)
batch = torch.cat(outs, 0)
graph = eng.read_batch(batch, max_new_tokens=12)
plain = eng.read_batch(batch, max_new_tokens=12, use_graph=False)
assert graph == plain, "CUDA-graph replay and step-by-step decode disagree"
alone = eng.read_batch(outs[1], max_new_tokens=12)[0]
assert alone == graph[1], "batch invariance"
# the opt-in persistent step (csrc/chain.cu): whole step as one plan launch, then chains with attention kernels between them
for fuse in (True, False):
    vlm.CHAIN_MAX_B, vlm.CHAIN_FUSE_ATTN = 128, fuse
    eng._states.clear()
    fused = eng.read_batch(batch, max_new_tokens=12)
    assert fused == eng.read_batch(batch, max_new_tokens=12, use_graph=False), "persistent step: graph vs step-by-step"
    assert [t[:3] for t in fused] == [t[:3] for t in graph], "persistent step diverges from the launch sequence at once"
vlm.CHAIN_MAX_B = 0
eng._states.clear()
# the long-sequence attention kernel on its own (the tiny config's sequences are short)
for hd, nq, nkv, causal, lens in ((80, 4, 4, 0, [300, 130]), (128, 4, 2, 1, [257, 64])):
    T = sum(lens)
    W_ = (nq + 2 * nkv) * hd
    qkv = (torch.randn(T, W_, device="cuda")).to(torch.bfloat16)
    cu = torch.tensor([0] + list(np.cumsum(lens)), dtype=torch.int32, device="cuda")
    out = torch.empty(T, nq * hd, device="cuda", dtype=torch.bfloat16)
    q, k, v = qkv[:, : nq * hd], qkv[:, nq * hd: (nq + nkv) * hd], qkv[:, (nq + nkv) * hd:]
    _lib.call("ocrb_flash_attention_bf16", q.data_ptr(), W_, k.data_ptr(), W_, v.data_ptr(), W_, out.data_ptr(), nq * hd,
              cu.data_ptr(), len(lens), T, max(lens), nq, nkv, hd, hd ** -0.5, causal, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
texts = [eng.detokenize(t) for t in graph]
assert textops.compare_versions(texts[0], texts[1]) == text_ref.compare_versions(texts[0], texts[1])
assert textops.merge_versions(texts) == text_ref.merge_versions(texts)
gt = synth.corrupt(texts[0], 3)
assert textops.tier1_metrics(gt, texts[2]) == text_ref.tier1_metrics(gt, texts[2])
torch.cuda.synchronize()
print(f"sanitize target ok: {_lib.launch_count()} kernel launches")
