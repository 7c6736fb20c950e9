"""SURVEY §8 rows f1 / f2 / a8 on the GPU (tiny VLM dimensions):
 * `folder.transcribe_folder`: N >= 8 pages primed in cross-page batches, the per-page read phase (the call sequence of
   nodes.py:76-134) answered from the cache; texts equal to reading every page alone; agreement / merge equal the oracle;
 * `folder.eval_folder`: one `ocrb_levenshtein_batch` launch for N files, JSON equal to what eval_final.main collects
   (restated with the oracle's text functions);
 * `tools._load_ocr_model()` itself (not an injected engine): random-init and a safetensors checkpoint directory with
   the legacy tensor names, both read the same text for the same weights."""
import json
import os

import numpy as np
import pytest
import torch
from PIL import Image

from oracle import text_ref

pytestmark = pytest.mark.gpu
S = [["deskew", "high_contrast", "binarize"], ["high_contrast", "binarize"], ["deskew", "high_contrast", "sharpen"]]


@pytest.fixture()
def tiny_tools(pkg):
    from handwritten_ocr_b200 import tools
    from handwritten_ocr_b200.vlm_config import VLMConfig
    saved = (tools._ocr_engine, dict(tools._options), tools.config.OCR_MAX_NEW_TOKENS, tools.config.PREPROCESSING_STRATEGIES,
             getattr(tools.config, "AGREEMENT_THRESHOLD", 80))
    tools.forget()
    tools._ocr_engine = None
    tools.configure(vlm_config=VLMConfig.tiny(), checkpoint=None, max_batch=32, speculative=True, seed=0, cache_pages=64)
    tools.config.OCR_MAX_NEW_TOKENS = 24
    tools.config.PREPROCESSING_STRATEGIES = S
    yield tools
    tools.forget()
    tools._ocr_engine, opts, tools.config.OCR_MAX_NEW_TOKENS, tools.config.PREPROCESSING_STRATEGIES, thr = saved
    tools.config.AGREEMENT_THRESHOLD = thr
    tools._options.clear()
    tools._options.update(opts)


def test_transcribe_folder_batches_pages_and_matches_page_by_page(tiny_tools, synth, tmp_path, capsys):
    from handwritten_ocr_b200 import folder
    tools = tiny_tools
    src = tmp_path / "in"
    src.mkdir()
    n_pages = 9
    for i in range(n_pages):
        Image.fromarray(synth.page(500 + i, 504, 392)).save(src / f"note_{i:02d}.png")
    (src / "readme.txt").write_text("not an image")
    eng = tools._load_ocr_model()                         # the real loader: random-init tiny VLM on the GPU
    assert tools._ocr_engine is eng and eng.max_batch == 32
    calls = []
    orig = eng.read_batch

    def counting(pages, **kw):
        calls.append(int(pages.shape[0]))
        return orig(pages, **kw)

    eng.read_batch = counting
    out_dir = tmp_path / "out"

    def page_fn(img, output_dir, gt_path, **kw):
        r = folder.initial_ocr_page(str(img), agreement_threshold=101, tools=tools)     # 101: always take the tiebreaker
        output_dir.mkdir(parents=True, exist_ok=True)
        (output_dir / f"{img.stem}_transcription.txt").write_text(r["current_best"], encoding="utf-8")
        return r

    try:
        res = folder.transcribe_folder(src, out_dir, pages_per_batch=4, page_fn=page_fn, tools=tools)
    finally:
        eng.read_batch = orig
    capsys.readouterr()
    # 9 pages x 3 strategies, 4 pages per batch: three batched reads, nothing page by page
    assert calls == [12, 12, 3], calls
    assert len(res) == n_pages and sorted(p.name for p in out_dir.iterdir()) == [f"note_{i:02d}_transcription.txt" for i in range(n_pages)]
    for i, r in enumerate(res):
        assert [c["source"] for c in r["candidates"]] == ["ocr_" + "+".join(s) for s in S]
        texts = [c["text"] for c in r["candidates"]]
        assert r["comparison"] == text_ref.compare_versions(texts[0], texts[1])
        assert r["current_best"] == text_ref.merge_versions(texts)
    # batch invariance at the folder level: page 5 read alone gives the texts it got inside its 12-sequence batch
    tools.forget()
    img = str(src / "note_05.png")
    alone = [tools.run_ocr(tools.preprocess_image(img, s)) for s in S]
    capsys.readouterr()
    assert alone == [c["text"] for c in res[5]["candidates"]]


def test_eval_folder_one_launch_equals_oracle(tiny_tools, synth, tmp_path):
    from handwritten_ocr_b200 import _lib, folder
    tools = tiny_tools
    res, gtd = tmp_path / "results", tmp_path / "gt"
    res.mkdir()
    gtd.mkdir()
    n = 12
    want = []
    for i in range(n):
        gt = synth.text(60 + i, 80 + 40 * i)
        ocr = synth.corrupt(gt, i, 0.02 * (i + 1))
        (res / f"f{i:02d}_transcription.txt").write_text(ocr, encoding="utf-8")
        if i != 7:
            (gtd / f"f{i:02d}.md").write_text(f"# notes\n## Ground Truth\n{gt}\n", encoding="utf-8")
            want.append({"tier1_raw_vs_gt": text_ref.tier1_metrics(gt.strip(), ocr), "file": str((res / f"f{i:02d}_transcription.txt").resolve())})
        else:
            want.append({"file": str((res / f"f{i:02d}_transcription.txt").resolve())})
    out = tmp_path / "eval.json"
    before = _lib.launch_count()
    got = folder.eval_folder(res, gtd, out, tools=tools)
    assert _lib.launch_count() - before == 1, "all 3 x 11 distances of the folder must go through ONE levenshtein launch"
    assert got == want
    assert json.loads(out.read_text(encoding="utf-8")) == want
    assert all(0 < r["tier1_raw_vs_gt"]["cer"] < 0.5 for r in got if "tier1_raw_vs_gt" in r)


def test_load_ocr_model_from_legacy_named_safetensors_dir(tiny_tools, synth, tmp_path, capsys):
    """tools._load_ocr_model (tools.py:683-711) on a LOCAL checkpoint directory: config.json + sharded safetensors with the
    tensor names published Qwen2.5-VL checkpoints carry (`visual.*`, `model.layers.*`).  Same weights as the random-init
    engine of the same seed => same greedy text."""
    from safetensors.torch import save_file
    from handwritten_ocr_b200 import vlm
    from handwritten_ocr_b200.vlm_config import VLMConfig
    tools = tiny_tools
    cfg = VLMConfig.tiny()
    img = str(tmp_path / "page.png")
    Image.fromarray(synth.page(640, 504, 392)).save(img)
    ref_text = tools.transcribe(img, S[1])                                  # random-init engine, seed 0
    sd = vlm.random_state_dict(cfg, "cpu", seed=0)

    def legacy(k):
        if k.startswith("model.visual."):
            return k[len("model."):]
        if k.startswith("model.language_model."):
            return "model." + k[len("model.language_model."):]
        return k

    ck = tmp_path / "ckpt"
    ck.mkdir()
    keys = sorted(sd)
    save_file({legacy(k): sd[k].contiguous() for k in keys[::2]}, str(ck / "model-00001-of-00002.safetensors"))
    save_file({legacy(k): sd[k].contiguous() for k in keys[1::2]}, str(ck / "model-00002-of-00002.safetensors"))
    tools.forget()
    tools._ocr_engine = None
    tools.configure(checkpoint=str(ck))
    eng = tools._load_ocr_model()
    assert eng is tools._ocr_engine
    got = tools.transcribe(img, S[1])
    capsys.readouterr()
    # torch.randn on "cpu" and on "cuda" are different generators, so compare against an engine built from the SAME dict
    from handwritten_ocr_b200 import engine as eng_mod
    w = vlm.VLMWeights.from_state_dict(cfg, {k: v.cuda() for k, v in sd.items()})
    direct = eng_mod.OcrEngine(w, max_batch=4, max_new_tokens=24, max_prompt=400)
    from handwritten_ocr_b200 import preprocess
    page = preprocess.apply_strategy(preprocess.to_device(np.array(Image.open(img))), S[1])
    want = direct.detokenize(direct.read_batch(page, max_new_tokens=24)[0])
    assert got == want and isinstance(ref_text, str) and ref_text
    tools.unload_ocr_model()
    assert tools._ocr_engine is eng, "weights stay resident by default (180 GB of HBM)"
    tools.configure(keep_resident=False)
    tools.unload_ocr_model()
    assert tools._ocr_engine is None
    capsys.readouterr()
