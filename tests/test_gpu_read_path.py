"""The drop-in surface end to end on the GPU (tiny VLM): the call order nodes.py makes (nodes.py:27-134 --
preprocess S0, read, preprocess S1, read, compare, tiebreaker S2, read, merge, unload; then a reocr read and
evaluate), checked for the reference's contract: real temp files with the input's suffix, one batched read serving
all candidates, texts equal to reading each candidate alone, agreement / merge / CER equal to the oracle."""
import os

import numpy as np
import pytest
import torch
from PIL import Image

from oracle import text_ref

pytestmark = pytest.mark.gpu
S = [["deskew", "high_contrast", "binarize"], ["high_contrast", "binarize"], ["deskew", "high_contrast", "sharpen"],
     ["deskew", "denoise", "high_contrast"]]


@pytest.fixture(scope="module")
def rig(pkg, synth, tmp_path_factory):
    from handwritten_ocr_b200 import tools, vlm, engine
    from handwritten_ocr_b200.vlm_config import VLMConfig
    cfg = VLMConfig.tiny()
    w = vlm.VLMWeights.random(cfg, torch.device("cuda"), seed=0)
    eng = engine.OcrEngine(w, max_batch=4, max_new_tokens=40, max_prompt=400)
    tools._ocr_engine = eng
    tools.configure(speculative=True, max_batch=4)
    tools.config.OCR_MAX_NEW_TOKENS = 40
    tools.config.PREPROCESSING_STRATEGIES = S
    d = tmp_path_factory.mktemp("pages")
    paths = []
    for i in range(2):
        p = str(d / f"page{i}.png")
        Image.fromarray(synth.page(300 + i, 504, 392)).save(p)
        paths.append(p)
    return tools, eng, paths


def test_node_call_order_contract(rig, capsys):
    tools, eng, paths = rig
    img = paths[0]
    calls = []
    orig = eng.read_batch

    def counting(pages, **kw):
        calls.append(pages.shape[0])
        return orig(pages, **kw)

    eng.read_batch = counting
    try:
        p0 = tools.preprocess_image(img, S[0])
        assert os.path.isfile(p0) and p0.endswith(".png") and os.path.basename(p0).startswith("ocr_deskew+high_contrast+binarize_")
        t0 = tools.run_ocr(p0)
        p1 = tools.preprocess_image(img, S[1])
        t1 = tools.run_ocr(p1)
        cmp_ = tools.compare_versions(t0, t1)
        p2 = tools.preprocess_image(img, S[2])
        t2 = tools.run_ocr(p2)
        merged = tools.merge_versions([t0, t1, t2])
        tools.unload_ocr_model()
    finally:
        eng.read_batch = orig
    out = capsys.readouterr().out
    assert "[preprocess] Applying deskew+high_contrast+binarize..." in out and "[ocr] Running OCR on" in out
    assert calls == [len(S)], f"all configured candidates of the page must share one batched read, got {calls}"
    assert len({p0, p1, p2}) == 3 and all(isinstance(t, str) and t for t in (t0, t1, t2))
    # the temp file holds exactly the preprocessed page (PNG round trip is lossless)
    from handwritten_ocr_b200 import preprocess
    want = preprocess.apply_strategy(preprocess.to_device(np.array(Image.open(img))), S[1])[0].cpu().numpy()
    assert np.array_equal(np.array(Image.open(p1)), want)
    # reading a candidate alone (fresh cache) gives the text it got in the batch: batch invariance at the API level
    tools.forget(img)
    alone = tools.run_ocr(p1)          # p1 is no longer a known handle: read from the file, batch of 1
    assert alone == t1
    # agreement / merge equal the oracle on the same strings
    assert cmp_ == text_ref.compare_versions(t0, t1)
    assert merged == text_ref.merge_versions([t0, t1, t2])


def test_reocr_and_evaluate(rig, synth):
    tools, eng, paths = rig
    img = paths[1]
    t = tools.transcribe(img, S[0])
    assert t == tools.run_ocr(tools.preprocess_image(img, S[0]))          # cached handle, same text
    gt = synth.corrupt(t, 9, 0.05)
    ev = tools.evaluate(t, gt)
    assert ev == text_ref.evaluate(t, gt) and 0 < ev["tier1_raw_vs_gt"]["cer"] < 0.2
    assert tools.evaluate(t) == {}
    # params override bypasses the cache (tools.py:741-742)
    short = tools.run_ocr(tools.preprocess_image(img, S[1]), {"max_new_tokens": 5})
    assert 0 < len(short.split()) <= 5


def test_original_unknown_and_denoise_transforms(rig, capsys):
    tools, eng, paths = rig
    img = paths[1]
    assert tools.preprocess_image(img, "original") == img and tools.preprocess_image(img, []) == img
    p = tools.preprocess_image(img, ["no_such_transform", "sharpen"])
    assert "Unknown transform 'no_such_transform', skipping" in capsys.readouterr().out and os.path.isfile(p)
    p3 = tools.preprocess_image(img, S[3])         # the configured denoise strategy (config.py:33)
    assert os.path.isfile(p3) and os.path.basename(p3).startswith("ocr_deskew+denoise+high_contrast_")
    with pytest.raises(Exception):
        tools.preprocess_image("/nonexistent/page.png", S[1])


def test_complete_reocr_sweep_config4(pkg, synth, tmp_path):
    """BASELINE configs[3]: every distinct configured strategy of a page (config.py:29-36, all five -- denoise and
    remove_lines included) preprocessed on the GPU, read in ONE paged-KV batch, each transcription scored against a
    synthetic ground truth.  The page is ruled, so remove_lines really inpaints.  Checks: preprocessed temp files equal
    the oracle's arrays bit for bit, one batched read of 5, evaluate() equal to the oracle's."""
    from handwritten_ocr_b200 import tools, vlm, engine
    from handwritten_ocr_b200.vlm_config import VLMConfig
    from oracle import image_ref
    strategies = [["deskew", "high_contrast", "binarize"], ["high_contrast", "binarize"],       # config.py:29-36:
                  ["deskew", "high_contrast", "sharpen"], ["deskew", "denoise", "high_contrast"],  # six entries,
                  ["deskew", "remove_lines", "high_contrast"], ["deskew", "high_contrast", "binarize"]]  # five distinct
    distinct = []
    for s in strategies:
        if s not in distinct:
            distinct.append(s)
    assert len(distinct) == 5
    cfg = VLMConfig.tiny()
    w = vlm.VLMWeights.random(cfg, torch.device("cuda"), seed=0)
    eng = engine.OcrEngine(w, max_batch=8, max_new_tokens=24, max_prompt=400)
    saved = (tools._ocr_engine, dict(tools._options), tools.config.OCR_MAX_NEW_TOKENS, tools.config.PREPROCESSING_STRATEGIES)
    tools._ocr_engine = eng
    tools.configure(speculative=True, max_batch=8)
    tools.config.OCR_MAX_NEW_TOKENS = 24
    tools.config.PREPROCESSING_STRATEGIES = strategies
    page = synth.rule_lines(synth.page(410, 504, 392))
    img = str(tmp_path / "ruled.png")
    Image.fromarray(page).save(img)
    calls = []
    orig = eng.read_batch

    def counting(pages, **kw):
        calls.append(pages.shape[0])
        return orig(pages, **kw)

    eng.read_batch = counting
    try:
        texts = [tools.run_ocr(tools.preprocess_image(img, s)) for s in strategies]
        assert calls == [5], calls
        assert texts[0] == texts[5]                               # the repeated strategy is served from the cache
        for s in distinct:
            got = np.array(Image.open(tools.preprocess_image(img, s)))
            assert np.array_equal(got, image_ref.apply_strategy(page, s)), s
        assert (image_ref.remove_lines(page) != page).any()
        gt = synth.corrupt(texts[0], 3, 0.05)
        for t in texts:
            assert tools.evaluate(t, gt) == text_ref.evaluate(t, gt)
    finally:
        eng.read_batch = orig
        tools.forget(img)
        tools._ocr_engine, opts, tools.config.OCR_MAX_NEW_TOKENS, tools.config.PREPROCESSING_STRATEGIES = saved
        tools._options.update(opts)
