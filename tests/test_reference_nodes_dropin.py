"""The UNMODIFIED reference nodes (`/root/reference/ocr_agent/nodes.py`) on top of this package's `tools` module
(SURVEY §8b): `install()` puts it at `ocr_agent.tools`, `nodes.py` binds its five names from it at import time, and
`node_initial_ocr` / `node_reocr` run through it.  CPU-only plumbing test: everything BELOW the tools surface is faked
(pages stay on the host and go through the oracle, the engine returns canned token ids, agreement / merge use the
oracle's text functions) -- the kernels themselves are covered by the `-m gpu` tests.  What is checked here is the
contract: names bound, call order, one batched read behind the sequential calls, temp files, candidate dicts, state keys.
Skipped where /root/reference does not exist (the GPU box)."""
import os
import sys
import types

import numpy as np
import pytest
import torch
from PIL import Image

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "ocr_agent")), reason="reference tree not present")

S = [["deskew", "high_contrast", "binarize"], ["high_contrast", "binarize"], ["deskew", "high_contrast", "sharpen"],
     ["deskew", "denoise", "high_contrast"], ["deskew", "remove_lines", "high_contrast"]]


class FakeEngine:
    """Stands where OcrEngine stands: batched read of uint8 pages -> token id lists."""
    max_batch = 8

    def __init__(self):
        self.calls = []

    def read_batch(self, pages, prompt=None, max_new_tokens=None):
        self.calls.append(int(pages.shape[0]))
        # ids derived from the page content, so different candidates give different (but reproducible) texts
        out = []
        for i in range(pages.shape[0]):
            seed = int(pages[i].to(torch.int64).sum()) % 9973
            rng = np.random.default_rng(seed % 7)            # few distinct seeds: candidates mostly agree
            out.append([int(t) for t in rng.integers(300, 5000, 30)])
        return out

    def detokenize(self, ids):
        from handwritten_ocr_b200.vlm_config import SyntheticTokenizer
        return SyntheticTokenizer().decode(ids)


@pytest.fixture()
def rig(pkg, synth, tmp_path, monkeypatch):
    from oracle import image_ref, text_ref
    if REF not in sys.path:
        monkeypatch.syspath_prepend(REF)
    monkeypatch.setitem(sys.modules, "ollama", types.ModuleType("ollama"))
    for m in [k for k in sys.modules if k == "ocr_agent" or k.startswith("ocr_agent.")]:
        monkeypatch.delitem(sys.modules, m)
    import handwritten_ocr_b200
    from handwritten_ocr_b200 import preprocess, tools
    # below the surface: host tensors, oracle arithmetic, canned reads
    monkeypatch.setattr(preprocess, "to_device", lambda a: torch.from_numpy(np.ascontiguousarray(a))[None])
    monkeypatch.setattr(preprocess, "apply_strategy",
                        lambda x, s: torch.from_numpy(image_ref.apply_strategy(x[0].numpy(), s))[None])
    monkeypatch.setattr(tools, "compare_versions", text_ref.compare_versions)
    monkeypatch.setattr(tools, "merge_versions", text_ref.merge_versions)
    eng = FakeEngine()
    monkeypatch.setattr(tools, "_ocr_engine", eng)
    saved = dict(tools._options)
    tools.configure(speculative=True, max_batch=8)
    tools.forget()
    handwritten_ocr_b200.install()
    import ocr_agent.nodes as nodes                          # the reference file, unmodified
    from ocr_agent import config as ref_config
    monkeypatch.setattr(tools.config, "PREPROCESSING_STRATEGIES", S, raising=False)
    img = str(tmp_path / "page.png")
    Image.fromarray(synth.rule_lines(synth.page(77, 200, 120))).save(img)
    state = {"image_path": img, "candidates": [], "critiques": [], "edits": [], "current_best": "", "current_score": 0.0,
             "iteration": 0, "max_iterations": 3, "status": "running", "reason": "", "strategies_used": [],
             "plateau_count": 0, "prev_score": 0.0, "prev_critique": None,
             "config": {"strategies": S, "agreement_threshold": 101, "accept_threshold": 85, "plateau_patience": 2},
             "trace_events": [], "start_time": 0.0}
    yield nodes, tools, eng, state, ref_config
    tools.forget()
    tools._options.update(saved)


def test_names_bound_from_our_module(rig):
    nodes, tools, eng, state, _ = rig
    assert nodes.__file__.startswith(REF)
    assert sys.modules["ocr_agent.tools"] is tools
    for name in ("compare_versions", "merge_versions", "preprocess_image", "run_ocr", "unload_ocr_model"):
        assert getattr(nodes, name) is getattr(tools, name), name


def test_node_initial_ocr_and_reocr_on_unmodified_nodes(rig, capsys):
    from oracle import image_ref, text_ref
    nodes, tools, eng, state, _ = rig
    upd = nodes.node_initial_ocr(state)
    out = capsys.readouterr().out
    assert "=== PHASE 1: Initial OCR Reads ===" in out and "[preprocess] Applying deskew+high_contrast+binarize..." in out
    assert set(upd) == {"candidates", "current_best", "strategies_used", "trace_events"}
    # agreement_threshold 101 forces the tiebreaker: three candidates, in strategy order, with the reference's dict shape
    labels = ["+".join(s) for s in S[:3]]
    assert [c["source"] for c in upd["candidates"]] == [f"ocr_{l}" for l in labels]
    assert all(c["ocr_params"] == {"strategy": l} and c["score"] is None for c, l in zip(upd["candidates"], labels))
    assert upd["strategies_used"] == labels
    assert upd["current_best"] == text_ref.merge_versions([c["text"] for c in upd["candidates"]])
    # the sequential calls were served by ONE batched read of all five configured candidates of the page
    assert eng.calls == [5], eng.calls
    # every returned path was a real PNG holding exactly the strategy's output
    page = np.array(Image.open(state["image_path"]))
    for s in S:
        p = tools.preprocess_image(state["image_path"], s)
        assert os.path.isfile(p) and p.endswith(".png") and os.path.basename(p).startswith("ocr_" + "+".join(s) + "_")
        assert np.array_equal(np.array(Image.open(p)), image_ref.apply_strategy(page, s)), s
    # reocr: the next unused strategy is served from the cached batch (no further read); the arbitrator is the LLM side
    state.update(upd)
    state["iteration"] = 1
    seen = {}
    nodes.run_arbitrator = lambda versions: seen.setdefault("v", versions) and types.SimpleNamespace(
        final_text=versions[1]["text"], confidence=50, uncertain_segments=[], decisions=[],
        model_dump=lambda: {"final_text": versions[1]["text"]})
    upd2 = nodes.node_reocr(state)
    assert eng.calls == [5]
    assert upd2["strategies_used"] == labels + ["+".join(S[3])]
    assert upd2["candidates"][-1]["source"] == "ocr_" + "+".join(S[3]) and upd2["current_best"] == upd2["candidates"][-1]["text"]
    assert seen["v"][0]["source"] == "current_best"
    # exhausted once every strategy has been used
    state.update(upd2)
    state.update(nodes.node_reocr(state))
    assert nodes.node_reocr(state) == {"reason": "exhausted", "trace_events": state["trace_events"]}
