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


# ───────────── the whole unmodified graph (graph.py) ─────────────
def _langgraph_shim():
    """`langgraph` is not installed here; graph.py uses only this much of it (graph.py:5,51-79)."""
    START, END = "__start__", "__end__"

    class _Compiled:
        def __init__(self, nodes, edges, cond):
            self.nodes, self.edges, self.cond = nodes, edges, cond

        def invoke(self, state):
            state = dict(state)
            cur = self.edges[START]
            for _ in range(200):
                if cur == END:
                    return state
                state.update(self.nodes[cur](state) or {})             # partial update, last write wins
                if cur in self.cond:
                    route, mapping = self.cond[cur]
                    cur = mapping[route(state)]
                else:
                    cur = self.edges[cur]
            raise RuntimeError("graph did not terminate")

    class StateGraph:
        def __init__(self, _state_type):
            self.nodes, self.edges, self.cond = {}, {}, {}

        def add_node(self, name, fn):
            self.nodes[name] = fn

        def add_edge(self, a, b):
            self.edges[a] = b

        def add_conditional_edges(self, a, route, mapping):
            self.cond[a] = (route, mapping)

        def compile(self):
            return _Compiled(self.nodes, self.edges, self.cond)

    pkg = types.ModuleType("langgraph")
    graph = types.ModuleType("langgraph.graph")
    graph.START, graph.END, graph.StateGraph = START, END, StateGraph
    pkg.graph = graph
    return pkg, graph


def test_full_graph_reocr_sweep_on_unmodified_graph(rig, monkeypatch, capsys):
    """graph.py + nodes.py + agents.py unmodified; the LLM side is a fake `call_llm_json` whose critic asks for a re-OCR
    with strictly rising confidence (a constant one trips the plateau exit: nodes.py:190-194), so the graph sweeps all
    five configured strategies (BASELINE configs[3]) and stops on 'exhausted'.  All reads of the page come out of ONE
    batched engine call."""
    nodes, tools, eng, state, _ = rig
    conf = iter(range(30, 84, 6))

    def fake_llm(system_prompt, user_msg, json_schema=None, **kw):
        title = (json_schema or {}).get("title", "")
        if title == "CriticResult":
            return {"overall_confidence": next(conf), "segments": [], "verdict": "needs_reocr", "reasoning": "fake"}
        if title == "ArbitratorResult":
            return {"final_text": "arbitrated " + user_msg[-40:], "decisions": [], "confidence": 55, "uncertain_segments": []}
        return {"corrected_text": "edited", "changes": [], "unresolved": []}

    import ocr_agent.agents as agents
    monkeypatch.setattr(agents, "call_llm_json", fake_llm)
    lg, lgg = _langgraph_shim()
    monkeypatch.setitem(sys.modules, "langgraph", lg)
    monkeypatch.setitem(sys.modules, "langgraph.graph", lgg)
    sys.modules.pop("ocr_agent.graph", None)
    import ocr_agent.graph as graph
    assert graph.__file__.startswith(REF)
    state["max_iterations"] = 10
    final = graph.build_ocr_graph().invoke(state)
    capsys.readouterr()
    assert final["status"] == "completed" and final["reason"] == "exhausted"
    assert final["strategies_used"] == ["+".join(s) for s in S]
    assert [c["source"] for c in final["candidates"]] == ["ocr_" + "+".join(s) for s in S]
    assert final["iteration"] == 3 and len(final["critiques"]) == 3 and final["current_best"].startswith("arbitrated ")
    assert eng.calls == [5]
    actions = [e["action"] for e in final["trace_events"]]
    assert actions.count("ocr") == 5 and actions.count("arbitrate") == 2 and actions[-1] == "strategies_exhausted"


def test_transcribe_single_unmodified_with_ground_truth(rig, monkeypatch, tmp_path, capsys):
    """The reference's own entry point (transcribe.py:21-116): graph run, transcription / trace / eval files, CER against a
    ground-truth file through the `evaluate` this package exports (oracle arithmetic here, the kernel on the GPU)."""
    import json
    from pathlib import Path
    from oracle import text_ref
    nodes, tools, eng, state, ref_config = rig
    monkeypatch.setattr(tools, "evaluate", text_ref.evaluate)
    conf = iter(range(30, 84, 6))

    def fake_llm(system_prompt, user_msg, json_schema=None, **kw):
        title = (json_schema or {}).get("title", "")
        if title == "CriticResult":
            return {"overall_confidence": next(conf), "segments": [], "verdict": "needs_reocr", "reasoning": "fake"}
        if title == "ArbitratorResult":
            return {"final_text": "kalo miren tusha veon darel", "decisions": [], "confidence": 55, "uncertain_segments": []}
        return {"corrected_text": "edited", "changes": [], "unresolved": []}

    import ocr_agent.agents as agents
    monkeypatch.setattr(agents, "call_llm_json", fake_llm)
    lg, lgg = _langgraph_shim()
    monkeypatch.setitem(sys.modules, "langgraph", lg)
    monkeypatch.setitem(sys.modules, "langgraph.graph", lgg)
    sys.modules.pop("ocr_agent.graph", None)
    monkeypatch.setattr(ref_config, "PREPROCESSING_STRATEGIES", S, raising=False)
    monkeypatch.setattr(ref_config, "AGREEMENT_THRESHOLD", 101, raising=False)
    import ocr_agent.transcribe as tr
    assert tr.__file__.startswith(REF)
    gt = tmp_path / "gt.md"
    gt.write_text("# page\n\n## Ground Truth\nkalo miren tusha veon darol\n", encoding="utf-8")
    out_dir = tmp_path / "out"
    path = tr.transcribe_single(Path(state["image_path"]), out_dir, ground_truth_path=gt, max_iterations=10)
    capsys.readouterr()
    assert path.read_text(encoding="utf-8") == "kalo miren tusha veon darel"
    ev = json.loads((out_dir / "page_eval.json").read_text(encoding="utf-8"))
    want = text_ref.tier1_metrics("kalo miren tusha veon darol", "kalo miren tusha veon darel")
    assert ev["tier1_raw_vs_gt"] == want and 0 < want["cer"] < 0.1
    assert ev["pipeline_status"] == "completed" and ev["iterations"] == 3
    assert (out_dir / "page_trace.json").exists() and (out_dir / "page_trace_summary.txt").exists()
    assert eng.calls == [5]


def test_eval_final_batch_mode_unmodified(rig, monkeypatch, tmp_path, capsys, synth):
    """eval_final.main() in directory mode (eval_final.py:94-134), unmodified: every file goes through the `evaluate` /
    `parse_ground_truth` this package exports under `ocr_agent.tools`."""
    import json
    from oracle import text_ref
    nodes, tools, eng, state, _ = rig
    monkeypatch.setattr(tools, "evaluate", text_ref.evaluate)
    res, gtd = tmp_path / "results", tmp_path / "gt"
    res.mkdir()
    gtd.mkdir()
    want = {}
    for i in range(3):
        gt = synth.text(20 + i, 60)
        ocr = synth.corrupt(gt, i, 0.04 * (i + 1))
        (res / f"p{i}_transcription.txt").write_text(ocr, encoding="utf-8")
        (gtd / f"p{i}.md").write_text(f"notes\n## Ground Truth\n{gt}\n", encoding="utf-8")
        want[f"p{i}"] = text_ref.tier1_metrics(gt.strip(), ocr)
    (res / "p9_transcription.txt").write_text("no ground truth for this one", encoding="utf-8")
    import ocr_agent.eval_final as ef
    assert ef.__file__.startswith(REF)
    out = tmp_path / "eval.json"
    monkeypatch.setattr(sys, "argv", ["eval_final", str(res), "--ground-truth-dir", str(gtd), "--output", str(out)])
    ef.main()
    text = capsys.readouterr().out
    got = json.loads(out.read_text(encoding="utf-8"))
    assert len(got) == 4 and "Batch Summary (3 files with GT)" in text
    for r in got:
        stem = os.path.basename(r["file"])[: -len("_transcription.txt")]
        if stem in want:
            assert r["tier1_raw_vs_gt"] == want[stem], stem
        else:
            assert "tier1_raw_vs_gt" not in r


def _fake_llm_factory():
    conf = iter(range(30, 10_000, 6))

    def fake_llm(system_prompt, user_msg, json_schema=None, **kw):
        title = (json_schema or {}).get("title", "")
        if title == "CriticResult":
            return {"overall_confidence": 90, "segments": [], "verdict": "accept", "reasoning": "fake"}
        if title == "ArbitratorResult":
            return {"final_text": "arbitrated", "decisions": [], "confidence": next(conf), "uncertain_segments": []}
        return {"corrected_text": "edited", "changes": [], "unresolved": []}

    return fake_llm


def test_transcribe_folder_runs_unmodified_transcribe_single_on_primed_batches(rig, monkeypatch, tmp_path, capsys, synth):
    """f1 (transcribe.py:185-210): the folder driver primes the cache for `pages_per_batch` pages with ONE batched read,
    then the reference's unmodified `transcribe_single` runs per page and every read of it is a cache hit."""
    import json
    from oracle import text_ref
    from handwritten_ocr_b200 import folder
    nodes, tools, eng, state, ref_config = rig
    eng.max_batch = 16
    monkeypatch.setattr(tools, "evaluate", text_ref.evaluate)
    import ocr_agent.agents as agents
    monkeypatch.setattr(agents, "call_llm_json", _fake_llm_factory())
    lg, lgg = _langgraph_shim()
    monkeypatch.setitem(sys.modules, "langgraph", lg)
    monkeypatch.setitem(sys.modules, "langgraph.graph", lgg)
    sys.modules.pop("ocr_agent.graph", None)
    monkeypatch.setattr(ref_config, "PREPROCESSING_STRATEGIES", S, raising=False)
    monkeypatch.setattr(ref_config, "AGREEMENT_THRESHOLD", 101, raising=False)
    src, gtd, out = tmp_path / "in", tmp_path / "gt", tmp_path / "out"
    src.mkdir()
    gtd.mkdir()
    for i in range(5):
        Image.fromarray(synth.page(100 + i, 200, 120)).save(src / f"p{i}.png")
        (gtd / f"p{i}.md").write_text(f"## Ground Truth\n{synth.text(i, 30)}\n", encoding="utf-8")
    (src / "notes.txt").write_text("not an image")
    res = folder.transcribe_folder(src, out, gtd, pages_per_batch=2, tools=tools, max_iterations=4)
    capsys.readouterr()
    # 5 pages x 5 strategies read in batches of 2 pages: 10 + 10 + 5 sequences, nothing else
    assert eng.calls == [10, 10, 5], eng.calls
    assert [p.name for p in res] == [f"p{i}_transcription.txt" for i in range(5)]
    for i in range(5):
        ev = json.loads((out / f"p{i}_eval.json").read_text(encoding="utf-8"))
        text = (out / f"p{i}_transcription.txt").read_text(encoding="utf-8")
        assert ev["tier1_raw_vs_gt"] == text_ref.tier1_metrics(synth.text(i, 30).strip(), text)
        assert ev["pipeline_status"] == "completed"
    assert not tools._pages, "the driver forgets a page once its per-page function returned"


def test_eval_folder_json_equals_eval_final_main(rig, monkeypatch, tmp_path, capsys, synth):
    """f2 (eval_final.py:94-134): one batched metrics call for the whole folder, JSON equal to `eval_final.main --output`."""
    import json
    from oracle import text_ref
    from handwritten_ocr_b200 import folder
    nodes, tools, eng, state, _ = rig
    monkeypatch.setattr(tools, "evaluate", text_ref.evaluate)
    res, gtd = tmp_path / "results", tmp_path / "gt"
    res.mkdir()
    gtd.mkdir()
    for i in range(9):
        gt = synth.text(40 + i, 50)
        (res / f"q{i}_transcription.txt").write_text(synth.corrupt(gt, i, 0.03 * (i + 1)), encoding="utf-8")
        if i != 4:
            ext = ".md" if i % 2 else ".txt"
            (gtd / f"q{i}{ext}").write_text(f"## Ground Truth\n{gt}\n" if ext == ".md" else gt, encoding="utf-8")
    import ocr_agent.eval_final as ef
    ref_out, our_out = tmp_path / "ref.json", tmp_path / "ours.json"
    monkeypatch.setattr(sys, "argv", ["eval_final", str(res), "--ground-truth-dir", str(gtd), "--output", str(ref_out)])
    ef.main()
    capsys.readouterr()
    calls = []

    class BatchOracle:                                          # stands where textops stands (the GPU kernel on the box)
        @staticmethod
        def tier1_metrics_batch(items, lower=False):
            calls.append(len(items))
            return [text_ref.tier1_metrics(g, o, lower) for g, o in items]

    got = folder.eval_folder(res, gtd, our_out, tools=tools, textops=BatchOracle)
    assert calls == [8], "all files with a ground truth go through ONE batched call"
    assert json.loads(our_out.read_text(encoding="utf-8")) == json.loads(ref_out.read_text(encoding="utf-8"))
    assert got == json.loads(ref_out.read_text(encoding="utf-8"))
    assert our_out.read_text(encoding="utf-8") == ref_out.read_text(encoding="utf-8")


def test_cache_signature_lru_and_resave(rig, tmp_path, synth, capsys):
    """the cache is keyed on the file signature, bounded (LRU), and a removed temp file is written again."""
    nodes, tools, eng, state, _ = rig
    img = state["image_path"]
    p0 = tools.preprocess_image(img, S[0])
    t0 = tools.run_ocr(p0)
    assert eng.calls == [5]
    os.unlink(p0)
    assert tools.preprocess_image(img, S[0]) == p0 and os.path.isfile(p0)       # re-saved from the cached page
    assert tools.run_ocr(p0) == t0 and eng.calls == [5]
    # the file changes on disk: the stale page and its texts are dropped
    Image.fromarray(synth.page(78, 200, 120)).save(img)
    os.utime(img, ns=(1, 1))
    p1 = tools.preprocess_image(img, S[0])
    assert p1 != p0 and not os.path.exists(p0)
    tools.run_ocr(p1)
    assert eng.calls == [5, 5]
    # bounded: with cache_pages=2 a third original evicts the oldest
    tools.configure(cache_pages=2)
    others = []
    for i in range(2):
        q = str(tmp_path / f"other{i}.png")
        Image.fromarray(synth.page(300 + i, 200, 120)).save(q)
        others.append(q)
        tools.preprocess_image(q, S[1])
    assert list(tools._pages) == others and not os.path.exists(p1)
    capsys.readouterr()


def test_speculative_failure_does_not_fail_the_requested_strategy(rig, monkeypatch, capsys):
    nodes, tools, eng, state, _ = rig
    from handwritten_ocr_b200 import preprocess
    real = preprocess.apply_strategy

    def flaky(x, s):
        if "remove_lines" in s:
            raise RuntimeError("workspace too small")
        return real(x, s)

    monkeypatch.setattr(preprocess, "apply_strategy", flaky)
    p = tools.preprocess_image(state["image_path"], S[0])
    assert os.path.isfile(p)
    assert "speculative deskew+remove_lines+high_contrast skipped" in capsys.readouterr().err
    with pytest.raises(RuntimeError, match="workspace too small"):
        tools.preprocess_image(state["image_path"], S[4])
