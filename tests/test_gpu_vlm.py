"""GPU parity of the VLM read (vision tower, prefill, batched greedy decode) against the HF
transformers implementation the reference calls (tools.py:705-709,764-765), on the same GPU, the same
random-init weights and the same synthetic pages.

Tolerances (bf16 path, fp32 accumulation, different summation order than cuBLAS/SDPA):
  * vision embeddings / prefill logits: max |delta| <= 4 % of the oracle's max |value| and cosine >= 0.999
  * greedy tokens: identical up to the first step whose ORACLE top-1/top-2 logit margin is below
    LOGIT_TOL; a divergence at a step with a larger margin fails (SURVEY §7 hard part 1, protocol v).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
BF = torch.bfloat16
LOGIT_TOL_REL = 0.04


@pytest.fixture(scope="module")
def ctx(pkg, synth):
    from transformers import Qwen2_5_VLForConditionalGeneration
    from handwritten_ocr_b200 import vlm, engine, preprocess
    from handwritten_ocr_b200.vlm_config import VLMConfig
    dev = torch.device("cuda")
    out = {}
    for name, kw in (("default", {}), ("peaked", {"lm_head_std": 0.5})):
        cfg = VLMConfig.tiny()
        sd = vlm.random_state_dict(cfg, dev, seed=0, **kw)
        with torch.device("cuda"):
            hf = Qwen2_5_VLForConditionalGeneration._from_config(cfg.to_hf(), dtype=BF).eval()
        hf.load_state_dict(sd, strict=True)
        w = vlm.VLMWeights.from_state_dict(cfg, sd)
        eng = engine.OcrEngine(w, max_batch=4, max_new_tokens=64, max_prompt=400)
        out[name] = (cfg, hf, eng)
    out["pages"] = [synth.page(100 + i, 504, 392) for i in range(3)]
    out["pp"] = preprocess
    return out


def hf_inputs(eng, pp, page):
    x = pp.to_device(page)
    pv, (gh, gw) = pp.pixel_values(x, dtype=torch.float32)
    plan = eng._plan((gh, gw), 1)
    ids, pos3, delta = eng.build_inputs(plan, "Extract and return all the text from this handwritten document.")
    t = torch.from_numpy(ids.astype(np.int64))[None].cuda()
    return dict(input_ids=t, attention_mask=torch.ones_like(t), pixel_values=pv,
                image_grid_thw=torch.tensor([[1, gh, gw]], device="cuda"),
                mm_token_type_ids=(t == 151655).int()), plan


def rel_err(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / b.abs().max()).item(), torch.nn.functional.cosine_similarity(a.flatten(), b.flatten(), dim=0).item()


def test_vision_tower_matches_hf(ctx):
    cfg, hf, eng = ctx["default"]
    pp = ctx["pp"]
    inp, plan = hf_inputs(eng, pp, ctx["pages"][0])
    with torch.no_grad():
        want = hf.model.visual(inp["pixel_values"].to(BF), grid_thw=inp["image_grid_thw"]).pooler_output
    merged, plan = eng.encode_images(pp.to_device(ctx["pages"][0]))
    got = torch.empty_like(merged)
    got[plan.group_perm.long()] = merged
    e, c = rel_err(got, want)
    print(f"vision tower: max rel err {e:.4f}, cosine {c:.6f}")
    assert e < 0.04 and c > 0.999


def test_prefill_logits_match_hf(ctx):
    cfg, hf, eng = ctx["default"]
    pp = ctx["pp"]
    inp, _ = hf_inputs(eng, pp, ctx["pages"][1])
    with torch.no_grad():
        want = hf(**inp).logits[0, -1]
    _, dbg = eng.read_batch(pp.to_device(ctx["pages"][1]), max_new_tokens=1, return_debug=True)
    got = dbg["prefill_logits"][0]
    e, c = rel_err(got, want)
    print(f"prefill logits: max rel err {e:.4f}, cosine {c:.6f}, oracle max {want.abs().max().item():.3f}")
    assert e < LOGIT_TOL_REL and c > 0.999


@pytest.mark.parametrize("which", ["default", "peaked"])
def test_greedy_tokens_vs_hf_generate(ctx, which):
    cfg, hf, eng = ctx[which]
    pp = ctx["pp"]
    n_new = 32
    flips = 0
    for page in ctx["pages"][:2]:
        inp, _ = hf_inputs(eng, pp, page)
        with torch.no_grad():
            gen = hf.generate(**inp, max_new_tokens=n_new, do_sample=False, output_scores=True,
                              return_dict_in_generate=True)
        want = gen.sequences[0, inp["input_ids"].shape[1]:].tolist()
        got = eng.read_batch(pp.to_device(page), max_new_tokens=n_new)[0]
        first_diff = next((i for i, (a, b) in enumerate(zip(got, want)) if a != b), None)
        if first_diff is None:
            assert len(got) == len(want)
            continue
        sc = gen.scores[first_diff][0].float()
        top2 = torch.topk(sc, 2).values
        margin = (top2[0] - top2[1]).item()
        tol = LOGIT_TOL_REL * sc.abs().max().item()
        print(f"[{which}] first divergence at step {first_diff}: oracle margin {margin:.5f}, tolerance {tol:.5f}")
        assert margin <= tol, f"token flip at step {first_diff} with oracle margin {margin} > tolerance {tol}"
        flips += 1
    print(f"[{which}] pages with a (below-tolerance) divergence: {flips} of 2")


def test_batch_invariance_and_graph(ctx):
    cfg, hf, eng = ctx["default"]
    pp = ctx["pp"]
    pages = ctx["pages"]
    batch = eng.read_batch(pp.to_device(pages), max_new_tokens=24)
    singles = [eng.read_batch(pp.to_device(p), max_new_tokens=24)[0] for p in pages]
    assert batch == singles
    nograph = eng.read_batch(pp.to_device(pages), max_new_tokens=24, use_graph=False)
    assert nograph == batch
    again = eng.read_batch(pp.to_device(pages), max_new_tokens=24)
    assert again == batch


@pytest.mark.parametrize("fuse_attention", [True, False])
def test_fused_step_plan_reads_the_same_text(ctx, monkeypatch, fuse_attention):
    """The decode step as ONE persistent plan launch (csrc/chain.cu; opt-in, OCRB_CHAIN_MAX_B) against the default
    one-launch-per-op step: the narrow linears add their split-K partials in a different (fixed) order, so tokens are
    compared up to the first step where the default path's own top-2 margin is a rounding tie; the fused path must be
    batch-invariant bit for bit like the default one."""
    from handwritten_ocr_b200 import vlm
    cfg, hf, eng = ctx["default"]
    pp, pages = ctx["pp"], ctx["pages"]
    base = eng.read_batch(pp.to_device(pages), max_new_tokens=24)
    monkeypatch.setattr(vlm, "CHAIN_MAX_B", 128)
    monkeypatch.setattr(vlm, "CHAIN_FUSE_ATTN", fuse_attention)
    eng._states.clear()                  # cached decode states carry a CUDA graph of the default step
    try:
        before = vlm._lib.launch_count()
        fused = eng.read_batch(pp.to_device(pages), max_new_tokens=24)
        singles = [eng.read_batch(pp.to_device(p), max_new_tokens=24)[0] for p in pages]
        assert fused == singles, "fused step is not batch-invariant"
        nograph = eng.read_batch(pp.to_device(pages), max_new_tokens=24, use_graph=False)
        assert nograph == fused
        assert vlm._lib.launch_count() > before
    finally:
        eng._states.clear()
    for a, b in zip(base, fused):
        same = next((i for i, (x, y) in enumerate(zip(a, b)) if x != y), len(a))
        assert same >= 4, f"fused step diverges from the default step at token {same}"


def test_eos_stops_and_pads(ctx):
    """EOS semantics of HF's greedy loop (generation/utils.py:2743-2806): a sequence stops at its first <|im_end|>, a
    finished row of a batch is padded with pad = eos while the others continue, the batch ends when every row is done.
    The EOS is forced with a wide margin: its lm_head row becomes 4x the row of the token HF picks at step k, so at
    step k the EOS logit is 4x the winning logit (rounding cannot flip that) -- our tokens must EQUAL HF's."""
    from handwritten_ocr_b200.vlm_config import EOS
    cfg, hf, eng = ctx["peaked"]
    pp = ctx["pp"]
    pages = ctx["pages"][:2]
    inputs = [hf_inputs(eng, pp, pg)[0] for pg in pages]
    n_new = 16
    with torch.no_grad():
        free = [hf.generate(**inp, max_new_tokens=n_new, do_sample=False)[0, inp["input_ids"].shape[1]:].tolist() for inp in inputs]
    k = 6
    tok_k = free[0][k]
    assert tok_k != EOS
    row_ours, row_hf = eng.w.lm_head[EOS].clone(), hf.lm_head.weight.data[EOS].clone()
    try:
        eng.w.lm_head[EOS] = 4.0 * eng.w.lm_head[tok_k]
        hf.lm_head.weight.data[EOS] = 4.0 * hf.lm_head.weight.data[tok_k]
        with torch.no_grad():
            want = [hf.generate(**inp, max_new_tokens=n_new, do_sample=False)[0, inp["input_ids"].shape[1]:].tolist() for inp in inputs]
        assert want[0][-1] == EOS and len(want[0]) <= k + 1, "the forced EOS must end page 0 by step k"
        # one page alone: exactly HF's tokens, EOS included, nothing after it
        got0 = eng.read_batch(pp.to_device(pages[0]), max_new_tokens=n_new)[0]
        assert got0 == want[0]
        # both pages in one batch: each row = HF's tokens, then eos padding up to the longest row of the batch
        got = eng.read_batch(pp.to_device(pages), max_new_tokens=n_new)
        longest = max(len(w_) for w_ in want)
        for g_, w_ in zip(got, want):
            assert len(g_) == longest
            assert g_[: len(w_)] == w_ and all(t_ == EOS for t_ in g_[len(w_):])
    finally:
        eng.w.lm_head[EOS] = row_ours
        hf.lm_head.weight.data[EOS] = row_hf
