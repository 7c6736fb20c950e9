"""Parity at the BASELINE dimensions: the 7B-class config, full-size synthetic pages (three 1024x768 and one
768x1024), against HF transformers (`Qwen2_5_VLForConditionalGeneration`, the class the reference's
AutoModelForImageTextToText resolves to -- tools.py:705-709) on the same GPU with the same random-init state dict, for
BOTH initialisations: HF's default (`initializer_range` 0.02 everywhere) and a peaked lm_head (std 0.5: large top-1 margins).

Protocol (SURVEY §7 hard part 1, v), 512 new tokens per page (BASELINE configs[0..3]):
  1. HF `generate(max_new_tokens=512, do_sample=False, output_logits=True)`: token ids and the logits of every step.
  2. TEACHER-FORCED pass of this repository's engine: HF's token ids are fed as inputs, so every step is compared on the
     same prefix and an early near-tie cannot hide later steps.  Per step: max |logit - HF logit| over the 152 064 entries.
     Stated bf16 tolerance: LOGIT_TOL_REL of the largest |logit| of that step's HF logits (60 bf16 layers deep,
     different summation order than cuBLAS / SDPA).
  3. A step "flips" when this engine's argmax differs from HF's token.  Every flip must sit at a step whose HF top-1/top-2
     margin is at most twice THAT step's measured max error (anything else would be an error in the engine, not rounding);
     the flip rate is printed and bounded.
  4. Free-running greedy decode (CUDA graph, what `run_ocr` executes): identical to HF up to the first step whose HF
     margin is at most twice the largest per-step error measured in 2.; a divergence at a larger margin fails.
  5. Batch invariance at full size: a candidate decoded inside a batch of 63 gives the tokens it gives alone.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
BF = torch.bfloat16
LOGIT_TOL_REL = 0.06
N_NEW = 512
PROMPT = "Extract and return all the text from this handwritten document."


def _hf_inputs(eng, preprocess, cand):
    pv, (gh, gw) = preprocess.pixel_values(cand, dtype=torch.float32)
    plan = eng._plan((gh, gw), 1)
    ids, _, _ = eng.build_inputs(plan, PROMPT)
    t = torch.from_numpy(ids.astype(np.int64))[None].cuda()
    return dict(input_ids=t, attention_mask=torch.ones_like(t), pixel_values=pv,
                image_grid_thw=torch.tensor([[1, gh, gw]], device="cuda"), mm_token_type_ids=(t == 151655).int()), (gh, gw)


@pytest.mark.parametrize("init", ["peaked", "hf_default"])
def test_7b_read_matches_hf_512_steps(pkg, synth, init):
    from transformers import Qwen2_5_VLForConditionalGeneration, initialization
    from handwritten_ocr_b200 import engine, preprocess, vlm
    from handwritten_ocr_b200.vlm_config import VLMConfig, EOS
    if torch.cuda.get_device_properties(0).total_memory < 100e9:
        pytest.skip("needs ~70 GB of device memory")
    dev = torch.device("cuda")
    cfg = VLMConfig.olmocr_7b()
    sd = vlm.random_state_dict(cfg, dev, seed=0, lm_head_std=0.5 if init == "peaked" else None)
    with torch.device("cuda"), initialization.no_init_weights():
        hf = Qwen2_5_VLForConditionalGeneration._from_config(cfg.to_hf(), dtype=BF).eval()
    hf.load_state_dict(sd, strict=True)
    w = vlm.VLMWeights.from_state_dict(cfg, sd)
    del sd
    eng = engine.OcrEngine(w, max_batch=63, max_new_tokens=N_NEW, max_prompt=1600)
    # four pages: three landscape (one per initial strategy) and one portrait
    pages = [(synth.page(0), ["high_contrast", "binarize"]), (synth.page(1), ["deskew", "high_contrast", "binarize"]),
             (synth.page(2), ["deskew", "high_contrast", "sharpen"]), (synth.page(3, 768, 1024), ["high_contrast", "binarize"])]
    cands = [preprocess.apply_strategy(preprocess.to_device(pg), st) for pg, st in pages]
    assert cands[0].shape[1:] == (768, 1024) and cands[3].shape[1:] == (1024, 768)

    # ---- 1. HF generate: tokens + per-step logits ----
    hf_tokens, hf_logits = [], []
    for cand in cands:
        inp, (gh, gw) = _hf_inputs(eng, preprocess, cand)
        assert gh * gw == 3996                                                   # 3 996 patches -> 999 tokens
        with torch.no_grad():
            gen = hf.generate(**inp, max_new_tokens=N_NEW, min_new_tokens=N_NEW, do_sample=False, output_logits=True,
                              return_dict_in_generate=True)
        hf_tokens.append(gen.sequences[0, inp["input_ids"].shape[1]:].to(torch.int32))
        hf_logits.append(torch.stack([l[0] for l in gen.logits]).to(BF))         # [512, V]: HF's logits are bf16 values
        assert hf_tokens[-1].shape[0] == N_NEW
        del gen

    # ---- 1b. the noise floor: HF against HF.  One causal forward over [prompt + generated tokens] evaluates the same
    # network on the same prefixes as the step-by-step `generate` did, only with other GEMM shapes (a 1 548-row prefill-style
    # pass instead of 512 single-row steps).  Where those two HF evaluations disagree on the argmax, no implementation can
    # be "token-identical to HF": this is the flip rate rounding alone produces with these weights.
    hf_self_flips, hf_self_err = 0, 0.0
    for pi, cand in enumerate(cands):
        inp, _ = _hf_inputs(eng, preprocess, cand)
        full = torch.cat([inp["input_ids"], hf_tokens[pi][None, :-1].long()], 1)
        kw = dict(inp, input_ids=full, attention_mask=torch.ones_like(full), mm_token_type_ids=(full == 151655).int())
        with torch.no_grad():
            one_shot = hf(**kw, logits_to_keep=N_NEW).logits[0]                    # positions T-1 .. T+510
        ref = hf_logits[pi].float()
        hf_self_flips += int((one_shot.float().argmax(-1) != ref.argmax(-1)).sum())
        hf_self_err = max(hf_self_err, ((one_shot.float() - ref).abs().amax(-1) / ref.abs().amax(-1)).max().item())
        del one_shot
    print(f"7B {init}: HF one-shot forward vs HF generate on the same prefixes: {hf_self_flips} argmax flips of {4 * N_NEW} "
          f"({100 * hf_self_flips / (4 * N_NEW):.2f} %), max |dlogit| {hf_self_err:.4f} of the step's max |logit|")

    # ---- 2./3. teacher-forced pass: per-step logit error, flips ----
    stats = {"max_rel": 0.0, "max_abs": 0.0, "flips": 0, "steps": 0, "bad_flips": [], "worst_step": None,
             "max_flip_margin_rel": 0.0, "sum_rel": 0.0}

    def run_group(idx):
        forced = torch.stack([hf_tokens[i] for i in idx])

        def on_step(i, logits):
            for r, pi in enumerate(idx):
                ref = hf_logits[pi][i].float()
                mine = logits[r].float()
                err = (mine - ref).abs().max().item()
                scale = ref.abs().max().item()
                stats["steps"] += 1
                if err / scale > stats["max_rel"]:
                    stats["max_rel"], stats["worst_step"] = err / scale, (pi, i)
                stats["max_abs"] = max(stats["max_abs"], err)
                tok = int(torch.argmax(mine))                 # first index on ties, like torch.argmax in HF's loop
                if tok != int(torch.argmax(ref)):             # (== HF's token unless min_new_tokens suppressed an EOS)
                    top2 = torch.topk(ref, 2).values
                    margin = (top2[0] - top2[1]).item()
                    stats["flips"] += 1
                    stats["max_flip_margin_rel"] = max(stats["max_flip_margin_rel"], margin / scale)
                    if margin > 2.0 * err:
                        stats["bad_flips"].append((pi, i, margin, err))
                stats["sum_rel"] += err / scale

        eng.teacher_forced_logits(torch.cat([cands[i] for i in idx]), forced, on_step, prompt=PROMPT)

    run_group([0, 1, 2])        # same page shape: one batch of three
    run_group([3])              # portrait page
    flip_rate = stats["flips"] / stats["steps"]
    print(f"7B {init}: teacher-forced {stats['steps']} steps over 4 pages: max |dlogit| {stats['max_abs']:.4f} = "
          f"{stats['max_rel']:.4f} of the step's max |logit| (page, step {stats['worst_step']}), mean per-step max "
          f"{stats['sum_rel'] / stats['steps']:.4f}; flips {stats['flips']} ({100 * flip_rate:.2f} %), largest HF margin at a "
          f"flip {stats['max_flip_margin_rel']:.4f} of max |logit|, flips beyond 2x the step error: {len(stats['bad_flips'])}")
    assert stats["steps"] == 4 * N_NEW
    assert stats["max_rel"] < LOGIT_TOL_REL, f"logit error {stats['max_rel']} of max |logit| exceeds the stated bf16 tolerance"
    assert not stats["bad_flips"], f"argmax differs at steps whose HF margin exceeds twice the measured error: {stats['bad_flips'][:5]}"
    # every flip sits inside the stated tolerance band: with random-init weights the top-2 logits of many steps are closer
    # than two bf16 evaluations of the same network can resolve (the flip RATE is a property of the weights, reported only)
    assert stats["max_flip_margin_rel"] < LOGIT_TOL_REL, "a flip at a margin wider than the stated logit tolerance"
    assert flip_rate < 0.5
    # ... and it is of the size of HF's own disagreement with itself (a real defect would add to it, not vanish in it)
    assert stats["flips"] <= 3 * hf_self_flips + 0.02 * stats["steps"], \
        f"{stats['flips']} flips against {hf_self_flips} between two HF evaluations of the same steps"

    # ---- 4. free-running greedy decode (CUDA graph) vs HF ----
    free = eng.read_batch(torch.cat(cands[:3]), prompt=PROMPT, max_new_tokens=N_NEW) + \
        eng.read_batch(cands[3], prompt=PROMPT, max_new_tokens=N_NEW)
    identical = 0
    for pi, got in enumerate(free):
        want = hf_tokens[pi].tolist()
        first = next((i for i, (a, b) in enumerate(zip(got, want)) if a != b), None)
        if first is None:
            identical += 1
            continue
        top2 = torch.topk(hf_logits[pi][first].float(), 2).values
        margin = (top2[0] - top2[1]).item()
        print(f"7B {init}: page {pi} free-running decode leaves HF at step {first} of {N_NEW}: HF margin {margin:.4f}, "
              f"2 x max step error {2 * stats['max_abs']:.4f}")
        assert margin <= 2.0 * stats["max_abs"], f"page {pi}: token flip at step {first} with HF margin {margin}"
    print(f"7B {init}: {identical} of 4 pages token-identical to HF generate over all {N_NEW} steps")

    # ---- 5. batch invariance at full size: batch of 63 ----
    alone = eng.read_batch(cands[0], prompt=PROMPT, max_new_tokens=96)[0]
    big = torch.cat([cands[i % 3] if i else cands[0] for i in range(63)])
    in_batch = eng.read_batch(big, prompt=PROMPT, max_new_tokens=96)
    assert in_batch[0] == alone and in_batch[3] == alone and in_batch[60] == alone, \
        "a candidate read in a batch of 63 must give the tokens it gives alone"

    if init != "peaked":
        return
    # ---- context for the tolerance: how far is HF's own bf16 result from the same model evaluated in fp32? ----
    inp, _ = _hf_inputs(eng, preprocess, cands[0])
    toks, dbg = eng.read_batch(cands[0], prompt=PROMPT, max_new_tokens=2, return_debug=True)
    mine = dbg["prefill_logits"][0].float()
    l0 = hf_logits[0][0].float()
    del eng, w
    torch.cuda.empty_cache()
    hf.float()
    with torch.no_grad():
        l32 = hf(**inp).logits[0, -1].float()

    def rms_rel(a, b):
        return ((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt()).item()

    e_hf, e_us = rms_rel(l0, l32), rms_rel(mine, l32)
    cos = torch.nn.functional.cosine_similarity(mine, l0, dim=0).item()
    print(f"7B prefill logits vs the fp32 model: HF bf16 rms rel err {e_hf:.4f}, this repo {e_us:.4f}; cosine(ours, HF) {cos:.6f}")
    assert cos > 0.999
    assert e_us < 2.0 * e_hf + 0.01, "our bf16 path is much further from the fp32 model than HF's bf16 path"
