"""SURVEY §8 row f4, host side: what `from_pretrained` reads next to the weights (tools.py:700-709) comes from a local
checkpoint directory -- dimensions from config.json (both layouts), EOS ids from generation_config.json (greedy only),
tokenizer files if present.  CPU only."""
import json
import os

import numpy as np
import pytest


def test_config_json_roundtrip_both_layouts(pkg, tmp_path):
    from handwritten_ocr_b200.vlm_config import VLMConfig
    for cfg in (VLMConfig.tiny(), VLMConfig.olmocr_7b(), VLMConfig.qwen72b()):
        d = tmp_path / cfg.name
        cfg.to_hf().save_pretrained(str(d))                       # nested layout (transformers 5.x)
        back = VLMConfig.from_pretrained_dir(str(d))
        assert back.text == cfg.text and back.vision == cfg.vision
    # the flat layout the published Qwen2.5-VL-7B checkpoints ship
    flat = {"model_type": "qwen2_5_vl", "hidden_size": 3584, "intermediate_size": 18944, "num_hidden_layers": 28,
            "num_attention_heads": 28, "num_key_value_heads": 4, "rms_norm_eps": 1e-6, "rope_theta": 1000000.0,
            "rope_scaling": {"type": "mrope", "mrope_section": [16, 24, 24]}, "vocab_size": 152064,
            "tie_word_embeddings": False,
            "vision_config": {"depth": 32, "hidden_size": 1280, "intermediate_size": 3420, "num_heads": 16,
                              "out_hidden_size": 3584, "fullatt_block_indexes": [7, 15, 23, 31], "window_size": 112,
                              "patch_size": 14, "spatial_merge_size": 2, "temporal_patch_size": 2, "tokens_per_second": 2,
                              "in_chans": 3}}
    got = VLMConfig.from_hf_dict(flat)
    ref = VLMConfig.olmocr_7b()
    assert got.text == ref.text and got.vision == ref.vision
    for bad in ({"model_type": "llava"}, {"tie_word_embeddings": True}, {"num_attention_heads": 16}):
        with pytest.raises(ValueError):
            VLMConfig.from_hf_dict({**flat, **bad})


def test_generation_config_greedy_only(pkg, tmp_path):
    from handwritten_ocr_b200 import tools
    from handwritten_ocr_b200.vlm_config import EOS, VLMConfig, greedy_generation_params
    d = tmp_path / "ckpt"
    VLMConfig.tiny().to_hf().save_pretrained(str(d))
    assert greedy_generation_params(str(d)) == {"eos_token_ids": [EOS]}         # no generation_config.json
    (d / "generation_config.json").write_text(json.dumps({"do_sample": False, "eos_token_id": [151645, 151643]}))
    cfg, tok, gen = tools.checkpoint_metadata(str(d))
    assert cfg.text == VLMConfig.tiny().text and tok is None and gen["eos_token_ids"] == [151645, 151643]
    (d / "generation_config.json").write_text(json.dumps({"do_sample": True, "temperature": 0.1, "repetition_penalty": 1.05,
                                                          "eos_token_id": 151645}))
    with pytest.raises(NotImplementedError, match="do_sample=true, repetition_penalty=1.05"):
        tools.checkpoint_metadata(str(d))
    assert tools.checkpoint_metadata(str(d), force_greedy=True)[2] == {"eos_token_ids": [EOS]}
    (d / "generation_config.json").write_text(json.dumps({"eos_token_id": 151643}))
    with pytest.raises(ValueError):
        tools.checkpoint_metadata(str(d))
    with pytest.raises(FileNotFoundError):
        tools.checkpoint_metadata(str(tmp_path / "missing"))
    # without a checkpoint: configured dimensions, synthetic tokenizer
    cfg, tok, gen = tools.checkpoint_metadata(None)
    assert cfg.text == VLMConfig.olmocr_7b().text and tok is None and gen == {"eos_token_ids": [EOS]}


def test_extra_eos_truncation_rule():
    """The host-side cut for further EOS ids (engine.read_batch): a row that emits one is padded with <|im_end|> from the
    next position on, and the batch is trimmed where the last row finishes -- what HF's stopping criteria produce."""
    EOS, EXTRA = 151645, (151643,)
    toks = np.array([[5, 6, 151643, 9, 9, 9], [7, 8, 9, 151645, 151645, 151645], [1, 2, 3, 4, 151643, 7]])
    for i in range(len(toks)):
        hit = np.isin(toks[i], EXTRA)
        if hit.any():
            toks[i, int(hit.argmax()) + 1:] = EOS
    is_eos = (toks == EOS) | np.isin(toks, EXTRA)
    first = np.where(is_eos.any(1), is_eos.argmax(1), toks.shape[1] - 1)
    keep = int(first.max()) + 1
    assert keep == 5 and toks[:, :keep].tolist() == [[5, 6, 151643, EOS, EOS], [7, 8, 9, EOS, EOS], [1, 2, 3, 4, 151643]]
