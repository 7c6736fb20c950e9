"""Dimensions of "the configured VLM" (reference: ocr_agent/config.py:16 names allenai/olmOCR-2-7B-1025,
a Qwen2.5-VL-7B fine-tune; the checkpoint is not available offline so the dimensions are pinned here --
SURVEY.md Appendix A.9) and the prompt / position-id builders that HF's processor and
`get_rope_index` would produce for one page.
"""
from __future__ import annotations

import re
import zlib
from dataclasses import dataclass, field

import numpy as np

# Qwen2.5-VL special token ids (HF configuration_qwen2_5_vl.py:123-124,182-185)
ENDOFTEXT, IM_START, IM_END = 151643, 151644, 151645
VISION_START, VISION_END, IMAGE_PAD, VIDEO_PAD = 151652, 151653, 151655, 151656
EOS = IM_END


@dataclass
class VisionCfg:
    depth: int = 32
    hidden: int = 1280
    heads: int = 16
    intermediate: int = 3420
    out_hidden: int = 3584
    patch: int = 14
    temporal_patch: int = 2
    merge: int = 2
    window: int = 112
    fullatt_blocks: tuple = (7, 15, 23, 31)
    tokens_per_second: int = 2
    in_channels: int = 3

    @property
    def head_dim(self):
        return self.hidden // self.heads

    @property
    def intermediate_padded(self):
        """SwiGLU packs [gate64|up64] row tiles and the down-proj K must be a multiple of 64."""
        return (self.intermediate + 63) // 64 * 64

    @property
    def patch_dim(self):
        return self.in_channels * self.temporal_patch * self.patch * self.patch


@dataclass
class TextCfg:
    hidden: int = 3584
    layers: int = 28
    heads: int = 28
    kv_heads: int = 4
    head_dim: int = 128
    intermediate: int = 18944
    vocab: int = 152064
    rms_eps: float = 1e-6
    rope_theta: float = 1e6
    mrope_section: tuple = (16, 24, 24)

    @property
    def intermediate_padded(self):
        return (self.intermediate + 63) // 64 * 64


@dataclass
class VLMConfig:
    vision: VisionCfg = field(default_factory=VisionCfg)
    text: TextCfg = field(default_factory=TextCfg)
    name: str = "olmocr-7b-class"

    @classmethod
    def olmocr_7b(cls):
        return cls()

    @classmethod
    def qwen72b(cls):
        return cls(VisionCfg(intermediate=3456, out_hidden=8192),
                   TextCfg(hidden=8192, layers=80, heads=64, kv_heads=8, intermediate=29568), "qwen2.5-vl-72b-class")

    @classmethod
    def tiny(cls, text_layers: int = 2, vision_depth: int = 3):
        """Small dims with the 7B model's structure (hd 80 vision, hd 128 text, GQA 7:1...) for tests."""
        return cls(VisionCfg(depth=vision_depth, hidden=320, heads=4, intermediate=600, out_hidden=512,
                             fullatt_blocks=(vision_depth - 1,)),
                   TextCfg(hidden=512, layers=text_layers, heads=4, kv_heads=2, head_dim=128, intermediate=1216,
                           vocab=152064), "tiny")

    @classmethod
    def from_hf_dict(cls, c: dict, name: str = "checkpoint"):
        """VLMConfig from a Qwen2.5-VL `config.json` (what `AutoProcessor / from_pretrained` read at tools.py:700-709).
        Both layouts are accepted: text fields nested under `text_config` with `rope_parameters` (transformers 5.x) or
        flat at the top level with `rope_theta` + `rope_scaling` (the layout the published checkpoints ship)."""
        if c.get("model_type") not in (None, "qwen2_5_vl"):
            raise ValueError(f"model_type {c.get('model_type')!r}: only Qwen2.5-VL-class checkpoints are supported")
        t = c.get("text_config") or c
        v = c["vision_config"]
        rp = t.get("rope_parameters") or t.get("rope_scaling") or c.get("rope_scaling") or {}
        theta = rp.get("rope_theta", t.get("rope_theta", c.get("rope_theta", 1e6)))
        heads = int(t["num_attention_heads"])
        text = TextCfg(hidden=int(t["hidden_size"]), layers=int(t["num_hidden_layers"]), heads=heads,
                       kv_heads=int(t.get("num_key_value_heads", heads)),
                       head_dim=int(t.get("head_dim") or int(t["hidden_size"]) // heads),
                       intermediate=int(t["intermediate_size"]), vocab=int(t["vocab_size"]),
                       rms_eps=float(t.get("rms_norm_eps", 1e-6)), rope_theta=float(theta),
                       mrope_section=tuple(int(x) for x in rp.get("mrope_section", (16, 24, 24))))
        if t.get("tie_word_embeddings", c.get("tie_word_embeddings", False)):
            raise ValueError("tied input / output embeddings are not supported by the weight layout")
        if t.get("use_sliding_window", False):
            raise ValueError("sliding-window text attention is not supported")
        vis = VisionCfg(depth=int(v["depth"]), hidden=int(v["hidden_size"]), heads=int(v["num_heads"]),
                        intermediate=int(v["intermediate_size"]), out_hidden=int(v["out_hidden_size"]),
                        patch=int(v.get("patch_size", v.get("spatial_patch_size", 14))),
                        temporal_patch=int(v.get("temporal_patch_size", 2)), merge=int(v.get("spatial_merge_size", 2)),
                        window=int(v.get("window_size", 112)),
                        fullatt_blocks=tuple(int(x) for x in v.get("fullatt_block_indexes", (7, 15, 23, 31))),
                        tokens_per_second=int(v.get("tokens_per_second", 2)),
                        in_channels=int(v.get("in_channels", v.get("in_chans", 3))))
        if vis.out_hidden != text.hidden:
            raise ValueError("vision out_hidden_size must equal the text hidden_size")
        if text.head_dim != 128 or sum(text.mrope_section) * 2 != text.head_dim:
            raise ValueError("the decode kernels are built for head_dim 128 with an mRoPE split of 64 pairs")
        return cls(vis, text, name)

    @classmethod
    def from_pretrained_dir(cls, path: str):
        import json
        import os
        with open(os.path.join(path, "config.json")) as f:
            return cls.from_hf_dict(json.load(f), name=os.path.basename(os.path.normpath(path)))

    def to_hf(self):
        """The equivalent transformers config (used by tests to build the HF oracle)."""
        from transformers import Qwen2_5_VLConfig
        t, v = self.text, self.vision
        return Qwen2_5_VLConfig(
            text_config=dict(vocab_size=t.vocab, hidden_size=t.hidden, intermediate_size=t.intermediate,
                             num_hidden_layers=t.layers, num_attention_heads=t.heads,
                             num_key_value_heads=t.kv_heads, rms_norm_eps=t.rms_eps,
                             max_position_embeddings=128000,
                             rope_parameters={"rope_type": "default", "rope_theta": t.rope_theta,
                                              "mrope_section": list(t.mrope_section)}),
            vision_config=dict(depth=v.depth, hidden_size=v.hidden, intermediate_size=v.intermediate,
                               num_heads=v.heads, out_hidden_size=v.out_hidden,
                               fullatt_block_indexes=list(v.fullatt_blocks), window_size=v.window,
                               patch_size=v.patch, spatial_merge_size=v.merge,
                               temporal_patch_size=v.temporal_patch, tokens_per_second=v.tokens_per_second))


# ───────────── tokenizer stand-in ─────────────
class SyntheticTokenizer:
    """The real tokenizer files are not available offline.  Parity is defined on token
    ids; this word-level stand-in turns prompts into ids < 151643 and ids back into pseudo text so
    the agreement / merge / CER stage has strings to work on.  A directory with real HF tokenizer
    files can be plugged in through `HFTokenizer`."""

    _SYL = ["ka", "lo", "mi", "ren", "tu", "sha", "ve", "on", "dar", "el", "qui", "st", "ar", "the", "ing", "pro"]

    _PIECE = re.compile(r"\s?\w+|\s?[^\w\s]+|\s+")

    def encode(self, text: str) -> list:
        """Word-piece-like split (about one id per word, as the real BPE does for English) with a
        stable hash into the non-special id range."""
        return [256 + zlib.crc32(p.encode("utf-8")) % 150000 for p in self._PIECE.findall(text)]

    def decode(self, ids, skip_special_tokens: bool = True) -> str:
        words = []
        for t in ids:
            t = int(t)
            if t >= ENDOFTEXT:
                if skip_special_tokens:
                    continue
                words.append(f"<|{t}|>")
                continue
            w = self._SYL[t % 16] + self._SYL[(t // 16) % 16]
            if (t // 256) % 7 == 0:
                w += self._SYL[(t // 1792) % 16]
            if t % 53 == 0:
                w += "\n"
            words.append(w)
        return " ".join(words).replace("\n ", "\n")


class HFTokenizer:
    def __init__(self, path: str):
        from transformers import AutoTokenizer
        self.tk = AutoTokenizer.from_pretrained(path)

    def encode(self, text):
        return self.tk.encode(text, add_special_tokens=False)

    def decode(self, ids, skip_special_tokens=True):
        return self.tk.decode(ids, skip_special_tokens=skip_special_tokens)


def greedy_generation_params(path: str) -> dict:
    """What `model.generate(**inputs, max_new_tokens=...)` (tools.py:764-765) takes from the checkpoint's
    `generation_config.json`: only greedy decoding exists here, so a checkpoint that asks for sampling, beams or a
    repetition penalty is refused loudly instead of being decoded differently from the reference.  Returns the EOS ids."""
    import json
    import os
    fn = os.path.join(path, "generation_config.json")
    if not os.path.exists(fn):
        return {"eos_token_ids": [EOS]}
    with open(fn) as f:
        g = json.load(f)
    unsupported = []
    if g.get("do_sample", False):
        unsupported.append("do_sample=true")
    if int(g.get("num_beams", 1)) != 1:
        unsupported.append(f"num_beams={g['num_beams']}")
    if float(g.get("repetition_penalty", 1.0)) != 1.0:
        unsupported.append(f"repetition_penalty={g['repetition_penalty']}")
    if unsupported:
        raise NotImplementedError("generation_config.json asks for " + ", ".join(unsupported) + "; this engine decodes "
                                  "greedily only (pass tools.configure(force_greedy=True) to decode greedily anyway)")
    eos = g.get("eos_token_id", EOS)
    return {"eos_token_ids": [int(e) for e in (eos if isinstance(eos, (list, tuple)) else [eos])]}


SYSTEM_PROMPT = "You are a helpful assistant."


def build_prompt_ids(tok, prompt: str, n_image_tokens: int) -> np.ndarray:
    """Qwen2.5-VL chat template for one user turn with one image (tools.py:744-762 ->
    HF processing_qwen2_5_vl.py:119-146: one <|image_pad|> expanded to grid.prod()/4 tokens)."""
    ids = [IM_START] + tok.encode("system\n" + SYSTEM_PROMPT) + [IM_END] + tok.encode("\n")
    ids += [IM_START] + tok.encode("user\n") + [VISION_START] + [IMAGE_PAD] * n_image_tokens + [VISION_END]
    ids += tok.encode(prompt) + [IM_END] + tok.encode("\n") + [IM_START] + tok.encode("assistant\n")
    return np.asarray(ids, np.int32)


def rope_index(input_ids: np.ndarray, grid_hw: tuple, merge: int = 2, tokens_per_second: int = 2):
    """HF `get_rope_index` (modeling_qwen2_5_vl.py:1024-1133, as installed: transformers 5.5.0) for one
    sequence with image tokens only: returns (position_ids int64 [3, T], rope_delta).  Text runs
    count up from `current_pos`; an image run of llm grid (h, w) gets t = current_pos *
    tokens_per_second, h = current_pos + row, w = current_pos + col and advances current_pos by
    max(grid_h, grid_w) // merge."""
    T = len(input_ids)
    is_img = input_ids == IMAGE_PAD
    pos = np.zeros((3, T), np.int64)
    cur = 0
    i = 0
    gh, gw = grid_hw[0] // merge, grid_hw[1] // merge
    while i < T:
        j = i
        while j < T and is_img[j] == is_img[i]:
            j += 1
        n = j - i
        if not is_img[i]:
            pos[:, i:j] = np.arange(n)[None, :] + cur
            cur += n
        else:
            if n != gh * gw:
                raise ValueError(f"image run of {n} tokens does not match grid {gh}x{gw}")
            pos[0, i:j] = cur * tokens_per_second
            pos[1, i:j] = cur + np.repeat(np.arange(gh), gw)
            pos[2, i:j] = cur + np.tile(np.arange(gw), gh)
            cur += max(grid_hw[0], grid_hw[1]) // merge
        i = j
    delta = int(pos.max()) + 1 - T
    return pos, delta


def window_index(grid_hw: tuple, merge: int = 2, window: int = 112, patch: int = 14):
    """HF `get_window_index` (modeling_qwen2_5_vl.py:411-451) for one image: returns
    (window_index over merged groups, cu_window_seqlens in patch units, de-duplicated)."""
    gh, gw = grid_hw[0] // merge, grid_hw[1] // merge
    vw = window // merge // patch
    idx = np.arange(gh * gw).reshape(gh, gw)
    pad_h = vw - gh % vw
    pad_w = vw - gw % vw
    nh, nw = (gh + pad_h) // vw, (gw + pad_w) // vw
    padded = np.full((gh + pad_h, gw + pad_w), -100, np.int64)
    padded[:gh, :gw] = idx
    padded = padded.reshape(nh, vw, nw, vw).transpose(0, 2, 1, 3).reshape(nh * nw, vw * vw)
    seqlens = (padded != -100).sum(1)
    flat = padded.reshape(-1)
    widx = flat[flat != -100]
    cu = np.concatenate([[0], np.cumsum(seqlens) * merge * merge])
    # torch.unique_consecutive drops the empty windows
    keep = np.concatenate([[True], cu[1:] != cu[:-1]])
    return widx.astype(np.int64), cu[keep].astype(np.int32)


def vision_rope_pos(grid_hw: tuple, merge: int = 2):
    """HF `rot_pos_emb` position ids (modeling_qwen2_5_vl.py:382-409): (h, w) per patch in merge-group order."""
    h, w = grid_hw
    hp = np.arange(h)[:, None].repeat(w, 1).reshape(h // merge, merge, w // merge, merge).transpose(0, 2, 1, 3).reshape(-1)
    wp = np.arange(w)[None, :].repeat(h, 0).reshape(h // merge, merge, w // merge, merge).transpose(0, 2, 1, 3).reshape(-1)
    return np.stack([hp, wp], -1)
