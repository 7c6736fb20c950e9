"""ORACLE / BASELINE INFRASTRUCTURE (not product code) -- the reference's CPU read path, timed.

Only bench.py's `cpu_baseline` leg and `bench.py --impl reference` import this module.

The reference (`/root/reference/ocr_agent`) is pure Python glue over third-party wheels, and
`/root/reference` does not exist on the GPU box, so the CPU arm is assembled from the very calls the
reference makes, on the box's host cores:

  * preprocessing: the same OpenCV calls with the same parameters as tools.py:503-573 (cv2 is the
    wheel the reference itself calls; when cv2 is missing the numpy restatement in image_ref.py is
    timed instead and the record says so);
  * the read: HF transformers' Qwen2.5-VL classes (what `AutoModelForImageTextToText` resolves to
    for the configured checkpoint, tools.py:705-709) in bf16 eager mode with `generate`
    (tools.py:764-765), at full width and (by default in bench.py) full depth; the vision tower,
    the prefill and a few decode steps are measured, only the NUMBER of decode steps is scaled to
    the 512-token read -- a complete read takes ~5 min on 8 cores, far beyond a
    bounded sample;
  * agreement / merge / CER: pure-Python dynamic programmes, line-for-line the algorithm of
    tools.py:69-100,465-493 (text_ref.levenshtein_py / align_to_backbone_py), on a bounded prefix and
    scaled by the cell count n*m.
"""
from __future__ import annotations

import os
import time

import numpy as np

STRATEGIES = (("deskew", "high_contrast", "binarize"), ("high_contrast", "binarize"),
              ("deskew", "high_contrast", "sharpen"))


# ───────────── preprocessing as the reference calls OpenCV ─────────────
def _cv_transforms():
    import cv2

    def gray_of(a):
        return cv2.cvtColor(a, cv2.COLOR_RGB2GRAY) if a.ndim == 3 else a

    def high_contrast(a):
        return cv2.createCLAHE(clipLimit=3.0, tileGridSize=(8, 8)).apply(gray_of(a))

    def binarize(a):
        return cv2.adaptiveThreshold(gray_of(a), 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, cv2.THRESH_BINARY, 21, 10)

    def sharpen(a):
        k = np.array([[0, -1, 0], [-1, 5, -1], [0, -1, 0]], np.float32)
        return cv2.filter2D(a, -1, k)

    def deskew(a):
        g = gray_of(a)
        pts = np.column_stack(np.where(g < 128))
        if len(pts) <= 100:
            return a
        ang = cv2.minAreaRect(pts)[-1]
        ang = -(90 + ang) if ang < -45 else -ang
        h, w = g.shape
        M = cv2.getRotationMatrix2D((w // 2, h // 2), ang, 1.0)
        return cv2.warpAffine(a, M, (w, h), flags=cv2.INTER_CUBIC, borderMode=cv2.BORDER_REPLICATE)

    def denoise(a):                                     # tools.py:582-587
        if a.ndim == 3:
            return cv2.fastNlMeansDenoisingColored(a, None, 10, 10, 7, 21)
        return cv2.fastNlMeansDenoising(a, None, 10, 7, 21)

    def remove_lines(a):                                # tools.py:598-617
        g = gray_of(a)
        hk = cv2.getStructuringElement(cv2.MORPH_RECT, (g.shape[1] // 4, 1))
        m = cv2.morphologyEx(cv2.adaptiveThreshold(cv2.bitwise_not(g), 255, cv2.ADAPTIVE_THRESH_MEAN_C,
                                                   cv2.THRESH_BINARY, 15, -2), cv2.MORPH_OPEN, hk, iterations=1)
        m = cv2.dilate(m, cv2.getStructuringElement(cv2.MORPH_RECT, (1, 3)))
        return cv2.inpaint(a, m, 3, cv2.INPAINT_TELEA)

    return {"high_contrast": high_contrast, "binarize": binarize, "sharpen": sharpen, "deskew": deskew,
            "denoise": denoise, "remove_lines": remove_lines}, "cv2"


def _np_transforms():
    from . import image_ref as R
    return dict(R.TRANSFORMS), "numpy-port"


def transform_times_cpu(page: np.ndarray, ruled: np.ndarray) -> dict:
    """Seconds of each single transform of tools.py:623-630 on one page with the reference's own cv2 calls (second of two
    runs; remove_lines on the ruled page).  Empty when cv2 is missing: the numpy port of NLM takes minutes per page."""
    try:
        tf, _ = _cv_transforms()
    except ImportError:
        return {}
    out = {}
    for name, fn in tf.items():
        a = ruled if name == "remove_lines" else page
        fn(a)
        t0 = time.perf_counter()
        fn(a)
        out[name] = time.perf_counter() - t0
    return out


def preprocess_cpu(page: np.ndarray, strategies=STRATEGIES):
    """All strategies of one page -> (list of uint8 arrays, seconds, backend name)."""
    try:
        tf, kind = _cv_transforms()
    except ImportError:
        tf, kind = _np_transforms()
    t0 = time.perf_counter()
    outs = []
    for s in strategies:
        a = page
        for step in s:
            a = tf[step](a)
        outs.append(a)
    return outs, time.perf_counter() - t0, kind


# ───────────── the VLM read with HF transformers on the CPU ─────────────
class HFCpuReader:
    """Full-width, reduced-depth HF model + per-layer timers.  Built once, reused per sample."""

    def __init__(self, cfg, text_layers: int = 2, vision_depth: int = 2, threads: int | None = None):
        import torch
        from transformers import Qwen2_5_VLForConditionalGeneration, initialization
        self.torch = torch
        if threads:
            torch.set_num_threads(threads)
        self.threads = torch.get_num_threads()
        self.full_cfg = cfg
        hf_cfg = cfg.to_hf()
        hf_cfg.text_config.num_hidden_layers = text_layers
        hf_cfg.text_config.layer_types = hf_cfg.text_config.layer_types[:text_layers] if getattr(
            hf_cfg.text_config, "layer_types", None) else None
        if vision_depth < cfg.vision.depth:
            hf_cfg.vision_config.depth = vision_depth
            hf_cfg.vision_config.fullatt_block_indexes = [vision_depth - 1]   # one windowed block + one full block
        self.full_blocks = set(int(i) for i in hf_cfg.vision_config.fullatt_block_indexes)
        self.text_layers, self.vision_depth = text_layers, vision_depth
        with initialization.no_init_weights():
            m = Qwen2_5_VLForConditionalGeneration._from_config(hf_cfg, dtype=torch.bfloat16)
        # cheap random fill (values only need to be finite and non-trivial for timing)
        g = torch.Generator().manual_seed(0)
        block = (torch.randn(1 << 20, generator=g) * 0.02).to(torch.bfloat16)
        for name, p in m.named_parameters():
            flat = p.data.view(-1)
            n = flat.numel()
            reps = (n + block.numel() - 1) // block.numel()
            flat.copy_(block.repeat(reps)[:n])
            if name.endswith("norm.weight") or "layernorm" in name or name.endswith("norm1.weight") or name.endswith(
                    "norm2.weight") or name.endswith("ln_q.weight"):
                p.data.fill_(1.0)
        self.model = m.eval()
        self._t = {"vis_win": [], "vis_full": [], "txt": []}
        self._hooks()

    def _hooks(self):
        vis = self.model.model.visual.blocks
        txt = self.model.model.language_model.layers
        starts = {}

        def pre(key):
            def f(mod, args, kwargs=None):
                starts[key] = time.perf_counter()
            return f

        def post(key, bucket):
            def f(mod, args, out):
                self._t[bucket].append(time.perf_counter() - starts[key])
            return f

        for i, b in enumerate(vis):
            b.register_forward_pre_hook(pre(("v", i)))
            b.register_forward_hook(post(("v", i), "vis_full" if i in self.full_blocks else "vis_win"))
        for i, l in enumerate(txt):
            l.register_forward_pre_hook(pre(("t", i)))
            l.register_forward_hook(post(("t", i), "txt"))

    def read(self, pixel_values, grid_hw, input_ids: np.ndarray, n_new: int, full_new: int):
        """One greedy read on the truncated model; returns measured and depth/step-extrapolated seconds."""
        torch = self.torch
        for v in self._t.values():
            v.clear()
        ids = torch.from_numpy(input_ids.astype(np.int64))[None]
        inp = dict(input_ids=ids, attention_mask=torch.ones_like(ids), pixel_values=pixel_values,
                   image_grid_thw=torch.tensor([[1, grid_hw[0], grid_hw[1]]]),
                   mm_token_type_ids=(ids == 151655).int())
        t0 = time.perf_counter()
        with torch.no_grad():
            out = self.model.generate(**inp, max_new_tokens=n_new, min_new_tokens=n_new, do_sample=False)
        total = time.perf_counter() - t0
        n_gen = out.shape[1] - ids.shape[1]
        fc = self.full_cfg
        L, Lm = fc.text.layers, self.text_layers
        txt = self._t["txt"]
        prefill_layer = float(np.mean(txt[:Lm]))
        dec_layer = float(np.mean(txt[Lm:])) if len(txt) > Lm else 0.0
        vis_win = float(np.mean(self._t["vis_win"]))
        vis_full = float(np.mean(self._t["vis_full"]))
        measured_layers = sum(txt) + sum(self._t["vis_win"]) + sum(self._t["vis_full"])
        # everything that is not a vision block or decoder layer: image->embeds, lm_head, sampling loop
        other = total - measured_layers
        n_dec_calls = max(n_gen - 1, 1)
        other_prefill_share = other / (n_dec_calls + 1)     # per forward call (lm_head etc.), rough
        n_full_blocks = len(fc.vision.fullatt_blocks)
        # (at full depth these sums are the measured totals: nothing is scaled)
        est_vision = (fc.vision.depth - n_full_blocks) * vis_win + n_full_blocks * vis_full
        est_prefill = L * prefill_layer + other_prefill_share
        est_step = L * dec_layer + other_prefill_share
        est_total = est_vision + est_prefill + (full_new - 1) * est_step
        return {"measured_s": total, "n_new_measured": int(n_gen), "vision_block_window_s": vis_win,
                "vision_block_full_s": vis_full, "prefill_layer_s": prefill_layer, "decode_layer_s": dec_layer,
                "per_call_other_s": other_prefill_share, "est_vision_s": est_vision, "est_prefill_s": est_prefill,
                "est_decode_step_s": est_step, "est_read_s": est_total, "full_new_tokens": full_new}


    def decode_sample(self, n_new: int = 4, prompt_tokens: int = 16):
        """`n_new` greedy decode steps on a short text-only prompt: seconds per decode step of the (full-depth) model.
        The per-step cost on the CPU is the 14 GB weight pass; the context length changes it by well under 1 %."""
        torch = self.torch
        for v in self._t.values():
            v.clear()
        ids = torch.arange(1000, 1000 + prompt_tokens, dtype=torch.int64)[None]
        t0 = time.perf_counter()
        with torch.no_grad():
            out = self.model.generate(input_ids=ids, attention_mask=torch.ones_like(ids), max_new_tokens=n_new + 1,
                                      min_new_tokens=n_new + 1, do_sample=False)
        total = time.perf_counter() - t0
        Lm = self.text_layers
        txt = self._t["txt"]
        n_calls = len(txt) // Lm                                 # 1 prefill call + n_new decode calls
        dec_layers = txt[Lm:]
        per_call_layers = sum(dec_layers) / max(n_calls - 1, 1)
        other = (total - sum(txt)) / max(n_calls, 1)             # lm_head, sampling loop: per forward call
        scale = self.full_cfg.text.layers / Lm
        return {"decode_step_s": per_call_layers * scale + other, "decode_steps_measured": int(n_calls - 1),
                "layers_run": Lm, "wall_s": total, "new_tokens": int(out.shape[1] - ids.shape[1])}


# ───────────── pure-Python text DP, as the reference runs it ─────────────
def text_ops_cpu(texts, prefix_chars: int = 700, prefix_words: int = 160):
    """compare_versions(t0, t1) + merge_versions(texts) cost in the reference's pure-Python form,
    measured on prefixes and scaled by DP cell counts.  Returns estimated seconds and the detail."""
    from . import text_ref as T
    n = [T.normalize_text(t) for t in texts]
    w = [x.split() for x in n]
    a, b = n[0][:prefix_chars], n[1][:prefix_chars]
    t0 = time.perf_counter()
    T.levenshtein_py(a, b)
    t_char = time.perf_counter() - t0
    cells_char = max(len(a) * len(b), 1)
    wa, wb = w[0][:prefix_words], w[1][:prefix_words]
    t0 = time.perf_counter()
    T.levenshtein_py(wa, wb)
    t_word = time.perf_counter() - t0
    cells_word = max(len(wa) * len(wb), 1)
    t0 = time.perf_counter()
    T.align_to_backbone_py(wa, wb)
    t_lcs = time.perf_counter() - t0
    full_char = len(n[0]) * len(n[1])
    full_word = len(w[0]) * len(w[1])
    backbone = max(w, key=len)
    full_lcs = sum(len(backbone) * len(x) for x in w)
    est = t_char * full_char / cells_char + t_word * full_word / cells_word + t_lcs * full_lcs / cells_word
    return est, {"char_cells": full_char, "word_cells": full_word, "lcs_cells": full_lcs,
                 "char_mcells_per_s": cells_char / t_char / 1e6, "lcs_mcells_per_s": cells_word / max(t_lcs, 1e-9) / 1e6}


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1
