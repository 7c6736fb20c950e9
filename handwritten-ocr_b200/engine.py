"""Batched OCR read engine: preprocessed pages in HBM -> token ids.

One `read_batch` call serves every (page, strategy) candidate it is given -- the 2 initial reads and
the tiebreaker of a page (nodes.py:86-110), or all strategies of a reocr sweep (nodes.py:239-302) --
in ONE vision-tower pass, ONE prefill and ONE batched greedy decode over the paged KV cache, where
the reference runs them strictly one after another with batch 1 (tools.py:744-765).
"""
from __future__ import annotations

import math
import os
import time

import numpy as np
import torch

from . import _lib, preprocess
from .vlm import BF, Decoder, DecodeState, PagedKV, VisionPlan, VLMWeights, text_rope_tables, vision_forward
from .vlm_config import EOS, IMAGE_PAD, SyntheticTokenizer, VLMConfig, build_prompt_ids, rope_index

OCR_PROMPT = "Extract and return all the text from this handwritten document."


class OcrEngine:
    def __init__(self, weights: VLMWeights, *, max_batch: int = 8, max_new_tokens: int = 2048,
                 max_prompt: int = 1600, page_size: int = 16, tokenizer=None, min_pixels: int = 256 * 256,
                 max_pixels: int = 1024 * 1024, tp=None, prefill_chunk: int = 6, vision_chunk: int | None = None):
        self.w = weights
        self.cfg: VLMConfig = weights.cfg
        self.dev = weights.device
        self.tok = tokenizer or SyntheticTokenizer()
        # further EOS ids of the checkpoint's generation_config (e.g. <|endoftext|>): the device loop stops on <|im_end|>
        # only; a row that emits one of these first is cut there on the host, which is where HF would have stopped it
        self.extra_eos: tuple = ()
        self.max_batch, self.max_new, self.page = max_batch, max_new_tokens, page_size
        self.min_pixels, self.max_pixels = min_pixels, max_pixels
        self.pages_per_seq = math.ceil((max_prompt + max_new_tokens) / page_size)
        self.kv = PagedKV(self.cfg, max_batch * self.pages_per_seq, page_size, self.dev)
        self.tp = tp
        # vision tower + prefill run over at most this many sequences at a time: their activation panels then stay in
        # the 126 MB L2 between the GEMMs' tile waves (measured: 21 -> 13 ms of prefill per read at 48 sequences)
        self.prefill_chunk = max(1, int(prefill_chunk))
        # the vision tower's activation panels are ~2.4x wider per token than the decoder's; it may use smaller chunks
        self.vision_chunk = max(1, int(vision_chunk if vision_chunk is not None else os.environ.get("OCRB_VISION_CHUNK", prefill_chunk)))
        self.dec = Decoder(weights, self.kv, max_batch, self.pages_per_seq * page_size, tp=tp)
        self._plans = {}
        self._scatter = {}
        self._states = {}
        self.timings = {}

    # ───────────── stages ─────────────
    def _plan(self, grid_hw, n):
        key = (grid_hw, n)
        if key not in self._plans:
            self._plans[key] = VisionPlan(self.cfg, grid_hw, n, self.dev)
        return self._plans[key]

    def encode_images(self, pages_u8: torch.Tensor):
        """uint8 pages [n,H,W,3] / [n,H,W] on the device -> (merged embeddings in window order, plan)."""
        n, H, W = pages_u8.shape[:3]
        rh, rw = preprocess.smart_resize(H, W, 28, self.min_pixels, self.max_pixels)
        plan = self._plan((rh // 14, rw // 14), n)
        pv, _ = preprocess.pixel_values(pages_u8, dtype=BF, group_perm=plan.group_perm, min_pixels=self.min_pixels,
                                        max_pixels=self.max_pixels)
        return vision_forward(self.w, plan, pv), plan

    def _scatter_index(self, plan: VisionPlan, img_pos: np.ndarray, T: int) -> torch.Tensor:
        """Destination row (inside the chunk's [n*T, hidden] embeddings) of every merged image token, cached per plan."""
        key = (id(plan), T, int(img_pos[0]))
        idx = self._scatter.get(key)
        if idx is None:
            gp = plan.group_perm.cpu().numpy()                              # group index inside the chunk (i*G + g)
            dst = (gp // plan.G) * T + img_pos[gp % plan.G]
            idx = torch.from_numpy(dst.astype(np.int32)).to(self.dev)
            self._scatter[key] = idx
        return idx

    def build_inputs(self, plan: VisionPlan, prompt: str):
        """Token ids, 3-D positions and rope delta for ONE sequence of this grid (identical for all
        candidates of a batch: same prompt, same grid)."""
        n_img_tok = plan.G
        ids = build_prompt_ids(self.tok, prompt, n_img_tok)
        pos3, delta = rope_index(ids, plan.grid_hw, self.cfg.vision.merge, self.cfg.vision.tokens_per_second)
        return ids, pos3, delta

    def _prefill_batch(self, pages_u8: torch.Tensor, prompt: str, max_new: int, ev=None):
        """Vision tower + prefill of every page of the batch into freshly allocated KV pages, logits of the last prompt
        token and the first greedy pick.  Returns the decode state and what the caller needs to finish / release."""
        n = pages_u8.shape[0]
        if n > self.max_batch:
            raise ValueError(f"batch of {n} exceeds max_batch={self.max_batch}")
        t = self.cfg.text
        H_, W_ = pages_u8.shape[1:3]
        rh, rw = preprocess.smart_resize(H_, W_, 28, self.min_pixels, self.max_pixels)
        plan0 = self._plan((rh // 14, rw // 14), min(n, self.prefill_chunk))
        ids, pos3, delta = self.build_inputs(plan0, prompt)
        T = len(ids)
        if T + max_new > self.pages_per_seq * self.page:
            raise ValueError("prompt + max_new_tokens exceeds the KV capacity this engine was built with")
        h = torch.empty((n * T, t.hidden), dtype=BF, device=self.dev)
        img_pos = np.nonzero(ids == IMAGE_PAD)[0]                       # positions of the G image tokens
        pos_d = torch.from_numpy(pos3).to(self.dev)
        cos1, sin1, inv_freq = text_rope_tables(self.cfg, pos_d)
        # paged KV: block table per sequence
        need = math.ceil((T + max_new) / self.page)
        pages = [self.kv.alloc(need) for _ in range(n)]
        try:
            bt = torch.full((n, self.pages_per_seq), 0, dtype=torch.int32)
            for i, pg in enumerate(pages):
                bt[i, :need] = torch.tensor(pg, dtype=torch.int32)
            bt = bt.to(self.dev)
            vis_ev = []
            merged = plan = None
            for i0 in range(0, n, self.prefill_chunk):
                c = min(self.prefill_chunk, n - i0)
                hc = h[i0 * T:(i0 + c) * T]
                ids_d = torch.from_numpy(np.tile(ids, c)).to(self.dev)
                _lib.call("ocrb_embed_gather", self.w.embed.data_ptr(), ids_d.data_ptr(), hc.data_ptr(), c * T, t.hidden,
                          _lib.stream_ptr())
                for j0 in range(0, c, self.vision_chunk):
                    cv = min(self.vision_chunk, c - j0)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    with _lib.nvtx_range("ocr.vision_tower"):
                        merged, plan = self.encode_images(pages_u8[i0 + j0:i0 + j0 + cv])
                    e1.record()
                    vis_ev.append((e0, e1))
                    # scatter image embeddings: window-order row r of image i is source group plan.group_perm[r]
                    dst_d = self._scatter_index(plan, img_pos, T)
                    hv = hc[j0 * T:(j0 + cv) * T]
                    _lib.call("ocrb_rows_copy", merged.data_ptr(), merged.stride(0), None, hv.data_ptr(), hv.stride(0),
                              dst_d.data_ptr(), merged.shape[0], t.hidden, _lib.stream_ptr())
                cos, sin = cos1.repeat(c, 1).contiguous(), sin1.repeat(c, 1).contiguous()
                cu = torch.arange(0, (c + 1) * T, T, dtype=torch.int32, device=self.dev)
                with _lib.nvtx_range("ocr.prefill"):
                    self.dec.prefill(hc, cos, sin, cu, c, T, bt[i0:i0 + c])
            last = h[T - 1::T]                                          # [n, hidden] last prompt token of each sequence
            key = (n, max_new)
            st = self._states.get(key)
            if st is None:
                st = DecodeState(self.dec, n, max_new, bt, [T] * n, [delta] * n, inv_freq)
                self._states[key] = st
            else:
                st.block_table.copy_(bt)
                st.ctx_len.fill_(T)
                st.rope_delta.fill_(delta)
                st.finished.zero_()
                st.step.zero_()
                st.out_tokens.fill_(EOS)
            last_c = last.contiguous()
            self.dec.logits_last(last_c, st.logits_local)
            if self.tp is not None:
                self.tp.gather_vocab(st.logits_local, st.logits)
            prefill_logits = st.logits.clone()
            _lib.call("ocrb_argmax_step", st.logits.data_ptr(), st.logits.stride(0), n, t.vocab, EOS, EOS, max_new,
                      st.out_tokens.data_ptr(), st.next_ids.data_ptr(), st.finished.data_ptr(), st.ctx_len.data_ptr(),
                      st.step.data_ptr(), 0, _lib.stream_ptr())
        except Exception:
            for pg in pages:
                self.kv.release(pg)
            raise
        return st, pages, {"T": T, "vis_ev": vis_ev, "prefill_logits": prefill_logits, "merged": merged, "plan": plan,
                           "ids": ids, "pos3": pos3, "delta": delta}

    def read_batch(self, pages_u8: torch.Tensor, *, prompt: str = OCR_PROMPT, max_new_tokens: int | None = None,
                   use_graph: bool = True, return_debug: bool = False):
        """Greedy transcription token ids for each page of the batch (HF `generate` semantics:
        tools.py:764-765 with the default GenerationConfig: greedy, eos 151645, pad = eos)."""
        max_new = self.max_new if max_new_tokens is None else int(max_new_tokens)
        n = pages_u8.shape[0]
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        st, pages, info = self._prefill_batch(pages_u8, prompt, max_new)
        T = info["T"]
        try:
            ev[2].record()
            with _lib.nvtx_range("ocr.decode"):
                self.dec.decode(st, max_new - 1, use_graph=use_graph)
            ev[3].record()
            toks = st.out_tokens.cpu().numpy()
            n_steps = int(st.step.cpu())
        finally:
            for pg in pages:
                self.kv.release(pg)
        torch.cuda.synchronize()
        vision_ms = sum(a.elapsed_time(b) for a, b in info["vis_ev"])
        self.timings = {"vision_ms": vision_ms, "prefill_ms": ev[0].elapsed_time(ev[2]) - vision_ms,
                        "decode_ms": ev[2].elapsed_time(ev[3]), "prompt_len": T, "steps": n_steps, "batch": n}
        # HF stops appending once every sequence is finished (finished rows are padded with eos until
        # then); the device loop only checks every 64 steps, so trim to HF's stopping point.
        toks = toks[:, :n_steps]
        if self.extra_eos:
            for i in range(n):
                hit = np.isin(toks[i], self.extra_eos)
                if hit.any():
                    toks[i, int(hit.argmax()) + 1:] = EOS          # HF pads a finished row with pad = eos
        is_eos = (toks == EOS) | (np.isin(toks, self.extra_eos) if self.extra_eos else False)
        first_eos = np.where(is_eos.any(1), is_eos.argmax(1), n_steps - 1)
        keep = int(first_eos.max()) + 1 if n_steps > 0 else 0
        out = [toks[i, :keep].tolist() for i in range(n)]
        if return_debug:
            return out, {k: info[k] for k in ("prefill_logits", "merged", "plan", "ids", "pos3", "delta")}
        return out

    def teacher_forced_logits(self, pages_u8: torch.Tensor, forced: torch.Tensor, on_step, *, prompt: str = OCR_PROMPT):
        """Parity instrument (SURVEY §7 hard part 1, protocol v): decode with the token ids of ANOTHER implementation
        (`forced`: int [n, steps], e.g. HF `generate` on the same weights) as inputs and hand this engine's logits of
        every position to `on_step(i, logits bf16 [n, V])`: i = 0 are the prefill logits (they pick forced[:, 0]), step
        i >= 1 is computed with forced[:, i - 1] as its input token.  Same kernels as read_batch, launched step by step
        (no CUDA graph) so the input token can be replaced between steps."""
        n, steps = forced.shape
        if n != pages_u8.shape[0]:
            raise ValueError("one row of forced token ids per page")
        forced = forced.to(device=self.dev, dtype=torch.int32).contiguous()
        st, pages, info = self._prefill_batch(pages_u8, prompt, steps)
        try:
            on_step(0, info["prefill_logits"])
            for i in range(1, steps):
                st.next_ids.copy_(forced[:, i - 1])
                st.finished.zero_()
                self.dec._step(st)
                on_step(i, st.logits)
        finally:
            for pg in pages:
                self.kv.release(pg)
        torch.cuda.synchronize()

    def close(self) -> None:
        """Drop the captured decode graphs and cached states (call before tearing down a torch.distributed group:
        a CUDA graph that captured NCCL collectives must not outlive its communicator)."""
        for st in self._states.values():
            st.graph = None
        self._states.clear()
        torch.cuda.synchronize()

    def detokenize(self, ids) -> str:
        return self.tok.decode(ids, skip_special_tokens=True)
