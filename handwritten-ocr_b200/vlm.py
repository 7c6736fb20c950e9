"""The configured VLM (Qwen2.5-VL class) on hand-written sm_100a kernels.

Host code here is plumbing only: it owns device buffers (torch tensors), packs weights into the
layouts the kernels want, and sequences C-ABI calls on the current CUDA stream.  All arithmetic on
activations happens in libocrb200 (csrc/): tcgen05 GEMMs with fused HF-rounding epilogues, flash
attention, weight-streaming GEMVs and paged decode attention.

Reference semantics: HF transformers modeling_qwen2_5_vl.py (vision tower :345-518, decoder
:672-942, generation/utils.py:2727-2806 greedy loop), reached from ocr_agent/tools.py:764-765.
"""
from __future__ import annotations

import ctypes
import math
import os

import numpy as np
import torch

from . import _lib
from .vlm_config import VLMConfig, rope_index, vision_rope_pos, window_index, EOS, IMAGE_PAD

BF = torch.bfloat16
EPI_NONE, EPI_RESIDUAL, EPI_SWIGLU, EPI_GELU = 0, 1, 2, 3


def _sp():
    return torch.cuda.current_stream().cuda_stream


# ───────────────────────── C-ABI call helpers ─────────────────────────
def gemm(A, W, out, *, bias=None, residual=None, epilogue=EPI_NONE, N=None, K=None):
    """out[M, N'] = epilogue(A[M,K] @ W[N,K]^T) on tcgen05 (A/out may be row-strided views)."""
    M = A.shape[0]
    K = A.shape[1] if K is None else K
    N = W.shape[0] if N is None else N
    _lib.call("ocrb_gemm_bf16", A.data_ptr(), A.stride(0), W.data_ptr(), W.stride(0), out.data_ptr(), out.stride(0),
              M, N, K, _lib.ptr(bias), _lib.ptr(residual), residual.stride(0) if residual is not None else 0,
              epilogue, _sp())
    return out


_SKINNY_WS = {}


def skinny_workspace(device) -> torch.Tensor:
    """Per-device stream-K workspace of ocrb_skinny_gemm_bf16 (zeroed once; launches are stream-ordered)."""
    key = torch.device(device).index or 0
    ws = _SKINNY_WS.get(key)
    if ws is None:
        n = int(_lib.load().ocrb_skinny_workspace_bytes())
        ws = torch.zeros(n, dtype=torch.uint8, device=device)
        _SKINNY_WS[key] = ws
    return ws


SKINNY_MAX_ROWS = 128
DECODE_KEYS_PER_CTA = int(os.environ.get("OCRB_ATTN_KEYS", "128"))    # decode attention: keys per CTA (4 warps x 16-key tiles)


def skinny(X, W, out, *, bias=None, residual=None, epilogue=EPI_NONE, norm_w=None, eps=1e-6):
    """out[B, N'] = epilogue(X[B,K] @ W[N,K]^T), B <= 128: every weight byte crosses HBM once (tcgen05 swap-AB)."""
    B, K = X.shape
    _lib.call("ocrb_skinny_gemm_bf16", X.data_ptr(), X.stride(0), W.data_ptr(), W.stride(0), out.data_ptr(),
              out.stride(0), B, W.shape[0], K, _lib.ptr(bias), _lib.ptr(residual),
              residual.stride(0) if residual is not None else 0, epilogue, _lib.ptr(norm_w), float(eps),
              skinny_workspace(X.device).data_ptr(), _sp())
    return out


_CHAIN_WS = {}
# Largest batch decoded through the persistent chain / plan kernels (csrc/chain.cu).  Default 0: one launch per op (skinny /
# cluster GEMMs + attention kernels under PDL) -- measured equal at B = 3 and faster above (profiles/r02_notes.md).
CHAIN_MAX_B = int(os.environ.get("OCRB_CHAIN_MAX_B", "0"))
CHAIN_FUSE_ATTN = os.environ.get("OCRB_CHAIN_ATTN", "1") == "1"   # the whole step (attention included) as one plan launch
TP_ARGMAX_GATHER = os.environ.get("OCRB_TP_ARGMAX_GATHER", "0") == "1"   # tensor parallel: all-gather the logits (NCCL) instead of the pair exchange


CHAIN_TRACE = None        # debugging: a list collects one timestamp buffer per chain launch


def chain_workspace(device) -> torch.Tensor:
    """Per-device workspace of ocrb_skinny_chain_bf16 (zeroed once; the kernel returns its counters to zero)."""
    key = torch.device(device).index or 0
    ws = _CHAIN_WS.get(key)
    if ws is None:
        ws = torch.zeros(int(_lib.load().ocrb_chain_workspace_bytes()), dtype=torch.uint8, device=device)
        _CHAIN_WS[key] = ws
    return ws


def chain_linear(X, W, out, *, bias=None, residual=None, epilogue=EPI_NONE, norm_w=None, eps=1e-6):
    """One linear of a `skinny_chain` call: same arguments and meaning as `skinny`."""
    return _lib.ChainLinear(X.data_ptr(), X.stride(0), W.data_ptr(), W.stride(0), out.data_ptr(), out.stride(0),
                            W.shape[0], X.shape[1], _lib.ptr(bias), _lib.ptr(residual),
                            residual.stride(0) if residual is not None else 0, epilogue, float(eps), _lib.ptr(norm_w))


def skinny_chain(linears, B: int, device):
    """Dependent skinny linears (each may read what the earlier ones wrote) in ONE persistent launch: the weight stream
    of the whole chain keeps HBM busy across the dependencies (csrc/chain.cu)."""
    arr = (_lib.ChainLinear * len(linears))(*linears)
    if CHAIN_TRACE is not None:       # scripts/trace_chain.py: one [grid][64] globaltimer buffer per launch
        buf = torch.zeros(296 * 64, dtype=torch.int64, device=device)
        CHAIN_TRACE.append(buf)
        _lib.load().ocrb_chain_set_trace(ctypes.c_void_p(buf.data_ptr()), 64)
    _lib.call("ocrb_skinny_chain_bf16", ctypes.addressof(arr), len(linears), B, chain_workspace(device).data_ptr(), _sp())


def linear_small_or_big(X, W, out, **kw):
    """Row count decides the datapath: <= 128 rows stream the weights once (skinny GEMM), else full tiles."""
    if X.shape[0] <= SKINNY_MAX_ROWS:
        return skinny(X, W, out, **kw)
    return gemm(X, W, out, **kw)


def residual_add(x, y):
    """x += y (bf16, rounded once) -- after a tensor-parallel all-reduce."""
    _lib.call("ocrb_residual_add_bf16", x.data_ptr(), x.stride(0), y.data_ptr(), y.stride(0), x.shape[0], x.shape[1], _sp())
    return x


def rmsnorm(x, w, out, eps=1e-6):
    _lib.call("ocrb_rmsnorm_bf16", x.data_ptr(), x.stride(0), w.data_ptr(), out.data_ptr(), out.stride(0),
              x.shape[0], x.shape[1], float(eps), _sp())
    return out


FLASH_TC_MIN_LEN = int(os.environ.get("OCRB_FLASH_TC_MIN_LEN", "256"))   # shorter sequences (vision windows) use mma.sync


def attention(q, k, v, out, cu, n_seq, max_len, n_q, n_kv, hd, causal):
    if max_len >= FLASH_TC_MIN_LEN and hd in (80, 128):
        # long sequences (vision full-attention blocks, prefill): tcgen05 + TMEM + TMA flash attention
        _lib.call("ocrb_flash_attention_bf16", q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(),
                  v.stride(0), out.data_ptr(), out.stride(0), cu.data_ptr(), n_seq, q.shape[0], max_len, n_q, n_kv, hd,
                  float(hd ** -0.5), int(causal), _sp())
        return out
    _lib.call("ocrb_attention_varlen", q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0),
              out.data_ptr(), out.stride(0), cu.data_ptr(), n_seq, max_len, n_q, n_kv, hd, float(hd ** -0.5),
              int(causal), _sp())
    return out


# ───────────────────────── weights ─────────────────────────
def random_state_dict(cfg: VLMConfig, device, seed: int = 0, std: float = 0.02, bias_std: float = 0.02,
                      lm_head_std: float | None = None, skip_tp_sharded: bool = False) -> dict:
    """Random-init weights in HF's state-dict naming (normal(0, 0.02) like HF `initializer_range`,
    norms = 1 + small noise so their multiply is exercised, non-zero biases so the bias paths are
    exercised).  The same dict is loaded into the HF oracle and into `VLMWeights.from_state_dict`."""
    g = torch.Generator(device=device).manual_seed(seed)
    sd = {}

    def w(name, *shape, s=std):
        sd[name] = (torch.randn(*shape, generator=g, device=device, dtype=torch.float32) * s).to(BF)

    def norm(name, dim):
        sd[name] = (1.0 + 0.05 * torch.randn(dim, generator=g, device=device, dtype=torch.float32)).to(BF)

    v, t = cfg.vision, cfg.text
    w("model.visual.patch_embed.proj.weight", v.hidden, v.in_channels, v.temporal_patch, v.patch, v.patch)
    for i in range(v.depth):
        p = f"model.visual.blocks.{i}."
        norm(p + "norm1.weight", v.hidden)
        norm(p + "norm2.weight", v.hidden)
        w(p + "attn.qkv.weight", 3 * v.hidden, v.hidden)
        w(p + "attn.qkv.bias", 3 * v.hidden, s=bias_std)
        w(p + "attn.proj.weight", v.hidden, v.hidden)
        w(p + "attn.proj.bias", v.hidden, s=bias_std)
        for nm, o, k in (("gate_proj", v.intermediate, v.hidden), ("up_proj", v.intermediate, v.hidden),
                         ("down_proj", v.hidden, v.intermediate)):
            w(p + f"mlp.{nm}.weight", o, k)
            w(p + f"mlp.{nm}.bias", o, s=bias_std)
    mh = v.hidden * v.merge * v.merge
    norm("model.visual.merger.ln_q.weight", v.hidden)
    w("model.visual.merger.mlp.0.weight", mh, mh)
    w("model.visual.merger.mlp.0.bias", mh, s=bias_std)
    w("model.visual.merger.mlp.2.weight", v.out_hidden, mh)
    w("model.visual.merger.mlp.2.bias", v.out_hidden, s=bias_std)
    w("model.language_model.embed_tokens.weight", t.vocab, t.hidden)
    for i in range(t.layers):
        p = f"model.language_model.layers.{i}."
        norm(p + "input_layernorm.weight", t.hidden)
        norm(p + "post_attention_layernorm.weight", t.hidden)
        if skip_tp_sharded:
            continue
        w(p + "self_attn.q_proj.weight", t.heads * t.head_dim, t.hidden)
        w(p + "self_attn.q_proj.bias", t.heads * t.head_dim, s=bias_std)
        w(p + "self_attn.k_proj.weight", t.kv_heads * t.head_dim, t.hidden)
        w(p + "self_attn.k_proj.bias", t.kv_heads * t.head_dim, s=bias_std)
        w(p + "self_attn.v_proj.weight", t.kv_heads * t.head_dim, t.hidden)
        w(p + "self_attn.v_proj.bias", t.kv_heads * t.head_dim, s=bias_std)
        w(p + "self_attn.o_proj.weight", t.hidden, t.heads * t.head_dim)
        w(p + "mlp.gate_proj.weight", t.intermediate, t.hidden)
        w(p + "mlp.up_proj.weight", t.intermediate, t.hidden)
        w(p + "mlp.down_proj.weight", t.hidden, t.intermediate)
    norm("model.language_model.norm.weight", t.hidden)
    if not skip_tp_sharded:
        w("lm_head.weight", t.vocab, t.hidden, s=std if lm_head_std is None else lm_head_std)
    return sd


def random_state_dict_text_only(cfg: VLMConfig, device, seed: int, vocab_rows: int | None = None, std: float = 0.02,
                                bias_std: float = 0.02, lm_head_std: float | None = None) -> dict:
    """Only the tensors tensor parallelism shards (q/k/v/o, gate/up/down, lm_head) at the shapes of `cfg` (a local
    config) -- used to give every rank its own random shard without building the full model."""
    g = torch.Generator(device=device).manual_seed(seed)
    sd = {}
    t = cfg.text

    def w(name, *shape, s=std):
        sd[name] = (torch.randn(*shape, generator=g, device=device, dtype=torch.float32) * s).to(BF)

    for i in range(t.layers):
        p = f"model.language_model.layers.{i}."
        w(p + "self_attn.q_proj.weight", t.heads * t.head_dim, t.hidden)
        w(p + "self_attn.q_proj.bias", t.heads * t.head_dim, s=bias_std)
        w(p + "self_attn.k_proj.weight", t.kv_heads * t.head_dim, t.hidden)
        w(p + "self_attn.k_proj.bias", t.kv_heads * t.head_dim, s=bias_std)
        w(p + "self_attn.v_proj.weight", t.kv_heads * t.head_dim, t.hidden)
        w(p + "self_attn.v_proj.bias", t.kv_heads * t.head_dim, s=bias_std)
        w(p + "self_attn.o_proj.weight", t.hidden, t.heads * t.head_dim)
        w(p + "mlp.gate_proj.weight", t.intermediate, t.hidden)
        w(p + "mlp.up_proj.weight", t.intermediate, t.hidden)
        w(p + "mlp.down_proj.weight", t.hidden, t.intermediate)
    w("lm_head.weight", vocab_rows or t.vocab, t.hidden, s=std if lm_head_std is None else lm_head_std)
    return sd


def normalize_checkpoint_key(key: str) -> str:
    """Checkpoint tensor name -> the name HF's Qwen2_5_VLForConditionalGeneration holds after `from_pretrained`
    (transformers conversion_mapping: `^visual` -> `model.visual`, `^model.(?!language_model|visual)` ->
    `model.language_model.`).  New-style names pass through unchanged."""
    if key.startswith("visual."):
        return "model." + key
    if key.startswith("model.") and not key.startswith(("model.language_model.", "model.visual.")):
        return "model.language_model." + key[len("model."):]
    return key


def expected_state_dict_keys(cfg: VLMConfig) -> list:
    """Every tensor name `VLMWeights.from_state_dict` consumes for this config (HF naming after load)."""
    v, t = cfg.vision, cfg.text
    keys = ["model.visual.patch_embed.proj.weight", "model.visual.merger.ln_q.weight"]
    keys += [f"model.visual.merger.mlp.{i}.{wb}" for i in (0, 2) for wb in ("weight", "bias")]
    for i in range(v.depth):
        p = f"model.visual.blocks.{i}."
        keys += [p + "norm1.weight", p + "norm2.weight"]
        keys += [p + f"attn.{nm}.{wb}" for nm in ("qkv", "proj") for wb in ("weight", "bias")]
        keys += [p + f"mlp.{nm}.{wb}" for nm in ("gate_proj", "up_proj", "down_proj") for wb in ("weight", "bias")]
    keys += ["model.language_model.embed_tokens.weight", "model.language_model.norm.weight", "lm_head.weight"]
    for i in range(t.layers):
        p = f"model.language_model.layers.{i}."
        keys += [p + "input_layernorm.weight", p + "post_attention_layernorm.weight", p + "self_attn.o_proj.weight"]
        keys += [p + f"self_attn.{nm}.{wb}" for nm in ("q_proj", "k_proj", "v_proj") for wb in ("weight", "bias")]
        keys += [p + f"mlp.{nm}.weight" for nm in ("gate_proj", "up_proj", "down_proj")]
    return keys


def _pack_swiglu(gate: torch.Tensor, up: torch.Tensor, ipad: int) -> torch.Tensor:
    """Rows interleaved per 64 as [gate64 | up64] (OCRB_EPI_SWIGLU); zero rows pad I up to `ipad`."""
    I = gate.shape[0]
    if ipad != I:
        pad = torch.zeros((ipad - I,) + tuple(gate.shape[1:]), dtype=gate.dtype, device=gate.device)
        gate, up = torch.cat([gate, pad]), torch.cat([up, pad])
    rest = tuple(gate.shape[1:])
    return torch.stack([gate.reshape((ipad // 64, 64) + rest), up.reshape((ipad // 64, 64) + rest)], 1).reshape(
        (2 * ipad,) + rest).contiguous()


def _pad_cols(w: torch.Tensor, kpad: int) -> torch.Tensor:
    if w.shape[1] == kpad:
        return w.contiguous()
    out = torch.zeros((w.shape[0], kpad), dtype=w.dtype, device=w.device)
    out[:, : w.shape[1]] = w
    return out


class VLMWeights:
    """Kernel-layout weights.  Packing: q|k|v fused, gate|up interleaved for the fused SwiGLU epilogue
    (intermediate zero-padded to a multiple of 64: 3420 -> 3456 for the 7B vision MLP), conv3d patch
    embed flattened to [hidden, 1176]."""

    def __init__(self, cfg: VLMConfig):
        self.cfg = cfg
        self.vis_blocks = []
        self.layers = []

    @classmethod
    def from_state_dict(cls, cfg: VLMConfig, sd: dict, free_source: bool = False) -> "VLMWeights":
        self = cls(cfg)
        v, t = cfg.vision, cfg.text
        # published Qwen2.5-VL checkpoints carry the legacy names (`visual.*`, `model.layers.*`); `from_pretrained`
        # renames them on load (HF conversion_mapping), so the same renaming happens here
        names = {normalize_checkpoint_key(k): k for k in sd}
        wanted = expected_state_dict_keys(cfg)
        missing = sorted(k for k in wanted if k not in names)
        if missing:
            raise KeyError(f"checkpoint lacks {len(missing)} of {len(wanted)} tensors this config needs, e.g. "
                           f"{missing[:6]} (has e.g. {sorted(sd)[:3]})")

        def take(name):
            src = names[name]
            x = sd[src]
            if free_source:
                del sd[src]
            return x.to(BF)

        self.patch_embed = take("model.visual.patch_embed.proj.weight").reshape(v.hidden, v.patch_dim).contiguous()
        vip = v.intermediate_padded
        for i in range(v.depth):
            p = f"model.visual.blocks.{i}."
            blk = dict(
                norm1=take(p + "norm1.weight").contiguous(), norm2=take(p + "norm2.weight").contiguous(),
                qkv_w=take(p + "attn.qkv.weight").contiguous(), qkv_b=take(p + "attn.qkv.bias").contiguous(),
                proj_w=take(p + "attn.proj.weight").contiguous(), proj_b=take(p + "attn.proj.bias").contiguous(),
                gu_w=_pack_swiglu(take(p + "mlp.gate_proj.weight"), take(p + "mlp.up_proj.weight"), vip),
                gu_b=_pack_swiglu(take(p + "mlp.gate_proj.bias"), take(p + "mlp.up_proj.bias"), vip),
                down_w=_pad_cols(take(p + "mlp.down_proj.weight"), vip), down_b=take(p + "mlp.down_proj.bias").contiguous())
            self.vis_blocks.append(blk)
        self.merger_ln = take("model.visual.merger.ln_q.weight").contiguous()
        self.merger_w0 = take("model.visual.merger.mlp.0.weight").contiguous()
        self.merger_b0 = take("model.visual.merger.mlp.0.bias").contiguous()
        self.merger_w2 = take("model.visual.merger.mlp.2.weight").contiguous()
        self.merger_b2 = take("model.visual.merger.mlp.2.bias").contiguous()
        self.embed = take("model.language_model.embed_tokens.weight").contiguous()
        tip = t.intermediate_padded
        for i in range(t.layers):
            p = f"model.language_model.layers.{i}."
            lay = dict(
                ln1=take(p + "input_layernorm.weight").contiguous(), ln2=take(p + "post_attention_layernorm.weight").contiguous(),
                qkv_w=torch.cat([take(p + "self_attn.q_proj.weight"), take(p + "self_attn.k_proj.weight"),
                                 take(p + "self_attn.v_proj.weight")]).contiguous(),
                qkv_b=torch.cat([take(p + "self_attn.q_proj.bias"), take(p + "self_attn.k_proj.bias"),
                                 take(p + "self_attn.v_proj.bias")]).contiguous(),
                o_w=take(p + "self_attn.o_proj.weight").contiguous(),
                gu_w=_pack_swiglu(take(p + "mlp.gate_proj.weight"), take(p + "mlp.up_proj.weight"), tip),
                down_w=_pad_cols(take(p + "mlp.down_proj.weight"), tip))
            self.layers.append(lay)
        self.final_norm = take("model.language_model.norm.weight").contiguous()
        self.lm_head = take("lm_head.weight").contiguous()
        self.device = self.lm_head.device
        return self

    @classmethod
    def random(cls, cfg: VLMConfig, device, seed: int = 0, **kw) -> "VLMWeights":
        return cls.from_state_dict(cfg, random_state_dict(cfg, device, seed, **kw), free_source=True)

    def decode_weight_bytes(self) -> int:
        """Bytes of weights every decode step streams (algorithmic HBM traffic of one step)."""
        n = self.lm_head.numel() + self.final_norm.numel()
        for lay in self.layers:
            n += sum(x.numel() for x in lay.values())
        return 2 * n


# ───────────────────────── vision tower ─────────────────────────
class VisionPlan:
    """Per-grid index tables (window permutation, cu_seqlens, fp32 rope cos/sin) for a batch of
    n same-sized images; cached by the engine."""

    def __init__(self, cfg: VLMConfig, grid_hw: tuple, n_img: int, device):
        v = cfg.vision
        gh, gw = grid_hw
        S = gh * gw
        widx, cu_win = window_index(grid_hw, v.merge, v.window, v.patch)
        G = S // (v.merge ** 2)
        self.S, self.G, self.n_img, self.grid_hw = S, G, n_img, grid_hw
        # group permutation over the whole batch: output group -> source group
        perm = np.concatenate([widx + i * G for i in range(n_img)]).astype(np.int32)
        self.group_perm = torch.from_numpy(perm).to(device)
        cu_all = np.concatenate([[0]] + [cu_win[1:] + i * S for i in range(n_img)]).astype(np.int32)
        self.cu_window = torch.from_numpy(cu_all).to(device)
        self.n_windows = len(cu_all) - 1
        self.max_window = int(np.max(np.diff(cu_all)))
        self.cu_full = torch.arange(0, (n_img + 1) * S, S, dtype=torch.int32, device=device)
        # rope tables exactly as HF builds them (fp32, on the device)
        hd = v.head_dim
        dim = hd // 2
        inv_freq = 1.0 / (10000.0 ** (torch.arange(0, dim, 2, dtype=torch.float) / dim))
        seq = torch.arange(max(gh, gw), dtype=inv_freq.dtype)
        freqs_full = torch.outer(seq, inv_freq).to(device)                      # [max_grid, dim/2]
        pos = torch.from_numpy(vision_rope_pos(grid_hw, v.merge)).to(device)    # [S, 2]
        rot = freqs_full[pos].flatten(1)                                        # [S, dim]
        rot = rot.reshape(G, v.merge ** 2, -1)[torch.from_numpy(widx).to(device)].reshape(S, -1)
        emb = torch.cat((rot, rot), dim=-1)
        self.cos = emb.cos().repeat(n_img, 1).contiguous()
        self.sin = emb.sin().repeat(n_img, 1).contiguous()
        # inverse permutation of the merged tokens: window-order row r holds source group perm[r]
        self.merged_src_group = self.group_perm


def vision_forward(w: VLMWeights, plan: VisionPlan, pixel_values: torch.Tensor) -> torch.Tensor:
    """pixel_values: bf16 [n*S, 1176] ALREADY in window order (normalize_patchify with plan.group_perm).
    Returns merged image embeddings bf16 [n*S/4, out_hidden] in WINDOW order (row r = source group
    plan.group_perm[r]); the caller scatters them into the token embeddings."""
    cfg = w.cfg.vision
    dev = pixel_values.device
    NS = pixel_values.shape[0]
    H, hd, nh = cfg.hidden, cfg.head_dim, cfg.heads
    h = torch.empty((NS, H), dtype=BF, device=dev)
    xn = torch.empty((NS, H), dtype=BF, device=dev)
    qkv = torch.empty((NS, 3 * H), dtype=BF, device=dev)
    att = torch.empty((NS, H), dtype=BF, device=dev)
    act = torch.empty((NS, cfg.intermediate_padded), dtype=BF, device=dev)
    gemm(pixel_values, w.patch_embed, h)
    for i, blk in enumerate(w.vis_blocks):
        rmsnorm(h, blk["norm1"], xn)
        gemm(xn, blk["qkv_w"], qkv, bias=blk["qkv_b"])
        _lib.call("ocrb_rope_vision", qkv.data_ptr(), NS, nh, hd, plan.cos.data_ptr(), plan.sin.data_ptr(), _sp())
        q, k, v = qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:]
        if i in cfg.fullatt_blocks:
            attention(q, k, v, att, plan.cu_full, plan.n_img, plan.S, nh, nh, hd, False)
        else:
            attention(q, k, v, att, plan.cu_window, plan.n_windows, plan.max_window, nh, nh, hd, False)
        gemm(att, blk["proj_w"], h, bias=blk["proj_b"], residual=h, epilogue=EPI_RESIDUAL)
        rmsnorm(h, blk["norm2"], xn)
        gemm(xn, blk["gu_w"], act, bias=blk["gu_b"], epilogue=EPI_SWIGLU)
        gemm(act, blk["down_w"], h, bias=blk["down_b"], residual=h, epilogue=EPI_RESIDUAL)
    rmsnorm(h, w.merger_ln, xn)
    mu = cfg.merge ** 2
    x4 = xn.view(NS // mu, H * mu)
    mid = torch.empty((NS // mu, H * mu), dtype=BF, device=dev)
    out = torch.empty((NS // mu, cfg.out_hidden), dtype=BF, device=dev)
    linear_small_or_big(x4, w.merger_w0, mid, bias=w.merger_b0, epilogue=EPI_GELU)
    linear_small_or_big(mid, w.merger_w2, out, bias=w.merger_b2)
    return out


# ───────────────────────── decoder ─────────────────────────
def text_rope_tables(cfg: VLMConfig, pos3: torch.Tensor):
    """HF Qwen2_5_VLRotaryEmbedding + the mrope section mix (modeling :596-608, :659-669): pos3 int64
    [3, T] on the device -> bf16 cos/sin [T, head_dim]."""
    t = cfg.text
    hd = t.head_dim
    inv_freq = (1.0 / (t.rope_theta ** (torch.arange(0, hd, 2, dtype=torch.int64).to(dtype=torch.float) / hd))).to(pos3.device)
    freqs = (inv_freq[None, :, None].float().expand(3, -1, 1) @ pos3[:, None, :].float()).transpose(1, 2)  # [3, T, hd/2]
    emb = torch.cat((freqs, freqs), dim=-1)
    cos, sin = emb.cos().to(BF), emb.sin().to(BF)
    sec = list(t.mrope_section) * 2
    cos = torch.cat([m[i % 3] for i, m in enumerate(cos.split(sec, dim=-1))], dim=-1)
    sin = torch.cat([m[i % 3] for i, m in enumerate(sin.split(sec, dim=-1))], dim=-1)
    return cos.contiguous(), sin.contiguous(), inv_freq.contiguous()


class PagedKV:
    """Paged KV cache for all layers: k/v [layers, n_pages, n_kv, page, hd] bf16 (the 16 tokens of a (page, kv head) are
    one contiguous 4 KiB block: one TMA box per attention tile) + a free list."""

    def __init__(self, cfg: VLMConfig, n_pages: int, page_size: int, device):
        t = cfg.text
        self.page = page_size
        self.n_pages = n_pages
        self.k = torch.zeros((t.layers, n_pages, t.kv_heads, page_size, t.head_dim), dtype=BF, device=device)
        self.v = torch.zeros_like(self.k)
        self.free = list(range(n_pages - 1, -1, -1))

    def alloc(self, n: int) -> list:
        if n > len(self.free):
            raise _lib.OcrbError(f"paged KV cache exhausted: need {n} pages, {len(self.free)} free")
        return [self.free.pop() for _ in range(n)]

    def release(self, pages) -> None:
        self.free.extend(int(p) for p in pages)


class Decoder:
    """Prefill + batched greedy decode over the paged KV cache.  One decode step is a fixed sequence
    of kernel launches whose per-step state (context lengths, next ids, step counter) lives on the
    device, so the step is captured once in a CUDA graph and replayed."""

    def __init__(self, w: VLMWeights, kv: PagedKV, max_batch: int, max_ctx: int, tp=None):
        self.w, self.kv, self.cfg = w, kv, w.cfg
        self.max_batch, self.max_ctx = max_batch, max_ctx
        self.replayed_launches = 0      # kernel launches issued through CUDA-graph replays (bench `gpu_launches`)
        self.tp = tp                    # tp.TPComm for the tensor-parallel large-VLM config, else None
        self._tp_tmp = {}

    def _row_parallel(self, X, W, h):
        """h += X @ W^T for a row-parallel linear (o_proj / down_proj).  Single GPU: residual fused in the GEMM
        epilogue.  Tensor parallel (HF base_model_tp_plan 'rowwise'): every rank holds a K-slice, the bf16 partial
        products are summed over ranks (NCCL all-reduce over NVLink), then the residual is added."""
        if self.tp is None:
            return linear_small_or_big(X, W, h, residual=h, epilogue=EPI_RESIDUAL)
        peer = getattr(self.tp, "peer", None)
        if peer is not None and X.shape[0] <= peer.MAX_ROWS:
            # decode: the GEMM writes its partial into the peer-mapped slot and, on the cluster kernel, exchanges it with
            # the peers and finishes the residual stream in its own epilogue (csrc/skinny.cu); tiny shapes: one more kernel
            self.tp.n_all_reduce += 1
            return peer.row_parallel(X, W, h, skinny_workspace(X.device))
        key = (X.shape[0], h.shape[1])
        tmp = self._tp_tmp.get(key)
        if tmp is None:
            tmp = torch.empty(key, dtype=BF, device=h.device)
            self._tp_tmp[key] = tmp
        linear_small_or_big(X, W, tmp)
        self.tp.all_reduce(tmp)
        return residual_add(h, tmp)

    # ---- prefill: T_total tokens of n_seq sequences (cu_seqlens), embeddings already assembled ----
    def prefill(self, h: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, cu: torch.Tensor, n_seq: int,
                max_len: int, block_table: torch.Tensor):
        t = self.cfg.text
        dev = h.device
        TT = h.shape[0]
        nq, nkv, hd, H = t.heads, t.kv_heads, t.head_dim, t.hidden
        qkv_dim = (nq + 2 * nkv) * hd
        xn = torch.empty_like(h)
        qkv = torch.empty((TT, qkv_dim), dtype=BF, device=dev)
        att = torch.empty((TT, nq * hd), dtype=BF, device=dev)
        act = torch.empty((TT, t.intermediate_padded), dtype=BF, device=dev)
        max_pages = block_table.shape[1]
        for li, lay in enumerate(self.w.layers):
            rmsnorm(h, lay["ln1"], xn, t.rms_eps)
            linear_small_or_big(xn, lay["qkv_w"], qkv, bias=lay["qkv_b"])
            q, k, v = qkv[:, : nq * hd], qkv[:, nq * hd: (nq + nkv) * hd], qkv[:, (nq + nkv) * hd:]
            _lib.call("ocrb_rope_text", q.data_ptr(), qkv.stride(0), k.data_ptr(), qkv.stride(0), TT, nq, nkv, hd,
                      cos.data_ptr(), sin.data_ptr(), _sp())
            _lib.call("ocrb_kv_write_prefill", k.data_ptr(), qkv.stride(0), v.data_ptr(), qkv.stride(0),
                      self.kv.k[li].data_ptr(), self.kv.v[li].data_ptr(), block_table.data_ptr(), max_pages,
                      cu.data_ptr(), n_seq, TT, self.kv.page, nkv, hd, _sp())
            attention(q, k, v, att, cu, n_seq, max_len, nq, nkv, hd, True)
            self._row_parallel(att, lay["o_w"], h)
            rmsnorm(h, lay["ln2"], xn, t.rms_eps)
            linear_small_or_big(xn, lay["gu_w"], act, epilogue=EPI_SWIGLU)
            self._row_parallel(act, lay["down_w"], h)
        return h

    def logits_last(self, h_last: torch.Tensor, out: torch.Tensor):
        """final norm + lm_head for <= 8 rows (fused in the weight-streaming kernel)."""
        B = h_last.shape[0]
        for b0 in range(0, B, SKINNY_MAX_ROWS):
            sl = slice(b0, b0 + SKINNY_MAX_ROWS)
            skinny(h_last[sl], self.w.lm_head, out[sl], norm_w=self.w.final_norm, eps=self.cfg.text.rms_eps)
        return out

    # ---- one decode step for B sequences (all state on the device) ----
    def _step_chained(self, st: "DecodeState"):
        """The step as 1 + layers persistent chain launches: [qkv 0] attn [o, gate/up, down, qkv 1] attn ... [o, gate/up,
        down, lm_head] argmax.  The RMSNorms run inside the chains; only the attention sits between two launches."""
        t = self.cfg.text
        nq, nkv, hd = t.heads, t.kv_heads, t.head_dim
        B, dev, eps = st.B, st.x.device, t.rms_eps
        L = self.w.layers
        _lib.call("ocrb_embed_gather", self.w.embed.data_ptr(), st.next_ids.data_ptr(), st.x.data_ptr(), B, t.hidden, _sp())
        _lib.call("ocrb_decode_rope_table", st.ctx_len.data_ptr(), st.rope_delta.data_ptr(), st.inv_freq.data_ptr(), B, hd,
                  st.cos.data_ptr(), st.sin.data_ptr(), _sp())
        max_pages = st.block_table.shape[1]

        def qkv_of(lay):
            return chain_linear(st.x, lay["qkv_w"], st.qkv, bias=lay["qkv_b"], norm_w=lay["ln1"], eps=eps)

        skinny_chain([qkv_of(L[0])], B, dev)
        for li, lay in enumerate(L):
            _lib.call("ocrb_decode_attention", st.qkv.data_ptr(), st.qkv.stride(0), self.kv.k[li].data_ptr(),
                      self.kv.v[li].data_ptr(), self.kv.n_pages, st.block_table.data_ptr(), max_pages, st.ctx_len.data_ptr(), B,
                      self.kv.page, nq, nkv, hd, st.cos.data_ptr(), st.sin.data_ptr(), float(hd ** -0.5),
                      st.att.data_ptr(), st.att.stride(0), st.split_ws.data_ptr(), st.n_splits, _sp())
            lins = [chain_linear(st.att, lay["o_w"], st.x, residual=st.x, epilogue=EPI_RESIDUAL),
                    chain_linear(st.x, lay["gu_w"], st.act, epilogue=EPI_SWIGLU, norm_w=lay["ln2"], eps=eps),
                    chain_linear(st.act, lay["down_w"], st.x, residual=st.x, epilogue=EPI_RESIDUAL)]
            if li + 1 < len(L):
                lins.append(qkv_of(L[li + 1]))
            else:
                lins.append(chain_linear(st.x, self.w.lm_head, st.logits_local, norm_w=self.w.final_norm, eps=eps))
            skinny_chain(lins, B, dev)
        _lib.call("ocrb_argmax_step", st.logits.data_ptr(), st.logits.stride(0), B, t.vocab, EOS, EOS, st.max_new,
                  st.out_tokens.data_ptr(), st.next_ids.data_ptr(), st.finished.data_ptr(), st.ctx_len.data_ptr(),
                  st.step.data_ptr(), 1, _sp())

    def _build_plan(self, st: "DecodeState"):
        """The whole step as one plan: per layer [RMSNorm + qkv, attention, o_proj + residual, RMSNorm + gate/up + SwiGLU,
        down_proj + residual], then final norm + lm_head.  Every pointer is baked into the device-side plan."""
        t = self.cfg.text
        nq, nkv, hd = t.heads, t.kv_heads, t.head_dim
        B, eps = st.B, t.rms_eps
        max_pages = st.block_table.shape[1]
        ops = []

        def lin(*a, **kw):
            op = _lib.ChainOp()
            op.kind = 0
            op.lin = chain_linear(*a, **kw)
            ops.append(op)

        for li, lay in enumerate(self.w.layers):
            lin(st.x, lay["qkv_w"], st.qkv, bias=lay["qkv_b"], norm_w=lay["ln1"], eps=eps)
            op = _lib.ChainOp()
            op.kind = 1
            op.att = _lib.ChainAttention(st.qkv.data_ptr(), st.qkv.stride(0), self.kv.k[li].data_ptr(), self.kv.v[li].data_ptr(),
                                         self.kv.n_pages, st.block_table.data_ptr(), max_pages, st.ctx_len.data_ptr(),
                                         self.kv.page, nq, nkv, hd, st.cos.data_ptr(), st.sin.data_ptr(), float(hd ** -0.5),
                                         st.att.data_ptr(), st.att.stride(0), st.split_ws.data_ptr(), st.n_splits)
            ops.append(op)
            lin(st.att, lay["o_w"], st.x, residual=st.x, epilogue=EPI_RESIDUAL)
            lin(st.x, lay["gu_w"], st.act, epilogue=EPI_SWIGLU, norm_w=lay["ln2"], eps=eps)
            lin(st.act, lay["down_w"], st.x, residual=st.x, epilogue=EPI_RESIDUAL)
        lin(st.x, self.w.lm_head, st.logits_local, norm_w=self.w.final_norm, eps=eps)
        n = len(ops)
        arr = (_lib.ChainOp * n)(*ops)
        plan = torch.zeros(int(_lib.load().ocrb_chain_plan_bytes(n)) + 64, dtype=torch.uint8, device=st.x.device)
        off = (-plan.data_ptr()) % 64
        _lib.call("ocrb_chain_plan_build", ctypes.addressof(arr), n, B, chain_workspace(st.x.device).data_ptr(),
                  plan.data_ptr() + off, _sp())
        st.plan, st.plan_ptr, st.plan_ops = plan, plan.data_ptr() + off, n

    def _step_fused(self, st: "DecodeState"):
        """One decode step = embedding gather, rope table, ONE persistent plan launch, argmax."""
        t = self.cfg.text
        B = st.B
        if st.plan is None:
            self._build_plan(st)
        _lib.call("ocrb_embed_gather", self.w.embed.data_ptr(), st.next_ids.data_ptr(), st.x.data_ptr(), B, t.hidden, _sp())
        _lib.call("ocrb_decode_rope_table", st.ctx_len.data_ptr(), st.rope_delta.data_ptr(), st.inv_freq.data_ptr(), B, t.head_dim,
                  st.cos.data_ptr(), st.sin.data_ptr(), _sp())
        dev = st.x.device
        if CHAIN_TRACE is not None:
            slots = 8 + 8 * st.plan_ops
            buf = torch.zeros(296 * slots, dtype=torch.int64, device=dev)
            CHAIN_TRACE.append(buf)
            _lib.load().ocrb_chain_set_trace(ctypes.c_void_p(buf.data_ptr()), slots)
        _lib.call("ocrb_chain_plan_run", st.plan_ptr, st.plan_ops, B, chain_workspace(dev).data_ptr(), _sp())
        _lib.call("ocrb_argmax_step", st.logits.data_ptr(), st.logits.stride(0), B, t.vocab, EOS, EOS, st.max_new,
                  st.out_tokens.data_ptr(), st.next_ids.data_ptr(), st.finished.data_ptr(), st.ctx_len.data_ptr(),
                  st.step.data_ptr(), 1, _sp())

    def _step(self, st: "DecodeState"):
        t = self.cfg.text
        nq, nkv, hd = t.heads, t.kv_heads, t.head_dim
        B = st.B
        if self.tp is None and B <= CHAIN_MAX_B and t.hidden <= 8192:
            if CHAIN_FUSE_ATTN and hd == 128 and 1 + 5 * len(self.w.layers) <= 192:
                return self._step_fused(st)
            return self._step_chained(st)
        _lib.call("ocrb_embed_gather", self.w.embed.data_ptr(), st.next_ids.data_ptr(), st.x.data_ptr(), B, t.hidden, _sp())
        _lib.call("ocrb_decode_rope_table", st.ctx_len.data_ptr(), st.rope_delta.data_ptr(), st.inv_freq.data_ptr(), B, hd,
                  st.cos.data_ptr(), st.sin.data_ptr(), _sp())
        max_pages = st.block_table.shape[1]
        for li, lay in enumerate(self.w.layers):
            for b0 in range(0, B, SKINNY_MAX_ROWS):
                sl = slice(b0, min(B, b0 + SKINNY_MAX_ROWS))
                skinny(st.x[sl], lay["qkv_w"], st.qkv[sl], bias=lay["qkv_b"], norm_w=lay["ln1"], eps=t.rms_eps)
            _lib.call("ocrb_decode_attention", st.qkv.data_ptr(), st.qkv.stride(0), self.kv.k[li].data_ptr(),
                      self.kv.v[li].data_ptr(), self.kv.n_pages, st.block_table.data_ptr(), max_pages, st.ctx_len.data_ptr(), B,
                      self.kv.page, nq, nkv, hd, st.cos.data_ptr(), st.sin.data_ptr(), float(hd ** -0.5),
                      st.att.data_ptr(), st.att.stride(0), st.split_ws.data_ptr(), st.n_splits, _sp())
            for b0 in range(0, B, SKINNY_MAX_ROWS):
                sl = slice(b0, min(B, b0 + SKINNY_MAX_ROWS))
                self._row_parallel(st.att[sl], lay["o_w"], st.x[sl])
                skinny(st.x[sl], lay["gu_w"], st.act[sl], epilogue=EPI_SWIGLU, norm_w=lay["ln2"], eps=t.rms_eps)
                self._row_parallel(st.act[sl], lay["down_w"], st.x[sl])
        self.logits_last(st.x, st.logits_local)
        if self.tp is not None:
            peer = getattr(self.tp, "peer", None)
            if peer is not None and B <= peer.MAX_ROWS and not TP_ARGMAX_GATHER:
                # vocab-split lm_head: (max, lowest index) pairs through peer memory instead of gathering B x V logits
                return peer.argmax_step(st.logits_local, B, EOS, EOS, st.max_new, st.out_tokens, st.next_ids, st.finished,
                                        st.ctx_len, st.step, 1)
            self.tp.gather_vocab(st.logits_local, st.logits)     # NCCL route: full [B, V] logits on every rank
        _lib.call("ocrb_argmax_step", st.logits.data_ptr(), st.logits.stride(0), B, t.vocab, EOS, EOS, st.max_new,
                  st.out_tokens.data_ptr(), st.next_ids.data_ptr(), st.finished.data_ptr(), st.ctx_len.data_ptr(),
                  st.step.data_ptr(), 1, _sp())

    def time_weight_stream(self, B: int, reps: int = 3) -> dict:
        """Time the weight-streaming launches of one decode step alone (every layer's qkv / o / gate-up /
        down GEMV + lm_head, back to back, CUDA events on the launching stream).  The 14 GB of weights
        far exceed L2, so every rep streams from HBM.  Returns achieved GB/s on the algorithmic bytes
        (each weight byte once per step)."""
        t = self.cfg.text
        dev = self.w.device
        x = torch.randn((B, t.hidden), device=dev).to(BF)
        qkv = torch.empty((B, (t.heads + 2 * t.kv_heads) * t.head_dim), dtype=BF, device=dev)
        att = torch.randn((B, t.heads * t.head_dim), device=dev).to(BF)
        act = torch.empty((B, t.intermediate_padded), dtype=BF, device=dev)
        y = torch.empty_like(x)
        logits = torch.empty((B, self.w.lm_head.shape[0]), dtype=BF, device=dev)

        def one():
            for lay in self.w.layers:
                for b0 in range(0, B, SKINNY_MAX_ROWS):
                    sl = slice(b0, min(B, b0 + SKINNY_MAX_ROWS))
                    skinny(x[sl], lay["qkv_w"], qkv[sl], bias=lay["qkv_b"], norm_w=lay["ln1"], eps=t.rms_eps)
                    skinny(att[sl], lay["o_w"], y[sl], residual=x[sl], epilogue=EPI_RESIDUAL)
                    skinny(x[sl], lay["gu_w"], act[sl], epilogue=EPI_SWIGLU, norm_w=lay["ln2"], eps=t.rms_eps)
                    skinny(act[sl], lay["down_w"], y[sl], residual=x[sl], epilogue=EPI_RESIDUAL)
            self.logits_last(x, logits)

        one()
        before = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            one()
        e1.record()
        torch.cuda.synchronize()
        launches = (_lib.launch_count() - before) // reps
        ms = e0.elapsed_time(e1) / reps
        nbytes = self.w.decode_weight_bytes()
        return {"ms": ms, "launches": launches, "avg_us": ms * 1e3 / max(launches, 1), "bytes": nbytes,
                "gbs": nbytes / (ms * 1e-3) / 1e9}

    def decode(self, st: "DecodeState", n_steps: int, use_graph: bool = True, check_every: int = 64):
        """Run up to n_steps decode steps; stops early when every sequence has emitted EOS."""
        if n_steps <= 0:
            return
        if not use_graph:
            for i in range(n_steps):
                self._step(st)
                if (i + 1) % check_every == 0 and bool(st.finished.all()):
                    break
            return
        g = st.graph
        if g is None:
            # warm-up launch outside capture is not possible without side effects on the state, so
            # snapshot and restore the small per-step state around it
            snap = st.snapshot()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._step(st)
            torch.cuda.current_stream().wait_stream(s)
            st.restore(snap)
            g = torch.cuda.CUDAGraph()
            before = _lib.launch_count()
            with torch.cuda.graph(g):
                self._step(st)
            st.graph_launches = _lib.launch_count() - before
            st.restore(snap)
            st.graph = g
            if self.tp is not None:
                # a rank that is still capturing must not be waited for inside the peer all-reduce of a replaying rank
                torch.cuda.synchronize()
                self.tp.dist.barrier(group=self.tp.group)
        for i in range(n_steps):
            g.replay()
            self.replayed_launches += st.graph_launches
            if (i + 1) % check_every == 0 and bool(st.finished.all()):
                break


class DecodeState:
    """Device-resident state of one batch of sequences being decoded."""

    def __init__(self, dec: Decoder, B: int, max_new: int, block_table: torch.Tensor, ctx_len, rope_delta,
                 inv_freq: torch.Tensor):
        t = dec.cfg.text
        dev = dec.w.device
        self.B, self.max_new = B, max_new
        self.block_table = block_table
        self.ctx_len = torch.as_tensor(ctx_len, dtype=torch.int32, device=dev).clone()
        self.rope_delta = torch.as_tensor(rope_delta, dtype=torch.int32, device=dev).clone()
        self.inv_freq = inv_freq
        self.next_ids = torch.zeros(B, dtype=torch.int32, device=dev)
        self.finished = torch.zeros(B, dtype=torch.int32, device=dev)
        self.step = torch.zeros(1, dtype=torch.int32, device=dev)
        self.out_tokens = torch.full((B, max_new), EOS, dtype=torch.int32, device=dev)
        qkv_dim = (t.heads + 2 * t.kv_heads) * t.head_dim
        self.x = torch.empty((B, t.hidden), dtype=BF, device=dev)
        self.qkv = torch.empty((B, qkv_dim), dtype=BF, device=dev)
        self.att = torch.empty((B, t.heads * t.head_dim), dtype=BF, device=dev)
        self.act = torch.empty((B, t.intermediate_padded), dtype=BF, device=dev)
        self.logits = torch.empty((B, t.vocab), dtype=BF, device=dev)
        v_local = dec.w.lm_head.shape[0]          # == vocab unless the lm_head is vocab-split (tensor parallel)
        self.logits_local = self.logits if v_local == t.vocab else torch.empty((B, v_local), dtype=BF, device=dev)
        self.cos = torch.empty((B, t.head_dim), dtype=BF, device=dev)
        self.sin = torch.empty((B, t.head_dim), dtype=BF, device=dev)
        max_ctx = block_table.shape[1] * dec.kv.page
        # split-KV: one CTA per (key range, kv head, sequence); ranges beyond the context exit at once.  The range length
        # is a property of the engine (never of B), so a sequence's bits do not depend on its batch.
        self.n_splits = int(max(1, math.ceil(max_ctx / DECODE_KEYS_PER_CTA)))
        self.split_ws = torch.empty(B * t.heads * self.n_splits * (t.head_dim + 2), dtype=torch.float32, device=dev)
        self.graph = None
        self.graph_launches = 0
        self.plan = None               # device-side plan of the fused step (built at the first step)
        self.plan_ptr, self.plan_ops = 0, 0

    def snapshot(self):
        return [x.clone() for x in (self.ctx_len, self.next_ids, self.finished, self.step, self.out_tokens)]

    def restore(self, snap):
        for dst, src in zip((self.ctx_len, self.next_ids, self.finished, self.step, self.out_tokens), snap):
            dst.copy_(src)
