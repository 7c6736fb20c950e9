"""Tensor parallelism for the large (72B-class) VLM config (BASELINE.json configs[4]).

The reference never shards anything (`device_map=device`, ocr_agent/tools.py:692-709); the plan here is the one
HF ships for the model's text layers and never activates (`base_model_tp_plan`,
HF:models/qwen2_5_vl/configuration_qwen2_5_vl.py:90-98):

    q_proj / k_proj / v_proj / gate_proj / up_proj : column-parallel (output rows split; heads stay whole,
                                                     q heads follow their KV head)
    o_proj / down_proj                             : row-parallel (input columns split) -> all-reduce(sum)
    lm_head                                        : vocab-split -> decode steps exchange one (max, lowest index) pair
                                                     per rank and sequence through peer memory (csrc/comm.cu);
                                                     the prefill step, whose logits the parity checks read, still
                                                     all-gathers them
    embeddings, norms, vision tower                : replicated

One process per GPU; NCCL over NVLink/NVSwitch carries the two all-reduces per layer (`[rows, hidden]` bf16) and
one small all-gather per step.  Everything else (weights-streaming GEMMs, attention over the rank's heads,
paged KV of the rank's KV heads) is the single-GPU code on local shapes.
"""
from __future__ import annotations

import copy
import os

import torch

from .vlm_config import TextCfg, VLMConfig

BF = torch.bfloat16


def local_config(cfg: VLMConfig, world: int) -> VLMConfig:
    """Per-rank dimensions: heads, KV heads and MLP width divided by `world`; hidden, vocab, vision unchanged."""
    t = cfg.text
    if t.heads % world or t.kv_heads % world or t.intermediate % world or t.vocab % world:
        raise ValueError(f"TP world {world} does not divide heads {t.heads} / kv {t.kv_heads} / "
                         f"intermediate {t.intermediate} / vocab {t.vocab}")
    lt = TextCfg(hidden=t.hidden, layers=t.layers, heads=t.heads // world, kv_heads=t.kv_heads // world,
                 head_dim=t.head_dim, intermediate=t.intermediate // world, vocab=t.vocab, rms_eps=t.rms_eps,
                 rope_theta=t.rope_theta, mrope_section=t.mrope_section)
    return VLMConfig(vision=copy.deepcopy(cfg.vision), text=lt, name=f"{cfg.name}/tp{world}")


_ROW_SPLIT = ("self_attn.q_proj.weight", "self_attn.q_proj.bias", "self_attn.k_proj.weight", "self_attn.k_proj.bias",
              "self_attn.v_proj.weight", "self_attn.v_proj.bias", "mlp.gate_proj.weight", "mlp.up_proj.weight")
_COL_SPLIT = ("self_attn.o_proj.weight", "mlp.down_proj.weight")


def shard_state_dict(sd: dict, rank: int, world: int) -> dict:
    """Slice a full HF-named state dict into rank `rank`'s shard (the text decoder and lm_head; the rest is
    replicated).  Equal contiguous slices: with heads laid out head-major this keeps whole heads together and
    gives rank r the KV heads its q heads attend to."""
    out = {}
    for name, w in sd.items():
        if name.startswith("model.language_model.layers.") and name.endswith(_ROW_SPLIT) or name == "lm_head.weight":
            n = w.shape[0] // world
            out[name] = w[rank * n:(rank + 1) * n].contiguous()
        elif name.startswith("model.language_model.layers.") and name.endswith(_COL_SPLIT):
            n = w.shape[1] // world
            out[name] = w[:, rank * n:(rank + 1) * n].contiguous()
        else:
            out[name] = w
    return out


class TPComm:
    """The collectives of the tensor-parallel decoder on a torch.distributed group (NCCL on GPUs, gloo in the
    CPU tests)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self._gather = {}
        self.n_all_reduce = 0
        self.peer = None            # PeerAllReduce for the decode step's small messages (enable_peer_all_reduce)

    def enable_peer_all_reduce(self, device, hidden: int):
        """Replace NCCL by the one-shot peer-memory kernel for messages of <= 64 rows (needs CUDA IPC between the
        ranks' processes on one NVLink-connected box).  OCRB_TP_NCCL=1 keeps NCCL for everything."""
        import os
        if os.environ.get("OCRB_TP_NCCL") == "1":
            return None
        self.peer = PeerAllReduce(self, device, hidden)
        return self.peer

    def all_reduce(self, t: torch.Tensor) -> torch.Tensor:
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        self.n_all_reduce += 1
        return t

    def gather_vocab(self, local: torch.Tensor, full: torch.Tensor) -> torch.Tensor:
        """local: [B, V/world] logits of this rank's vocab slice -> full: [B, V] (slices in rank order)."""
        B, vl = local.shape
        buf = self._gather.get((B, vl))
        if buf is None:
            buf = torch.empty((self.world * B, vl), dtype=local.dtype, device=local.device)   # rank-major concatenation
            self._gather[(B, vl)] = buf
        self.dist.all_gather_into_tensor(buf, local.contiguous(), group=self.group)
        full.view(B, self.world, vl).copy_(buf.view(self.world, B, vl).permute(1, 0, 2))
        return full


class PeerAllReduce:
    """One-shot all-reduce + residual add over NVLink peer memory for the small [B, hidden] messages of the decode
    step (csrc/comm.cu).  Every rank allocates two partial-sum slots and a flag array, shares them through CUDA IPC
    (handles exchanged with all_gather_object), and maps its peers'.  The row-parallel GEMM writes its partial
    straight into the current slot; `all_reduce_residual` then does flags + reads + sum + residual in one kernel.
    Calls alternate between the two slots, which is safe without a second barrier (see comm.cu)."""

    MAX_ROWS = 64

    def __init__(self, comm: "TPComm", device, hidden: int):
        import ctypes
        from . import _lib
        self._ct, self._lib = ctypes, _lib
        self.comm, self.world, self.rank = comm, comm.world, comm.rank
        self.hidden = hidden
        self.local = torch.zeros((2, self.MAX_ROWS, hidden), dtype=BF, device=device)
        self.flags = torch.zeros(16 * 8, dtype=torch.int32, device=device)
        self.seq = torch.zeros(16, dtype=torch.int32, device=device)
        # (max, index) exchange of the vocab-split lm_head: two pair slots per sequence, one flag per (sequence, rank)
        self.am_pairs = torch.zeros(2 * self.MAX_ROWS, dtype=torch.int64, device=device)
        self.am_flags = torch.zeros(self.MAX_ROWS * 8, dtype=torch.int32, device=device)
        self.am_seq = torch.zeros(self.MAX_ROWS, dtype=torch.int32, device=device)
        # all-reduce fused into the row-parallel GEMM's epilogue: one flag row per (tile, cluster CTA) exchange unit
        self.fz_flags = torch.zeros(512 * 8, dtype=torch.int32, device=device)
        self.fz_seq = torch.zeros(512, dtype=torch.int32, device=device)
        # LL route of the fused exchange: cells of (2 x bf16, call index) pushed into every peer's receive buffer
        self.ll_buf = torch.zeros(2 * 8 * 64 * (hidden // 2), dtype=torch.int64, device=device)
        self.ll_seq = torch.zeros(2, dtype=torch.int32, device=device)
        # 0: GEMM + all-reduce kernel; 1: exchange in the GEMM epilogue, flag + pull; 2: exchange in the epilogue, LL push
        # default 2: measured 17.2-17.5 ms per decode step against 17.8-18.0 (route 0) and 18.3 (route 1) at TP-2 on 72B-class shards
        self.fused = int(os.environ.get("OCRB_TP_FUSED", "2"))
        torch.cuda.synchronize()

        def handle(t):
            h = ctypes.create_string_buffer(64)
            off = ctypes.c_int64()
            _lib.call("ocrb_comm_ipc_handle", t.data_ptr(), h, ctypes.byref(off))
            return (h.raw, off.value)

        mine = {"data": handle(self.local), "flags": handle(self.flags), "am_pairs": handle(self.am_pairs),
                "am_flags": handle(self.am_flags), "fz_flags": handle(self.fz_flags), "ll_buf": handle(self.ll_buf)}
        everyone = [None] * self.world
        comm.dist.all_gather_object(everyone, mine, group=comm.group)
        slot_bytes = self.MAX_ROWS * hidden * 2
        data_base, flag_ptr, am_pair_ptr, am_flag_ptr, fz_flag_ptr, ll_ptr = [], [], [], [], [], []
        for r, item in enumerate(everyone):
            if r == self.rank:
                data_base.append(self.local.data_ptr())
                flag_ptr.append(self.flags.data_ptr())
                am_pair_ptr.append(self.am_pairs.data_ptr())
                am_flag_ptr.append(self.am_flags.data_ptr())
                fz_flag_ptr.append(self.fz_flags.data_ptr())
                ll_ptr.append(self.ll_buf.data_ptr())
                continue
            out = []
            for key in ("data", "flags", "am_pairs", "am_flags", "fz_flags", "ll_buf"):
                raw, off = item[key]
                p = ctypes.c_void_p()
                _lib.call("ocrb_comm_ipc_open", ctypes.create_string_buffer(raw, 64), off, ctypes.byref(p))
                out.append(p.value)
            data_base.append(out[0])
            flag_ptr.append(out[1])
            am_pair_ptr.append(out[2])
            am_flag_ptr.append(out[3])
            fz_flag_ptr.append(out[4])
            ll_ptr.append(out[5])
        arr = ctypes.c_void_p * self.world
        self._data_ptrs = [arr(*[b + s * slot_bytes for b in data_base]) for s in range(2)]
        self._flag_ptrs = arr(*flag_ptr)
        self._am_pair_ptrs = arr(*am_pair_ptr)
        self._am_flag_ptrs = arr(*am_flag_ptr)
        self._fz_flag_ptrs = arr(*fz_flag_ptr)
        self._ll_ptrs = arr(*ll_ptr)
        self.calls = 0
        comm.dist.barrier(group=comm.group)          # nobody launches before every mapping exists

    def next_slot(self) -> int:
        s = self.calls & 1
        self.calls += 1
        return s

    def all_reduce_residual(self, x: torch.Tensor, rows: int, slot: int) -> torch.Tensor:
        """x[rows, hidden] += sum over ranks of local[slot][:rows] (every rank's own partial in its slot)."""
        self._lib.call("ocrb_allreduce_residual_bf16", x.data_ptr(), x.stride(0), self._data_ptrs[slot], self._flag_ptrs,
                       self.world, self.rank, self.seq.data_ptr(), rows, self.hidden, self.hidden,
                       torch.cuda.current_stream().cuda_stream)
        return x


    def row_parallel(self, X: torch.Tensor, W: torch.Tensor, x: torch.Tensor, workspace: torch.Tensor) -> torch.Tensor:
        """x[rows, hidden] += sum over ranks of bf16(X @ W^T) for a row-parallel linear, ONE kernel when the shape runs on the
        cluster GEMM (the exchange is its epilogue), GEMM + `all_reduce_residual` kernel otherwise -- same bits."""
        slot = self.next_slot()
        rows = X.shape[0]
        self._lib.call("ocrb_skinny_rowparallel_tp_bf16", X.data_ptr(), X.stride(0), W.data_ptr(), W.stride(0), rows,
                       W.shape[0], X.shape[1], x.data_ptr(), x.stride(0), self._data_ptrs[slot], self.hidden,
                       self._fz_flag_ptrs, self.fz_seq.data_ptr(), self._flag_ptrs, self.seq.data_ptr(), self._ll_ptrs,
                       self.ll_seq.data_ptr(), self.world, self.rank, workspace.data_ptr(), int(self.fused),
                       torch.cuda.current_stream().cuda_stream)
        return x

    def argmax_step(self, logits_local: torch.Tensor, B: int, eos: int, pad: int, max_new: int, out_tokens, next_ids,
                    finished, ctx_len, step, advance_ctx: int):
        """Greedy token of a vocab-split lm_head from every rank's [B, V / world] slice: one (max, lowest global index) pair
        per rank and sequence through peer memory, then the step bookkeeping -- no all-gather of the logits."""
        self._lib.call("ocrb_tp_argmax_step", logits_local.data_ptr(), logits_local.stride(0), B, logits_local.shape[1],
                       self._am_pair_ptrs, self._am_flag_ptrs, self.world, self.rank, self.am_seq.data_ptr(), self.MAX_ROWS,
                       eos, pad, max_new, out_tokens.data_ptr(), next_ids.data_ptr(), finished.data_ptr(),
                       ctx_len.data_ptr(), step.data_ptr(), advance_ctx, torch.cuda.current_stream().cuda_stream)


def random_weights_tp(cfg: VLMConfig, device, rank: int, world: int, seed: int = 0, **kw):
    """Random-init weights of rank `rank` WITHOUT materialising the full model (a 72B-class model is 146 GB):
    replicated tensors come from the common seed, sharded ones from a rank-specific seed.  Returns
    (VLMWeights on local shapes, local VLMConfig)."""
    from . import vlm
    lcfg = local_config(cfg, world)
    sd = vlm.random_state_dict(lcfg, device, seed, skip_tp_sharded=True, **kw)   # replicated tensors: same on all ranks
    sharded = vlm.random_state_dict_text_only(lcfg, device, seed * 7919 + 1 + rank, vocab_rows=cfg.text.vocab // world, **kw)
    sd.update(sharded)
    w = vlm.VLMWeights.from_state_dict(lcfg, sd, free_source=True)
    return w, lcfg


def sharded_weights_from_full(cfg: VLMConfig, sd_full: dict, rank: int, world: int):
    """Rank shard of a full state dict (small configs / real checkpoints that fit one device)."""
    from . import vlm
    lcfg = local_config(cfg, world)
    return vlm.VLMWeights.from_state_dict(lcfg, shard_state_dict(sd_full, rank, world)), lcfg
