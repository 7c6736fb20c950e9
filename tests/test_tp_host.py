"""Host logic of the tensor-parallel large-VLM config on CPU: the sharding plan reproduces the full linears,
local dimensions are right for the 72B-class config, and the TP collectives work across two gloo ranks."""
import os
import socket

import pytest
import torch


@pytest.fixture(scope="module")
def mods(pkg):
    from handwritten_ocr_b200 import tp, vlm, vlm_config
    return tp, vlm, vlm_config


def test_local_config_72b(mods):
    tp, vlm, vc = mods
    cfg = vc.VLMConfig.qwen72b()
    l8 = tp.local_config(cfg, 8)
    assert (l8.text.heads, l8.text.kv_heads, l8.text.intermediate, l8.text.hidden) == (8, 1, 3696, 8192)
    assert l8.text.intermediate_padded == 3712 and l8.text.vocab == 152064
    # bytes streamed per decode step per rank at TP-8 (17.86 GB)
    t = l8.text
    per_layer = ((t.heads + 2 * t.kv_heads) * t.head_dim + t.heads * t.head_dim) * t.hidden + 3 * t.intermediate * t.hidden
    total = 2 * (t.layers * per_layer + (t.vocab // 8) * t.hidden)
    assert abs(total / 1e9 - 17.86) < 0.3
    with pytest.raises(ValueError):
        tp.local_config(cfg, 3)


def test_shards_reproduce_full_linears(mods):
    tp, vlm, vc = mods
    cfg = vc.VLMConfig.tiny()
    sd = vlm.random_state_dict(cfg, "cpu", seed=3)
    t = cfg.text
    world = 2
    shards = [tp.shard_state_dict(sd, r, world) for r in range(world)]
    p = "model.language_model.layers.1."
    x = torch.randn(5, t.hidden)
    # column-parallel q/k/v: concatenating the rank outputs gives the full projection; heads stay whole
    for nm in ("q_proj", "k_proj", "v_proj"):
        full = x @ sd[p + f"self_attn.{nm}.weight"].float().t() + sd[p + f"self_attn.{nm}.bias"].float()
        parts = [x @ s[p + f"self_attn.{nm}.weight"].float().t() + s[p + f"self_attn.{nm}.bias"].float() for s in shards]
        assert torch.allclose(torch.cat(parts, 1), full, atol=1e-5)
        assert shards[0][p + f"self_attn.{nm}.weight"].shape[0] % t.head_dim == 0
    # q heads of rank r attend the KV heads of rank r (GQA groups are contiguous)
    G = t.heads // t.kv_heads
    hl, kl = t.heads // world, t.kv_heads // world
    for r in range(world):
        assert {h // G for h in range(r * hl, (r + 1) * hl)} == set(range(r * kl, (r + 1) * kl))
    # row-parallel o_proj / down_proj: the sum of the rank partials is the full product
    a = torch.randn(5, t.heads * t.head_dim)
    full = a @ sd[p + "self_attn.o_proj.weight"].float().t()
    n = a.shape[1] // world
    part = sum(a[:, r * n:(r + 1) * n] @ shards[r][p + "self_attn.o_proj.weight"].float().t() for r in range(world))
    assert torch.allclose(part, full, atol=1e-4)
    # SwiGLU MLP: column-parallel gate/up, row-parallel down
    g = x @ sd[p + "mlp.gate_proj.weight"].float().t()
    u = x @ sd[p + "mlp.up_proj.weight"].float().t()
    full = (torch.nn.functional.silu(g) * u) @ sd[p + "mlp.down_proj.weight"].float().t()
    part = 0
    for s in shards:
        gl = x @ s[p + "mlp.gate_proj.weight"].float().t()
        ul = x @ s[p + "mlp.up_proj.weight"].float().t()
        part = part + (torch.nn.functional.silu(gl) * ul) @ s[p + "mlp.down_proj.weight"].float().t()
    assert torch.allclose(part, full, atol=1e-4)
    # vocab-split lm_head; replicated tensors untouched
    assert torch.equal(torch.cat([s["lm_head.weight"] for s in shards]), sd["lm_head.weight"])
    assert shards[1]["model.language_model.norm.weight"] is sd["model.language_model.norm.weight"]
    assert shards[1]["model.visual.blocks.0.attn.qkv.weight"] is sd["model.visual.blocks.0.attn.qkv.weight"]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import handwritten_ocr_b200  # noqa: F401
    from handwritten_ocr_b200.tp import TPComm
    c = TPComm()
    x = torch.full((3, 8), float(rank + 1))
    c.all_reduce(x)
    local = torch.arange(3 * 4, dtype=torch.float32).view(3, 4) + 100 * rank          # [B, V/world]
    full = torch.empty(3, 4 * world)
    c.gather_vocab(local, full)
    q.put((rank, x.tolist(), full.tolist(), c.n_all_reduce))
    dist.barrier()
    dist.destroy_process_group()


def test_tp_collectives_two_ranks_gloo():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, x, full, n in got:
        assert x == [[3.0] * 8] * 3 and n == 1
        want = [[float(b * 4 + j) for j in range(4)] + [float(100 + b * 4 + j) for j in range(4)] for b in range(3)]
        assert full == want          # row b: rank 0's vocab slice, then rank 1's
