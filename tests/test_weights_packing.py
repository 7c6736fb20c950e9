"""Host-side weight handling on CPU: kernel-layout packing (fused qkv, [gate64|up64] interleave, zero padding of the
SwiGLU width) reproduces the plain HF linears, and a safetensors checkpoint directory (the real-checkpoint path of
tools._load_ocr_model, OCRB_CHECKPOINT) round-trips into the same packed weights."""
import torch


def test_packing_reproduces_hf_linears(pkg):
    from handwritten_ocr_b200 import vlm
    from handwritten_ocr_b200.vlm_config import VLMConfig
    cfg = VLMConfig.tiny()
    sd = vlm.random_state_dict(cfg, "cpu", seed=5)
    w = vlm.VLMWeights.from_state_dict(cfg, dict(sd))
    t, v = cfg.text, cfg.vision
    p = "model.language_model.layers.1."
    x = torch.randn(4, t.hidden)
    lay = w.layers[1]
    # fused qkv rows = q | k | v
    q = x @ sd[p + "self_attn.q_proj.weight"].float().t() + sd[p + "self_attn.q_proj.bias"].float()
    fused = x @ lay["qkv_w"].float().t() + lay["qkv_b"].float()
    assert torch.allclose(fused[:, : t.heads * t.head_dim], q, atol=1e-5)
    assert fused.shape[1] == (t.heads + 2 * t.kv_heads) * t.head_dim
    # packed gate/up: tile i holds gate rows [64i, 64i+64) then up rows [64i, 64i+64); padding rows are zero
    ip = t.intermediate_padded
    assert lay["gu_w"].shape == (2 * ip, t.hidden) and lay["down_w"].shape == (t.hidden, ip)
    gu = (x @ lay["gu_w"].float().t()).view(4, ip // 64, 2, 64)
    g, u = gu[:, :, 0].reshape(4, ip), gu[:, :, 1].reshape(4, ip)
    g_ref = x @ sd[p + "mlp.gate_proj.weight"].float().t()
    u_ref = x @ sd[p + "mlp.up_proj.weight"].float().t()
    assert torch.allclose(g[:, : t.intermediate], g_ref, atol=1e-5) and torch.allclose(u[:, : t.intermediate], u_ref, atol=1e-5)
    assert (g[:, t.intermediate:] == 0).all() and (u[:, t.intermediate:] == 0).all()
    act = torch.nn.functional.silu(g) * u
    mlp = act @ lay["down_w"].float().t()
    mlp_ref = (torch.nn.functional.silu(g_ref) * u_ref) @ sd[p + "mlp.down_proj.weight"].float().t()
    assert torch.allclose(mlp, mlp_ref, atol=1e-4)
    # vision: 600 -> 640 padded, conv3d patch embed flattened to [hidden, 1176]
    assert w.vis_blocks[0]["gu_w"].shape == (2 * v.intermediate_padded, v.hidden) and v.intermediate_padded % 64 == 0
    assert w.patch_embed.shape == (v.hidden, 3 * 2 * 14 * 14)
    assert w.decode_weight_bytes() == 2 * (w.lm_head.numel() + w.final_norm.numel() + sum(
        sum(x.numel() for x in lay.values()) for lay in w.layers))


def test_safetensors_checkpoint_roundtrip(pkg, tmp_path):
    from safetensors.torch import load_file, save_file
    from handwritten_ocr_b200 import vlm
    from handwritten_ocr_b200.vlm_config import VLMConfig
    cfg = VLMConfig.tiny()
    sd = vlm.random_state_dict(cfg, "cpu", seed=6)
    keys = sorted(sd)
    half = len(keys) // 2
    save_file({k: sd[k].contiguous() for k in keys[:half]}, str(tmp_path / "model-00001-of-00002.safetensors"))
    save_file({k: sd[k].contiguous() for k in keys[half:]}, str(tmp_path / "model-00002-of-00002.safetensors"))
    loaded = {}
    for f in sorted(tmp_path.glob("*.safetensors")):          # what tools._load_ocr_model does with OCRB_CHECKPOINT
        loaded.update(load_file(str(f)))
    a = vlm.VLMWeights.from_state_dict(cfg, dict(sd))
    b = vlm.VLMWeights.from_state_dict(cfg, loaded, free_source=True)
    assert not loaded, "free_source must consume the source dict as it packs"
    assert torch.equal(a.lm_head, b.lm_head) and torch.equal(a.embed, b.embed)
    for la, lb in zip(a.layers, b.layers):
        assert all(torch.equal(la[k], lb[k]) for k in la)
    for va, vb in zip(a.vis_blocks, b.vis_blocks):
        assert all(torch.equal(va[k], vb[k]) for k in va)


def test_legacy_checkpoint_key_names(pkg):
    """Published Qwen2.5-VL safetensors use `visual.*` / `model.layers.*` / `model.embed_tokens.*` / `model.norm.*`
    (HF renames them in from_pretrained); packing must accept them and give the same weights."""
    import pytest
    from handwritten_ocr_b200 import vlm
    from handwritten_ocr_b200.vlm_config import VLMConfig
    cfg = VLMConfig.tiny()
    sd = vlm.random_state_dict(cfg, "cpu", seed=7)

    def legacy(k):
        if k.startswith("model.visual."):
            return k[len("model."):]
        if k.startswith("model.language_model."):
            return "model." + k[len("model.language_model."):]
        return k

    old = {legacy(k): x for k, x in sd.items()}
    assert "visual.patch_embed.proj.weight" in old and "model.layers.0.self_attn.q_proj.weight" in old
    assert "model.embed_tokens.weight" in old and "model.norm.weight" in old and "lm_head.weight" in old
    assert all(vlm.normalize_checkpoint_key(legacy(k)) == k for k in sd)
    assert sorted(vlm.expected_state_dict_keys(cfg)) == sorted(sd)
    a = vlm.VLMWeights.from_state_dict(cfg, dict(sd))
    b = vlm.VLMWeights.from_state_dict(cfg, old, free_source=True)
    assert not old
    assert torch.equal(a.lm_head, b.lm_head) and torch.equal(a.embed, b.embed) and torch.equal(a.final_norm, b.final_norm)
    assert torch.equal(a.patch_embed, b.patch_embed) and torch.equal(a.merger_w2, b.merger_w2)
    for la, lb in zip(a.layers, b.layers):
        assert all(torch.equal(la[k], lb[k]) for k in la)
    for va, vb in zip(a.vis_blocks, b.vis_blocks):
        assert all(torch.equal(va[k], vb[k]) for k in va)
    # a truncated checkpoint fails with the list of what is missing, not a bare KeyError on the first name
    broken = {k: x for k, x in sd.items() if "layers.1.mlp.down_proj" not in k}
    with pytest.raises(KeyError, match="down_proj"):
        vlm.VLMWeights.from_state_dict(cfg, broken)
