"""Host-side plumbing of the VLM path vs the installed HF implementation (CPU, no weights):
position ids (get_rope_index), window index, vision rope ids."""
import types

import numpy as np
import pytest
import torch


@pytest.fixture(scope="module")
def vc(pkg):
    from handwritten_ocr_b200 import vlm_config
    return vlm_config


def _hf_stub(tokens_per_second=2, window=112):
    from transformers.models.qwen2_5_vl import modeling_qwen2_5_vl as M
    stub = types.SimpleNamespace()
    stub.config = types.SimpleNamespace(vision_config=types.SimpleNamespace(spatial_merge_size=2,
                                                                            tokens_per_second=tokens_per_second))
    stub.get_vision_position_ids = types.MethodType(M.Qwen2_5_VLModel.get_vision_position_ids, stub)
    stub.window_size = window
    stub.spatial_merge_size = 2
    stub.patch_size = 14
    stub.spatial_merge_unit = 4
    return M, stub


@pytest.mark.parametrize("grid,tps", [((54, 74), 2), ((74, 54), 2), ((28, 36), 4), ((18, 20), 2)])
def test_rope_index_matches_hf(vc, grid, tps):
    M, stub = _hf_stub(tps)
    tok = vc.SyntheticTokenizer()
    ids = vc.build_prompt_ids(tok, "Extract and return all the text from this handwritten document.", grid[0] * grid[1] // 4)
    t = torch.from_numpy(ids.astype(np.int64))[None]
    mm = (t == vc.IMAGE_PAD).int()
    pos, delta = M.Qwen2_5_VLModel.get_rope_index(stub, t, mm, image_grid_thw=torch.tensor([[1, grid[0], grid[1]]]))
    mine, d = vc.rope_index(ids, grid, 2, tps)
    assert np.array_equal(pos[:, 0].numpy(), mine)
    assert int(delta[0, 0]) == d


@pytest.mark.parametrize("grid", [(54, 74), (74, 54), (28, 36), (18, 20), (8, 8)])
def test_window_index_matches_hf(vc, grid):
    M, stub = _hf_stub()
    widx, cu = M.Qwen2_5_VisionTransformerPretrainedModel.get_window_index(stub, torch.tensor([[1, grid[0], grid[1]]]))
    cu = torch.unique_consecutive(torch.tensor(cu, dtype=torch.int32))
    mine_idx, mine_cu = vc.window_index(grid)
    assert np.array_equal(widx.numpy(), mine_idx)
    assert np.array_equal(cu.numpy(), mine_cu)


@pytest.mark.parametrize("grid", [(54, 74), (28, 36)])
def test_vision_rope_pos_matches_hf(vc, grid):
    M, stub = _hf_stub()
    got = {}

    class Rot:
        def __call__(self, n):
            got["n"] = int(n)
            return torch.arange(int(n))[:, None].float()

    stub.rotary_pos_emb = Rot()
    out = M.Qwen2_5_VisionTransformerPretrainedModel.rot_pos_emb(stub, torch.tensor([[1, grid[0], grid[1]]]))
    mine = vc.vision_rope_pos(grid)
    assert np.array_equal(out.numpy().astype(np.int64), mine)


def test_prompt_length_and_specials(vc):
    tok = vc.SyntheticTokenizer()
    ids = vc.build_prompt_ids(tok, "Extract and return all the text from this handwritten document.", 999)
    assert (ids == vc.IMAGE_PAD).sum() == 999
    assert ids[0] == vc.IM_START and 1025 <= len(ids) <= 1045
    text = tok.decode([5, 151645, 777, 9000])
    assert "<|" not in text and len(text.split()) == 3
