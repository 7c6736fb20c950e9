"""One rank's share of the TP-8 72B-class decode step on ONE GPU, without the collectives: the single-GPU kernels on the
local shapes (8 q heads + 1 KV head, 3 696 MLP columns, hidden 8 192, 80 layers, 19 008 vocab rows = 17.9 GB of weights).
Separates kernel time from all-reduce time in the TP-8 step (profiles/r02_notes.md)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import handwritten_ocr_b200  # noqa
from handwritten_ocr_b200 import tp, vlm
from handwritten_ocr_b200.vlm_config import VLMConfig

B = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ctx = 1100
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
cfg = VLMConfig.qwen72b()
cfg.vision.depth = 1
cfg.vision.fullatt_blocks = (0,)
w, lcfg = tp.random_weights_tp(cfg, dev, 0, 8, seed=0)
pages = (ctx + 600) // 16 + 1
kv = vlm.PagedKV(lcfg, n_pages=B * pages, page_size=16, device=dev)
dec = vlm.Decoder(w, kv, max_batch=B, max_ctx=ctx + 600)
bt = torch.arange(B * pages, dtype=torch.int32, device=dev).view(B, pages)
cos, sin, inv = vlm.text_rope_tables(lcfg, torch.zeros((3, 1), dtype=torch.int64, device=dev))
st = vlm.DecodeState(dec, B, 64, bt, [ctx] * B, [0] * B, inv)
for _ in range(3):
    dec._step(st)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    dec._step(st)
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    g.replay()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
wb = w.decode_weight_bytes()
print(f"TP-8 rank share, B={B}: {ms:.3f} ms per step, {wb / 1e9:.2f} GB of weights -> {wb / ms / 1e6:.0f} GB/s "
      f"(OCRB_SK_CLUSTER={os.environ.get('OCRB_SK_CLUSTER', '1')}, OCRB_CHAIN_MAX_B={os.environ.get('OCRB_CHAIN_MAX_B', '0')})")
