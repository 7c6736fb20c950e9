"""In-kernel timeline (globaltimer stamps) of the chained decode step: one CUDA-graph replay of a 7B-class decode step at
batch B, every skinny_chain_kernel launch stamping [CTA][slot].  Prints, per launch, when each linear's inputs were
complete, when its first k-block was ready, when its last segment was accumulated and when it was announced -- relative to
the end of the previous chain launch -- so the bubbles between dependent linears and around the attention are visible."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import handwritten_ocr_b200  # noqa
from handwritten_ocr_b200 import _lib, engine, vlm
from handwritten_ocr_b200.vlm_config import VLMConfig

B = int(sys.argv[1]) if len(sys.argv) > 1 else 3
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 6
ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 1100
cfg = VLMConfig.olmocr_7b()
cfg.text.layers = layers
cfg.vision.depth = 1
cfg.vision.fullatt_blocks = (0,)
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
w = vlm.VLMWeights.random(cfg, dev, seed=0)
kv = vlm.PagedKV(cfg, n_pages=B * ((ctx + 600) // 16 + 1), page_size=16, device=dev)
dec = vlm.Decoder(w, kv, max_batch=B, max_ctx=ctx + 600)
pages = (ctx + 600) // 16 + 1
bt = torch.arange(B * pages, dtype=torch.int32, device=dev).view(B, pages)
cos, sin, inv = vlm.text_rope_tables(cfg, torch.zeros((3, 1), dtype=torch.int64, device=dev))
st = vlm.DecodeState(dec, B, 64, bt, [ctx] * B, [0] * B, inv)
L = _lib.load()
import ctypes
L.ocrb_chain_set_trace.argtypes = [ctypes.c_void_p]; L.ocrb_chain_set_trace.restype = None
for _ in range(3):
    dec._step(st)
torch.cuda.synchronize()
vlm.CHAIN_TRACE = []
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    dec._step(st)
bufs = vlm.CHAIN_TRACE
vlm.CHAIN_TRACE = None
L.ocrb_chain_set_trace(None)
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    g.replay()
e1.record(); torch.cuda.synchronize()
print(f"B={B} layers={layers} ctx={ctx}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per step = {e0.elapsed_time(e1) / 20 / layers * 1e3:.1f} us per layer (traced)")
t = torch.stack(bufs).cpu().numpy().reshape(len(bufs), 296, 64)[:, :148, :].astype(np.float64)
t[t == 0] = np.nan
names = ["inputs", "normed", "kb0_ready", "last_acc", "announced"]
for i in range(1, len(bufs)):
    prev_end = np.nanmax(t[i - 1, :, 63])
    r = t[i] - prev_end
    nd = 4
    print(f"-- chain launch {i}: start {np.nanmin(r[:, 0]):.0f}..{np.nanmax(r[:, 0]):.0f}  setup {np.nanmax(r[:, 1]):.0f}  end {np.nanmin(r[:, 63]):.0f}..{np.nanmax(r[:, 63]):.0f} ns after the previous chain launch ended")
    for gidx in range(nd):
        row = []
        for k, nm in enumerate(names):
            v = r[:, 8 + gidx * 8 + k]
            if np.all(np.isnan(v)):
                continue
            row.append(f"{nm} {np.nanmin(v):.0f}/{np.nanmedian(v):.0f}/{np.nanmax(v):.0f}")
        print(f"   linear {gidx}: " + "  ".join(row))
