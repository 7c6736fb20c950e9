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
L.ocrb_chain_set_trace.argtypes = [ctypes.c_void_p, ctypes.c_int32]; L.ocrb_chain_set_trace.restype = None
for _ in range(3):
    dec._step(st)
torch.cuda.synchronize()
vlm.CHAIN_TRACE = []
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    dec._step(st)
bufs = vlm.CHAIN_TRACE
vlm.CHAIN_TRACE = None
L.ocrb_chain_set_trace(None, 0)
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    g.replay()
e1.record(); torch.cuda.synchronize()
print(f"B={B} layers={layers} ctx={ctx}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per step = {e0.elapsed_time(e1) / 20 / layers * 1e3:.1f} us per layer (traced)")
fused = vlm.CHAIN_FUSE_ATTN and vlm.CHAIN_MAX_B >= B
names = ["inputs", "normed", "kb0_ready", "last_acc", "announced"]
if fused:
    n_ops = 5 * layers + 1
    slots = 8 + 8 * n_ops
    t = bufs[-1].cpu().numpy().reshape(296, slots)[:148].astype(np.float64)
    t[t == 0] = np.nan
    t0 = np.nanmin(t[:, 0])
    print(f"plan launch: start {0:.0f}..{np.nanmax(t[:, 0]) - t0:.0f}  setup {np.nanmax(t[:, 1]) - t0:.0f}  end {np.nanmin(t[:, 7]) - t0:.0f}..{np.nanmax(t[:, 7]) - t0:.0f} ns")
    opn = ["qkv", "attn", "o_proj", "gate_up", "down"]
    ref = t0
    for g in range(n_ops):
        row = []
        for k, nm in enumerate(names):
            v = t[:, 8 + g * 8 + k] - ref
            if np.all(np.isnan(v)):
                continue
            if opn[g % 5] == "attn" and g < n_ops - 1:
                nm = {"inputs": "qkv_seen", "normed": "q_frags", "kb0_ready": "kv0_ready", "last_acc": "partials", "announced": "combined"}.get(nm, nm)
            row.append(f"{nm} {np.nanmin(v):.0f}/{np.nanmedian(v):.0f}/{np.nanmax(v):.0f}")
        name = "lm_head" if g == n_ops - 1 else f"L{g // 5} {opn[g % 5]}"
        print(f"   {name:12s} " + "  ".join(row))
        if g % 5 == 4:
            ref = np.nanmax(t[:, 8 + g * 8 + 4])       # next layer relative to the end of this layer's down_proj
            print(f"   -- layer {g // 5} ends {ref - t0:.0f} ns after the launch started")
else:
    t = torch.stack(bufs).cpu().numpy().reshape(len(bufs), 296, 64)[:, :148, :].astype(np.float64)
    t[t == 0] = np.nan
    for i in range(1, len(bufs)):
        prev_end = np.nanmax(t[i - 1, :, 7])
        r = t[i] - prev_end
        print(f"-- chain launch {i}: start {np.nanmin(r[:, 0]):.0f}..{np.nanmax(r[:, 0]):.0f}  setup {np.nanmax(r[:, 1]):.0f}  end {np.nanmin(r[:, 7]):.0f}..{np.nanmax(r[:, 7]):.0f} ns after the previous chain launch ended")
        for gidx in range(4):
            row = []
            for k, nm in enumerate(names):
                v = r[:, 8 + gidx * 8 + k]
                if np.all(np.isnan(v)):
                    continue
                row.append(f"{nm} {np.nanmin(v):.0f}/{np.nanmedian(v):.0f}/{np.nanmax(v):.0f}")
            print(f"   linear {gidx}: " + "  ".join(row))
