"""Experiment: how fast can ONE SM stream weights through the 5-stage TMA ring?  N = 128*g rows, grid = g CTAs,
one full tile per CTA (no stream-K fix-up)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import handwritten_ocr_b200
from handwritten_ocr_b200 import vlm
BF = torch.bfloat16; dev = torch.device("cuda")
g = int(os.environ["OCRB_SK_GRID"]); K = 3584; N = 128 * g
Ws = [(torch.randn(N, K, device=dev) * K ** -0.5).to(BF) for _ in range(max(2, 600_000_000 // (N * K * 2)))]
X = torch.randn(3, K, device=dev).to(BF); D = torch.empty(3, N, device=dev, dtype=BF)
for i in range(3): vlm.skinny(X, Ws[i % len(Ws)], D)
torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    for i in range(20): vlm.skinny(X, Ws[i % len(Ws)], D)
gr.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / 20
print(f"grid {g:3d} CTAs, one 128x{K} tile each: {us:7.2f} us per launch, {N*K*2/us/1e3:7.1f} GB/s total, {128*K*2/us/1e3:6.1f} GB/s per SM")
