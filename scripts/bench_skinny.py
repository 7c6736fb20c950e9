"""Per-shape timing of the decode weight-streaming kernel (7B decoder shapes), CUDA events, weights rotated so
every launch streams from HBM (working set >> L2)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import handwritten_ocr_b200
from handwritten_ocr_b200 import _lib, vlm

BF = torch.bfloat16
dev = torch.device("cuda")
shapes = [("qkv+norm", 4608, 3584, 0, True), ("o_proj+res", 3584, 3584, 1, False), ("gate_up+norm+swiglu", 37888, 3584, 2, True),
          ("down+res", 3584, 18944, 1, False), ("lm_head+norm", 152064, 3584, 0, True)]
Bs = [int(b) for b in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["3", "16", "32", "64"])]
for name, N, K, epi, norm in shapes:
    copies = max(2, int(600e6 // (N * K * 2)) + 1)        # rotate over > 600 MB of weights
    Ws = [(torch.randn(N, K, device=dev) * K ** -0.5).to(BF) for _ in range(copies)]
    nw = torch.ones(K, device=dev, dtype=BF)
    for B in Bs:
        X = torch.randn(B, K, device=dev).to(BF)
        Nout = N // 2 if epi == 2 else N
        D = torch.empty(B, Nout, device=dev, dtype=BF)
        R = torch.randn(B, Nout, device=dev).to(BF)
        def run(i):
            vlm.skinny(X, Ws[i % copies], D, residual=R if epi == 1 else None, epilogue=epi, norm_w=nw if norm else None)
        for i in range(3): run(i)
        torch.cuda.synchronize()
        reps = 20
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(reps): run(i)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        print(f"{name:22s} N={N:6d} K={K:5d} B={B:2d}: {us:8.2f} us  {N*K*2/us/1e3:7.1f} GB/s", flush=True)
    del Ws
