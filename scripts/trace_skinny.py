"""In-kernel timeline of skinny_gemm_kernel (globaltimer stamps per CTA) for the small decode GEMMs, PDL chain in a graph."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import handwritten_ocr_b200
from handwritten_ocr_b200 import _lib, vlm
BF = torch.bfloat16; dev = torch.device("cuda")
L = _lib.load()
L.ocrb_skinny_set_trace.argtypes = [ctypes.c_void_p]; L.ocrb_skinny_set_trace.restype = None
names = ["start", "setup_done", "wait_ret", "w_first", "x_first", "seg0_acc", "last_acc", "published", "fixup_done", "end", "fix_in_smem", "epi_stored"]
BATCH = int(sys.argv[1]) if len(sys.argv) > 1 else 3
SHAPES = {"o": ("o_proj+res", 3584, 3584, 1, False), "qkv": ("qkv+norm", 4608, 3584, 0, True), "gu": ("gate_up", 37888, 3584, 2, True),
          "down": ("down+res", 3584, 18944, 1, False)}
for name, N, K, epi, norm in [SHAPES[k] for k in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["o"])]:
    Ws = [(torch.randn(N, K, device=dev) * K ** -0.5).to(BF) for _ in range(12)]
    nw = torch.ones(K, device=dev, dtype=BF); B = BATCH
    X = torch.randn(B, K, device=dev).to(BF); D = torch.empty(B, N // 2 if epi == 2 else N, device=dev, dtype=BF); R = torch.randn(B, N, device=dev).to(BF)
    traces = [torch.zeros(296 * 64, dtype=torch.int64, device=dev) for _ in range(12)]
    def run(i):
        L.ocrb_skinny_set_trace(traces[i % 12].data_ptr())
        vlm.skinny(X, Ws[i % 12], D, residual=R if epi == 1 else None, epilogue=epi, norm_w=nw if norm else None)
    for i in range(3): run(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(12): run(i)
    g.replay(); torch.cuda.synchronize(); g.replay(); torch.cuda.synchronize()
    L.ocrb_skinny_set_trace(None)
    t = torch.stack(traces).cpu().numpy().reshape(12, 296, 64)[:, :148, :].astype(np.float64)
    # kernel i in the chain: reference time = end of kernel i-1 (max over CTAs)
    rows = []
    for i in range(2, 11):
        prev_end = t[i - 1, :, 9].max()
        cur = t[i] - prev_end
        cur[t[i] == 0] = np.nan
        rows.append(cur)
    cur = np.stack(rows)   # [kernels, ctas, stamps]
    print(f"== {name} N={N} K={K} B={B}: ns relative to the end of the previous kernel in the chain (median over kernels; min / median / max over CTAs)")
    for s, nm in enumerate(names):
        v = cur[:, :, s]
        if np.all(np.isnan(v)): continue
        print(f"   {nm:12s} min {np.nanmedian(np.nanmin(v, 1)):8.0f}  med {np.nanmedian(np.nanmedian(v, 1)):8.0f}  max {np.nanmedian(np.nanmax(v, 1)):8.0f}")
    print(f"   kernel period (end to end): {np.median(np.diff(t[1:, :, 9].max(1))):.0f} ns")
    k = cur[4]          # one kernel of the chain, CTA 5 and CTA 100: per-unit (w_ready, x_ready) in the first segment
    for cta in (5, 100):
        print(f"   CTA {cta} units (w_ready, x_ready):", " ".join(f"({k[cta, 16 + 2 * i]:.0f},{k[cta, 17 + 2 * i]:.0f})" for i in range(12) if not np.isnan(k[cta, 16 + 2 * i])))
        print(f"   CTA {cta} epilogue chunk starts:", " ".join(f"{k[cta, 40 + i]:.0f}" for i in range(12) if not np.isnan(k[cta, 40 + i])))
    own = [c for c in range(148) if not np.isnan(k[c, 8])][:3]
    for cta in own:
        print(f"   owner CTA {cta}: last_acc {k[cta, 6]:.0f} fix_in_smem {k[cta, 10]:.0f} chunk starts", " ".join(f"{k[cta, 40 + i]:.0f}" for i in range(12) if not np.isnan(k[cta, 40 + i])), f"epi_stored {k[cta, 11]:.0f} fixup_done {k[cta, 8]:.0f} end {k[cta, 9]:.0f}")
