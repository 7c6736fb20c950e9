"""Timing of the paged decode attention (partial + combine) at 7B dims, CUDA graph replay."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import handwritten_ocr_b200
from handwritten_ocr_b200 import _lib
BF = torch.bfloat16
nq, nkv, hd, page = 28, 4, 128, 16
CASES = [(3, 1300)] if len(sys.argv) > 1 and sys.argv[1] == "one" else [(3, 1300), (3, 1500), (24, 1300), (63, 1300), (96, 1300)]
KEYS = [int(k) for k in os.environ.get("KEYS", "64,128,256").split(",")]        # keys per CTA
for B, ctx in [(b, c) for b, c in CASES for _ in KEYS]:
    max_ctx = 2112
    max_pages = max_ctx // page
    n_pages = B * max_pages
    layers = 8   # rotate over several layers' caches so L2 does not hold everything
    kc = [torch.randn(n_pages, nkv, page, hd, device="cuda").to(BF) for _ in range(layers)]
    vc = [torch.randn(n_pages, nkv, page, hd, device="cuda").to(BF) for _ in range(layers)]
    bt = torch.arange(n_pages, device="cuda", dtype=torch.int32).view(B, max_pages).contiguous()
    qkv = torch.randn(B, (nq + 2 * nkv) * hd, device="cuda").to(BF)
    cos = torch.randn(B, hd, device="cuda").to(BF); sin = torch.randn(B, hd, device="cuda").to(BF)
    ctx_d = torch.full((B,), ctx, dtype=torch.int32, device="cuda")
    KEYS.append(KEYS.pop(0))
    keys = KEYS[-1]
    n_splits = math.ceil(max_ctx / keys)
    ws = torch.empty(B * nq * n_splits * (hd + 2), device="cuda", dtype=torch.float32)
    out = torch.empty(B, nq * hd, device="cuda", dtype=BF)
    sp = lambda: torch.cuda.current_stream().cuda_stream
    def run(i):
        _lib.call("ocrb_decode_attention", qkv.data_ptr(), qkv.stride(0), kc[i % layers].data_ptr(), vc[i % layers].data_ptr(),
                  n_pages, bt.data_ptr(), max_pages, ctx_d.data_ptr(), B, page, nq, nkv, hd, cos.data_ptr(), sin.data_ptr(), hd ** -0.5,
                  out.data_ptr(), nq * hd, ws.data_ptr(), n_splits, sp())
    for i in range(3): run(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    reps = 16
    with torch.cuda.graph(g):
        for i in range(reps): run(i)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    kvb = B * ctx * 2 * nkv * hd * 2
    print(f"decode attention B={B:2d} ctx={ctx} keys/CTA={keys:3d}: {us:7.2f} us per layer (partial+combine), KV bytes {kvb/1e6:.1f} MB -> {kvb/us/1e3:.0f} GB/s", flush=True)
