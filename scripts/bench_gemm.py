"""TFLOP/s of the tcgen05 GEMM on the vision-tower and prefill shapes of the 7B config (B = 3 candidates)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import handwritten_ocr_b200
from handwritten_ocr_b200 import vlm
BF = torch.bfloat16; dev = torch.device("cuda")
S, T = 3 * 3996, 3 * 1036
shapes = [("vis patch_embed", S, 1280, 1176, 0), ("vis qkv", S, 3840, 1280, 0), ("vis proj+res", S, 1280, 1280, 1),
          ("vis gate_up+swiglu", S, 6912, 1280, 2), ("vis down+res", S, 1280, 3456, 1), ("merger fc1+gelu", 2997, 5120, 5120, 3),
          ("txt qkv", T, 4608, 3584, 0), ("txt o+res", T, 3584, 3584, 1), ("txt gate_up+swiglu", T, 37888, 3584, 2),
          ("txt down+res", T, 3584, 18944, 1)]
tot_f, tot_t = 0.0, 0.0
for name, M, N, K, epi in shapes:
    A = torch.randn(M, K, device=dev).to(BF); W = (torch.randn(N, K, device=dev) * K ** -0.5).to(BF)
    No = N // 2 if epi == 2 else N
    D = torch.empty(M, No, device=dev, dtype=BF); R = torch.randn(M, No, device=dev).to(BF); b = torch.randn(N, device=dev).to(BF)
    run = lambda: vlm.gemm(A, W, D, bias=b if epi == 3 else None, residual=R if epi == 1 else None, epilogue=epi)
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 2.0 * M * N * K
    print(f"{name:22s} M={M:6d} N={N:6d} K={K:6d}: {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TFLOP/s", flush=True)
