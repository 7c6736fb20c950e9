"""GPU numerics of the dense VLM kernels through the C ABI vs plain PyTorch references that mirror
HF's rounding points (bf16 tensors between ops, fp32 accumulation inside them)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

BF = torch.bfloat16


@pytest.fixture(scope="module")
def L(pkg):
    from handwritten_ocr_b200 import _lib
    _lib.load()
    return _lib


def sp():
    return torch.cuda.current_stream().cuda_stream


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda") * scale).to(BF)


def ref_linear(A, W, bias=None):
    out = A.float() @ W.float().t()
    if bias is not None:
        out = out + bias.float()
    return out.to(BF)


def pack_swiglu(Wg, Wu):
    """[gate64|up64] row interleave expected by OCRB_EPI_SWIGLU."""
    I = Wg.shape[0]
    assert I % 64 == 0
    return torch.stack([Wg.view(I // 64, 64, -1), Wu.view(I // 64, 64, -1)], 1).reshape(2 * I, -1).contiguous()


def close_bf16(got, want, what, ulps=2.0, frac_exact=0.97, mag=None):
    """|got - want| <= ulps bf16 ulp of the larger of |want| and `mag` (the magnitude of the
    intermediate a rounding happened at, e.g. the pre-residual linear output)."""
    g, w = got.float(), want.float()
    ref_mag = w.abs() if mag is None else torch.maximum(w.abs(), mag.float().abs())
    true_ulp = torch.exp2(torch.floor(torch.log2(ref_mag.clamp_min(2.0 ** -6))) - 7)   # bf16: 8 significant bits
    tol = ulps * true_ulp
    bad = (g - w).abs() > tol
    assert not bad.any(), f"{what}: {int(bad.sum())} of {bad.numel()} beyond {ulps} bf16 ulp; max err {(g - w).abs().max().item()}"
    ex = (g == w).float().mean().item()
    assert ex >= frac_exact, f"{what}: only {ex:.4f} bit-equal"


GEMM_SHAPES = [(128, 128, 64), (256, 256, 128), (300, 384, 1176), (130, 200, 72), (1000, 1280, 1280),
               (4096, 3584, 512), (3108, 4608, 3584), (999, 5120, 5120)]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_plain_and_bias(L, M, N, K):
    A, W, b = rnd(M, K, seed=1), rnd(N, K, scale=K ** -0.5, seed=2), rnd(N, seed=3)
    for bias in (None, b):
        D = torch.full((M, N), float("nan"), device="cuda", dtype=BF)
        L.call("ocrb_gemm_bf16", A.data_ptr(), K, W.data_ptr(), K, D.data_ptr(), N, M, N, K,
               L.ptr(bias), None, 0, 0, sp())
        torch.cuda.synchronize()
        close_bf16(D, ref_linear(A, W, bias), f"gemm {M}x{N}x{K} bias={bias is not None}")


def test_gemm_strided_operands(L):
    M, N, K = 333, 256, 320
    Abig, Wbig = rnd(M, K + 64, seed=4), rnd(N, K + 128, scale=K ** -0.5, seed=5)
    A, W = Abig[:, 8:8 + K], Wbig[:, 16:16 + K]
    Dbig = torch.zeros((M, N + 40), device="cuda", dtype=BF)
    L.call("ocrb_gemm_bf16", A.data_ptr(), Abig.stride(0), W.data_ptr(), Wbig.stride(0), Dbig[:, 8:].data_ptr(),
           Dbig.stride(0), M, N, K, None, None, 0, 0, sp())
    torch.cuda.synchronize()
    close_bf16(Dbig[:, 8:8 + N], ref_linear(A, W), "strided gemm")
    assert (Dbig[:, :8] == 0).all() and (Dbig[:, 8 + N:] == 0).all()


@pytest.mark.parametrize("M,N,K", [(300, 384, 256), (3108, 3584, 1024)])
def test_gemm_residual(L, M, N, K):
    A, W, R = rnd(M, K, seed=6), rnd(N, K, scale=K ** -0.5, seed=7), rnd(M, N, seed=8)
    D = torch.empty((M, N), device="cuda", dtype=BF)
    L.call("ocrb_gemm_bf16", A.data_ptr(), K, W.data_ptr(), K, D.data_ptr(), N, M, N, K, None, R.data_ptr(), N, 1, sp())
    torch.cuda.synchronize()
    lin = ref_linear(A, W)
    close_bf16(D, (lin.float() + R.float()).to(BF), "gemm+residual", mag=lin)


@pytest.mark.parametrize("M,I,K,bias", [(200, 128, 256, False), (1000, 3456, 1280, True), (700, 2048, 512, False)])
def test_gemm_swiglu(L, M, I, K, bias):
    A = rnd(M, K, seed=9)
    Wg, Wu = rnd(I, K, scale=K ** -0.5, seed=10), rnd(I, K, scale=K ** -0.5, seed=11)
    bg, bu = rnd(I, seed=12), rnd(I, seed=13)
    Wp = pack_swiglu(Wg, Wu)
    bp = pack_swiglu(bg.view(-1, 1), bu.view(-1, 1)).view(-1).contiguous() if bias else None
    D = torch.empty((M, I), device="cuda", dtype=BF)
    L.call("ocrb_gemm_bf16", A.data_ptr(), K, Wp.data_ptr(), K, D.data_ptr(), I, M, 2 * I, K, L.ptr(bp), None, 0, 2, sp())
    torch.cuda.synchronize()
    g = ref_linear(A, Wg, bg if bias else None)
    u = ref_linear(A, Wu, bu if bias else None)
    want = torch.nn.functional.silu(g) * u
    close_bf16(D, want, "gemm+swiglu", ulps=4.0, frac_exact=0.9, mag=g.float().abs() * u.float().abs())


def test_gemm_gelu(L):
    M, N, K = 999, 512, 640
    A, W, b = rnd(M, K, seed=14), rnd(N, K, scale=K ** -0.5, seed=15), rnd(N, seed=16)
    D = torch.empty((M, N), device="cuda", dtype=BF)
    L.call("ocrb_gemm_bf16", A.data_ptr(), K, W.data_ptr(), K, D.data_ptr(), N, M, N, K, b.data_ptr(), None, 0, 3, sp())
    torch.cuda.synchronize()
    lin = ref_linear(A, W, b)
    close_bf16(D, torch.nn.functional.gelu(lin), "gemm+gelu", ulps=3.0, frac_exact=0.9, mag=lin)


def hf_rmsnorm(x, w, eps):
    xf = x.float()
    var = xf.pow(2).mean(-1, keepdim=True)
    return w * (xf * torch.rsqrt(var + eps)).to(x.dtype)


# ───────────── skinny GEMM (tcgen05 swap-AB, stream-K): the decode weight-streaming kernel ─────────────
@pytest.fixture(scope="module")
def skws(L):
    n = int(L.load().ocrb_skinny_workspace_bytes())
    return torch.zeros(n, dtype=torch.uint8, device="cuda")


def skinny_call(L, ws, X, W, D, B, N, K, bias=None, res=None, epi=0, norm_w=None, eps=0.0, ldx=None, ldd=None):
    L.call("ocrb_skinny_gemm_bf16", X.data_ptr(), ldx or X.stride(0), W.data_ptr(), W.stride(0), D.data_ptr(),
           ldd or D.stride(0), B, N, K, L.ptr(bias), L.ptr(res), res.stride(0) if res is not None else 0, epi,
           L.ptr(norm_w), eps, ws.data_ptr(), sp())


@pytest.mark.parametrize("B", [1, 3, 8, 16, 17, 32, 40, 64, 65, 96, 97, 128])
@pytest.mark.parametrize("N,K", [(512, 256), (4608, 3584), (3584, 3584), (3584, 18944), (1000, 328), (128, 64), (152064, 512),
                                 (1280, 1024), (9216, 8192)])
def test_skinny_plain_bias_residual(L, skws, B, N, K):
    X, W, b, R = rnd(B, K, seed=20), rnd(N, K, scale=K ** -0.5, seed=21), rnd(N, seed=22), rnd(B, N, seed=23)
    for epi, bias in [(0, None), (0, b), (1, None)]:
        D = torch.full((B, N), float("nan"), device="cuda", dtype=BF)
        skinny_call(L, skws, X, W, D, B, N, K, bias=bias, res=R if epi == 1 else None, epi=epi)
        torch.cuda.synchronize()
        lin = ref_linear(X, W, bias)
        want = (lin.float() + R.float()).to(BF) if epi == 1 else lin
        close_bf16(D, want, f"skinny B={B} {N}x{K} epi={epi}", mag=lin)
    flags_at = 296 * 128 * 128 * 4          # the flags follow SK_MAX_GRID x SK_MAXBP x 128 fp32 partials (skinny.cu)
    assert int(skws[flags_at:flags_at + 296 * 4].view(torch.int32).abs().sum()) == 0, "stream-K flags must return to zero"


@pytest.mark.parametrize("B", [1, 3, 24, 63, 96, 128])
def test_skinny_fused_norm_swiglu_gelu(L, skws, B):
    K, I = 3584, 1024
    X, nw = rnd(B, K, seed=30), (1 + 0.1 * rnd(K, seed=31).float()).to(BF)
    Wg, Wu = rnd(I, K, scale=K ** -0.5, seed=32), rnd(I, K, scale=K ** -0.5, seed=33)
    Wp = pack_swiglu(Wg, Wu)
    D = torch.full((B, I), float("nan"), device="cuda", dtype=BF)
    skinny_call(L, skws, X, Wp, D, B, 2 * I, K, epi=2, norm_w=nw, eps=1e-6)
    torch.cuda.synchronize()
    xn = hf_rmsnorm(X, nw, 1e-6)
    g, u = ref_linear(xn, Wg), ref_linear(xn, Wu)
    # silu(g) and u are each rounded to bf16 before the product is rounded again: 3 roundings, of which the reference
    # (a different summation order inside the two linears) can flip the first two -- up to ~5 ulp in rare elements
    close_bf16(D, torch.nn.functional.silu(g) * u, "skinny norm+swiglu", ulps=8.0, frac_exact=0.9,
               mag=g.float().abs() * u.float().abs())
    bias = rnd(I, seed=34)
    D2 = torch.full((B, I), float("nan"), device="cuda", dtype=BF)
    skinny_call(L, skws, X, Wg, D2, B, I, K, bias=bias, epi=3)
    torch.cuda.synchronize()
    lin = ref_linear(X, Wg, bias)
    close_bf16(D2, torch.nn.functional.gelu(lin), "skinny gelu", ulps=3.0, frac_exact=0.9, mag=lin)


def test_skinny_strided_and_inplace_residual(L, skws):
    B, N, K = 5, 3584, 3584
    Xbig, W = rnd(B, K + 64, seed=35), rnd(N, K, scale=K ** -0.5, seed=36)
    X = Xbig[:, 8:8 + K]
    H = rnd(B, N, seed=37)
    H0 = H.clone()
    skinny_call(L, skws, X, W, H, B, N, K, res=H, epi=1)          # x += o_proj(att), in place as the decoder does
    torch.cuda.synchronize()
    lin = ref_linear(X, W)
    close_bf16(H, (lin.float() + H0.float()).to(BF), "skinny in-place residual", mag=lin)


def test_skinny_batch_invariance(L, skws):
    """A sequence decoded in a batch of 3, 16, 17 (BP=32), 64 (BP=64), 96 or 128 produces the bits it produces alone."""
    N, K = 4608, 3584
    X, W, nw = rnd(128, K, seed=40), rnd(N, K, scale=K ** -0.5, seed=41), (1 + 0.1 * rnd(K, seed=42).float()).to(BF)
    alone = []
    for b in range(4):
        D1 = torch.empty((1, N), device="cuda", dtype=BF)
        skinny_call(L, skws, X[b:b + 1], W, D1, 1, N, K, norm_w=nw, eps=1e-6)
        alone.append(D1[0].clone())
    for Bb in (3, 16, 17, 64, 96, 128):
        D = torch.empty((Bb, N), device="cuda", dtype=BF)
        skinny_call(L, skws, X[:Bb], W, D, Bb, N, K, norm_w=nw, eps=1e-6)
        torch.cuda.synchronize()
        for b in range(3):
            assert torch.equal(alone[b], D[b]), f"row {b} differs between B=1 and B={Bb}"
    # and run-to-run determinism of the stream-K fix-up
    D_a = torch.empty((16, N), device="cuda", dtype=BF)
    D_b = torch.empty((16, N), device="cuda", dtype=BF)
    skinny_call(L, skws, X[:16], W, D_a, 16, N, K)
    skinny_call(L, skws, X[:16], W, D_b, 16, N, K)
    assert torch.equal(D_a, D_b)


# ───────────── chain of dependent skinny linears in one persistent launch (csrc/chain.cu) ─────────────
@pytest.fixture(scope="module")
def chws(L):
    return torch.zeros(int(L.load().ocrb_chain_workspace_bytes()), dtype=torch.uint8, device="cuda")


def chain_call(L, ws, lins, B):
    import ctypes
    arr = (L.ChainLinear * len(lins))(*lins)
    L.call("ocrb_skinny_chain_bf16", ctypes.addressof(arr), len(lins), B, ws.data_ptr(), sp())


def chain_lin(L, X, W, D, bias=None, res=None, epi=0, norm_w=None, eps=1e-6):
    return L.ChainLinear(X.data_ptr(), X.stride(0), W.data_ptr(), W.stride(0), D.data_ptr(), D.stride(0), W.shape[0],
                         X.shape[1], L.ptr(bias), L.ptr(res), res.stride(0) if res is not None else 0, epi, eps, L.ptr(norm_w))


def layer_tensors(B, H, I, Q, seed):
    t = {"att": rnd(B, H, seed=seed), "x": rnd(B, H, seed=seed + 1),
         "o_w": rnd(H, H, scale=H ** -0.5, seed=seed + 2), "ln2": (1 + 0.1 * rnd(H, seed=seed + 3).float()).to(BF),
         "gu_w": pack_swiglu(rnd(I, H, scale=H ** -0.5, seed=seed + 4), rnd(I, H, scale=H ** -0.5, seed=seed + 5)),
         "down_w": rnd(H, I, scale=I ** -0.5, seed=seed + 6), "ln1": (1 + 0.1 * rnd(H, seed=seed + 7).float()).to(BF),
         "qkv_w": rnd(Q, H, scale=H ** -0.5, seed=seed + 8), "qkv_b": rnd(Q, seed=seed + 9)}
    return t


def run_layer_chain(L, ws, t, B, I, Q):
    x = t["x"][:B].clone()
    act = torch.full((B, I), float("nan"), device="cuda", dtype=BF)
    qkv = torch.full((B, Q), float("nan"), device="cuda", dtype=BF)
    att = t["att"][:B]
    chain_call(L, ws, [chain_lin(L, att, t["o_w"], x, res=x, epi=1),
                       chain_lin(L, x, t["gu_w"], act, epi=2, norm_w=t["ln2"]),
                       chain_lin(L, act, t["down_w"], x, res=x, epi=1),
                       chain_lin(L, x, t["qkv_w"], qkv, bias=t["qkv_b"], norm_w=t["ln1"])], B)
    torch.cuda.synchronize()
    return x, act, qkv


@pytest.mark.parametrize("B", [1, 3, 8, 16, 24, 48, 64, 96, 128])
@pytest.mark.parametrize("H,I,Q", [(3584, 18944, 4608), (1024, 2816, 1536)])
def test_chain_layer_matches_reference(L, chws, B, H, I, Q):
    """o_proj + residual -> RMSNorm -> gate/up + SwiGLU -> down_proj + residual -> RMSNorm -> qkv + bias in one launch."""
    t = layer_tensors(B, H, I, Q, seed=100)
    x, act, qkv = run_layer_chain(L, chws, t, B, I, Q)
    # the first linear alone (a chain of one gives the bits it gives inside the longer chain): every later check then
    # starts from the kernel's own activations and isolates one linear
    x1 = t["x"][:B].clone()
    chain_call(L, chws, [chain_lin(L, t["att"][:B], t["o_w"], x1, res=x1, epi=1)], B)
    torch.cuda.synchronize()
    lin_o = ref_linear(t["att"][:B], t["o_w"])
    close_bf16(x1, (lin_o.float() + t["x"][:B].float()).to(BF), "chain o_proj + residual", mag=lin_o)
    xn = hf_rmsnorm(x1, t["ln2"], 1e-6)
    Wg = t["gu_w"].view(I // 64, 2, 64, H)[:, 0].reshape(I, H)
    Wu = t["gu_w"].view(I // 64, 2, 64, H)[:, 1].reshape(I, H)
    g, u = ref_linear(xn, Wg), ref_linear(xn, Wu)
    close_bf16(act, torch.nn.functional.silu(g) * u, "chain norm + gate/up + swiglu", ulps=8.0, frac_exact=0.9,
               mag=g.float().abs() * u.float().abs())
    lin_d = ref_linear(act, t["down_w"])                       # from the kernel's own activations: isolates each linear
    x2 = (lin_d.float() + x1.float()).to(BF)
    close_bf16(x, x2, "chain down + residual (after o_proj + residual)", ulps=3.0, frac_exact=0.9, mag=lin_d)
    qkv_ref = ref_linear(hf_rmsnorm(x, t["ln1"], 1e-6), t["qkv_w"], t["qkv_b"])
    close_bf16(qkv, qkv_ref, "chain qkv", mag=qkv_ref)
    ctr_at = 296 * 128 * 128 * 4 + 192 * 296 * 4          # done / stage2 / exit counters follow the epoch-valued flags (chain.cu)
    assert int(chws[ctr_at:ctr_at + 4 * (2 * 192 + 1)].view(torch.int32).abs().sum()) == 0, "chain counters must return to zero"


def test_chain_batch_invariance_and_determinism(L, chws):
    H, I, Q = 3584, 18944, 4608
    t = layer_tensors(128, H, I, Q, seed=200)
    alone = [run_layer_chain(L, chws, {**t, "att": t["att"][b:b + 1], "x": t["x"][b:b + 1]}, 1, I, Q) for b in range(3)]
    for Bb in (3, 16, 17, 64, 96, 128):
        x, act, qkv = run_layer_chain(L, chws, t, Bb, I, Q)
        x_b, act_b, qkv_b = run_layer_chain(L, chws, t, Bb, I, Q)
        assert torch.equal(x, x_b) and torch.equal(act, act_b) and torch.equal(qkv, qkv_b), f"run-to-run difference at B={Bb}"
        for b in range(3):
            assert torch.equal(alone[b][0][0], x[b]) and torch.equal(alone[b][1][0], act[b]) and \
                torch.equal(alone[b][2][0], qkv[b]), f"row {b} differs between B=1 and B={Bb}"


def test_chain_of_one_equals_skinny_stream_k(L, skws, chws):
    """A linear wide enough for the stream-K skinny kernel (more than 74 tiles) gives the same bits through the chain."""
    B, N, K = 5, 152064, 3584
    X, W, nw = rnd(B, K, seed=50), rnd(N, K, scale=K ** -0.5, seed=51), (1 + 0.1 * rnd(K, seed=52).float()).to(BF)
    D1 = torch.empty((B, N), device="cuda", dtype=BF)
    D2 = torch.empty((B, N), device="cuda", dtype=BF)
    skinny_call(L, skws, X, W, D1, B, N, K, norm_w=nw, eps=1e-6)
    chain_call(L, chws, [chain_lin(L, X, W, D2, norm_w=nw)], B)
    torch.cuda.synchronize()
    assert torch.equal(D1, D2)


@pytest.mark.parametrize("rows,dim", [(7, 1280), (3, 3584), (999, 5120)])
def test_rmsnorm(L, rows, dim):
    x, w = rnd(rows, dim, seed=50), (1 + 0.1 * rnd(dim, seed=51).float()).to(BF)
    y = torch.empty_like(x)
    L.call("ocrb_rmsnorm_bf16", x.data_ptr(), dim, w.data_ptr(), y.data_ptr(), dim, rows, dim, 1e-6, sp())
    close_bf16(y, hf_rmsnorm(x, w, 1e-6), "rmsnorm", ulps=1.0, frac_exact=0.995)


def rotate_half(x):
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def test_rope_vision(L):
    S, H, hd = 500, 16, 80
    qkv = rnd(S, 3 * H * hd, seed=60)
    ang = torch.rand(S, hd // 2, device="cuda") * 50
    emb = torch.cat((ang, ang), -1)
    cos, sin = emb.cos().contiguous(), emb.sin().contiguous()
    q, k, v = qkv.view(S, 3, H, hd).unbind(1)
    qf, kf = q.float(), k.float()
    c, s = cos.unsqueeze(-2), sin.unsqueeze(-2)
    wq = (qf * c + rotate_half(qf) * s).to(BF)
    wk = (kf * c + rotate_half(kf) * s).to(BF)
    out = qkv.clone()
    L.call("ocrb_rope_vision", out.data_ptr(), S, H, hd, cos.data_ptr(), sin.data_ptr(), sp())
    o = out.view(S, 3, H, hd)
    assert torch.equal(o[:, 0], wq) and torch.equal(o[:, 1], wk) and torch.equal(o[:, 2], v)


def test_rope_text_bf16(L):
    T, nq, nkv, hd = 77, 28, 4, 128
    q, k = rnd(T, nq * hd, seed=61), rnd(T, nkv * hd, seed=62)
    ang = torch.rand(T, hd // 2, device="cuda") * 100
    emb = torch.cat((ang, ang), -1)
    cos, sin = emb.cos().to(BF).contiguous(), emb.sin().to(BF).contiguous()
    wq = (q.view(T, nq, hd) * cos[:, None]) + (rotate_half(q.view(T, nq, hd)) * sin[:, None])
    wk = (k.view(T, nkv, hd) * cos[:, None]) + (rotate_half(k.view(T, nkv, hd)) * sin[:, None])
    q2, k2 = q.clone(), k.clone()
    L.call("ocrb_rope_text", q2.data_ptr(), nq * hd, k2.data_ptr(), nkv * hd, T, nq, nkv, hd, cos.data_ptr(),
           sin.data_ptr(), sp())
    assert torch.equal(q2.view(T, nq, hd), wq) and torch.equal(k2.view(T, nkv, hd), wk)


def sdpa_ref(q, k, v, causal, scale):
    # q: [T, H, hd] one sequence, fp32 reference
    qf, kf, vf = (t.float().transpose(0, 1) for t in (q, k, v))
    if kf.shape[0] != qf.shape[0]:
        rep = qf.shape[0] // kf.shape[0]
        kf, vf = kf.repeat_interleave(rep, 0), vf.repeat_interleave(rep, 0)
    s = (qf @ kf.transpose(1, 2)) * scale
    if causal:
        T = s.shape[-1]
        s = s.masked_fill(torch.triu(torch.ones(T, T, device=s.device, dtype=torch.bool), 1), float("-inf"))
    return (torch.softmax(s, -1) @ vf).transpose(0, 1)


@pytest.mark.parametrize("hd,nq,nkv,causal,lens", [
    (80, 16, 16, 0, [64, 48, 16, 12, 64, 1]),
    (80, 16, 16, 0, [700]),
    (128, 28, 4, 1, [300, 65, 1]),
    (128, 8, 8, 1, [1036]),
    (64, 4, 2, 1, [130]),
])
def test_attention_varlen(L, hd, nq, nkv, causal, lens):
    T = sum(lens)
    q, k, v = rnd(T, nq * hd, seed=70), rnd(T, nkv * hd, seed=71), rnd(T, nkv * hd, seed=72)
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device="cuda")
    out = torch.full((T, nq * hd), float("nan"), device="cuda", dtype=BF)
    scale = hd ** -0.5
    L.call("ocrb_attention_varlen", q.data_ptr(), nq * hd, k.data_ptr(), nkv * hd, v.data_ptr(), nkv * hd,
           out.data_ptr(), nq * hd, cu.data_ptr(), len(lens), max(lens), nq, nkv, hd, scale, causal, sp())
    torch.cuda.synchronize()
    off = 0
    for n in lens:
        want = sdpa_ref(q[off:off + n].view(n, nq, hd), k[off:off + n].view(n, nkv, hd), v[off:off + n].view(n, nkv, hd),
                        bool(causal), scale)
        got = out[off:off + n].view(n, nq, hd).float()
        err = (got - want).abs().max().item()
        assert err < 2e-2, f"attention len {n}: max err {err}"
        off += n


@pytest.mark.parametrize("hd", [80, 64])
def test_window_attention_equals_general_kernel(L, hd):
    """Sequences of <= 64 tokens (the vision tower's windows) go through the persistent double-buffered window kernel:
    bit-equal to the general one-CTA-per-tile kernel (reached by declaring max_seqlen = 65) on 630 windows of mixed length."""
    nq = 16
    g = torch.Generator().manual_seed(3)
    lens = [64] * 500 + [int(x) for x in torch.randint(1, 65, (130,), generator=g)]
    T = sum(lens)
    q, k, v = rnd(T, nq * hd, seed=73), rnd(T, nq * hd, seed=74), rnd(T, nq * hd, seed=75)
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device="cuda")
    outs = []
    for max_len in (64, 65):
        out = torch.full((T, nq * hd), float("nan"), device="cuda", dtype=BF)
        L.call("ocrb_attention_varlen", q.data_ptr(), nq * hd, k.data_ptr(), nq * hd, v.data_ptr(), nq * hd, out.data_ptr(),
               nq * hd, cu.data_ptr(), len(lens), max_len, nq, nq, hd, hd ** -0.5, 0, sp())
        torch.cuda.synchronize()
        outs.append(out)
    assert not torch.isnan(outs[0].float()).any()
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("hd,nq,nkv,causal,lens", [
    (80, 16, 16, 0, [700]),
    (80, 16, 16, 0, [256, 257, 511, 64, 1]),
    (80, 16, 16, 0, [3996, 3996]),
    (128, 28, 4, 1, [1036, 300, 65, 1]),
    (128, 28, 4, 1, [1036] * 6),
    (128, 8, 8, 1, [1036]),
    (128, 4, 4, 0, [513, 129]),
    (128, 4, 2, 1, [128, 256, 384]),
])
def test_flash_attention_tcgen05(L, hd, nq, nkv, causal, lens):
    """tcgen05 / TMEM / TMA flash attention (vision full-attention blocks and prefill) against an fp32 softmax reference,
    through strided views of ONE fused qkv buffer (as the model calls it); rows of other heads / the padding of the
    last tile must not leak in; the output of every token outside [0, T) stays untouched."""
    T = sum(lens)
    W = (nq + 2 * nkv) * hd
    qkv = rnd(T, W, seed=75)
    q, k, v = qkv[:, : nq * hd], qkv[:, nq * hd: (nq + nkv) * hd], qkv[:, (nq + nkv) * hd:]
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device="cuda")
    out = torch.full((T + 8, nq * hd), float("nan"), device="cuda", dtype=BF)
    scale = hd ** -0.5
    L.call("ocrb_flash_attention_bf16", q.data_ptr(), W, k.data_ptr(), W, v.data_ptr(), W, out.data_ptr(), nq * hd,
           cu.data_ptr(), len(lens), T, max(lens), nq, nkv, hd, scale, causal, sp())
    torch.cuda.synchronize()
    assert torch.isnan(out[T:]).all(), "rows past the last token were written"
    off = 0
    worst = 0.0
    for n in lens:
        want = sdpa_ref(q[off:off + n].reshape(n, nq, hd), k[off:off + n].reshape(n, nkv, hd), v[off:off + n].reshape(n, nkv, hd),
                        bool(causal), scale)
        got = out[off:off + n].view(n, nq, hd).float()
        assert not torch.isnan(got).any(), f"len {n}: NaN in the output"
        err = (got - want).abs().max().item()
        worst = max(worst, err)
        assert err < 2e-2, f"flash attention (tcgen05) len {n}: max err {err}"
        off += n
    # the legacy mma.sync kernel on the same inputs agrees to bf16 rounding of the output
    out2 = torch.empty((T, nq * hd), device="cuda", dtype=BF)
    L.call("ocrb_attention_varlen", q.data_ptr(), W, k.data_ptr(), W, v.data_ptr(), W, out2.data_ptr(), nq * hd,
           cu.data_ptr(), len(lens), max(lens), nq, nkv, hd, scale, causal, sp())
    torch.cuda.synchronize()
    assert (out[:T].float() - out2.float()).abs().max().item() < 2e-2


@pytest.mark.parametrize("n_splits,max_pages,nq,nkv,hd,ctx", [
    (6, 20, 28, 4, 128, [100, 37, 250]), (1, 20, 28, 4, 128, [100, 37, 250]), (3, 24, 28, 4, 128, [100, 37, 250]),
    (4, 20, 4, 2, 128, [100, 37, 250]), (2, 20, 8, 2, 64, [100, 37, 250]), (20, 20, 28, 4, 128, [0, 15, 16, 17, 319]),
    (17, 132, 28, 4, 128, [1036, 1547, 2111, 127, 128]), (9, 132, 64, 8, 128, [1036, 2000]),
    (5, 40, 16, 1, 128, [639, 1, 63, 64, 65, 300, 301])])
def test_decode_attention_paged(L, n_splits, max_pages, nq, nkv, hd, ctx):
    B, page = len(ctx), 16
    n_pages = B * max_pages
    kc = rnd(n_pages, nkv, page, hd, seed=80)            # cache layout [page][kv head][token][hd]
    vc = rnd(n_pages, nkv, page, hd, seed=81)
    perm = torch.randperm(n_pages, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    bt = perm.view(B, max_pages).to(torch.int32).contiguous()
    qkv = rnd(B, (nq + 2 * nkv) * hd, seed=82)
    ang = torch.rand(B, hd // 2, device="cuda") * 100
    emb = torch.cat((ang, ang), -1)
    cos, sin = emb.cos().to(BF).contiguous(), emb.sin().to(BF).contiguous()
    ctx_d = torch.tensor(ctx, dtype=torch.int32, device="cuda")
    ws = torch.empty(B * nq * n_splits * (hd + 2), device="cuda", dtype=torch.float32)
    out = torch.empty(B, nq * hd, device="cuda", dtype=BF)
    kc0, vc0 = kc.clone(), vc.clone()
    L.call("ocrb_decode_attention", qkv.data_ptr(), qkv.stride(0), kc.data_ptr(), vc.data_ptr(), n_pages, bt.data_ptr(), max_pages,
           ctx_d.data_ptr(), B, page, nq, nkv, hd, cos.data_ptr(), sin.data_ptr(), hd ** -0.5, out.data_ptr(), nq * hd,
           ws.data_ptr(), n_splits, sp())
    torch.cuda.synchronize()
    for b in range(B):
        q = qkv[b, :nq * hd].view(nq, hd)
        kn = qkv[b, nq * hd:(nq + nkv) * hd].view(nkv, hd)
        vn = qkv[b, (nq + nkv) * hd:].view(nkv, hd)
        qr = q * cos[b] + rotate_half(q) * sin[b]
        kr = kn * cos[b] + rotate_half(kn) * sin[b]
        pos = torch.arange(ctx[b], device="cuda")
        pg = bt[b, pos // page].long()
        K = torch.cat([kc0[pg, :, pos % page], kr[None]], 0)   # [ctx+1, nkv, hd]
        V = torch.cat([vc0[pg, :, pos % page], vn[None]], 0)
        want = sdpa_ref(qr[None], K, V, False, hd ** -0.5)[0]
        err = (out[b].view(nq, hd).float() - want).abs().max().item()
        assert err < 2e-2, f"decode attention b={b}: {err}"
        # the new token was appended at position ctx[b]
        p_new = bt[b, ctx[b] // page].long()
        assert torch.equal(kc[p_new, :, ctx[b] % page], kr)
        assert torch.equal(vc[p_new, :, ctx[b] % page], vn)


@pytest.mark.parametrize("n_splits,max_pages,nq,nkv,ctx", [
    (6, 20, 28, 4, [100, 37, 250]), (1, 20, 28, 4, [100, 37, 250]), (3, 24, 28, 4, [100, 37, 250]), (4, 20, 4, 2, [100, 37, 250]),
    (20, 20, 28, 4, [0, 15, 16, 17, 319]), (17, 132, 28, 4, [1036, 1547, 2111, 127, 128]), (9, 132, 64, 8, [1036, 2000]),
    (5, 40, 8, 1, [639, 1, 63, 64, 65, 300, 301]), (13, 97, 28, 4, [1036 + 7 * i for i in range(24)]),
    (13, 97, 28, 4, [1100 + 3 * i for i in range(96)])])
def test_plan_attention_equals_decode_attention(L, chws, n_splits, max_pages, nq, nkv, ctx):
    """The attention op of a plan (csrc/chain.cu) = ocrb_decode_attention bit for bit: output, cache append and all."""
    import ctypes
    hd, B, page = 128, len(ctx), 16
    n_pages = B * max_pages
    kc = rnd(n_pages, nkv, page, hd, seed=80)
    vc = rnd(n_pages, nkv, page, hd, seed=81)
    perm = torch.randperm(n_pages, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    bt = perm.view(B, max_pages).to(torch.int32).contiguous()
    qkv = rnd(B, (nq + 2 * nkv) * hd, seed=82)
    ang = torch.rand(B, hd // 2, device="cuda") * 100
    emb = torch.cat((ang, ang), -1)
    cos, sin = emb.cos().to(BF).contiguous(), emb.sin().to(BF).contiguous()
    ctx_d = torch.tensor(ctx, dtype=torch.int32, device="cuda")
    res = []
    for fused in (False, True):
        k, v = kc.clone(), vc.clone()
        ws = torch.zeros(B * nq * n_splits * (hd + 2), device="cuda", dtype=torch.float32)
        out = torch.full((B, nq * hd), float("nan"), device="cuda", dtype=BF)
        if not fused:
            L.call("ocrb_decode_attention", qkv.data_ptr(), qkv.stride(0), k.data_ptr(), v.data_ptr(), n_pages, bt.data_ptr(),
                   max_pages, ctx_d.data_ptr(), B, page, nq, nkv, hd, cos.data_ptr(), sin.data_ptr(), hd ** -0.5, out.data_ptr(),
                   nq * hd, ws.data_ptr(), n_splits, sp())
        else:
            op = L.ChainOp()
            op.kind = 1
            op.att = L.ChainAttention(qkv.data_ptr(), qkv.stride(0), k.data_ptr(), v.data_ptr(), n_pages, bt.data_ptr(), max_pages,
                                      ctx_d.data_ptr(), page, nq, nkv, hd, cos.data_ptr(), sin.data_ptr(), hd ** -0.5,
                                      out.data_ptr(), nq * hd, ws.data_ptr(), n_splits)
            arr = (L.ChainOp * 1)(op)
            plan = torch.zeros(int(L.load().ocrb_chain_plan_bytes(1)) + 64, dtype=torch.uint8, device="cuda")
            pp = plan.data_ptr() + (-plan.data_ptr()) % 64
            L.call("ocrb_chain_plan_build", ctypes.addressof(arr), 1, B, chws.data_ptr(), pp, sp())
            for _ in range(2):          # twice: the counters must come back to zero, the second run sees the appended row again
                L.call("ocrb_chain_plan_run", pp, 1, B, chws.data_ptr(), sp())
        torch.cuda.synchronize()
        res.append((out, k, v))
    assert torch.equal(res[0][0], res[1][0]), "attention output differs"
    assert torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][2], res[1][2]), "KV append differs"


def test_argmax_step_first_index_and_eos(L):
    B, V, max_new = 4, 152064, 8
    logits = rnd(B, V, seed=90)
    logits[0, 777] = 50.0
    logits[0, 90000] = 50.0          # tie -> lowest index
    logits[1, 151645] = 60.0         # eos
    logits[2, V - 1] = 70.0
    out = torch.full((B, max_new), -1, dtype=torch.int32, device="cuda")
    nxt = torch.zeros(B, dtype=torch.int32, device="cuda")
    fin = torch.tensor([0, 0, 0, 1], dtype=torch.int32, device="cuda")
    ctx = torch.tensor([10, 20, 30, 40], dtype=torch.int32, device="cuda")
    step = torch.zeros(1, dtype=torch.int32, device="cuda")
    L.call("ocrb_argmax_step", logits.data_ptr(), V, B, V, 151645, 151645, max_new, out.data_ptr(), nxt.data_ptr(),
           fin.data_ptr(), ctx.data_ptr(), step.data_ptr(), 1, sp())
    torch.cuda.synchronize()
    assert nxt.tolist() == [777, 151645, V - 1, 151645]
    assert out[:, 0].tolist() == nxt.tolist() and fin.tolist() == [0, 1, 0, 1]
    assert ctx.tolist() == [11, 21, 31, 41] and step.item() == 1
    assert nxt[0].item() == int(torch.argmax(logits[0].float()))
