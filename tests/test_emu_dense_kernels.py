"""The small SIMT kernels of the VLM glue (handwritten-ocr_b200/csrc/dense_kernels.cuh: RMSNorm, vision / text RoPE, row copies,
paged-KV prefill write, greedy argmax + bookkeeping, residual add) on the CPU through tests/emu/cuda_emu.h.  Copies and the
argmax are checked exactly against numpy; the vectorised RoPE kernels against their scalar twins (bit-equal); RMSNorm against
a float64 restatement of HF's rounding points within one bf16 ulp (rsqrtf and the reduction order are not bit-pinned on the
host)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "emu")
SRC = [os.path.join(EMU, "emu_dense.cpp"), os.path.join(EMU, "cuda_emu.h"),
       os.path.join(HERE, "..", "handwritten-ocr_b200", "csrc", "dense_kernels.cuh")]
LL = ctypes.c_longlong


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(EMU, "_build", "libemu_dense.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(s) for s in SRC):
        subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, SRC[0]],
                       check=True)
    return ctypes.CDLL(so)


def P(a):
    return ctypes.c_void_p(a.ctypes.data)


def to_bf16(x):
    """float32 array -> bf16 bit patterns (uint16), round to nearest even."""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    u = u + 0x7FFF + ((u >> 16) & 1)
    return (u >> 16).astype(np.uint16)


def from_bf16(b):
    return (b.astype(np.uint32) << 16).view(np.float32)


def aligned_u16(shape):
    n = int(np.prod(shape))
    raw = np.zeros(n * 2 + 64, np.uint8)
    off = (-raw.ctypes.data) % 64
    return raw[off:off + 2 * n].view(np.uint16).reshape(shape)


def test_rope_text_vectorised_equals_scalar(emu):
    rng = np.random.default_rng(1)
    T, nq, nkv, hd = 5, 4, 2, 32
    q0 = to_bf16(rng.normal(0, 1, (T, nq * hd)))
    k0 = to_bf16(rng.normal(0, 1, (T, nkv * hd)))
    cosT, sinT = aligned_u16((T, hd)), aligned_u16((T, hd))
    cosT[...] = to_bf16(np.cos(rng.uniform(0, 6, (T, hd))))
    sinT[...] = to_bf16(np.sin(rng.uniform(0, 6, (T, hd))))
    outs = []
    for variant in (0, 1):
        q, k = aligned_u16(q0.shape), aligned_u16(k0.shape)
        q[...], k[...] = q0, k0
        assert emu.emu_rope_text(P(q), LL(nq * hd), P(k), LL(nkv * hd), T, nq, nkv, hd, P(cosT), P(sinT), variant) == 0
        outs.append((q.copy(), k.copy()))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    assert not np.array_equal(outs[0][0], q0)


def test_rope_vision_vectorised_equals_scalar(emu):
    rng = np.random.default_rng(2)
    S, heads, hd = 7, 2, 32
    x0 = to_bf16(rng.normal(0, 1, (S, 3 * heads * hd)))
    cosT = np.ascontiguousarray(np.cos(rng.uniform(0, 6, (S, hd))), np.float32)          # tables are [S, hd]
    sinT = np.ascontiguousarray(np.sin(rng.uniform(0, 6, (S, hd))), np.float32)
    outs = []
    for variant in (0, 1):
        x = aligned_u16(x0.shape)
        x[...] = x0
        assert emu.emu_rope_vision(P(x), S, heads, hd, P(cosT), P(sinT), variant) == 0
        outs.append(x.copy())
    assert np.array_equal(outs[0], outs[1])
    assert np.array_equal(outs[0][:, 2 * heads * hd:], x0[:, 2 * heads * hd:]), "v must be untouched"
    assert not np.array_equal(outs[0][:, :heads * hd], x0[:, :heads * hd])


def test_rmsnorm_both_kernels(emu):
    rng = np.random.default_rng(3)
    rows, dim, eps = 11, 256, 1e-6
    x, w, = aligned_u16((rows, dim)), aligned_u16((dim,))
    x[...] = to_bf16(rng.normal(0, 2, (rows, dim)))
    w[...] = to_bf16(rng.normal(1, 0.2, dim))
    xf, wf = from_bf16(x).astype(np.float64), from_bf16(w).astype(np.float64)
    rstd = 1.0 / np.sqrt((xf * xf).mean(-1, keepdims=True) + eps)
    normed = from_bf16(to_bf16((xf * rstd).astype(np.float32))).astype(np.float64)        # HF: .to(bf16) before the weight
    want = from_bf16(to_bf16((wf * normed).astype(np.float32)))
    got = []
    for variant in (0, 1):
        y = aligned_u16((rows, dim))
        assert emu.emu_rmsnorm(P(x), LL(dim), P(w), P(y), LL(dim), rows, dim, ctypes.c_float(eps), variant) == 0
        g = from_bf16(y)
        got.append(y.copy())
        ulp = np.maximum(np.abs(want), 1e-30) * 2.0 ** -7
        assert (np.abs(g - want) <= ulp).all(), variant
        assert (y == to_bf16(want)).mean() > 0.98, variant
    assert (got[0] == got[1]).mean() > 0.98


def test_rows_copy_and_kv_write_and_residual(emu):
    rng = np.random.default_rng(4)
    src = aligned_u16((9, 64))
    src[...] = rng.integers(0, 65536, (9, 64), dtype=np.uint16)
    dst = aligned_u16((12, 64))
    si = np.array([8, 0, 3, 3], np.int32)
    di = np.array([1, 11, 5, 6], np.int32)
    assert emu.emu_rows_copy(P(src), LL(64), P(si), P(dst), LL(64), P(di), 4, 64) == 0
    want = np.zeros_like(dst)
    want[di] = src[si]
    assert np.array_equal(dst, want)
    # paged KV write: 2 sequences (5 and 3 tokens), pages of 4 tokens, layout [page][kv head][token][hd]
    n_kv, hd, page = 2, 16, 4
    T, cu = 8, np.array([0, 5, 8], np.int32)
    k, v = aligned_u16((T, n_kv * hd)), aligned_u16((T, n_kv * hd))
    k[...] = rng.integers(0, 65536, k.shape, dtype=np.uint16)
    v[...] = rng.integers(0, 65536, v.shape, dtype=np.uint16)
    bt = np.array([[3, 1], [0, 7]], np.int32)
    kc, vc = aligned_u16((8, n_kv, page, hd)), aligned_u16((8, n_kv, page, hd))
    assert emu.emu_kv_write_prefill(P(k), LL(n_kv * hd), P(v), LL(n_kv * hd), P(kc), P(vc), P(bt), 2, P(cu), 2, T, page, n_kv, hd) == 0
    wk, wv = np.zeros_like(kc), np.zeros_like(vc)
    for s in range(2):
        for t in range(cu[s], cu[s + 1]):
            pos = t - cu[s]
            pg = bt[s, pos // page]
            wk[pg, :, pos % page] = k[t].reshape(n_kv, hd)
            wv[pg, :, pos % page] = v[t].reshape(n_kv, hd)
    assert np.array_equal(kc, wk) and np.array_equal(vc, wv)
    # residual add: bf16(float(x) + float(y))
    x, y = aligned_u16((3, 32)), aligned_u16((3, 32))
    x[...] = to_bf16(rng.normal(0, 1, (3, 32)))
    y[...] = to_bf16(rng.normal(0, 1, (3, 32)))
    want = to_bf16(from_bf16(x) + from_bf16(y))
    assert emu.emu_residual_add(P(x), LL(32), P(y), LL(32), 3, 32) == 0
    assert np.array_equal(x, want)


def test_argmax_step_first_index_and_bookkeeping(emu):
    rng = np.random.default_rng(5)
    B, V, max_new, eos = 4, 1003, 6, 7
    logits = aligned_u16((B, 1008))
    vals = rng.normal(0, 1, (B, 1008)).astype(np.float32)
    vals[0, 500] = vals[0, 20] = 9.0           # tie: the first index wins
    vals[1, 1002] = 9.0                        # in the scalar tail (V % 8 != 0)
    vals[2, eos] = 9.0                         # EOS: the sequence finishes
    vals[3, 5] = 9.0                           # already finished: pad is written
    logits[...] = to_bf16(vals)
    out = np.full((B, max_new), -1, np.int32)
    nxt, fin = np.zeros(B, np.int32), np.array([0, 0, 0, 1], np.int32)
    ctx, step = np.full(B, 10, np.int32), np.array([2], np.int32)
    assert emu.emu_argmax_step(P(logits), LL(1008), B, V, eos, eos, max_new, P(out), P(nxt), P(fin), P(ctx), P(step), 1) == 0
    assert nxt.tolist() == [20, 1002, eos, eos]
    assert out[:, 2].tolist() == [20, 1002, eos, eos] and (out[:, [0, 1, 3, 4, 5]] == -1).all()
    assert fin.tolist() == [0, 0, 1, 1] and ctx.tolist() == [11] * B and step[0] == 3
