"""The Levenshtein wavefront and LCS-align kernels (handwritten-ocr_b200/csrc/textops_kernels.cuh) run thread by thread on
the CPU through tests/emu/cuda_emu.h, with the dispatch of the product's C ABI, against the C oracle (oracle/text_ref.c):
strip hand-over between threads, band boundaries of every instantiation, ragged / empty pairs, the LCS tie rule."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import text_ref as T

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "emu")
SRC = [os.path.join(EMU, "emu_text.cpp"), os.path.join(EMU, "cuda_emu.h"),
       os.path.join(HERE, "..", "handwritten-ocr_b200", "csrc", "textops_kernels.cuh")]


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(EMU, "_build", "libemu_text.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(s) for s in SRC):
        subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-fPIC", "-shared", "-o", so, SRC[0]], check=True)
    return ctypes.CDLL(so)


def pack(seqs):
    off = np.zeros(len(seqs) + 1, np.int32)
    off[1:] = np.cumsum([len(s) for s in seqs])
    flat = np.concatenate([np.asarray(s, np.int32) for s in seqs] + [np.zeros(1, np.int32)])
    return np.ascontiguousarray(flat), off


def P(a):
    return ctypes.c_void_p(a.ctypes.data)


def lev_batch(emu, pairs, force=0):
    fa, oa = pack([a for a, _ in pairs])
    fb, ob = pack([b for _, b in pairs])
    out = np.full(len(pairs), -99, np.int32)
    mlb = max(len(b) for _, b in pairs)
    rc = emu.emu_levenshtein_batch(P(fa), P(oa), P(fb), P(ob), len(pairs), mlb, P(out), force)
    assert rc == 0
    return out.tolist()


def test_levenshtein_wavefront_every_emulable_instantiation(emu):
    rng = np.random.default_rng(0)
    # <1,32>: ragged edge cases around the warp-wide band
    pairs = [(rng.integers(0, 4, n).astype(np.int32), rng.integers(0, 4, m).astype(np.int32))
             for n, m in [(1, 1), (32, 32), (33, 31), (31, 32), (64, 1), (1, 32), (65, 17), (0, 5), (5, 0), (0, 0), (7, 30)]]
    assert lev_batch(emu, pairs) == [T._lev_ids(a, b) for a, b in pairs]
    # <4,64>: strips of 4 columns, b lengths on both sides of a strip boundary
    pairs = [(rng.integers(0, 6, n).astype(np.int32), rng.integers(0, 6, m).astype(np.int32))
             for n, m in [(40, 33), (200, 255), (3, 256), (256, 4), (129, 128), (90, 131)]]
    assert lev_batch(emu, pairs) == [T._lev_ids(a, b) for a, b in pairs]
    # <8,128>: one pair that needs it + short pairs riding along in the same (wider) instantiation
    pairs = [(rng.integers(0, 20, 300).astype(np.int32), rng.integers(0, 20, 700).astype(np.int32)),
             (rng.integers(0, 3, 9).astype(np.int32), rng.integers(0, 3, 8).astype(np.int32)),
             (rng.integers(0, 3, 50).astype(np.int32), rng.integers(0, 3, 257).astype(np.int32))]
    assert lev_batch(emu, pairs) == [T._lev_ids(a, b) for a, b in pairs]
    # <16,256> forced on modest pairs (its own size class is too slow to emulate): same recurrence, 16-column strips
    pairs = [(rng.integers(0, 5, 70).astype(np.int32), rng.integers(0, 5, 100).astype(np.int32)),
             (rng.integers(0, 5, 33).astype(np.int32), rng.integers(0, 5, 16).astype(np.int32))]
    assert lev_batch(emu, pairs, force=3) == [T._lev_ids(a, b) for a, b in pairs]


def test_levenshtein_text_like_pairs(emu):
    """Near-identical sequences (what compare_versions sees): long diagonals of zeros, few edits."""
    rng = np.random.default_rng(1)
    pairs = []
    for n in (60, 180, 240):
        a = rng.integers(0, 40, n).astype(np.int32)
        b = a.copy()
        for _ in range(5):
            k = int(rng.integers(0, len(b)))
            op = int(rng.integers(0, 3))
            if op == 0:
                b[k] = 99
            elif op == 1:
                b = np.delete(b, k)
            else:
                b = np.insert(b, k, 98)
        pairs.append((a, b.astype(np.int32)))
    assert lev_batch(emu, pairs) == [T._lev_ids(a, b) for a, b in pairs]


def test_lcs_align_tie_rule_and_ragged_pairs(emu):
    rng = np.random.default_rng(2)
    shapes = [(1, 1), (5, 9), (40, 37), (300, 280), (17, 1), (1, 23), (64, 64)]
    bbs = [rng.integers(0, 5, n).astype(np.int32) for n, _ in shapes]       # small alphabet: many ties
    ws = [rng.integers(0, 5, m).astype(np.int32) for _, m in shapes]
    fb, ob = pack(bbs)
    fw, ow = pack(ws)
    sizes = np.array([len(b) * len(w) for b, w in zip(bbs, ws)], np.int64)
    ws_off = np.zeros(len(shapes), np.int64)
    ws_off[1:] = np.cumsum(sizes)[:-1]
    work = np.zeros(int(sizes.sum()) + 1, np.uint8)
    aligned = np.full(int(ob[-1]) + 1, -77, np.int32)
    rc = emu.emu_lcs_align_batch(P(fb), P(ob), P(fw), P(ow), len(shapes), max(len(b) for b in bbs), P(aligned), P(work), P(ws_off))
    assert rc == 0
    L = T._lib()
    for k, (b, w) in enumerate(zip(bbs, ws)):
        want = np.empty(len(b), np.int32)
        b, w = np.ascontiguousarray(b), np.ascontiguousarray(w)
        L.oracle_lcs_align_i32(b.ctypes.data, len(b), w.ctypes.data, len(w), want.ctypes.data)
        assert np.array_equal(aligned[ob[k]:ob[k + 1]], want), (k, shapes[k])
