"""ORACLE (test infrastructure, not product code) -- restatement of the reference's
first-party text arithmetic: normalize_text, levenshtein, cer/wer/tier1_metrics,
compare_versions, merge_versions (`/root/reference/ocr_agent/tools.py:51-139,326-493`).

The O(n*m) loops run in oracle/text_ref.c (gcc); small cases also have a pure-Python
form (`levenshtein_py`) that follows tools.py:69-83 line by line.  Pinned by
tests/golden/text_golden.json, generated from the unmodified reference functions.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build() -> str:
    so = os.path.join(_HERE, "_build", "liboracle_text.so")
    src = os.path.join(_HERE, "text_ref.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        L.oracle_levenshtein_i32.restype = ctypes.c_int32
        L.oracle_levenshtein_i32.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32]
        L.oracle_lcs_align_i32.restype = None
        L.oracle_lcs_align_i32.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32,
                                           ctypes.c_void_p]
        _LIB = L
    return _LIB


# tools.py:51-63.  `\s` for str patterns = the 29 code points of SURVEY A.8.
_WS = set([0x09, 0x0A, 0x0B, 0x0C, 0x0D, 0x1C, 0x1D, 0x1E, 0x1F, 0x20, 0x85, 0xA0, 0x1680,
           0x2028, 0x2029, 0x202F, 0x205F, 0x3000] + list(range(0x2000, 0x200B)))
_MAP = {0x2018: "'", 0x2019: "'", 0x201C: '"', 0x201D: '"', 0x2013: "-", 0x2014: "-"}


def normalize_text(text: str, lower: bool = False) -> str:
    out = []
    pending_space = False
    for ch in text:
        cp = ord(ch)
        if cp in _WS:
            pending_space = True
            continue
        if pending_space and out:
            out.append(" ")
        pending_space = False
        out.append(_MAP.get(cp, ch))
    t = "".join(out)
    return t.lower() if lower else t


def _codes(s: str) -> np.ndarray:
    return np.fromiter((ord(c) for c in s), dtype=np.int32, count=len(s))


def _word_ids(*lists, key=lambda w: w):
    table: dict = {}
    res = []
    for wl in lists:
        res.append(np.fromiter((table.setdefault(key(w), len(table)) for w in wl), dtype=np.int32, count=len(wl)))
    return res


def _lev_ids(a: np.ndarray, b: np.ndarray) -> int:
    a = np.ascontiguousarray(a, np.int32)
    b = np.ascontiguousarray(b, np.int32)
    return int(_lib().oracle_levenshtein_i32(a.ctypes.data, len(a), b.ctypes.data, len(b)))


def levenshtein(a: str, b: str) -> int:
    return _lev_ids(_codes(a), _codes(b))


def levenshtein_py(a, b) -> int:
    """tools.py:69-83 verbatim in structure (small cases only)."""
    n, m = len(a), len(b)
    dp = list(range(m + 1))
    for i in range(1, n + 1):
        prev, dp[0] = dp[0], i
        for j in range(1, m + 1):
            cur = dp[j]
            dp[j] = min(dp[j] + 1, dp[j - 1] + 1, prev + (0 if a[i - 1] == b[j - 1] else 1))
            prev = cur
    return dp[m]


def levenshtein_words(a: list, b: list) -> int:
    ia, ib = _word_ids(a, b)
    return _lev_ids(ia, ib)


def tier1_metrics(ground_truth: str, ocr_output: str, lower: bool = False) -> dict:
    gt = normalize_text(ground_truth, lower)
    ocr = normalize_text(ocr_output, lower)
    cer_val = levenshtein(gt, ocr) / max(len(gt), 1)
    gt_words, ocr_words = gt.split(), ocr.split()
    jg, jo = " ".join(gt_words), " ".join(ocr_words)
    wer_char = levenshtein(jg, jo) / max(len(jg), 1)
    wer_tok = levenshtein_words(gt_words, ocr_words) / max(len(gt_words), 1)
    return {"input": ocr_output, "cer": round(cer_val, 4), "wer": round(wer_char, 4),
            "wer_token": round(wer_tok, 4), "exact_match": gt == ocr, "gt_chars": len(gt),
            "ocr_chars": len(ocr)}


def evaluate(transcription: str, ground_truth=None, lower: bool = False) -> dict:
    res = {}
    if ground_truth is not None:
        res["tier1_raw_vs_gt"] = tier1_metrics(ground_truth, transcription, lower)
    return res


def find_differing_segments(w1: list, w2: list) -> list:
    """tools.py:353-405 (greedy 9-word look-ahead; O(n), host only)."""
    segs = []
    i = j = 0
    n1, n2 = len(w1), len(w2)
    while i < n1 and j < n2:
        if w1[i] == w2[j]:
            i += 1
            j += 1
            continue
        si, sj = i, j
        found = False
        for look in range(1, min(10, max(n1 - i, n2 - j) + 1)):
            if i + look < n1 and w1[i + look] == w2[j]:
                segs.append({"position": si, "v1_text": " ".join(w1[si:i + look]), "v2_text": ""})
                i += look
                found = True
                break
            if j + look < n2 and w2[j + look] == w1[i]:
                segs.append({"position": si, "v1_text": "", "v2_text": " ".join(w2[sj:j + look])})
                j += look
                found = True
                break
        if not found:
            segs.append({"position": si, "v1_text": w1[i], "v2_text": w2[j]})
            i += 1
            j += 1
    if i < n1 or j < n2:
        segs.append({"position": i, "v1_text": " ".join(w1[i:]), "v2_text": " ".join(w2[j:])})
    return segs


def compare_versions(v1: str, v2: str) -> dict:
    n1, n2 = normalize_text(v1), normalize_text(v2)
    d = levenshtein(n1, n2)
    mx = max(len(n1), len(n2), 1)
    w1, w2 = n1.split(), n2.split()
    return {"agreement_rate": round((1 - d / mx) * 100, 1), "char_edit_distance": d,
            "word_edit_distance": levenshtein_words(w1, w2),
            "differing_segments": find_differing_segments(w1, w2)}


def align_to_backbone(backbone: list, words: list) -> list:
    ib, iw = _word_ids(backbone, words, key=lambda w: w.lower())
    out = np.empty(len(backbone), np.int32)
    ib = np.ascontiguousarray(ib)
    iw = np.ascontiguousarray(iw)
    _lib().oracle_lcs_align_i32(ib.ctypes.data, len(ib), iw.ctypes.data, len(iw), out.ctypes.data)
    return [words[j] if j >= 0 else None for j in out]


def align_to_backbone_py(backbone: list, words: list) -> list:
    """tools.py:465-493 in pure Python (full (n+1)x(m+1) table, `.lower()` per cell as the reference
    does): used for small cases and as the timed CPU baseline of the merge step."""
    n, m = len(backbone), len(words)
    tab = [[0] * (m + 1) for _ in range(n + 1)]
    for i in range(n):
        row, nxt = tab[i], tab[i + 1]
        bi = backbone[i]
        for j in range(m):
            if bi.lower() == words[j].lower():
                nxt[j + 1] = row[j] + 1
            else:
                up, left = row[j + 1], nxt[j]
                nxt[j + 1] = up if up >= left else left
    out = [None] * n
    i, j = n, m
    while i > 0 and j > 0:
        if backbone[i - 1].lower() == words[j - 1].lower():
            out[i - 1] = words[j - 1]
            i, j = i - 1, j - 1
        elif tab[i - 1][j] >= tab[i][j - 1]:
            i -= 1
        else:
            j -= 1
    return out


def merge_versions(versions: list) -> str:
    if not versions:
        return ""
    if len(versions) == 1:
        return versions[0]
    wls = [normalize_text(v).split() for v in versions]
    bi = max(range(len(wls)), key=lambda i: len(wls[i]))
    backbone = wls[bi]
    aligned = [align_to_backbone(backbone, wl) for wl in wls]
    merged = []
    for pos in range(len(backbone)):
        cands = [a[pos] for a in aligned if pos < len(a) and a[pos] is not None]
        if not cands:
            merged.append(backbone[pos])
            continue
        votes: dict = {}
        for c in cands:
            votes[c] = votes.get(c, 0) + 1
        mv = max(votes.values())
        winners = [w for w, c in votes.items() if c == mv]
        if len(winners) == 1:
            merged.append(winners[0])
        else:
            uniq = list(dict.fromkeys(cands))
            merged.append(uniq[0] if len(uniq) == 1 else "[" + "|".join(uniq) + "]")
    return " ".join(merged)
