"""Candidate agreement, majority-vote merge and CER/WER (reference: ocr_agent/tools.py:51-139,
305-350, 411-493) with the O(n*m) dynamic programmes on the GPU.

Host side (this file): text normalisation, symbol tables, rounding of ratios, vote counting --
all O(n).  Device side (csrc/textops.cu through the C ABI): `ocrb_levenshtein_batch` (one warp per
pair, anti-diagonal wavefront) and `ocrb_lcs_align_batch` (one CTA per pair).  Distances and
alignments are integer-exact, so every dict/string returned here equals the reference's.
"""
from __future__ import annotations

import re
from typing import Sequence

import numpy as np
import torch

from . import _lib

_QUOTE_DASH = str.maketrans({"‘": "'", "’": "'", "“": '"', "”": '"',
                             "–": "-", "—": "-"})
_WS_RUN = re.compile(r"\s+")


def normalize_text(text: str, lower: bool = False) -> str:
    """tools.py:51-63: curly quotes/dashes -> ASCII, whitespace runs -> one space, strip, optional lower."""
    t = _WS_RUN.sub(" ", text.translate(_QUOTE_DASH)).strip()
    return t.lower() if lower else t


def _dev():
    if not torch.cuda.is_available():
        raise _lib.OcrbError("handwritten-ocr_b200 text ops need a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _pack(seqs: Sequence[np.ndarray]):
    off = np.zeros(len(seqs) + 1, np.int32)
    if seqs:
        off[1:] = np.cumsum([len(s) for s in seqs])
    flat = np.concatenate(seqs).astype(np.int32, copy=False) if off[-1] else np.zeros(1, np.int32)
    return flat, off


def levenshtein_ids_batch(pairs: Sequence[tuple]) -> list:
    """Edit distances of many (int32 array, int32 array) pairs in ONE kernel launch."""
    if not pairs:
        return []
    dev = _dev()
    fa, oa = _pack([p[0] for p in pairs])
    fb, ob = _pack([p[1] for p in pairs])
    max_b = int(max(len(p[1]) for p in pairs))
    host = torch.from_numpy(np.concatenate([fa, oa, fb, ob])).pin_memory()
    d = host.to(dev, non_blocking=True)
    a_, oa_, b_, ob_ = torch.split(d, [len(fa), len(oa), len(fb), len(ob)])
    out = torch.empty(len(pairs), dtype=torch.int32, device=dev)
    ws = torch.empty(len(pairs) * (max_b + 1), dtype=torch.int32, device=dev)
    _lib.call("ocrb_levenshtein_batch", _lib.ptr(a_), _lib.ptr(oa_), _lib.ptr(b_), _lib.ptr(ob_),
              len(pairs), max_b, _lib.ptr(out), _lib.ptr(ws), _lib.stream_ptr())
    return [int(v) for v in out.cpu().tolist()]


def _codes(s: str) -> np.ndarray:
    return np.frombuffer(s.encode("utf-32-le", "surrogatepass"), dtype=np.int32)


def _ids(*lists, key=None):
    table: dict = {}
    out = []
    for wl in lists:
        arr = np.empty(len(wl), np.int32)
        for i, w in enumerate(wl):
            k = w if key is None else key(w)
            arr[i] = table.setdefault(k, len(table))
        out.append(arr)
    return out


def levenshtein(a: str, b: str) -> int:
    """tools.py:69-83 (character level)."""
    return levenshtein_ids_batch([(_codes(a), _codes(b))])[0]


def _levenshtein_words(a: list, b: list) -> int:
    """tools.py:86-100 (word-token level)."""
    ia, ib = _ids(a, b)
    return levenshtein_ids_batch([(ia, ib)])[0]


def cer(ground_truth: str, ocr_output: str, lower: bool = False) -> float:
    gt = normalize_text(ground_truth, lower)
    ocr = normalize_text(ocr_output, lower)
    return levenshtein(gt, ocr) / max(len(gt), 1)


def wer(ground_truth: str, ocr_output: str, lower: bool = False) -> float:
    gw = normalize_text(ground_truth, lower).split()
    ow = normalize_text(ocr_output, lower).split()
    return _levenshtein_words(gw, ow) / max(len(gw), 1)


def tier1_metrics_batch(items: Sequence[tuple], lower: bool = False) -> list:
    """[(ground_truth, ocr_output), ...] -> list of tier-1 dicts (tools.py:119-139), with all
    3*len(items) distances computed in one launch.  Used by evaluate() and the eval_final caller."""
    prep = []
    pairs = []
    for gt_raw, ocr_raw in items:
        gt = normalize_text(gt_raw, lower)
        ocr = normalize_text(ocr_raw, lower)
        gw, ow = gt.split(), ocr.split()
        jg, jo = " ".join(gw), " ".join(ow)
        ig, io = _ids(gw, ow)
        pairs += [(_codes(gt), _codes(ocr)), (_codes(jg), _codes(jo)), (ig, io)]
        prep.append((ocr_raw, gt, ocr, gw, jg))
    d = levenshtein_ids_batch(pairs)
    res = []
    for k, (ocr_raw, gt, ocr, gw, jg) in enumerate(prep):
        dc, dj, dw = d[3 * k: 3 * k + 3]
        res.append({
            "input": ocr_raw,
            "cer": round(dc / max(len(gt), 1), 4),
            "wer": round(dj / max(len(jg), 1), 4),
            "wer_token": round(dw / max(len(gw), 1), 4),
            "exact_match": gt == ocr,
            "gt_chars": len(gt),
            "ocr_chars": len(ocr),
        })
    return res


def tier1_metrics(ground_truth: str, ocr_output: str, lower: bool = False) -> dict:
    return tier1_metrics_batch([(ground_truth, ocr_output)], lower)[0]


def evaluate(transcription: str, ground_truth: str | None = None, lower: bool = False) -> dict:
    """tools.py:305-320."""
    result = {}
    if ground_truth is not None:
        print("  [eval] Computing CER/WER against ground truth...")
        result["tier1_raw_vs_gt"] = tier1_metrics(ground_truth, transcription, lower)
    return result


def _find_differing_segments(w1: list, w2: list) -> list:
    """tools.py:353-405: greedy resynchronisation with a 9-word look-ahead (O(n), host)."""
    segs = []
    n1, n2 = len(w1), len(w2)
    i = j = 0
    while i < n1 and j < n2:
        if w1[i] == w2[j]:
            i, j = i + 1, j + 1
            continue
        start = i
        horizon = min(10, max(n1 - i, n2 - j) + 1)
        hit = None
        for look in range(1, horizon):
            if i + look < n1 and w1[i + look] == w2[j]:
                hit = ("v1", look)
                break
            if j + look < n2 and w2[j + look] == w1[i]:
                hit = ("v2", look)
                break
        if hit is None:
            segs.append({"position": start, "v1_text": w1[i], "v2_text": w2[j]})
            i, j = i + 1, j + 1
        elif hit[0] == "v1":
            segs.append({"position": start, "v1_text": " ".join(w1[i:i + hit[1]]), "v2_text": ""})
            i += hit[1]
        else:
            segs.append({"position": start, "v1_text": "", "v2_text": " ".join(w2[j:j + hit[1]])})
            j += hit[1]
    if i < n1 or j < n2:
        segs.append({"position": i, "v1_text": " ".join(w1[i:]), "v2_text": " ".join(w2[j:])})
    return segs


def compare_versions_batch(pairs: Sequence[tuple]) -> list:
    """[(v1, v2), ...] -> list of tools.py:326-350 dicts; the 2 * len(pairs) char and word distances share ONE launch
    (a folder batch compares the candidates of all its pages at once)."""
    prep, lev = [], []
    for v1, v2 in pairs:
        n1, n2 = normalize_text(v1), normalize_text(v2)
        w1, w2 = n1.split(), n2.split()
        i1, i2 = _ids(w1, w2)
        lev += [(_codes(n1), _codes(n2)), (i1, i2)]
        prep.append((n1, n2, w1, w2))
    d = levenshtein_ids_batch(lev)
    out = []
    for k, (n1, n2, w1, w2) in enumerate(prep):
        dc, dw = d[2 * k], d[2 * k + 1]
        out.append({
            "agreement_rate": round((1 - dc / max(len(n1), len(n2), 1)) * 100, 1),
            "char_edit_distance": dc,
            "word_edit_distance": dw,
            "differing_segments": _find_differing_segments(w1, w2),
        })
    return out


def compare_versions(v1: str, v2: str) -> dict:
    """tools.py:326-350; char and word distances share one launch."""
    return compare_versions_batch([(v1, v2)])[0]


def lcs_align_pairs(pairs: Sequence[tuple]) -> list:
    """[(backbone ids, version ids), ...] -> per pair an int32 array [len(backbone)] of indices into the version (-1 = gap):
    tools.py:465-493 for every pair in ONE launch (one CTA per pair)."""
    dev = _dev()
    out = [np.full(len(b), -1, np.int32) for b, _ in pairs]
    live = [k for k, (b, v) in enumerate(pairs) if len(b) and len(v)]
    if not live:
        return out
    fb, ob = _pack([pairs[k][0] for k in live])
    fw, ow = _pack([pairs[k][1] for k in live])
    sizes = np.array([len(pairs[k][0]) * len(pairs[k][1]) for k in live], np.int64)
    ws_off = np.zeros(len(live), np.int64)
    ws_off[1:] = np.cumsum(sizes)[:-1]
    ints = torch.from_numpy(np.concatenate([fb, ob, fw, ow])).to(dev)
    b_, ob_, w_, ow_ = torch.split(ints, [len(fb), len(ob), len(fw), len(ow)])
    wsoff_d = torch.from_numpy(ws_off).to(dev)
    ws = torch.empty(max(int(sizes.sum()), 1), dtype=torch.uint8, device=dev)
    aligned = torch.empty(int(ob[-1]), dtype=torch.int32, device=dev)
    _lib.call("ocrb_lcs_align_batch", _lib.ptr(b_), _lib.ptr(ob_), _lib.ptr(w_), _lib.ptr(ow_), len(live),
              int(max(len(pairs[k][0]) for k in live)), _lib.ptr(aligned), _lib.ptr(ws), _lib.ptr(wsoff_d), _lib.stream_ptr())
    al = aligned.cpu().numpy()
    for j, k in enumerate(live):
        out[k] = al[ob[j]:ob[j + 1]]
    return out


def lcs_align_batch(backbone_ids: np.ndarray, version_ids: Sequence[np.ndarray]) -> list:
    """Align every version to the backbone (tools.py:465-493) in one launch.
    Returns, per version, an int32 array [len(backbone)] of indices into that version (-1 = gap)."""
    return lcs_align_pairs([(backbone_ids, v) for v in version_ids])


def merge_versions_batch(version_lists: Sequence[list]) -> list:
    """tools.py:411-462 for many pages at once: every candidate of every page is aligned to its page's first-longest one
    (case-insensitive LCS) in ONE launch, then the per-position vote runs on the host; ties keep all variants as `[a|b]` in
    candidate order."""
    results = [None] * len(version_lists)
    jobs, pairs = [], []
    for p, versions in enumerate(version_lists):
        if not versions:
            results[p] = ""
            continue
        if len(versions) == 1:
            results[p] = versions[0]
            continue
        word_lists = [normalize_text(v).split() for v in versions]
        backbone_idx = max(range(len(word_lists)), key=lambda i: len(word_lists[i]))
        backbone = word_lists[backbone_idx]
        ids = _ids(backbone, *word_lists, key=str.lower)
        jobs.append((p, word_lists, backbone, len(pairs)))
        pairs += [(ids[0], v) for v in ids[1:]]
    aligned_all = lcs_align_pairs(pairs) if pairs else []
    for p, word_lists, backbone, first in jobs:
        aligned = aligned_all[first:first + len(word_lists)]
        merged = []
        for pos, bw in enumerate(backbone):
            cands = [wl[a[pos]] for wl, a in zip(word_lists, aligned) if a[pos] >= 0]
            if not cands:
                merged.append(bw)
                continue
            votes: dict = {}
            for c in cands:
                votes[c] = votes.get(c, 0) + 1
            top = max(votes.values())
            winners = [w for w, c in votes.items() if c == top]
            if len(winners) == 1:
                merged.append(winners[0])
            else:
                uniq = list(dict.fromkeys(cands))
                merged.append(uniq[0] if len(uniq) == 1 else "[" + "|".join(uniq) + "]")
        results[p] = " ".join(merged)
    return results


def merge_versions(versions: list) -> str:
    """tools.py:411-462: align every candidate to the first-longest one (case-insensitive LCS), then
    vote per backbone position; ties keep all variants as `[a|b]` in candidate order."""
    return merge_versions_batch([versions])[0]
