"""Seed-indexed synthetic inputs for tests and bench.

There are no sample pages or ground-truth files in the reference
(`data/input/.gitkeep` only), so every measurement and parity test runs on
pages and texts generated here.  numpy only: the generator must give the same
bytes on the CPU box and on the GPU box.

  page(seed)          RGB uint8 handwritten-note-like page, W=1024 x H=768
  text(seed, n_words) pseudo transcription (ASCII words + a few curly quotes)
  corrupt(text, ...)  seeded character corruption (for candidate variants / GT)
"""
from __future__ import annotations

import numpy as np

PAGE_W = 1024
PAGE_H = 768


def page(seed: int, width: int = PAGE_W, height: int = PAGE_H, *, ruled: bool = False,
         skew_deg: float | None = None, gray: bool = False) -> np.ndarray:
    """One synthetic page.  Paper 228+-10 uniform noise, ~28 lines of 2-px
    pseudo-strokes of intensity 20..90 at 65 % fill, optional ruled lines,
    global skew drawn from U(-3, 3) degrees (strokes are generated rotated)."""
    rng = np.random.default_rng(seed)
    img = rng.integers(218, 239, size=(height, width, 3), dtype=np.int32)
    if skew_deg is None:
        skew_deg = float(rng.uniform(-3.0, 3.0))
    th = np.deg2rad(skew_deg)
    c, s = np.cos(th), np.sin(th)
    cx, cy = width / 2.0, height / 2.0
    n_lines = max(4, int(28 * height / PAGE_H))
    margin_x = int(0.06 * width)
    line_h = (height * 0.88) / n_lines
    segs = []
    for li in range(n_lines):
        y_base = 0.06 * height + (li + 0.6) * line_h
        x = float(margin_x)
        if ruled:
            segs.append((margin_x * 0.5, y_base + 0.35 * line_h, width - margin_x * 0.5,
                         y_base + 0.35 * line_h, 150.0))
        while x < width - margin_x:
            wlen = rng.uniform(25, 90)
            if rng.uniform() < 0.65:
                n_st = int(wlen / 6) + 1
                px = x + np.sort(rng.uniform(0, wlen, size=n_st))
                for k in range(n_st):
                    x0 = px[k]
                    y0 = y_base + rng.uniform(-0.3, 0.3) * line_h
                    x1 = x0 + rng.uniform(-5, 7)
                    y1 = y_base + rng.uniform(-0.3, 0.3) * line_h
                    segs.append((x0, y0, x1, y1, float(rng.integers(20, 91))))
            x += wlen + rng.uniform(8, 18)
    if segs:
        sg = np.asarray(segs, dtype=np.float64)
        t = np.linspace(0.0, 1.0, 24)[None, :]
        xs = sg[:, 0:1] + (sg[:, 2:3] - sg[:, 0:1]) * t
        ys = sg[:, 1:2] + (sg[:, 3:4] - sg[:, 1:2]) * t
        inten = np.repeat(sg[:, 4:5], t.shape[1], axis=1)
        # rotate about the page centre
        xr = cx + (xs - cx) * c - (ys - cy) * s
        yr = cy + (xs - cx) * s + (ys - cy) * c
        xi = np.floor(xr).astype(np.int64).ravel()
        yi = np.floor(yr).astype(np.int64).ravel()
        iv = inten.ravel().astype(np.int32)
        for dy in (0, 1):
            for dx in (0, 1):
                xx, yy = xi + dx, yi + dy
                ok = (xx >= 0) & (xx < width) & (yy >= 0) & (yy < height)
                img[yy[ok], xx[ok], :] = iv[ok, None]
    out = img.astype(np.uint8)
    if gray:
        return np.ascontiguousarray(out[:, :, 1])
    return out


_SYL = ["ka", "lo", "mi", "ren", "tu", "sha", "ve", "on", "dar", "el", "qui", "st", "ar", "the",
        "ing", "pro", "un", "ly", "ex", "co", "ba", "fi", "zu", "wh", "ight", "ou", "ea", "nd"]


def text(seed: int, n_words: int = 350) -> str:
    """Pseudo transcription with punctuation, capitals, line breaks and a few
    non-ASCII quotes/dashes so `normalize_text` has work to do."""
    rng = np.random.default_rng(1_000_003 + seed)
    words = []
    for i in range(n_words):
        k = int(rng.integers(1, 4))
        w = "".join(_SYL[int(j)] for j in rng.integers(0, len(_SYL), size=k))
        r = rng.uniform()
        if r < 0.08:
            w = w.capitalize()
        elif r < 0.10:
            w = "“" + w + "”"
        elif r < 0.12:
            w = w + "’s"
        elif r < 0.14:
            w = w + " —"
        if rng.uniform() < 0.10:
            w += ","
        elif rng.uniform() < 0.07:
            w += "."
        words.append(w)
        if rng.uniform() < 0.08:
            words.append("\n")
    return " ".join(words).replace(" \n ", "\n")


def corrupt(s: str, seed: int, rate: float = 0.05) -> str:
    """Seeded character-level corruption: substitute / delete / insert."""
    rng = np.random.default_rng(2_000_003 + seed)
    out = []
    alphabet = "abcdefghijklmnopqrstuvwxyz "
    for ch in s:
        r = rng.uniform()
        if r < rate / 3:
            continue
        if r < 2 * rate / 3:
            out.append(alphabet[int(rng.integers(0, len(alphabet)))])
            continue
        out.append(ch)
        if r < rate:
            out.append(alphabet[int(rng.integers(0, len(alphabet)))])
    return "".join(out)


def rule_lines(page: np.ndarray, *, first: int = 60, pitch: int = 57, margin: int = 30, color=(70, 70, 90),
               thick: int = 2) -> np.ndarray:
    """Copy of `page` with notebook ruling: dark `thick`-px lines every `pitch` rows from x = margin to W - margin, line i
    dropping i % 3 pixels over its length (so the lines are not all perfectly horizontal).  Pure numpy."""
    out = page.copy()
    h, w = out.shape[:2]
    x = np.arange(margin, w - margin)
    for i, y0 in enumerate(range(first, h - 40, pitch)):
        y = y0 + ((x - margin) * (i % 3) * 2 + (w - 2 * margin)) // (2 * (w - 2 * margin))
        for k in range(thick):
            if out.ndim == 3:
                out[y + k, x] = np.asarray(color, out.dtype)
            else:
                out[y + k, x] = (int(color[0]) * 9798 + int(color[1]) * 19235 + int(color[2]) * 3735 + 16384) >> 15
    return out
