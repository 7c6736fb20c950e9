"""The throughput image kernels (handwritten-ocr_b200/csrc/image_fast.cuh) run thread by thread on the CPU through
tests/emu/cuda_emu.h and are compared bit for bit with the oracle -- no GPU needed.  This pins the indexing, byte-permute
selectors, packed 16-bit-lane arithmetic, dp2a tap pairing, CLAHE cell boundaries and the hull tree before any GPU time
is spent; the `-m gpu` tests remain the parity tests of the compiled kernels."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import image_ref as R

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "emu")
SRC = [os.path.join(EMU, "emu_image.cpp"), os.path.join(EMU, "cuda_emu.h"),
       os.path.join(HERE, "..", "handwritten-ocr_b200", "csrc", "image_fast.cuh")]


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(EMU, "_build", "libemu_image.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(s) for s in SRC):
        subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, SRC[0]],
                       check=True)
    return ctypes.CDLL(so)


def aligned(shape, dtype=np.uint8, fill=None):
    """A numpy array whose data pointer is 64-byte aligned (the kernels' vector paths require 16)."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    raw = np.empty(n + 64, np.uint8)
    off = (-raw.ctypes.data) % 64
    a = raw[off:off + n].view(dtype).reshape(shape)
    if fill is not None:
        a[...] = fill
    return a


def P(a):
    return ctypes.c_void_p(a.ctypes.data)


def page(rng, H, W, C, kind):
    """Mostly-paper pages with dark strokes (long equal runs, as real pages have) or plain noise."""
    shape = (H, W, 3) if C == 3 else (H, W)
    if kind == "noise":
        a = rng.integers(0, 256, shape, dtype=np.uint8)
    else:
        a = np.full(shape, 235, np.uint8)
        a += rng.integers(0, 12, shape, dtype=np.uint8)
        for _ in range(max(4, H * W // 600)):
            y, x = int(rng.integers(0, H)), int(rng.integers(0, W))
            h, w = int(rng.integers(1, 4)), int(rng.integers(2, 14))
            a[y:y + h, x:x + w] = rng.integers(0, 90)
    out = aligned(shape)
    out[...] = a
    return out


@pytest.mark.parametrize("C", [1, 3])
@pytest.mark.parametrize("shape", [(5, 16), (9, 32), (33, 64), (2, 48)])
def test_sharpen_vec16(emu, C, shape):
    rng = np.random.default_rng(11 + C)
    H, W = shape
    for kind in ("noise", "paper"):
        imgs = [page(rng, H, W, C, kind) for _ in range(2)]
        src = aligned((2,) + imgs[0].shape)
        src[0], src[1] = imgs
        dst = aligned(src.shape, fill=7)
        assert emu.emu_sharpen(P(src), P(dst), 2, H, W, C) == 0
        for i in range(2):
            assert np.array_equal(dst[i], R.sharpen(imgs[i])), (shape, C, kind, i)


@pytest.mark.parametrize("C,shape", [(3, (64, 128)), (1, (64, 128)), (3, (96, 256)), (1, (40, 96)), (3, (40, 96)), (3, (50, 70)),
                                     (1, (50, 70))])
def test_high_contrast_fused(emu, C, shape):
    """(64,128), (96,256): vector histogram path (tile width 16 / 32) + cell kernel; (40,96): scalar histogram path
    (tile width 12) + cell kernel refused (boundaries not multiples of 4) -> LUTs only; (50,70): reflect-101 extension."""
    rng = np.random.default_rng(5)
    H, W = shape
    for kind in (("paper",) if shape in ((64, 128), (40, 96)) else ("noise",)):     # the emulated scan is slow: one kind each
        imgs = [page(rng, H, W, C, kind), page(rng, H, W, C, "noise" if kind == "paper" else "paper")]
        src = aligned((2,) + imgs[0].shape)
        src[0], src[1] = imgs
        dst = aligned((2, H, W), fill=9)
        gray = aligned((2, H, W), fill=3)
        lut = aligned((2, 64, 256))
        vec = ctypes.c_int(0)
        rc = emu.emu_high_contrast(P(src), P(dst), P(gray), P(lut), 2, H, W, C, ctypes.byref(vec))
        assert rc in (0, 1)
        if shape in ((64, 128), (96, 256)):
            assert rc == 0 and vec.value == 1, "expected the vector histogram path and the cell kernel"
        for i in range(2):
            g = R.rgb2gray(imgs[i]) if C == 3 else imgs[i]
            if C == 3:
                assert np.array_equal(gray[i], g), "gray page written by the histogram pass"
            luts = R.clahe_luts(g)
            luts = luts[0] if isinstance(luts, tuple) else luts
            assert np.array_equal(lut[i].reshape(-1), np.asarray(luts, np.uint8).reshape(-1)), (shape, C, kind)
            if rc == 0:
                assert np.array_equal(dst[i], R.clahe(g)), (shape, C, kind, i)


@pytest.mark.parametrize("C", [1, 3])
@pytest.mark.parametrize("shape", [(96, 64), (100, 132), (37, 61), (200, 24)])
def test_binarize_tile(emu, C, shape):
    rng = np.random.default_rng(3)
    H, W = shape
    for kind in ("paper", "noise"):
        img = page(rng, H, W, C, kind)
        src = aligned((1,) + img.shape)
        src[0] = img
        dst = aligned((1, H, W), fill=5)
        assert emu.emu_binarize(P(src), P(dst), 1, H, W, C) == 0
        g = R.rgb2gray(img) if C == 3 else img
        assert np.array_equal(dst[0], R.adaptive_threshold(g)), (shape, C, kind)


def skewed_page(rng, H, W, C, angle_deg):
    """Dark text lines on paper, rotated by shearing the line positions."""
    a = np.full((H, W), 240, np.uint8)
    t = np.tan(np.deg2rad(angle_deg))
    for y0 in range(H // 8, H - H // 8, max(6, H // 12)):
        for x in range(W // 10, W - W // 10):
            if rng.random() < 0.7:
                y = int(round(y0 + (x - W / 2) * t))
                if 0 <= y < H:
                    a[y, x] = rng.integers(0, 100)
    if C == 3:
        a = np.stack([a, a, a], -1)
    out = aligned(a.shape)
    out[...] = a
    return out


@pytest.mark.parametrize("C", [1, 3])
def test_deskew_angle_tree_and_warp(emu, C):
    rng = np.random.default_rng(17)
    for (H, W), ang in ([((96, 128), 2.0), ((130, 160), -3.5), ((72, 64), 7.0)] if C == 3 else [((96, 128), -2.0), ((72, 64), 5.0)]):
        imgs = [skewed_page(rng, H, W, C, ang), skewed_page(rng, H, W, C, -ang / 2)]
        blank = aligned(imgs[0].shape, fill=255)               # <= 100 dark pixels -> NaN, page unchanged
        src = aligned((3,) + imgs[0].shape)
        src[0], src[1], src[2] = imgs[0], imgs[1], blank
        angle = np.zeros(3, np.float64)
        M = np.zeros((3, 6), np.float64)
        ext = np.zeros((3, H, 3), np.int32)
        assert emu.emu_deskew_angle(P(src), 3, H, W, C, P(angle), P(M), P(ext)) == 0
        for i in range(2):
            g = R.rgb2gray(imgs[i]) if C == 3 else imgs[i]
            cnt, mn, mx = R.dark_extents(g)
            assert np.array_equal(ext[i, :, 0], cnt)
            rows = cnt > 0
            assert np.array_equal(ext[i, rows, 1], mn[rows]) and np.array_equal(ext[i, rows, 2], mx[rows])
            ref = R.deskew_angle(g)
            assert ref is not None and angle[i] == ref, ((H, W), C, i, angle[i], ref)
            assert np.array_equal(M[i].reshape(2, 3), R.rotation_matrix(W // 2, H // 2, ref))
        assert np.isnan(angle[2]) and np.isnan(M[2]).all()
        dst = aligned(src.shape, fill=1)
        assert emu.emu_warp(P(src), P(dst), 3, H, W, C, P(M)) == 0
        for i in range(2):
            assert np.array_equal(dst[i], R.warp_affine_cubic(imgs[i], M[i].reshape(2, 3))), ((H, W), C, i)
        assert np.array_equal(dst[2], blank)


def test_hull_tree_equals_single_scan_on_random_extents(emu):
    """Extent patterns a page would never produce (ragged, sparse, single-column) through the 64 -> 8 -> 1 hull tree:
    the angle must equal the oracle's, which runs one monotone chain over all points."""
    rng = np.random.default_rng(23)
    for trial in range(12):
        H, W = int(rng.integers(70, 260)), 16 * int(rng.integers(3, 12))
        a = aligned((1, H, W), fill=255)
        mode = trial % 4
        for y in range(H):
            if mode == 0 and rng.random() < 0.5:
                continue
            if mode == 3 and not (H // 3 < y < H // 2):
                continue
            lo = int(rng.integers(0, W - 1))
            hi = int(rng.integers(lo, W)) if mode != 2 else lo
            a[0, y, lo] = 0
            a[0, y, hi] = 0
            if mode == 1:
                a[0, y, lo:hi + 1] = 0
        angle = np.zeros(1, np.float64)
        M = np.zeros((1, 6), np.float64)
        ext = np.zeros((1, H, 3), np.int32)
        assert emu.emu_deskew_angle(P(a), 1, H, W, 1, P(angle), P(M), P(ext)) == 0
        try:
            ref = R.deskew_angle(a[0])
        except Exception:       # degenerate hull in the oracle (collinear points): the kernel reports NaN
            ref = None
        if ref is None:
            assert np.isnan(angle[0]), (trial, angle[0])
        else:
            assert angle[0] == ref, (trial, H, W, angle[0], ref)


def test_clahe_cell_boundaries_match_the_per_pixel_formula(emu):
    """clahe_apply_cells_kernel trusts the host to say where the LUT quadruple changes.  For many page sizes the boundaries
    must reproduce, coordinate by coordinate, the tile index the per-pixel kernel (and OpenCV) derive in fp32:
    floor(x * inv_tw - 0.5)."""
    rng = np.random.default_rng(31)
    sizes = [(768, 1024), (1024, 768), (389, 517), (8, 8), (4000, 3000), (2481, 3508)] + \
            [(int(rng.integers(8, 4200)), int(rng.integers(8, 4200))) for _ in range(300)]
    f32 = np.float32
    for H, W in sizes:
        xb = np.zeros(10, np.int32)
        yb = np.zeros(10, np.int32)
        itw, ith = ctypes.c_float(), ctypes.c_float()
        ok = emu.emu_clahe_cells(H, W, P(xb), P(yb), ctypes.byref(itw), ctypes.byref(ith))
        for n, b, inv in ((W, xb, f32(itw.value)), (H, yb, f32(ith.value))):
            assert b[0] == 0 and b[9] == n and (np.diff(b) >= 0).all(), (H, W, b)
            idx = np.arange(n, dtype=f32)
            t1 = np.floor((idx * inv).astype(f32) - f32(0.5)).astype(np.int64)      # raw tile index, -1 .. 7
            cell = np.clip(t1 + 1, 0, 8)
            want = np.searchsorted(cell, np.arange(10), side="left")             # first coordinate of each cell
            want[9] = n
            assert np.array_equal(b, want), (H, W, n, b, want)
        if W % 4 == 0 and all(int(v) % 4 == 0 for v in xb[:9]) and np.diff(xb).max() <= 1024 and np.diff(yb).max() <= 1024:
            assert ok == 1, (H, W)
        else:
            assert ok == 0, (H, W)


def test_sharpen_packed_lanes_saturate_both_ways(emu):
    """Pixels drawn from the extremes: 5c - (u + d + l + r) reaches -1020 and +1275, so both clamps of the 16-bit-lane
    arithmetic and the bias of 1020 are exercised in every lane position."""
    rng = np.random.default_rng(41)
    vals = np.array([0, 1, 2, 127, 128, 253, 254, 255], np.uint8)
    for C, (H, W) in [(1, (12, 64)), (3, (12, 32))]:
        shape = (H, W, 3) if C == 3 else (H, W)
        img = aligned(shape)
        img[...] = vals[rng.integers(0, len(vals), shape)]
        src = aligned((1,) + shape)
        src[0] = img
        dst = aligned(src.shape, fill=9)
        assert emu.emu_sharpen(P(src), P(dst), 1, H, W, C) == 0
        ref = R.sharpen(img)
        assert np.array_equal(dst[0], ref)
        assert (ref == 0).any() and (ref == 255).any()


def test_round_half_even_by_magic_add():
    """The kernels round fp32 results in [0, 256) with `x + 1.5 * 2^23` and read the low mantissa bits (no F2I).  numpy's
    float32 add rounds to nearest-even exactly like FADD.RN: the trick must equal rint on ties and on random values."""
    f32 = np.float32
    rng = np.random.default_rng(7)
    x = np.concatenate([np.arange(0, 256, dtype=f32), np.arange(0, 256, dtype=f32) + f32(0.5),
                        np.nextafter(np.arange(0, 256, dtype=f32) + f32(0.5), f32(0)),
                        np.nextafter(np.arange(0, 256, dtype=f32) + f32(0.5), f32(300)),
                        rng.uniform(0, 255.49, 200000).astype(f32)])
    magic = f32(12582912.0)
    got = (x + magic).view(np.int32) - np.int32(0x4B400000)
    assert np.array_equal(got, np.rint(x).astype(np.int32))
    # bytes as floats without a conversion: 2^23 + b has b in its low mantissa bits
    b = np.arange(256, dtype=np.uint32)
    assert np.array_equal(((b | np.uint32(0x4B000000)).view(f32) - f32(8388608.0)), b.astype(f32))


def test_fuzz_odd_sizes_binarize_and_clahe_luts(emu):
    """Random small page sizes (odd widths, pages smaller than one tile, tiles cut by every edge) through the new kernels'
    scalar / clamped paths."""
    rng = np.random.default_rng(53)
    for _ in range(12):
        H, W, C = int(rng.integers(1, 150)), int(rng.integers(1, 150)), int(rng.choice([1, 3]))
        img = page(rng, H, W, C, "noise" if rng.random() < 0.5 else "paper")
        src = aligned((1,) + img.shape)
        src[0] = img
        dst = aligned((1, H, W), fill=5)
        assert emu.emu_binarize(P(src), P(dst), 1, H, W, C) == 0
        g = R.rgb2gray(img) if C == 3 else img
        assert np.array_equal(dst[0], R.adaptive_threshold(g)), (H, W, C)
    for _ in range(4):
        H, W, C = int(rng.integers(8, 90)), int(rng.integers(8, 90)), int(rng.choice([1, 3]))
        img = page(rng, H, W, C, "paper")
        src = aligned((1,) + img.shape)
        src[0] = img
        dst = aligned((1, H, W), fill=9)
        gray = aligned((1, H, W), fill=3)
        lut = aligned((1, 64, 256))
        vec = ctypes.c_int(0)
        rc = emu.emu_high_contrast(P(src), P(dst), P(gray), P(lut), 1, H, W, C, ctypes.byref(vec))
        assert rc in (0, 1), (H, W, C)
        g = R.rgb2gray(img) if C == 3 else img
        if C == 3:
            assert np.array_equal(gray[0], g)
        assert np.array_equal(lut[0].reshape(-1), np.asarray(R.clahe_luts(g)[0], np.uint8).reshape(-1)), (H, W, C)
        if rc == 0:
            assert np.array_equal(dst[0], R.clahe(g)), (H, W, C)


# ───────────── the general-shape kernels (csrc/image_general.cuh: any width, any alignment) ─────────────
def test_general_kernels_odd_widths(emu):
    """rgb2gray (16-pixel path + scalar tail), both general sharpen kernels, both general CLAHE apply kernels, the
    general extent kernel + one-thread hull scan, and the ruled-line mask, on sizes the fast paths refuse."""
    rng = np.random.default_rng(61)
    # rgb2gray: pixel counts around the 16-pixel vector
    for H, W in [(3, 5), (4, 16), (7, 37)]:
        img = page(rng, H, W, 3, "noise")
        src = aligned((2, H, W, 3))
        src[0], src[1] = img, img[::-1]
        dst = aligned((2, H, W), fill=1)
        assert emu.emu_rgb2gray(P(src), P(dst), 2, H, W) == 0
        assert np.array_equal(dst[0], R.rgb2gray(img)) and np.array_equal(dst[1], R.rgb2gray(img[::-1]))
    # sharpen: byte kernel on any width, word kernel where the row bytes are a multiple of 4
    for (H, W), C in [((5, 7), 3), ((4, 9), 1), ((6, 12), 3), ((3, 8), 1), ((2, 2), 1)]:
        img = page(rng, H, W, C, "noise")
        src = aligned((1,) + img.shape)
        src[0] = img
        for variant in (0, 1):
            dst = aligned(src.shape, fill=3)
            rc = emu.emu_sharpen_general(P(src), P(dst), 1, H, W, C, variant)
            if variant == 1 and ((W * C) % 4 or W < 4):
                assert rc == -1
                continue
            assert rc == 0 and np.array_equal(dst[0], R.sharpen(img)), (H, W, C, variant)
    # CLAHE apply: per pixel (odd width, reflect-101 extension) and four pixels per thread
    for (H, W), variant in [((23, 27), 0), ((24, 36), 1), ((16, 16), 1), ((31, 36), 0)]:
        img = page(rng, H, W, 1, "paper")
        src = aligned((1, H, W))
        src[0] = img
        dst = aligned((1, H, W), fill=4)
        lut = aligned((1, 64, 256))
        assert emu.emu_clahe_general(P(src), P(dst), P(lut), 1, H, W, variant) == 0
        assert np.array_equal(dst[0], R.clahe(img)), (H, W, variant)
    # deskew angle by the general extent kernel + the sequential scan (global-memory hull): odd width, RGB and gray
    for (H, W), C in [((90, 75), 3), ((70, 101), 1)]:
        img = skewed_page(rng, H, W, C, 3.0)
        src = aligned((1,) + img.shape)
        src[0] = img
        angle = np.zeros(1, np.float64)
        M = np.zeros((1, 6), np.float64)
        ext = np.zeros((1, H, 3), np.int32)
        hull = np.zeros((1, (4 * H + 8) * 2), np.int32)
        assert emu.emu_deskew_angle_general(P(src), 1, H, W, C, P(angle), P(M), P(ext), P(hull)) == 0
        g = R.rgb2gray(img) if C == 3 else img
        assert angle[0] == R.deskew_angle(g), (H, W, C)
    # ruled-line mask of remove_lines
    for (H, W), C in [((40, 64), 3), ((33, 50), 1)]:
        img = page(rng, H, W, C, "paper")
        img[10, 3:W - 3] = 60
        img[25, 5:W - 8] = 70                                       # two ruled lines
        src = aligned((1,) + img.shape)
        src[0] = img
        mask = aligned((1, H, W), fill=9)
        tmp = aligned((1, H, W), fill=9)
        nz = np.zeros(1, np.int32)
        assert emu.emu_remove_lines_mask(P(src), P(mask), P(nz), P(tmp), 1, H, W, C) == 0
        g = R.rgb2gray(img) if C == 3 else img
        want = R.lines_mask(g)
        assert np.array_equal(mask[0], want), (H, W, C)
        assert bool(nz[0]) == bool(want.any()) and want.any()
