"""GPU parity tests (run on the B200 box): CUDA path through the C ABI vs the oracle, vs the
committed golden vectors of the reference, and vs cv2 itself when importable."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import image_ref as R
from oracle import text_ref as T

pytestmark = pytest.mark.gpu

CHAINS = ["deskew+high_contrast+binarize", "high_contrast+binarize", "deskew+high_contrast+sharpen"]
SMALL = ["rgb_256x192", "rgb_203x157", "gray_256x192", "rgb_blank_128x96", "rgb_ruled_320x240"]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def pp(pkg):
    from handwritten_ocr_b200 import preprocess
    return preprocess


@pytest.fixture(scope="module")
def tx(pkg):
    from handwritten_ocr_b200 import textops
    return textops


def run1(pp, arr, fn):
    x = pp.to_device(arr)
    return fn(x)[0].cpu().numpy()


# ───────────── text ─────────────
def test_levenshtein_golden(tx, text_golden):
    for c in text_golden["levenshtein"]:
        assert tx.levenshtein(c["a"], c["b"]) == c["d"], (c["a"][:20], c["b"][:20])
    for c in text_golden["levenshtein_words"]:
        assert tx._levenshtein_words(c["a"], c["b"]) == c["d"]


def test_levenshtein_random_vs_oracle(tx):
    rng = np.random.default_rng(0)
    pairs = []
    for k in range(300):
        n = int(rng.integers(0, 140)); m = int(rng.integers(0, 140))
        alpha = int(rng.integers(2, 30))
        pairs.append((rng.integers(0, alpha, n).astype(np.int32), rng.integers(0, alpha, m).astype(np.int32)))
    # ragged edge cases: band boundaries of the 32-row wavefront
    for n, m in [(1, 1), (32, 32), (33, 31), (31, 33), (64, 1), (1, 64), (65, 97), (0, 5), (5, 0), (0, 0)]:
        pairs.append((rng.integers(0, 4, n).astype(np.int32), rng.integers(0, 4, m).astype(np.int32)))
    got = tx.levenshtein_ids_batch(pairs)
    want = [T._lev_ids(a, b) for a, b in pairs]
    assert got == want


def test_levenshtein_dispatch_buckets(tx):
    """Every strip-width / thread-count instantiation of the wavefront kernel, at its boundaries, in one batch."""
    rng = np.random.default_rng(5)
    sizes = [(40, 32), (40, 33), (300, 256), (300, 257), (1100, 1024), (900, 1025), (4200, 4096), (3000, 4097),
             (2500, 8192), (2000, 8193), (1500, 16384), (1200, 16500), (8071, 8103)]
    for n, m in sizes:
        a = rng.integers(0, 6, n).astype(np.int32)
        b = rng.integers(0, 6, m).astype(np.int32)
        assert tx.levenshtein_ids_batch([(a, b)]) == [T._lev_ids(a, b)], (n, m)
    # mixed lengths in one launch (the widest pair selects the instantiation; short pairs ride along)
    pairs = [(rng.integers(0, 5, n).astype(np.int32), rng.integers(0, 5, m).astype(np.int32))
             for n, m in [(3, 2000), (2000, 3), (700, 700), (1, 1), (0, 9), (2100, 1900)]]
    assert tx.levenshtein_ids_batch(pairs) == [T._lev_ids(a, b) for a, b in pairs]


def test_levenshtein_full_size_properties(tx, synth):
    """BASELINE sizes (2k and 8k chars): oracle equality + symmetry + identity + triangle bound."""
    a = synth.text(1, 350); b = synth.corrupt(a, 1, 0.06)
    big_a = synth.text(2, 1400); big_b = synth.corrupt(big_a, 2, 0.05)
    ca, cb, cA, cB = (tx._codes(s) for s in (a, b, big_a, big_b))
    d = tx.levenshtein_ids_batch([(ca, cb), (cb, ca), (ca, ca), (cA, cB), (cB, cA)])
    assert d[0] == d[1] == T.levenshtein(a, b)
    assert d[2] == 0
    assert d[3] == d[4] == T.levenshtein(big_a, big_b)
    assert abs(len(big_a) - len(big_b)) <= d[3] <= max(len(big_a), len(big_b))


def test_compare_merge_tier1_golden(tx, text_golden):
    for c in text_golden["compare_versions"]:
        assert tx.compare_versions(c["v1"], c["v2"]) == c["out"]
    for c in text_golden["merge_versions"]:
        assert tx.merge_versions(c["versions"]) == c["out"]
    for c in text_golden["tier1_metrics"]:
        assert tx.tier1_metrics(c["gt"], c["ocr"], c["lower"]) == c["out"]


def test_merge_vs_oracle_large(tx, synth):
    base = synth.text(7, 1400)
    vs = [synth.corrupt(base, s, 0.03 + 0.01 * s) for s in range(3)]
    assert tx.merge_versions(vs) == T.merge_versions(vs)
    assert tx.compare_versions(vs[0], vs[1]) == T.compare_versions(vs[0], vs[1])


def test_compare_merge_batches_equal_singles_and_goldens(tx, synth, text_golden):
    """A folder batch's agreement / merge step in two launches: the goldens of the unmodified reference, 32 synthetic pages
    of ragged length (empty and single-candidate pages included), all equal to the one-page calls and the oracle."""
    cmp_cases = text_golden["compare_versions"]
    assert tx.compare_versions_batch([(c["v1"], c["v2"]) for c in cmp_cases]) == [c["out"] for c in cmp_cases]
    mrg_cases = text_golden["merge_versions"]
    assert tx.merge_versions_batch([c["versions"] for c in mrg_cases]) == [c["out"] for c in mrg_cases]
    pages = []
    for p in range(32):
        base = synth.text(100 + p, 40 + 37 * (p % 9))
        pages.append([synth.corrupt(base, 3 * p + k, 0.02 + 0.01 * k) for k in range(3)])
    pages += [[], ["only one candidate"], ["", "two words"], ["same", "same", "same"]]
    merged = tx.merge_versions_batch(pages)
    assert merged == [T.merge_versions(v) for v in pages]
    pairs = [(v[0], v[1]) for v in pages if len(v) >= 2]
    assert tx.compare_versions_batch(pairs) == [T.compare_versions(a, b) for a, b in pairs]
    assert tx.compare_versions_batch([]) == [] and tx.merge_versions_batch([]) == []


# ───────────── image ─────────────
@pytest.mark.parametrize("name", SMALL)
def test_small_transforms_golden(pp, image_small, name):
    arr = image_small[f"{name}/input"]
    assert np.array_equal(run1(pp, arr, pp.high_contrast), image_small[f"{name}/high_contrast"])
    assert np.array_equal(run1(pp, arr, pp.binarize), image_small[f"{name}/binarize"])
    assert np.array_equal(run1(pp, arr, pp.sharpen), image_small[f"{name}/sharpen"])


@pytest.mark.parametrize("name", SMALL)
def test_small_deskew_golden(pp, image_small, name):
    arr = image_small[f"{name}/input"]
    ref_ang = float(image_small[f"{name}/angle"][0])
    x = pp.to_device(arr)
    ang, M = pp.deskew_angle(x)
    ang = float(ang[0].cpu())
    if np.isnan(ref_ang):
        assert np.isnan(ang)
        assert np.array_equal(pp.deskew(x)[0].cpu().numpy(), arr)
        return
    assert ang == ref_ang, "deskew angle must be bit-equal to cv2.minAreaRect's (golden, from the reference)"
    # warp is bit-exact given the reference's angle
    Mref = torch.from_numpy(R.rotation_matrix(arr.shape[1] // 2, arr.shape[0] // 2, ref_ang).reshape(1, 6))
    out = pp.warp_affine(x, Mref)[0].cpu().numpy()
    assert np.array_equal(out, image_small[f"{name}/deskew"])
    # and the device angle equals the oracle's restatement exactly
    assert ang == R.deskew_angle(R.rgb2gray(arr))


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_full_page_hashes(pp, synth, image_hashes, seed):
    ent = image_hashes[f"seed{seed}"]
    arr = synth.page(seed, ent["w"], ent["h"])
    assert sha(arr) == ent["input"]
    x = pp.to_device(arr)
    for step in ["high_contrast", "binarize", "sharpen"]:
        assert sha(pp.apply_transform(x, step)[0].cpu().numpy()) == ent[step], step
    ang, M = pp.deskew_angle(x)
    assert float(ang[0].cpu()) == ent["angle"], "deskew angle must be bit-equal to the reference's"
    Mref = torch.from_numpy(R.rotation_matrix(ent["w"] // 2, ent["h"] // 2, ent["angle"]).reshape(1, 6))
    assert sha(pp.warp_affine(x, Mref)[0].cpu().numpy()) == ent["deskew"]
    for ch in CHAINS:
        out = pp.apply_strategy(x, ch.split("+"))
        assert sha(out[0].cpu().numpy()) == ent[ch], ch
        pv, (gh, gw) = pp.pixel_values(out, dtype=torch.float32)
        assert [1, gh, gw] == ent[f"grid:{ch}"][0]
        assert sha(pv.cpu().numpy()) == ent[f"pv:{ch}"], ch
    pv, (gh, gw) = pp.pixel_values(x, dtype=torch.float32)
    assert sha(pv.cpu().numpy()) == ent["pv:original"]


def test_deskew_angle_vs_cv2_many_pages(pp, synth):
    """The rotating-calipers restatement against the installed cv2 (the wheel the reference calls) on 120 pages,
    both orientations, odd sizes and ruled paper: bit-equal angles (skipped where cv2 is not installed)."""
    cv2 = pytest.importorskip("cv2")
    pages = []
    for seed in range(120):
        if seed < 80:
            pages.append(synth.page(500 + seed, 1024, 768) if seed % 2 == 0 else synth.page(500 + seed, 768, 1024))
        else:
            pages.append(synth.page(500 + seed, 504 + (seed % 7) * 13, 392 + (seed % 5) * 11, ruled=(seed % 3 == 0)))
    bad = []
    for i, page in enumerate(pages):
        gray = cv2.cvtColor(page, cv2.COLOR_RGB2GRAY)
        coords = np.column_stack(np.where(gray < 128))
        if len(coords) <= 100:
            continue
        a = cv2.minAreaRect(coords)[-1]
        want = float(-(90 + a) if a < -45 else -a)
        ang, _ = pp.deskew_angle(pp.to_device(page))
        if float(ang[0].cpu()) != want:
            bad.append((i, want, float(ang[0].cpu())))
    assert not bad, f"{len(bad)} of {len(pages)} angles differ from cv2: {bad[:3]}"


def test_pixel_values_resize_paths(pp, synth, image_hashes):
    for seed in (5, 6):
        ent = image_hashes[f"seed{seed}"]
        arr = synth.page(seed, ent["w"], ent["h"])
        pv, (gh, gw) = pp.pixel_values(pp.to_device(arr), dtype=torch.float32)
        assert [1, gh, gw] == ent["grid:original"][0]
        assert sha(pv.cpu().numpy()) == ent["pv:original"]
        pvb, _ = pp.pixel_values(pp.to_device(arr), dtype=torch.bfloat16)
        assert torch.equal(pvb, pv.to(torch.bfloat16))


def test_batch_equals_single_and_oracle(pp, synth):
    pages = [synth.page(40 + i, 384, 288) for i in range(5)]
    x = pp.to_device(pages)
    for fn, ofn in [(pp.high_contrast, lambda a: R.clahe(R.rgb2gray(a))),
                    (pp.binarize, lambda a: R.adaptive_threshold(R.rgb2gray(a))),
                    (pp.sharpen, R.sharpen), (pp.deskew, R.deskew)]:
        out = fn(x).cpu().numpy()
        for i, p in enumerate(pages):
            assert np.array_equal(out[i], ofn(p)), (fn.__name__, i)


def test_vs_cv2_direct_odd_sizes(pp, synth):
    cv2 = pytest.importorskip("cv2")
    for seed, (w, h) in [(50, (517, 389)), (51, (130, 71)), (52, (1000, 37))]:
        arr = synth.page(seed, w, h)
        g = cv2.cvtColor(arr, cv2.COLOR_RGB2GRAY)
        x = pp.to_device(arr)
        assert np.array_equal(pp.to_gray(x)[0].cpu().numpy(), g)
        assert np.array_equal(pp.high_contrast(x)[0].cpu().numpy(),
                              cv2.createCLAHE(clipLimit=3.0, tileGridSize=(8, 8)).apply(g))
        assert np.array_equal(pp.binarize(x)[0].cpu().numpy(),
                              cv2.adaptiveThreshold(g, 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, cv2.THRESH_BINARY, 21, 10))
        k = np.array([[0, -1, 0], [-1, 5, -1], [0, -1, 0]], np.float32)
        assert np.array_equal(pp.sharpen(x)[0].cpu().numpy(), cv2.filter2D(arr, -1, k))
        M = cv2.getRotationMatrix2D((w // 2, h // 2), 1.75, 1.0)
        assert np.array_equal(pp.warp_affine(x, torch.from_numpy(M.reshape(1, 6)))[0].cpu().numpy(),
                              cv2.warpAffine(arr, M, (w, h), flags=cv2.INTER_CUBIC, borderMode=cv2.BORDER_REPLICATE))


def test_unknown_transforms(pp, synth, capsys):
    x = pp.to_device(synth.page(60, 128, 96))
    out = pp.apply_strategy(x, ["nonsense", "original"])
    assert out is x
    assert "Unknown transform 'nonsense'" in capsys.readouterr().out


def test_vectorised_paths_small_widths(pp):
    """The 4-bytes-per-thread sharpen / CLAHE kernels (row bytes % 4 == 0) at widths where most words touch a border,
    gray and RGB, against the oracle; the register-blocked threshold kernel on tiles cut by the image edge."""
    rng = np.random.default_rng(17)
    for (h, w) in [(5, 4), (3, 8), (7, 12), (9, 64), (33, 100), (6, 5), (4, 7)]:
        for shape in ((h, w), (h, w, 3)):
            a = rng.integers(0, 256, shape, dtype=np.uint8)
            assert np.array_equal(pp.sharpen(pp.to_device(a))[0].cpu().numpy(), R.sharpen(a)), shape
    for (h, w) in [(16, 16), (24, 40), (50, 68), (31, 36)]:
        a = rng.integers(0, 256, (h, w), dtype=np.uint8)
        assert np.array_equal(pp.high_contrast(pp.to_device(a))[0].cpu().numpy(), R.clahe(a)), (h, w)
    for (h, w) in [(33, 72), (70, 130), (8, 8), (40, 64)]:
        a = rng.integers(0, 256, (h, w), dtype=np.uint8)
        assert np.array_equal(pp.binarize(pp.to_device(a))[0].cpu().numpy(), R.adaptive_threshold(a)), (h, w)


def test_fused_transforms_full_pages_batched(pp, synth):
    """Round-2 throughput kernels at page size (RGB -> gray fused into the CLAHE histogram / threshold staging, cell-wise
    CLAHE apply, packed-lane sharpen, hull tree, dp2a warp): a batch of different pages equals the pages one by one
    (whose results the golden hashes pin), and the one-call transforms equal the two-pass C-ABI route
    (ocrb_rgb2gray_u8 -> ocrb_clahe_u8 / ocrb_adaptive_gauss_thresh_u8)."""
    from handwritten_ocr_b200 import _lib
    for (w, h) in [(1024, 768), (768, 1024)]:
        pages = [synth.page(700 + i, w, h) for i in range(3)]
        x = pp.to_device(pages)
        g = pp.to_gray(x)
        n = x.shape[0]
        for name in ("high_contrast", "binarize", "sharpen", "deskew"):
            fn = getattr(pp, name)
            out = fn(x)
            for i in range(n):
                assert torch.equal(out[i], fn(x[i:i + 1].contiguous())[0]), (name, w, h, i)
            if name in ("high_contrast", "binarize"):
                assert torch.equal(out, fn(g)), (name, "gray input", w, h)
        two = torch.empty_like(g)
        lut = torch.empty((n, 64, 256), dtype=torch.uint8, device=x.device)
        _lib.call("ocrb_clahe_u8", _lib.ptr(g), _lib.ptr(two), n, h, w, _lib.ptr(lut), _lib.stream_ptr())
        assert torch.equal(two, pp.high_contrast(x))
        _lib.call("ocrb_adaptive_gauss_thresh_u8", _lib.ptr(g), _lib.ptr(two), n, h, w, _lib.stream_ptr())
        assert torch.equal(two, pp.binarize(x))
        # oracle on one page of the batch (numpy, a few seconds)
        assert np.array_equal(pp.high_contrast(x)[2].cpu().numpy(), R.clahe(R.rgb2gray(pages[2])))
        assert np.array_equal(pp.sharpen(x)[1].cpu().numpy(), R.sharpen(pages[1]))
        assert np.array_equal(pp.deskew(x)[1].cpu().numpy(), R.deskew(pages[1]))
