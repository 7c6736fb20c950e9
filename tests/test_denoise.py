"""denoise (tools.py:576-589): cv2.fastNlMeansDenoising(10, 7, 21) / fastNlMeansDenoisingColored(10, 10, 7, 21).

CPU: the oracle against the golden outputs of the unmodified reference (tests/golden/make_golden_f3.py) and, where cv2 is
importable, against cv2 itself (Lab conversions on the whole 2^24 domain, the filter on random and tiny images); the
host-built integer tables of libocrb200 against the oracle's.  GPU: the kernels against the oracle, the golden pages and the
full-size golden hashes, bit-exact."""
import ctypes
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import image_ref as R

HERE = os.path.dirname(__file__)
GOLDEN = json.load(open(os.path.join(HERE, "golden", "denoise.json")))
SMALL = dict(np.load(os.path.join(HERE, "golden", "denoise_small.npz")))
CHAIN = ["deskew", "denoise", "high_contrast"]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_oracle_golden_small():
    for name in ("rgb_259x197", "gray_131x97"):
        page = SMALL[f"{name}/input"]
        assert sha(page) == GOLDEN[name]["input"]
        out = R.denoise(page)
        assert np.array_equal(out, SMALL[f"{name}/denoise"]) and sha(out) == GOLDEN[name]["denoise"]


def test_oracle_inputs_are_the_synthetic_pages(synth):
    for name, g in GOLDEN.items():
        assert sha(synth.page(g["seed"], g["w"], g["h"], gray=g["gray"])) == g["input"], name


def test_oracle_vs_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    # both 8-bit Lab conversions on every possible input triple
    v = np.arange(1 << 24, dtype=np.uint32)
    allrgb = np.stack([v & 255, (v >> 8) & 255, v >> 16], -1).astype(np.uint8).reshape(4096, 4096, 3)
    assert np.array_equal(R.lbgr2lab_u8(allrgb), cv2.cvtColor(allrgb, cv2.COLOR_LBGR2Lab))
    assert np.array_equal(R.lab2lbgr_u8(allrgb), cv2.cvtColor(allrgb, cv2.COLOR_Lab2LBGR))
    for shape in ((37, 53), (9, 5), (1, 40), (40, 1), (2, 2)):
        x = rng.integers(0, 256, shape, dtype=np.uint8)
        assert np.array_equal(R.nlm_denoise(x), cv2.fastNlMeansDenoising(x, None, 10, 7, 21)), shape
    for shape in ((31, 44, 3), (6, 7, 3)):
        x = rng.integers(0, 256, shape, dtype=np.uint8)
        assert np.array_equal(R.denoise(x), cv2.fastNlMeansDenoisingColored(x, None, 10, 10, 7, 21)), shape
    # smooth + noise: weights that are not all-or-nothing
    yy, xx = np.mgrid[0:48, 0:64]
    x = np.clip(120 + 60 * np.sin(xx / 9.0) * np.cos(yy / 7.0) + rng.normal(0, 6, (48, 64)), 0, 255).astype(np.uint8)
    assert np.array_equal(R.nlm_denoise(x), cv2.fastNlMeansDenoising(x, None, 10, 7, 21))


def test_library_tables_match_oracle(pkg):
    """The integer tables the CUDA path uploads are built on the host by libocrb200: equal to the oracle's."""
    from handwritten_ocr_b200 import _lib
    L = _lib.load()
    cb, yf, cf = np.zeros(3072, np.int32), np.zeros(512, np.int32), np.zeros(18, np.int32)
    w1, w2 = np.zeros(2048, np.int32), np.zeros(2048, np.int32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)   # noqa: E731
    assert L.ocrb_denoise_tables_host(p(cb), p(yf), p(cf), p(w1), p(w2)) == 0
    assert np.array_equal(cb, R.lab_cbrt_tab()) and np.array_equal(yf.reshape(256, 2), R.lab_inv_tabs())
    assert np.array_equal(cf[:9], R.lab_fwd_coef()) and np.array_equal(cf[9:], R.lab_inv_coef())
    for cn, w in ((1, w1), (2, w2)):
        wt, shift, fpm = R.nlm_weights(cn)
        assert shift == 6 and fpm == 19096 and np.array_equal(w, wt[:2048]) and not wt[2047:].any()


@pytest.mark.gpu
def test_gpu_small_and_edges(pkg, synth):
    from handwritten_ocr_b200 import preprocess as pp
    for name in ("rgb_259x197", "gray_131x97"):
        out = pp.denoise(pp.to_device(SMALL[f"{name}/input"]))[0].cpu().numpy()
        assert np.array_equal(out, SMALL[f"{name}/denoise"]), name
    rng = np.random.default_rng(11)
    # ragged sizes around the 26 x 32 tile, images smaller than the 13-pixel border (multiple reflections), batches
    for shape in ((1, 33, 27), (2, 32, 26), (1, 5, 9), (1, 1, 40), (1, 40, 1), (3, 64, 53), (1, 2, 2)):
        x = rng.integers(0, 256, shape, dtype=np.uint8)
        out = pp.denoise(pp.to_device(list(x))).cpu().numpy()
        for i in range(shape[0]):
            assert np.array_equal(out[i], R.nlm_denoise(x[i])), (shape, i)
    for shape in ((2, 31, 44, 3), (1, 6, 7, 3), (1, 45, 27, 3)):
        x = rng.integers(0, 256, shape, dtype=np.uint8)
        out = pp.denoise(pp.to_device(list(x))).cpu().numpy()
        for i in range(shape[0]):
            assert np.array_equal(out[i], R.denoise(x[i])), (shape, i)
    # smooth + noise (partial weights) and a constant page (all weights maximal: the accumulators' upper bound)
    yy, xx = np.mgrid[0:70, 0:90]
    x = np.clip(120 + 60 * np.sin(xx / 9.0) * np.cos(yy / 7.0) + rng.normal(0, 6, (70, 90)), 0, 255).astype(np.uint8)
    assert np.array_equal(pp.denoise(pp.to_device(x))[0].cpu().numpy(), R.nlm_denoise(x))
    c = np.full((40, 60), 255, np.uint8)
    assert np.array_equal(pp.denoise(pp.to_device(c))[0].cpu().numpy(), c)


@pytest.mark.gpu
def test_gpu_lab_roundtrip_full_domain(pkg):
    """Colored denoise of a constant-colour page = Lab -> (NLM of a constant = identity) -> back: every one of the 2^24
    colours, 4096 one-pixel... pages would be slow, so the colours are laid out as 64 x 64 constant blocks."""
    from handwritten_ocr_b200 import preprocess as pp
    rng = np.random.default_rng(3)
    cols = rng.integers(0, 256, (12, 16, 3), dtype=np.uint8)
    page = np.repeat(np.repeat(cols, 64, 0), 64, 1)                 # 768 x 1024, blocks wider than the 21 + 7 window
    out = pp.denoise(pp.to_device(page))[0].cpu().numpy()
    want = R.lab2lbgr_u8(R.lbgr2lab_u8(cols))
    assert np.array_equal(out[32::64, 32::64], want)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GOLDEN))
def test_gpu_golden_hashes(pkg, synth, name):
    from handwritten_ocr_b200 import preprocess as pp
    g = GOLDEN[name]
    x = pp.to_device(synth.page(g["seed"], g["w"], g["h"], gray=g["gray"]))
    assert sha(pp.denoise(x)[0].cpu().numpy()) == g["denoise"]
    assert sha(pp.apply_strategy(x, CHAIN)[0].cpu().numpy()) == g["+".join(CHAIN)]
