"""Differential pinning of the image oracle against the installed OpenCV wheel (4.13.0.92, the version the reference
locks: poetry.lock:3175) on random sizes and contents -- beyond the golden pages: every transform of tools.py:503-619
whose arithmetic the oracle restates.  Skipped where cv2 is not importable.  CPU only."""
import numpy as np
import pytest

from oracle import image_ref as R

cv2 = pytest.importorskip("cv2")


def _contents(rng, shape, k):
    x = rng.integers(0, 256, shape, dtype=np.uint8)
    if k == 1:
        x = (x // 32 * 32 + 7).astype(np.uint8)                                   # flat plateaus: exact ties
    elif k == 2:
        x = np.clip(228 + rng.integers(-10, 11, shape), 0, 255).astype(np.uint8)  # paper
    elif k == 3 and shape[0] > 4 and shape[1] > 4:
        x = cv2.GaussianBlur(x, (5, 5), 0)
    elif k == 4:
        x[:] = rng.integers(0, 256)
    return x


def test_clahe_threshold_sharpen_warp_random_sizes():
    rng = np.random.default_rng(21)
    k3 = np.array([[0, -1, 0], [-1, 5, -1], [0, -1, 0]], np.float32)
    for it in range(60):
        H, W = int(rng.integers(8, 90)), int(rng.integers(8, 130))
        if it % 2:
            W = max(W // 8 * 8, 8)
        g = _contents(rng, (H, W), it % 5)
        rgb = _contents(rng, (H, W, 3), (it + 2) % 5)
        assert np.array_equal(R.clahe(g), cv2.createCLAHE(clipLimit=3.0, tileGridSize=(8, 8)).apply(g)), (it, H, W)
        if W % 8 == 0:      # cv2's last W % 8 columns take an unfused scalar tail: defined on W % 8 == 0
            assert np.array_equal(R.adaptive_threshold(g), cv2.adaptiveThreshold(
                g, 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, cv2.THRESH_BINARY, 21, 10)), (it, H, W)
        assert np.array_equal(R.sharpen(rgb), cv2.filter2D(rgb, -1, k3)), (it, H, W)
        assert np.array_equal(R.sharpen(g), cv2.filter2D(g, -1, k3)), (it, H, W)
        M = cv2.getRotationMatrix2D((W // 2, H // 2), float(rng.uniform(-3, 3)), 1.0)
        assert np.array_equal(R.warp_affine_cubic(rgb, M), cv2.warpAffine(
            rgb, M, (W, H), flags=cv2.INTER_CUBIC, borderMode=cv2.BORDER_REPLICATE)), (it, H, W)


def _cv_lines_mask(gray):
    hk = cv2.getStructuringElement(cv2.MORPH_RECT, (gray.shape[1] // 4, 1))
    m = cv2.morphologyEx(cv2.adaptiveThreshold(cv2.bitwise_not(gray), 255, cv2.ADAPTIVE_THRESH_MEAN_C,
                                               cv2.THRESH_BINARY, 15, -2), cv2.MORPH_OPEN, hk, iterations=1)
    return cv2.dilate(m, cv2.getStructuringElement(cv2.MORPH_RECT, (1, 3)))


def test_lines_mask_random(synth):
    rng = np.random.default_rng(3)
    for it in range(80):
        H, W = int(rng.integers(4, 120)), int(rng.integers(4, 200))
        k = it % 4
        if k == 0:
            g = rng.integers(0, 256, (H, W), dtype=np.uint8)
        elif k == 1:
            g = np.clip(228 + rng.integers(-10, 11, (H, W)), 0, 255).astype(np.uint8)
            for y in range(5, H - 2, 11):
                g[y:y + 2, int(rng.integers(0, W // 3 + 1)):W - int(rng.integers(0, W // 3 + 1))] = 70
        elif k == 2:
            g = R.rgb2gray(synth.rule_lines(synth.page(it, max(W, 70), max(H, 110))))
        else:
            g = (rng.integers(0, 2, (H, W)) * 255).astype(np.uint8)
        assert np.array_equal(R.lines_mask(g), _cv_lines_mask(g)), (it, g.shape)


def test_denoise_random():
    rng = np.random.default_rng(11)
    for it in range(60):
        H, W = int(rng.integers(1, 45)), int(rng.integers(1, 55))
        col = it % 2 == 0
        x = _contents(rng, (H, W, 3) if col else (H, W), it % 5)
        want = cv2.fastNlMeansDenoisingColored(x, None, 10, 10, 7, 21) if col else cv2.fastNlMeansDenoising(x, None, 10, 7, 21)
        assert np.array_equal(R.denoise(x), want), (it, x.shape)
