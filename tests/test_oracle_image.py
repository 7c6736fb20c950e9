"""The numpy oracle (oracle/image_ref.py) pinned against golden outputs of the
unmodified reference (`ocr_agent.tools._apply_*`, tests/golden/make_golden.py) and,
where the wheel is importable, against cv2 / the HF image processor directly."""
import hashlib

import numpy as np
import pytest

from oracle import image_ref as R

CHAINS = ["deskew+high_contrast+binarize", "high_contrast+binarize", "deskew+high_contrast+sharpen"]
SMALL = ["rgb_256x192", "rgb_203x157", "gray_256x192", "rgb_blank_128x96", "rgb_ruled_320x240"]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def oracle_step(arr, step, angle="auto"):
    if step == "high_contrast":
        return R.clahe(R.rgb2gray(arr))
    if step == "binarize":
        return R.adaptive_threshold(R.rgb2gray(arr))
    if step == "sharpen":
        return R.sharpen(arr)
    if step == "deskew":
        return R.deskew(arr, angle)
    raise KeyError(step)


def oracle_chain(arr, chain, angle="auto"):
    for s in chain.split("+"):
        arr = oracle_step(arr, s, angle)
    return arr


@pytest.mark.parametrize("name", SMALL)
@pytest.mark.parametrize("step", ["high_contrast", "binarize", "sharpen"])
def test_small_transforms_bit_exact(image_small, name, step):
    out = oracle_step(image_small[f"{name}/input"], step)
    assert np.array_equal(out, image_small[f"{name}/{step}"])


@pytest.mark.parametrize("name", SMALL)
def test_small_deskew_warp_bit_exact_given_reference_angle(image_small, name):
    arr = image_small[f"{name}/input"]
    ang = float(image_small[f"{name}/angle"][0])
    out = R.deskew(arr, None if np.isnan(ang) else ang)
    assert np.array_equal(out, image_small[f"{name}/deskew"])


@pytest.mark.parametrize("name", SMALL)
def test_small_deskew_angle(image_small, name):
    arr = image_small[f"{name}/input"]
    ang = float(image_small[f"{name}/angle"][0])
    mine = R.deskew_angle(R.rgb2gray(arr))
    if np.isnan(ang):
        assert mine is None
    else:
        # angle parity: <= 2 fp32 ulp (last ulp of cv2.minAreaRect is not pinned)
        assert abs(np.float32(mine) - np.float32(ang)) <= 4 * np.spacing(np.float32(abs(ang) + 90.0))


@pytest.mark.parametrize("name", SMALL)
@pytest.mark.parametrize("chain", CHAINS)
def test_small_chains_given_reference_angle(image_small, name, chain):
    arr = image_small[f"{name}/input"]
    ang = float(image_small[f"{name}/angle"][0])
    out = oracle_chain(arr, chain, None if np.isnan(ang) else ang)
    assert np.array_equal(out, image_small[f"{name}/{chain}"])


@pytest.mark.parametrize("seed", [0, 1])
def test_full_page_hashes(synth, image_hashes, seed):
    ent = image_hashes[f"seed{seed}"]
    arr = synth.page(seed, ent["w"], ent["h"])
    assert sha(arr) == ent["input"], "synthetic page generator drifted"
    for step in ["high_contrast", "binarize", "sharpen"]:
        assert sha(oracle_step(arr, step)) == ent[step], step
    assert sha(R.deskew(arr, ent["angle"])) == ent["deskew"]
    ch = "high_contrast+binarize"
    out = oracle_chain(arr, ch)
    assert sha(out) == ent[ch]
    rgb = np.repeat(out[:, :, None], 3, axis=2)
    rh, rw = R.smart_resize(ent["h"], ent["w"])
    pv, grid = R.normalize_patchify(R.resize_bicubic_aa_u8(rgb, rh, rw))
    assert list(grid) == ent[f"grid:{ch}"][0]
    assert sha(pv.astype(np.float32)) == ent[f"pv:{ch}"]


def test_pixel_values_downscale_hash(synth, image_hashes):
    ent = image_hashes["seed6"]
    arr = synth.page(6, ent["w"], ent["h"])
    rh, rw = R.smart_resize(ent["h"], ent["w"])
    pv, grid = R.normalize_patchify(R.resize_bicubic_aa_u8(arr, rh, rw))
    assert list(grid) == ent["grid:original"][0]
    assert sha(pv.astype(np.float32)) == ent["pv:original"]


def test_oracle_vs_cv2_direct(synth):
    cv2 = pytest.importorskip("cv2")
    arr = synth.page(21, 333, 211)
    g = cv2.cvtColor(arr, cv2.COLOR_RGB2GRAY)
    assert np.array_equal(g, R.rgb2gray(arr))
    assert np.array_equal(cv2.createCLAHE(clipLimit=3.0, tileGridSize=(8, 8)).apply(g), R.clahe(g))
    assert np.array_equal(
        cv2.adaptiveThreshold(g, 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, cv2.THRESH_BINARY, 21, 10),
        R.adaptive_threshold(g))
    k = np.array([[0, -1, 0], [-1, 5, -1], [0, -1, 0]], np.float32)
    assert np.array_equal(cv2.filter2D(arr, -1, k), R.sharpen(arr))
    M = cv2.getRotationMatrix2D((333 // 2, 211 // 2), 2.37, 1.0)
    assert np.array_equal(M, R.rotation_matrix(333 // 2, 211 // 2, 2.37))
    assert np.array_equal(
        cv2.warpAffine(arr, M, (333, 211), flags=cv2.INTER_CUBIC, borderMode=cv2.BORDER_REPLICATE),
        R.warp_affine_cubic(arr, M))


def test_smart_resize_cases():
    assert R.smart_resize(768, 1024) == (756, 1036)
    assert R.smart_resize(1024, 768) == (1036, 756)
    assert R.smart_resize(100, 100) == (280, 280) or R.smart_resize(100, 100)[0] % 28 == 0
    h, w = R.smart_resize(3000, 4000)
    assert h * w <= 1024 * 1024 and h % 28 == 0 and w % 28 == 0


def test_pixel_values_vs_hf_random_sizes():
    """The image-processor oracle (smart_resize -> uint8 bicubic-antialias resize -> normalize -> patchify) against
    HF's Qwen2VLImageProcessor itself on random sizes: upscaled, downscaled, extreme aspect ratios, no resize."""
    from PIL import Image
    tr = pytest.importorskip("transformers")
    ip = tr.Qwen2VLImageProcessor(min_pixels=256 * 256, max_pixels=1024 * 1024)
    rng = np.random.default_rng(4)
    sizes = [(768, 1024), (56, 56), (761, 105), (117, 704), (1105, 911), (1133, 1041), (44, 1436), (1092, 532)]
    for it, (H, W) in enumerate(sizes):
        a = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        if it % 3 == 0:
            a = np.repeat(np.repeat(rng.integers(0, 256, (H // 8 + 1, W // 8 + 1, 3), dtype=np.uint8), 8, 0), 8, 1)[:H, :W]
        r = ip(images=[Image.fromarray(a)], return_tensors="np")
        rh, rw = R.smart_resize(H, W)
        x = a if (rh, rw) == (H, W) else R.resize_bicubic_aa_u8(a, rh, rw)
        pv, grid = R.normalize_patchify(x)
        assert list(grid) == [int(v) for v in r["image_grid_thw"][0]], (H, W)
        assert np.array_equal(pv.astype(np.float32), r["pixel_values"].astype(np.float32)), (H, W)
