"""The HF image-processor kernels (handwritten-ocr_b200/csrc/resize_kernels.cuh: uint8 bicubic-antialias resize in two passes,
normalize + patchify) on the CPU through tests/emu/cuda_emu.h, launch sequence of the C ABI, against the oracle."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import image_ref as R

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "emu")
SRC = [os.path.join(EMU, "emu_resize.cpp"), os.path.join(EMU, "cuda_emu.h"),
       os.path.join(HERE, "..", "handwritten-ocr_b200", "csrc", "resize_kernels.cuh")]


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(EMU, "_build", "libemu_resize.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(s) for s in SRC):
        subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, SRC[0]],
                       check=True)
    return ctypes.CDLL(so)


def P(a):
    return ctypes.c_void_p(a.ctypes.data)


@pytest.mark.parametrize("shape,out", [((37, 53, 3), (28, 56)), ((30, 20), (56, 28)), ((40, 64, 3), (40, 28)),
                                        ((25, 28, 3), (56, 28))])
def test_resize_bicubic_aa(emu, shape, out):
    """Down- and up-scaling in both axes, one axis unchanged (pass skipped), gray and RGB."""
    rng = np.random.default_rng(sum(shape))
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    H, W = shape[:2]
    C = 3 if len(shape) == 3 else 1
    oh, ow = out
    src = np.ascontiguousarray(np.stack([img, img[::-1]]))
    dst = np.zeros((2, oh, ow) + ((3,) if C == 3 else ()), np.uint8)
    tmp = np.zeros(2 * H * ow * C, np.uint8)
    assert emu.emu_resize_bicubic_aa(P(src), P(dst), P(tmp), 2, H, W, C, oh, ow) == 0
    assert np.array_equal(dst[0], R.resize_bicubic_aa_u8(img, oh, ow)), (shape, out)
    assert np.array_equal(dst[1], R.resize_bicubic_aa_u8(np.ascontiguousarray(img[::-1]), oh, ow)), (shape, out)


def test_normalize_patchify(emu):
    rng = np.random.default_rng(9)
    for (H, W), C in [((28, 56), 3), ((56, 28), 1)]:
        shape = (H, W, 3) if C == 3 else (H, W)
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        src = np.ascontiguousarray(img[None])
        S = (H // 14) * (W // 14)
        dst = np.zeros((S, 1176), np.float32)
        assert emu.emu_normalize_patchify_f32(P(src), P(dst), 1, H, W, C, None) == 0
        rgb = img if C == 3 else np.stack([img] * 3, -1)           # PIL convert("RGB") of an "L" page
        want, grid = R.normalize_patchify(np.ascontiguousarray(rgb))
        assert grid == (1, H // 14, W // 14)
        assert np.array_equal(dst.view(np.uint32), want.view(np.uint32)), (H, W, C)
        # group permutation: output group k comes from source group perm[k]
        perm = rng.permutation(S // 4).astype(np.int32)
        dst2 = np.zeros_like(dst)
        assert emu.emu_normalize_patchify_f32(P(src), P(dst2), 1, H, W, C, P(perm)) == 0
        assert np.array_equal(dst2.reshape(S // 4, 4, 1176), want.reshape(S // 4, 4, 1176)[perm])
