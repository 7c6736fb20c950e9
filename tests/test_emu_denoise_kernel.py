"""The non-local-means kernels and the two 8-bit Lab conversions (handwritten-ocr_b200/csrc/denoise_kernels.cuh) on the CPU
through tests/emu/cuda_emu.h, launch sequence of ocrb_nlm_denoise_u8, against the oracle: the per-warp split of the 441
displacements, the shuffle-built 7-column patch sums, the packed own-column registers, tile edges and reflect-101."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import image_ref as R

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "emu")
SRC = [os.path.join(EMU, "emu_denoise.cpp"), os.path.join(EMU, "cuda_emu.h"),
       os.path.join(HERE, "..", "handwritten-ocr_b200", "csrc", "denoise_kernels.cuh")]


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(EMU, "_build", "libemu_denoise.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(s) for s in SRC):
        subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, SRC[0]],
                       check=True)
    return ctypes.CDLL(so)


def P(a):
    return ctypes.c_void_p(a.ctypes.data)


@pytest.mark.parametrize("shape", [(30, 28), (20, 40, 3)])
def test_nlm_denoise_emulated(emu, shape):
    """(30, 28): one tile row, two tile columns (26 + 2) of a gray page; (20, 40, 3): the colored route (Lab, L and (a, b)
    planes, back)."""
    rng = np.random.default_rng(len(shape))
    img = np.clip(rng.normal(200, 25, shape), 0, 255).astype(np.uint8)
    img[5:9, 4:20] = 40                                            # a stroke, so that weights differ across the page
    H, W = shape[:2]
    C = 3 if len(shape) == 3 else 1
    src = np.ascontiguousarray(img[None])
    dst = np.zeros_like(src)
    ws = np.zeros(H * W * 6 + 8, np.uint8)
    ws = ws[(-ws.ctypes.data) % 2:]
    assert emu.emu_nlm_denoise(P(src), P(dst), P(ws), 1, H, W, C) == 0
    assert np.array_equal(dst[0], R.denoise(img)), shape
