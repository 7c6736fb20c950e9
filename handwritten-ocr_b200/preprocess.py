"""Preprocessing strategies of the read path on the GPU (reference: ocr_agent/tools.py:496-673).

Pages live in HBM as uint8 tensors [n, H, W, 3] (RGB) or [n, H, W] (gray, PIL mode "L"); every
transform is one or two hand-written kernels called through the C ABI and is bit-exact against the
reference's OpenCV result.  A whole batch of same-sized pages goes through each kernel in one launch.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

# name -> implemented on the GPU?  (tools.py:623-630 registers these six)
TRANSFORMS = ("high_contrast", "binarize", "sharpen", "deskew", "denoise", "remove_lines")


def _check(x: torch.Tensor):
    if not (x.is_cuda and x.dtype == torch.uint8 and x.is_contiguous()):
        raise ValueError("expected a contiguous CUDA uint8 tensor [n,H,W,3] or [n,H,W]")
    if x.dim() == 4 and x.shape[-1] != 3:
        raise ValueError("only RGB (3 channels) or gray pages are supported (reference: tools.py:656 keeps the file's mode)")
    if x.dim() not in (3, 4):
        raise ValueError("expected [n,H,W,3] or [n,H,W]")


def to_gray(x: torch.Tensor) -> torch.Tensor:
    """cv2.cvtColor(RGB2GRAY) when 3-D, identity when already gray (tools.py:510)."""
    _check(x)
    if x.dim() == 3:
        return x
    n, H, W, _ = x.shape
    out = torch.empty((n, H, W), dtype=torch.uint8, device=x.device)
    _lib.call("ocrb_rgb2gray_u8", _lib.ptr(x), _lib.ptr(out), n, H, W, _lib.stream_ptr())
    return out


def high_contrast(x: torch.Tensor) -> torch.Tensor:
    """tools._apply_high_contrast (tools.py:503-516): gray + CLAHE(3.0, 8x8) -> "L".  One C-ABI call; for RGB pages the
    gray conversion rides on the tile-histogram pass."""
    _check(x)
    n, H, W = x.shape[:3]
    C = 3 if x.dim() == 4 else 1
    out = torch.empty((n, H, W), dtype=torch.uint8, device=x.device)
    gray = torch.empty_like(out) if C == 3 else None
    lut = torch.empty((n, 64, 256), dtype=torch.uint8, device=x.device)
    _lib.call("ocrb_high_contrast_u8", _lib.ptr(x), _lib.ptr(out), _lib.ptr(gray), _lib.ptr(lut), n, H, W, C,
              _lib.stream_ptr())
    return out


def binarize(x: torch.Tensor) -> torch.Tensor:
    """tools._apply_binarize (tools.py:519-531): gray + adaptive Gaussian threshold 21/10 -> "L".  One kernel; RGB pages
    become gray while their tiles are staged."""
    _check(x)
    n, H, W = x.shape[:3]
    C = 3 if x.dim() == 4 else 1
    out = torch.empty((n, H, W), dtype=torch.uint8, device=x.device)
    _lib.call("ocrb_binarize_u8", _lib.ptr(x), _lib.ptr(out), n, H, W, C, _lib.stream_ptr())
    return out


def sharpen(x: torch.Tensor) -> torch.Tensor:
    """tools._apply_sharpen (tools.py:534-546): 3x3 [0,-1,0;-1,5,-1;0,-1,0] on every channel."""
    _check(x)
    n, H, W = x.shape[:3]
    C = 3 if x.dim() == 4 else 1
    out = torch.empty_like(x)
    _lib.call("ocrb_sharpen3x3_u8", _lib.ptr(x), _lib.ptr(out), n, H, W, C, _lib.stream_ptr())
    return out


def deskew_angle(x: torch.Tensor):
    """Rotation angle (degrees, NaN = leave unchanged) and forward matrix per page (tools.py:556-567)."""
    _check(x)
    n, H, W = x.shape[:3]
    C = 3 if x.dim() == 4 else 1
    dev = x.device
    angle = torch.empty(n, dtype=torch.float64, device=dev)
    M = torch.empty((n, 6), dtype=torch.float64, device=dev)
    ext = torch.empty((n, H, 3), dtype=torch.int32, device=dev)
    hull = torch.empty((n, (4 * H + 8) * 2), dtype=torch.int32, device=dev)
    _lib.call("ocrb_deskew_angle", _lib.ptr(x), n, H, W, C, _lib.ptr(angle), _lib.ptr(M), _lib.ptr(ext),
              _lib.ptr(hull), _lib.stream_ptr())
    return angle, M


def warp_affine(x: torch.Tensor, M: torch.Tensor) -> torch.Tensor:
    """cv2.warpAffine(INTER_CUBIC, BORDER_REPLICATE) with per-page forward matrices M [n,6] float64."""
    _check(x)
    n, H, W = x.shape[:3]
    C = 3 if x.dim() == 4 else 1
    M = M.to(device=x.device, dtype=torch.float64).contiguous()
    out = torch.empty_like(x)
    _lib.call("ocrb_warp_affine_cubic_u8", _lib.ptr(x), _lib.ptr(out), n, H, W, C, _lib.ptr(M), _lib.stream_ptr())
    return out


def deskew(x: torch.Tensor, M: torch.Tensor | None = None) -> torch.Tensor:
    """tools._apply_deskew (tools.py:549-573); same mode/shape as the input."""
    if M is None:
        _, M = deskew_angle(x)
    return warp_affine(x, M)


def denoise(x: torch.Tensor) -> torch.Tensor:
    """tools._apply_denoise (tools.py:576-589): fastNlMeansDenoising(10, 7, 21) on gray pages,
    fastNlMeansDenoisingColored(10, 10, 7, 21) on RGB pages; same mode/shape as the input, bit-exact."""
    _check(x)
    n, H, W = x.shape[:3]
    C = 3 if x.dim() == 4 else 1
    out = torch.empty_like(x)
    ws = torch.empty((n * H * W * 6,), dtype=torch.uint8, device=x.device) if C == 3 else None
    _lib.call("ocrb_nlm_denoise_u8", _lib.ptr(x), _lib.ptr(out), _lib.ptr(ws), n, H, W, C, _lib.stream_ptr())
    return out


def remove_lines_mask(x: torch.Tensor):
    """Ruled-line mask of tools._apply_remove_lines (tools.py:598-614), bit-exact: (mask uint8 [n,H,W], nonzero int32 [n])."""
    _check(x)
    n, H, W = x.shape[:3]
    C = 3 if x.dim() == 4 else 1
    mask = torch.empty((n, H, W), dtype=torch.uint8, device=x.device)
    tmp = torch.empty_like(mask)
    nz = torch.empty(n, dtype=torch.int32, device=x.device)
    _lib.call("ocrb_remove_lines_mask_u8", _lib.ptr(x), _lib.ptr(mask), _lib.ptr(nz), _lib.ptr(tmp), n, H, W, C,
              _lib.stream_ptr())
    return mask, nz


def inpaint_telea(x: torch.Tensor, mask: torch.Tensor, radius: int = 3) -> torch.Tensor:
    """cv2.inpaint(page, mask, radius, cv2.INPAINT_TELEA) per page (tools.py:617), bit-exact; mask uint8 [n,H,W]."""
    _check(x)
    n, H, W = x.shape[:3]
    C = 3 if x.dim() == 4 else 1
    if not (mask.is_cuda and mask.dtype == torch.uint8 and mask.is_contiguous() and tuple(mask.shape) == (n, H, W)):
        raise ValueError("mask must be a contiguous CUDA uint8 tensor [n,H,W]")
    out = torch.empty_like(x)
    nbytes = _lib.load().ocrb_inpaint_workspace_bytes(n, H, W)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=x.device)
    _lib.call("ocrb_inpaint_telea_u8", _lib.ptr(x), _lib.ptr(mask), _lib.ptr(out), _lib.ptr(ws), n, H, W, C, radius,
              _lib.stream_ptr())
    return out


def remove_lines(x: torch.Tensor) -> torch.Tensor:
    """tools._apply_remove_lines (tools.py:592-619): ruled-line mask, then Telea inpaint with radius 3.  A batch without
    any ruled line comes back as is (cv2.inpaint copies its input when the mask is empty)."""
    mask, nz = remove_lines_mask(x)
    if not bool(nz.any()):
        return x
    return inpaint_telea(x, mask, 3)


def apply_transform(x: torch.Tensor, name: str) -> torch.Tensor:
    if name == "high_contrast":
        return high_contrast(x)
    if name == "binarize":
        return binarize(x)
    if name == "sharpen":
        return sharpen(x)
    if name == "deskew":
        return deskew(x)
    if name == "remove_lines":
        return remove_lines(x)
    if name == "denoise":
        return denoise(x)
    raise KeyError(name)


def apply_strategy(x: torch.Tensor, strategy) -> torch.Tensor:
    """Chain of transforms, left to right (tools.py:645-665). Unknown names are skipped with the
    reference's message; "original" is a no-op."""
    steps = [strategy] if isinstance(strategy, str) else list(strategy)
    for step in steps:
        if step == "original":
            continue
        if step not in TRANSFORMS:
            print(f"  [preprocess] Unknown transform '{step}', skipping")
            continue
        x = apply_transform(x, step)
    return x


# ───────────── HF image-processor stage (run_ocr side) ─────────────
def smart_resize(H: int, W: int, factor: int = 28, min_pixels: int = 256 * 256, max_pixels: int = 1024 * 1024):
    import ctypes
    oh, ow = ctypes.c_int32(), ctypes.c_int32()
    _lib.call("ocrb_smart_resize_host", H, W, factor, min_pixels, max_pixels, ctypes.byref(oh), ctypes.byref(ow))
    return oh.value, ow.value


def resize(x: torch.Tensor, out_h: int, out_w: int) -> torch.Tensor:
    """torchvision uint8 bicubic-antialias resize semantics (HF image_processing_backends.py:200-251)."""
    _check(x)
    n, H, W = x.shape[:3]
    C = 3 if x.dim() == 4 else 1
    shape = (n, out_h, out_w, 3) if C == 3 else (n, out_h, out_w)
    out = torch.empty(shape, dtype=torch.uint8, device=x.device)
    tmp = torch.empty((n * H * out_w * C,), dtype=torch.uint8, device=x.device)
    _lib.call("ocrb_resize_bicubic_aa_u8", _lib.ptr(x), _lib.ptr(out), _lib.ptr(tmp), n, H, W, C, out_h, out_w,
              _lib.stream_ptr())
    return out


def pixel_values(x: torch.Tensor, *, dtype=torch.bfloat16, group_perm: torch.Tensor | None = None,
                 min_pixels: int = 256 * 256, max_pixels: int = 1024 * 1024):
    """HF Qwen2VLImageProcessor on device pages: smart_resize -> resize -> normalize -> patchify.
    Gray pages are replicated to three channels (PIL convert("RGB"), HF image_utils.py:462-501).
    Returns (pixel_values [n*gh*gw, 1176], (gh, gw))."""
    _check(x)
    n, H, W = x.shape[:3]
    C = 3 if x.dim() == 4 else 1
    rh, rw = smart_resize(H, W, 28, min_pixels, max_pixels)
    r = x if (rh, rw) == (H, W) else resize(x, rh, rw)
    gh, gw = rh // 14, rw // 14
    out = torch.empty((n * gh * gw, 1176), dtype=dtype, device=x.device)
    code = {torch.float32: 0, torch.bfloat16: 1}[dtype]
    _lib.call("ocrb_normalize_patchify", _lib.ptr(r), _lib.ptr(out), n, rh, rw, C, _lib.ptr(group_perm), code,
              _lib.stream_ptr())
    return out, (gh, gw)


def to_device(pages) -> torch.Tensor:
    """numpy page(s) [H,W,3]/[H,W] or a list of same-sized pages -> CUDA uint8 batch."""
    if isinstance(pages, np.ndarray):
        pages = [pages]
    arr = np.stack([np.ascontiguousarray(p) for p in pages], 0)
    return torch.from_numpy(arr).pin_memory().cuda(non_blocking=True)
