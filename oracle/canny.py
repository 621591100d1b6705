"""Oracle for depth -> edge extraction (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates the body of ``edge_from_depth`` (``edge.py:73-93``, twin
``packnet_code/packnet_sfm/utils/edge.py:64-89``) and the algorithm of the
third-party call it makes, ``cv2.Canny(u8, t1, t2)`` (aperture 3, L1 gradient;
opencv-python, version unpinned by the reference, 4.13.0 in this image).

Pinned: ``canny_np`` is checked bit-for-bit against ``cv2.Canny`` itself in
tests/test_oracle_canny.py and against goldens from the reference
``edge_from_depth`` (tests/golden/canny_*.npz).
"""
from __future__ import annotations

import numpy as np
from scipy import ndimage

TG22 = int(0.4142135623730950488016887242097 * (1 << 15) + 0.5)  # 13573


def quantise_depth(depth: np.ndarray, min_depth: float = 0.0, max_depth: float = 80.0) -> np.ndarray:
    """edge.py:81-87: clamp, scale by 255/max_depth in the array's own float
    type, truncate to uint8."""
    d = np.array(depth, copy=True)
    d[d < min_depth] = min_depth
    d[d > max_depth] = max_depth
    factor = 255.0 / max_depth
    return (d * factor).astype(np.uint8)


def sobel3_replicate(img_u8: np.ndarray):
    p = np.pad(img_u8.astype(np.int32), 1, mode="edge")
    H, W = img_u8.shape
    s = lambda a, b: p[a:a + H, b:b + W]
    dx = (s(0, 2) + 2 * s(1, 2) + s(2, 2)) - (s(0, 0) + 2 * s(1, 0) + s(2, 0))
    dy = (s(2, 0) + 2 * s(2, 1) + s(2, 2)) - (s(0, 0) + 2 * s(0, 1) + s(0, 2))
    return dx, dy


def nms_magnitude(img_u8: np.ndarray) -> np.ndarray:
    """Threshold-independent part of Canny: L1 magnitude where the pixel
    survives non-maximum suppression, else 0 (int32 plane)."""
    dx, dy = sobel3_replicate(img_u8)
    mag = np.abs(dx) + np.abs(dy)
    mp = np.pad(mag, 1)  # magnitude outside the image is 0
    H, W = mag.shape
    nb = lambda a, b: mp[1 + a:1 + a + H, 1 + b:1 + b + W]
    x = np.abs(dx).astype(np.int64)
    y = np.abs(dy).astype(np.int64) << 15
    tg22x = x * TG22
    tg67x = tg22x + (x << 16)
    horiz = y < tg22x
    vert = y > tg67x
    diag = ~horiz & ~vert
    s = np.where((dx ^ dy) < 0, -1, 1)
    keep_h = (mag > nb(0, -1)) & (mag >= nb(0, 1))
    keep_v = (mag > nb(-1, 0)) & (mag >= nb(1, 0))
    d_a = np.where(s < 0, nb(-1, 1), nb(-1, -1))
    d_b = np.where(s < 0, nb(1, -1), nb(1, 1))
    keep_d = (mag > d_a) & (mag > d_b)
    keep = (horiz & keep_h) | (vert & keep_v) | (diag & keep_d)
    return np.where(keep, mag, 0).astype(np.int32)


def hysteresis_from_nms(nms: np.ndarray, low: int, high: int) -> np.ndarray:
    if low > high:
        low, high = high, low
    cand = nms > low
    strong = nms > high
    lab, n = ndimage.label(cand, structure=np.ones((3, 3), dtype=bool))
    if n == 0:
        return np.zeros(nms.shape, dtype=np.uint8)
    has_strong = np.zeros(n + 1, dtype=bool)
    has_strong[np.unique(lab[strong])] = True
    has_strong[0] = False
    return np.where(has_strong[lab], 255, 0).astype(np.uint8)


def canny_np(img_u8: np.ndarray, low: int, high: int) -> np.ndarray:
    return hysteresis_from_nms(nms_magnitude(img_u8), int(np.floor(low)), int(np.floor(high)))


def edges_from_depth_np(depth, min_depth=0.0, max_depth=80.0, thresh_1=20, thresh_2=40):
    """Array-level body of edge_from_depth (no file IO, no resize)."""
    return canny_np(quantise_depth(depth, min_depth, max_depth), thresh_1, thresh_2)


def canny_birth_levels(img_u8: np.ndarray, pairs) -> np.ndarray:
    """For nested threshold pairs (sorted strictest first) the index of the
    first pair at which each pixel is an edge, 255 = never.  Definition used by
    the fused multi-threshold kernel; built here from independent Canny runs."""
    nms = nms_magnitude(img_u8)
    out = np.full(nms.shape, 255, dtype=np.uint8)
    for k in range(len(pairs) - 1, -1, -1):
        lo, hi = pairs[k]
        e = hysteresis_from_nms(nms, int(lo), int(hi))
        out[e > 0] = k
    return out
