"""Oracle for the DEE annotation post-process (TEST INFRASTRUCTURE).

Restates, vectorised, the pure-Python loops of
``packnet_code/packnet_sfm/utils/tools.py``:

* ``non_max_suppression``  tools.py:9-46
* ``hysteresis`` + ``DFS`` tools.py:49-92
* the normals block of ``infer_edge_estimation.py:193-199`` (= :244-250)

and the third-party ``cv2.Sobel(img, CV_64F, dx, dy, ksize=5)`` they call
(separable [1,4,6,4,1] x [-1,-2,0,2,1], BORDER_REFLECT_101, row pass first,
accumulated in fp64 in tap order).

Pinned against goldens from the reference ``tools.py`` loops
(tests/golden/dee_*.npz) and against cv2.Sobel directly.
"""
from __future__ import annotations

import numpy as np
from scipy import ndimage

SMOOTH5 = np.array([1.0, 4.0, 6.0, 4.0, 1.0])
DERIV5 = np.array([-1.0, -2.0, 0.0, 2.0, 1.0])


def _reflect101(i: np.ndarray, n: int) -> np.ndarray:
    if n == 1:
        return np.zeros_like(i)
    period = 2 * (n - 1)
    i = np.mod(i, period)
    return np.where(i >= n, period - i, i)


def sobel5(img: np.ndarray):
    """(sobel_x, sobel_y) as cv2.Sobel(img, CV_64F, 1,0 / 0,1, ksize=5)."""
    src = np.asarray(img)
    H, W = src.shape
    cols = _reflect101(np.arange(-2, W + 2), W)
    rows = _reflect101(np.arange(-2, H + 2), H)
    s64 = src.astype(np.float64)

    def row_pass(k):
        acc = k[0] * s64[:, cols[0:W]]
        for t in range(1, 5):
            acc = acc + k[t] * s64[:, cols[t:t + W]]
        return acc

    def col_pass(buf, k, symmetric):
        r = lambda t: buf[rows[t:t + H], :]
        if symmetric:
            acc = k[2] * r(2)
            acc = acc + k[3] * (r(3) + r(1))
            acc = acc + k[4] * (r(4) + r(0))
        else:
            acc = k[3] * (r(3) - r(1))
            acc = acc + k[4] * (r(4) - r(0))
        return acc

    sx = col_pass(row_pass(DERIV5), SMOOTH5, True)
    sy = col_pass(row_pass(SMOOTH5), DERIV5, False)
    return sx, sy


def normals_u8(prob: np.ndarray) -> np.ndarray:
    """infer_edge_estimation.py:193-199: quantised atan2(-sy, sx)."""
    sx, sy = sobel5(prob)
    ang = np.arctan2(-sy, sx)
    return (((ang * (180 / np.pi) + 180) / 360) * 255).astype("uint8")


def nms_bins(sx: np.ndarray, sy: np.ndarray) -> np.ndarray:
    """0:E/W 1:NW/SE 2:S/N 3:SW/NE 4:none (NaN) -- tools.py:12-38."""
    a = np.rad2deg(np.arctan2(sy, sx))
    a[a < 0] += 180
    b = np.full(a.shape, 4, dtype=np.int8)
    b[((0 <= a) & (a < 22.5)) | ((157.5 <= a) & (a <= 180))] = 0
    b[(22.5 <= a) & (a < 67.5)] = 1
    b[(67.5 <= a) & (a < 112.5)] = 2
    b[(112.5 <= a) & (a < 157.5)] = 3
    return b


def non_max_suppression(img: np.ndarray) -> np.ndarray:
    sx, sy = sobel5(img)
    b = nms_bins(sx, sy)
    H, W = img.shape
    out = np.zeros((H, W))
    if H < 3 or W < 3:
        return out
    c = img[1:-1, 1:-1]
    sh = lambda di, dj: img[1 + di:H - 1 + di, 1 + dj:W - 1 + dj]
    one = np.ones_like(c)
    q = np.select([b[1:-1, 1:-1] == k for k in range(4)],
                  [sh(0, 1), sh(-1, -1), sh(1, 0), sh(1, -1)], one)
    r = np.select([b[1:-1, 1:-1] == k for k in range(4)],
                  [sh(0, -1), sh(1, 1), sh(-1, 0), sh(-1, 1)], one)
    keep = (c >= q) & (c >= r)
    out[1:-1, 1:-1] = np.where(keep, c, 0)
    return out


def hysteresis(img: np.ndarray, t_low=0.3, t_high=0.7) -> np.ndarray:
    img = np.asarray(img)
    H, W = img.shape
    temp = np.copy(img)
    if H > 2 and W > 2:
        c = img[1:-1, 1:-1]
        strong = c > t_high
        weak = ~strong & ~(c < t_low)
        lab, n = ndimage.label(strong | weak, structure=np.ones((3, 3), dtype=bool))
        keep_lab = np.zeros(n + 1, dtype=bool)
        keep_lab[np.unique(lab[strong])] = True
        keep_lab[0] = False
        temp[1:-1, 1:-1] = np.where(keep_lab[lab], 2, 0)
    with np.errstate(invalid="ignore", divide="ignore"):
        temp = temp / np.max(temp)
        return img * temp
