"""Oracle for the precision/recall count sweep (TEST INFRASTRUCTURE).

Restates ``eval_depth_edges.py``:

* ``evaluate_boundaries``              :67-145  (threshold grid :102-104, counts :120-143)
* ``_pred_eval`` binarise + crop       :191-215
* ``pr_evaluation`` Canny sweep + sums :243-301, :344-345
* ``compute_rec_prec_f1``              :147-161
* ``mean_recall_at_precision_range``   :365-375

The matcher the reference imports (``bsds_metric.bsds.correspond_pixels``,
py-bsds500, not vendored, no version pin) is restated in oracle/csrc/match.c as
a maximum-cardinality matching within ``max_dist * diagonal``.
PARITY UNPINNED for the matcher; ``match_count_scipy`` is the independent
cross-check (scipy's Hopcroft-Karp on the explicit pair graph).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.oracle_correspond_pixels.restype = ctypes.c_int64
        _LIB.oracle_correspond_pixels.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double,
            ctypes.c_void_p, ctypes.c_void_p]
    return _LIB


def match_radius(shape, max_dist: float) -> float:
    h, w = shape
    return max_dist * np.sqrt(h * h + w * w)


def correspond_pixels(bmap1, bmap2, max_dist=0.0075, outlier_cost=100):
    """Signature of the py-bsds500 function -> (match1, match2, cost, oc).
    Only ``match > 0`` is meaningful (that is all the reference reads)."""
    a = np.ascontiguousarray(np.asarray(bmap1) != 0, dtype=np.uint8)
    b = np.ascontiguousarray(np.asarray(bmap2) != 0, dtype=np.uint8)
    h, w = a.shape
    m1 = np.zeros((h, w), np.uint8)
    m2 = np.zeros((h, w), np.uint8)
    _lib().oracle_correspond_pixels(a.ctypes.data, b.ctypes.data, h, w,
                                    float(match_radius((h, w), max_dist)),
                                    m1.ctypes.data, m2.ctypes.data)
    return m1.astype(np.float64), m2.astype(np.float64), 0.0, float(outlier_cost)


def match_count(bmap1, bmap2, max_dist) -> int:
    a = np.ascontiguousarray(np.asarray(bmap1) != 0, dtype=np.uint8)
    b = np.ascontiguousarray(np.asarray(bmap2) != 0, dtype=np.uint8)
    h, w = a.shape
    return int(_lib().oracle_correspond_pixels(a.ctypes.data, b.ctypes.data, h, w,
                                               float(match_radius((h, w), max_dist)), None, None))


def match_count_scipy(bmap1, bmap2, max_dist) -> int:
    """Independent cardinality check."""
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import maximum_bipartite_matching
    a = np.asarray(bmap1) != 0
    b = np.asarray(bmap2) != 0
    h, w = a.shape
    R = match_radius((h, w), max_dist)
    r = int(R)
    if a.sum() == 0 or b.sum() == 0:
        return 0
    ia = -np.ones((h, w), np.int64)
    ia[a] = np.arange(a.sum())
    ib = -np.ones((h, w), np.int64)
    ib[b] = np.arange(b.sum())
    rows, cols = [], []
    ya, xa = np.nonzero(a)
    for dy in range(-r, r + 1):
        for dx in range(-r, r + 1):
            if dy * dy + dx * dx > R * R:
                continue
            y, x = ya + dy, xa + dx
            ok = (y >= 0) & (y < h) & (x >= 0) & (x < w)
            j = ib[y[ok], x[ok]]
            sel = j >= 0
            rows.append(ia[ya[ok], xa[ok]][sel])
            cols.append(j[sel])
    rows = np.concatenate(rows)
    cols = np.concatenate(cols)
    if len(rows) == 0:
        return 0
    g = csr_matrix((np.ones(len(rows), np.int8), (rows, cols)), shape=(int(a.sum()), int(b.sum())))
    return int((maximum_bipartite_matching(g, perm_type="column") >= 0).sum())


def threshold_grid(thresholds):
    if isinstance(thresholds, (int, np.integer)):
        t = int(thresholds)
        return np.linspace(1.0 / (t + 1), 1.0 - 1.0 / (t + 1), t)
    return np.asarray(thresholds, dtype=np.float64)


def evaluate_boundaries(pred, gts, thresholds=99, max_dist=0.0075, apply_thinning=True):
    """-> int64[T,4] columns (count_r, sum_r, count_p, sum_p) and the thresholds."""
    from . import thin as _thin
    thr = threshold_grid(thresholds)
    out = np.zeros((len(thr), 4), dtype=np.int64)
    for i, t in enumerate(thr):
        b = pred >= t
        if apply_thinning:
            b = _thin.binary_thin(b)
        acc = np.zeros(b.shape, bool)
        for gt in gts:
            m1, m2, _, _ = correspond_pixels(b, gt, max_dist=max_dist)
            acc |= m1 > 0
            out[i, 1] += int(np.asarray(gt).sum())
            out[i, 0] += int((m2 > 0).sum())
        out[i, 3] = int(b.sum())
        out[i, 2] = int(acc.sum())
    return out, thr


def binarise_and_crop(img_u8, crop):
    """_pred_eval's treatment of an 8-bit edge image (eval_depth_edges.py:191-197)."""
    v = img_u8 / 255
    v[v > 0.5] = 1.0
    v[v < 0.5] = 0.0
    if crop is not None and len(crop) > 0:
        v = v[crop[2]:crop[3], crop[0]:crop[1]]
    return v


def pr_sweep_counts(depths, gts_u8, thresh_range=None, gt_crop=(44, 1197, 153, 371),
                    min_depth=0.0, max_depth=80.0, max_dist=0.002):
    """Array-level pr_evaluation: for every Canny setting t -> (t//2, t) the
    counts summed over images.  -> int64[len(range), 4]."""
    from .canny import edges_from_depth_np
    if thresh_range is None:
        thresh_range = list(range(20, 241, 20))
    out = np.zeros((len(thresh_range), 4), dtype=np.int64)
    for k, t in enumerate(thresh_range):
        for d, g in zip(depths, gts_u8):
            e = edges_from_depth_np(d, min_depth, max_depth, int(t / 2), int(t))
            p = binarise_and_crop(e.astype(np.float64), gt_crop)
            q = binarise_and_crop(g.astype(np.float64), gt_crop)
            c, _ = evaluate_boundaries(p, [q], thresholds=1, max_dist=max_dist, apply_thinning=False)
            out[k] += c[0]
    return out


def rec_prec_f1(count_r, sum_r, count_p, sum_p):
    count_r, sum_r, count_p, sum_p = (np.asarray(v, dtype=np.float64) for v in (count_r, sum_r, count_p, sum_p))
    rec = count_r / (sum_r + (sum_r == 0))
    prec = count_p / (sum_p + (sum_p == 0))
    f1 = 2.0 * prec * rec / (prec + rec + ((prec + rec) == 0))
    return rec, prec, f1


def mean_recall_at_precision_range(pr, small_lim=0.0, large_lim=1.0):
    pr = np.asarray(pr, dtype=np.float64)
    xs = np.array(range(int(small_lim * 100), int(large_lim * 100))) / 100
    ys = np.clip(np.interp(xs, pr[:, 0], pr[:, 1]), 0, 1)
    return float(np.mean(ys))
