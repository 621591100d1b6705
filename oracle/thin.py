"""Oracle for ``bsds_metric.bsds.thin.binary_thin`` (TEST INFRASTRUCTURE).

The reference calls it at eval_depth_edges.py:45 and :125; the implementation
lives in py-bsds500 (github.com/Britefury/py-bsds500, no version pin, NOT
vendored, absent from this image).  PARITY UNPINNED: restated from its
published algorithm -- MATLAB ``bwmorph(x,'thin',Inf)`` (Lam, Lee & Suen 1992):
two alternating sub-iterations, each deleting the pixels whose 3x3
neighbourhood satisfies G1 & G2 & G3 (first) or G1 & G2 & G3' (second), the
image zero-padded, until a sub-iteration deletes nothing.

Neighbour numbering x1..x8 starts east and runs counter-clockwise:
    x4 x3 x2
    x5  p x1
    x6 x7 x8
"""
from __future__ import annotations

import numpy as np

# (dy, dx) of x1..x8
NEIGH = [(0, 1), (-1, 1), (-1, 0), (-1, -1), (0, -1), (1, -1), (1, 0), (1, 1)]


def neighbour_code(x: np.ndarray) -> np.ndarray:
    """bit k-1 set iff neighbour x_k is set (zero padding)."""
    p = np.pad(x.astype(np.int32), 1)
    H, W = x.shape
    code = np.zeros((H, W), np.int32)
    for k, (dy, dx) in enumerate(NEIGH):
        code |= p[1 + dy:1 + dy + H, 1 + dx:1 + dx + W] << k
    return code


def build_luts():
    """256-entry deletion tables for the two sub-iterations."""
    lut1 = np.zeros(256, bool)
    lut2 = np.zeros(256, bool)
    for c in range(256):
        x = [0] + [(c >> k) & 1 for k in range(8)]  # x[1..8]
        x.append(x[1])                               # x9 = x1
        b = sum(1 for i in range(1, 5) if x[2 * i - 1] == 0 and (x[2 * i] or x[2 * i + 1]))
        n1 = sum(1 for k in range(1, 5) if x[2 * k - 1] or x[2 * k])
        n2 = sum(1 for k in range(1, 5) if x[2 * k] or x[2 * k + 1])
        g1 = b == 1
        g2 = 2 <= min(n1, n2) <= 3
        g3 = ((x[2] or x[3] or not x[8]) and x[1]) == 0
        g3p = ((x[6] or x[7] or not x[4]) and x[5]) == 0
        lut1[c] = g1 and g2 and g3
        lut2[c] = g1 and g2 and g3p
    return lut1, lut2


_LUT1, _LUT2 = build_luts()


def binary_thin(x: np.ndarray, max_iter=None) -> np.ndarray:
    x = np.asarray(x) != 0
    it = 0
    while max_iter is None or it < max_iter:
        kill = x & _LUT1[neighbour_code(x)]
        if not kill.any():
            break
        x = x & ~kill
        kill = x & _LUT2[neighbour_code(x)]
        if not kill.any():
            break
        x = x & ~kill
        it += 1
    return x
