"""Seeded synthetic inputs shared by tests, smoke() and bench.py."""
import numpy as np


def scene_with_gt(H, W, seed, n_rect=40, noise=0.3, shift=2):
    """A piecewise-constant depth scene.  GT edges = its region boundaries; the "predicted" depth is
    the scene slightly shifted plus noise, so predicted Canny edges sit near (not on) the GT edges and
    the matcher has real work (config 2 of BASELINE.json with synthetic GT)."""
    r = np.random.default_rng(seed)
    d = np.full((H, W), 40.0, np.float32)
    for _ in range(n_rect):
        y0, x0 = r.integers(0, H), r.integers(0, W)
        h, w = r.integers(8, max(9, H // 2)), r.integers(8, max(9, W // 2))
        d[y0:y0 + h, x0:x0 + w] = r.uniform(3, 80)
    gt = np.zeros((H, W), bool)
    gt[:, 1:] |= d[:, 1:] != d[:, :-1]
    gt[1:, :] |= d[1:, :] != d[:-1, :]
    pred = np.roll(d, (int(r.integers(-shift, shift + 1)), int(r.integers(-shift, shift + 1))), (0, 1))
    pred = pred + r.normal(0, noise, (H, W)).astype(np.float32)
    return gt.astype(np.uint8) * 255, pred.astype(np.float32)


def random_boundary_maps(h, w, seed, density=0.01, jitter=3):
    r = np.random.default_rng(seed)
    gt = np.zeros((h, w), bool)
    for _ in range(max(2, h * w // 4000)):
        y, x = r.integers(0, h), r.integers(0, w)
        for _ in range(int(r.integers(20, 200))):
            gt[y, x] = 1
            y = int(np.clip(y + r.integers(-1, 2), 0, h - 1))
            x = int(np.clip(x + r.integers(0, 2), 0, w - 1))
    ys, xs = np.nonzero(gt)
    pred = np.zeros((h, w), bool)
    j = r.integers(-jitter, jitter + 1, size=(2, len(ys)))
    pred[np.clip(ys + j[0], 0, h - 1), np.clip(xs + j[1], 0, w - 1)] = 1
    pred |= r.random((h, w)) < density
    return pred.astype(np.uint8), gt.astype(np.uint8)


def prob_map(H, W, seed):
    """DEE-like probability map (config 4): blurred sigmoid noise plus a few ridges."""
    import cv2
    r = np.random.default_rng(seed)
    p = 1 / (1 + np.exp(-r.normal(-3, 2, (H, W))))
    p = cv2.blur(p, (7, 7))
    for _ in range(max(3, H // 40)):
        p[r.integers(0, H), :] += 0.6
        p[:, r.integers(0, W)] += 0.5
    return np.clip(p, 0, 1).astype(np.float32)
