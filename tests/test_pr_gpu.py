"""PR-count parity: the GPU matcher / counts against goldens produced by the reference
``eval_depth_edges.py`` (run with the oracle stand-in for py-bsds500) and against the C oracle."""
import os
import tempfile

import cv2
import numpy as np
import pytest
import torch

from conftest import GOLDEN
from synth import random_boundary_maps, scene_with_gt

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,max_dist", [((60, 90), 0.0075), ((218, 1153), 0.002), ((40, 40), 0.05),
                                            ((100, 300), 0.0075), ((218, 1153), 0.0075), ((30, 200), 0.002),
                                            ((1, 1), 0.5), ((3, 50), 0.1)])
def test_matcher_vs_oracle(shape, max_dist):
    from mindtheedge_b200.eval_depth_edges import correspond_pixels_batch
    from oracle import pr_counts as opr
    h, w = shape
    preds, gts = zip(*[random_boundary_maps(h, w, 10 * h + k) for k in range(5)])
    a = torch.from_numpy(np.stack(preds)).cuda()
    b = torch.from_numpy(np.stack(gts)).cuda()
    ma, mb, cnt = correspond_pixels_batch(a, b, max_dist)
    ma, mb, cnt = ma.cpu().numpy(), mb.cpu().numpy(), cnt.cpu().numpy()
    for k in range(5):
        ref = opr.match_count(preds[k], gts[k], max_dist)
        assert cnt[k] == ref, (k, cnt[k], ref)
        assert ma[k].sum() == ref and mb[k].sum() == ref
        assert not (ma[k] & ~(preds[k] != 0)).any() and not (mb[k] & ~(gts[k] != 0)).any()


def test_matcher_worst_cases():
    """Dense blobs (every pixel set) and a long two-chain component (long augmenting paths)."""
    from mindtheedge_b200.eval_depth_edges import correspond_pixels_batch
    from oracle import pr_counts as opr
    h, w = 64, 256
    a = np.zeros((3, h, w), np.uint8)
    b = np.zeros((3, h, w), np.uint8)
    a[0, 10:40, 20:120] = 1
    b[0, 12:45, 30:100] = 1
    a[1, 30, 5:250] = 1          # one pred chain, GT chain shifted so the greedy start is sub-optimal
    b[1, 31, 4:249] = 1
    b[1, 29, 100:140] = 1
    r = np.random.default_rng(0)
    a[2] = r.random((h, w)) < 0.3
    b[2] = r.random((h, w)) < 0.3
    _, _, cnt = correspond_pixels_batch(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), 0.008)
    for k in range(3):
        assert cnt[k].item() == opr.match_count(a[k], b[k], 0.008) == opr.match_count_scipy(a[k], b[k], 0.008)


def test_golden_evaluate_boundaries():
    from mindtheedge_b200.eval_depth_edges import evaluate_boundaries
    z = np.load(os.path.join(GOLDEN, "pr.npz"))
    for thin_flag in (0, 1):
        c_r, s_r, c_p, s_p, thr = evaluate_boundaries(z["soft"], [z["soft_gt"].astype(np.float64)], thresholds=9,
                                                      max_dist=0.0075, apply_thinning=bool(thin_flag))
        got = np.stack([c_r, s_r, c_p, s_p], 1).astype(np.int64)
        assert np.array_equal(got, z[f"soft_counts_thin{thin_flag}"]), thin_flag
        assert np.array_equal(thr, z["soft_thr"])


def test_golden_pr_evaluation_files():
    """The file-level drop-in against precision/recall/AUC produced by the reference pr_evaluation."""
    from mindtheedge_b200.eval_depth_edges import mean_recall_at_precision_range, pr_evaluation
    z = np.load(os.path.join(GOLDEN, "pr.npz"))
    H, W = z["pr_shape"]
    with tempfile.TemporaryDirectory() as tmp:
        gts, preds = [], []
        for i in range(3):
            g = np.unpackbits(z[f"pr_gt{i}"])[: H * W].reshape(H, W)
            gp = os.path.join(tmp, f"gt{i}.png")
            cv2.imwrite(gp, g.astype(np.uint8) * 255)
            dp = os.path.join(tmp, f"pred{i}.npy")
            np.save(dp, (z[f"pr_depth_u16_{i}"] / 256).astype(np.float32))
            gts.append(gp)
            preds.append(dp)
        pv, rv = pr_evaluation(gts, preds, edge_thresh_range=[int(v) for v in z["pr_range"]],
                               gt_crop=[int(v) for v in z["pr_crop"]], save_folder=os.path.join(tmp, "o"))
    assert np.array_equal(np.array(pv), z["pr_precision"])
    assert np.array_equal(np.array(rv), z["pr_recall"])
    pr = np.vstack((pv, rv)).transpose()
    assert mean_recall_at_precision_range(pr) == z["pr_auc_full"]
    assert mean_recall_at_precision_range(pr, 0.12, 0.65) == z["pr_auc_part"]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_soft_map_99_threshold_sweep_vs_oracle(dtype):
    """evaluate_boundaries' default threshold grid (99 ascending thresholds on a strength map): the incremental
    sweep matcher must give the count of every threshold exactly as 99 independent matchings do."""
    from mindtheedge_b200.eval_depth_edges import pr_counts
    from oracle import pr_counts as opr
    h, w = 120, 300
    r = np.random.default_rng(7)
    thr = np.linspace(0.01, 0.99, 99)
    preds, gts = [], []
    for k in range(3):
        pb, gb = random_boundary_maps(h, w, 900 + k)
        soft = (cv2.GaussianBlur((pb != 0).astype(np.float32), (5, 5), 1.2) * 2.5).clip(0, 1)
        soft = (soft * (0.4 + 0.6 * r.random((h, w)))).astype(dtype)
        soft[5, 7] = thr[40]          # values sitting exactly on thresholds
        soft[9, 11] = np.nan
        preds.append(soft)
        gts.append((gb != 0).astype(np.uint8))
    c = pr_counts(torch.from_numpy(np.stack(preds)).cuda(), torch.from_numpy(np.stack(gts)).cuda(), thr,
                  max_dist=0.0075).cpu().numpy()
    ref = np.zeros((99, 4), np.int64)
    for k in range(3):
        for t in range(99):
            b = (preds[k].astype(np.float64) >= thr[t])
            m = opr.match_count(b.astype(np.uint8), gts[k], 0.0075)
            ref[t] += (m, int(gts[k].sum()), m, int(b.sum()))
    assert np.array_equal(c, ref)


@pytest.mark.parametrize("pred_density,gt_density", [(0.06, 0.015), (0.02, 0.08), (0.07, 0.04)])
def test_sweep_matcher_capacity_paths_vs_oracle(pred_density, gt_density):
    """KITTI crop window with more boundary pixels than the sweep matcher keeps in shared memory (~9.3 k per side):
    many predicted pixels -> its compact mode (only matched pixels + a chunk resident); many GT pixels -> the image's
    problems overflow to the per-problem kernels.  Counts of every threshold must still equal independent matchings."""
    from mindtheedge_b200.eval_depth_edges import pr_counts
    from oracle import pr_counts as opr
    H, W, crop = 384, 1280, [44, 1197, 153, 371]
    r = np.random.default_rng(int(1000 * pred_density + 100000 * gt_density))
    gt = np.zeros((H, W), np.uint8)
    for _ in range(int(gt_density * H * W / 150)):          # GT as random polylines (contours), not salt noise
        y, x = int(r.integers(0, H)), int(r.integers(0, W))
        for _ in range(150):
            gt[y, x] = 1
            y = int(np.clip(y + r.integers(-1, 2), 0, H - 1)); x = int(np.clip(x + r.integers(0, 2), 0, W - 1))
    near = cv2.dilate(gt, np.ones((5, 5), np.uint8)) > 0
    strength = np.where(near, r.random((H, W)), 0.0) * (r.random((H, W)) < pred_density / max(near.mean(), 1e-6) * 0.7)
    strength = np.maximum(strength, (r.random((H, W)) < pred_density * 0.3) * r.random((H, W))).astype(np.float32)
    thr = np.array([0.15, 0.3, 0.45, 0.6, 0.75, 0.9])
    win = (slice(crop[2], crop[3]), slice(crop[0], crop[1]))
    n_pred, n_gt = int((strength[win] >= thr[0]).sum()), int(gt[win].sum())
    assert max(n_pred, n_gt) > 9500, (n_pred, n_gt)           # the case really exceeds the resident capacity
    c = pr_counts(torch.from_numpy(strength)[None].cuda(), torch.from_numpy(gt)[None].cuda(), thr, max_dist=0.002,
                  crop=crop).cpu().numpy()
    ref = np.zeros((len(thr), 4), np.int64)
    for t, v in enumerate(thr):
        b = (strength[win].astype(np.float64) >= v).astype(np.uint8)
        m = opr.match_count(b, np.ascontiguousarray(gt[win]), 0.002)
        ref[t] = (m, n_gt, m, int(b.sum()))
    assert np.array_equal(c, ref), (c.tolist(), ref.tolist())


def test_kitti_size_sweep_vs_oracle():
    """Config 2 shape: 384x1280 planes, 12 Canny settings, KITTI crop, max_dist 0.002 -- bit-exact counts."""
    from mindtheedge_b200.eval_depth_edges import pr_evaluation_arrays
    from oracle import pr_counts as opr
    gts, depths = zip(*[scene_with_gt(384, 1280, 100 + k) for k in range(4)])
    pv, rv, counts = pr_evaluation_arrays(depths, gts)
    ref = opr.pr_sweep_counts(depths, gts)
    assert np.array_equal(counts.cpu().numpy(), ref)
    assert ref[:, 0].min() > 1000  # the matcher has real work


def test_bench_kitti_set_counts_vs_oracle():
    """The workload bench.py times (BASELINE.json config 2): the reference's bundled KITTI-DE GT maps against
    bench.kitti_like_set's synthetic predicted depth, shipped wiring (12 Canny settings, KITTI crop, max_dist 0.002).
    26 of the 102 images incl. the two heaviest (#19, #29): counts bit-exact against the oracle."""
    import bench
    from mindtheedge_b200.eval_depth_edges import sweep_counts
    from oracle import pr_counts as opr
    depths, gts = bench.kitti_like_set(102, 7000)
    idx = sorted(set([19, 29] + list(range(0, 102, 4))))
    assert len(idx) >= 24
    rng = list(range(20, 241, 20))
    d = torch.from_numpy(depths[idx]).cuda()
    g = torch.from_numpy(gts[idx]).cuda()
    c = sweep_counts(d, g, rng, bench.KITTI_CROP, 0.0, 80.0, max_dist=0.002).cpu().numpy()
    ref = opr.pr_sweep_counts([depths[i] for i in idx], [gts[i] * 255 for i in idx], rng, tuple(bench.KITTI_CROP))
    assert np.array_equal(c, ref), (c.tolist(), ref.tolist())
    assert ref[:, 0].min() > 5000


def test_ddad_size_uncropped_vs_oracle():
    """Config 5 shape (1216x1936, NO crop, matching radius 4.57 px = 69 offsets, ~26 k GT + ~27 k predicted pixels
    per window -- more than one SM's shared memory holds): counts bit-exact against the C oracle for all 12
    settings, plus the size-independent properties."""
    from mindtheedge_b200.eval_depth_edges import correspond_pixels_batch, sweep_counts
    from oracle import pr_counts as opr
    gt, depth = scene_with_gt(1216, 1936, 5, n_rect=120)
    d = torch.from_numpy(depth)[None].cuda()
    g = torch.from_numpy((gt > 127).astype(np.uint8))[None].cuda()
    rng = list(range(20, 241, 20))
    c = sweep_counts(d, g, rng, None, 0.0, 80.0, max_dist=0.002).cpu().numpy()
    ref = opr.pr_sweep_counts([depth], [gt], rng, gt_crop=None)
    assert np.array_equal(c, ref), (c.tolist(), ref.tolist())
    assert (c[:, 0] == c[:, 2]).all() and (c[:, 0] <= c[:, 1]).all() and (c[:, 0] <= c[:, 3]).all()
    assert (np.diff(c[:, 3]) <= 0).all() and (np.diff(c[:, 0]) <= 0).all()  # higher threshold -> fewer edges
    _, _, cnt = correspond_pixels_batch(g, g, 0.002, want_maps=False)
    assert cnt.item() == int((gt > 127).sum())


def _mask_fixture(tmp_path):
    import cv2
    z = np.load(os.path.join(GOLDEN, "pr_mask.npz"))
    H, W = (int(v) for v in z["shape"])
    mp = str(tmp_path / "mask.png")
    cv2.imwrite(mp, z["mask"])
    gts, preds = [], []
    for i in range(2):
        g = np.unpackbits(z[f"gt{i}"])[:H * W].reshape(H, W)
        gp, dp = str(tmp_path / f"gt{i}.png"), str(tmp_path / f"pred{i}.npy")
        cv2.imwrite(gp, g.astype(np.uint8) * 255)
        np.save(dp, (z[f"depth_u16_{i}"] / 256).astype(np.float32))
        gts.append(gp)
        preds.append(dp)
    return z, mp, gts, preds


def test_pr_evaluation_mask_image_golden(tmp_path):
    """The mask-image branch of _pred_eval (eval_depth_edges.py:182-186, 198-200, 209-210) through pr_evaluation:
    precision / recall equal to the unmodified reference run (fractional sum_r, full-plane matching radius)."""
    from mindtheedge_b200.eval_depth_edges import pr_evaluation
    z, mp, gts, preds = _mask_fixture(tmp_path)
    pv, rv = pr_evaluation(gts, preds, edge_thresh_range=[int(v) for v in z["range"]], gt_crop=mp)
    assert np.array_equal(np.array(pv), z["precision"]), (pv, z["precision"])
    assert np.array_equal(np.array(rv), z["recall"]), (rv, z["recall"])


def test_pred_eval_golden(tmp_path):
    """_pred_eval on one predicted edge image: mask image, list crop and empty crop -- EvalResult fields equal to the
    reference's."""
    import cv2
    from mindtheedge_b200.eval_depth_edges import _pred_eval
    z, mp, gts, _ = _mask_fixture(tmp_path)
    pe = str(tmp_path / "pred_edge0.png")
    cv2.imwrite(pe, z["pred_edge0"])   # lossless copy of the JPEG-decoded plane the reference read
    for tag, crop in (("mask", mp), ("crop", str([10, 250, 8, 112])), ("nocrop", "[]")):
        r = _pred_eval(pe, gts[0], crop)
        got = np.array([r.count_r_overall[0], r.sum_r_overall[0], r.count_p_overall[0], r.sum_p_overall[0],
                        r.recall[0], r.precision[0]])
        assert np.array_equal(got, z[f"pe_{tag}"]), (tag, got, z[f"pe_{tag}"])
        assert r.count_r_best == r.count_r_overall[0] and r.used_thresholds[0] == 0.5


def test_sum_r_is_gt_sum_not_count():
    """evaluate_boundaries: sum_r is gt.sum() (eval_depth_edges.py:138) on both code paths, also for GT maps that are
    not 0/1 (255-valued PNG planes, soft maps)."""
    from mindtheedge_b200.eval_depth_edges import evaluate_boundaries
    pred, gt = random_boundary_maps(60, 90, 7)
    g255 = gt.astype(np.float64) * 255
    for thin_flag in (False, True):
        c_r, s_r, c_p, s_p, _ = evaluate_boundaries(pred.astype(np.float64), [g255], thresholds=1,
                                                    apply_thinning=thin_flag)
        assert s_r[0] == g255.sum()


def test_thin_vs_oracle():
    from mindtheedge_b200.bsds import thin
    from oracle import thin as othin
    r = np.random.default_rng(1)
    for shape, dens in [((50, 80), 0.4), ((218, 1153), 0.1), ((7, 7), 0.9), ((1, 5), 1.0), ((64, 64), 1.0)]:
        x = r.random(shape) < dens
        x[: shape[0] // 2, : shape[1] // 3] = True  # a thick blob needs several iterations
        assert np.array_equal(thin.binary_thin(x), othin.binary_thin(x)), shape
        assert np.array_equal(thin.binary_thin(x, max_iter=1), othin.binary_thin(x, max_iter=1))


def test_unmodified_reference_script_seam():
    """The bsds stand-in has the module layout and call signatures eval_depth_edges.py:7,50,125 expect."""
    from mindtheedge_b200 import bsds
    bsds.install_as_bsds_metric()
    from bsds_metric.bsds import correspond_pixels, thin  # noqa: F401  (the reference's import line)
    pred, gt = random_boundary_maps(80, 120, 3)
    m1, m2, cost, oc = correspond_pixels.correspond_pixels(pred.astype(bool), gt.astype(np.float64), max_dist=0.0075)
    assert m1.shape == pred.shape and (m1 > 0).sum() == (m2 > 0).sum() > 0
    assert thin.binary_thin(pred.astype(bool)).dtype == bool
