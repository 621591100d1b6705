"""CPU suite, part 1: the oracle against the golden vectors generated from the reference
(tests/golden/make_golden.py) and against the third-party calls the reference makes."""
import os

import cv2
import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_cases
from synth import random_boundary_maps, scene_with_gt

LOSS_CASES = load_cases("edge_loss.npz")


def _t(a):
    return None if a.size == 0 else torch.from_numpy(a)


@pytest.mark.parametrize("name", sorted(LOSS_CASES))
def test_edge_loss_oracle_vs_reference_golden(name):
    from oracle.edge_loss import edge_loss_np64, edge_loss_torch
    c = LOSS_CASES[name]
    is_grad, is_sigmoid, thresh, weight, p2n = c["attrs"]
    x = _t(c["depth"]).requires_grad_(True)
    loss, gmap = edge_loss_torch(x, _t(c["edge"]), _t(c["mask"]), bool(is_grad), bool(is_sigmoid), float(thresh),
                                 _t(c["normal"]), weight=float(weight), pos_to_neg=float(p2n))
    loss.backward()
    ref = float(c["loss"])
    if np.isnan(ref):
        assert np.isnan(loss.item())
        return
    assert abs(loss.item() - ref) <= 1e-6 * abs(ref)
    assert np.abs(gmap.numpy() - c["grad_map"]).max() <= 1e-5 * max(np.abs(c["grad_map"]).max(), 1e-30)
    assert np.abs(x.grad.numpy() - c["dgrad"]).max() <= 1e-5 * np.abs(c["dgrad"]).max()
    # the independent fp64 restatement with the analytic backward agrees with both
    l64, g64, dx64 = edge_loss_np64(c["depth"], c["edge"], None if c["mask"].size == 0 else c["mask"], bool(is_grad),
                                    bool(is_sigmoid), float(thresh), None if c["normal"].size == 0 else c["normal"],
                                    weight=float(weight), pos_to_neg=float(p2n))
    assert abs(l64 - ref) <= 2e-6 * abs(ref)
    assert np.abs(dx64 - c["dgrad"]).max() <= 5e-5 * np.abs(c["dgrad"]).max()


def test_direction_bands_at_the_fp32_limits():
    from oracle.edge_loss import DIR_H, DIR_LR, DIR_RL, DIR_V, direction_index
    b = [np.float32(k * np.pi / 8) for k in range(9)]
    up = lambda v: np.nextafter(np.float32(v), np.float32(10))
    dn = lambda v: np.nextafter(np.float32(v), np.float32(-10))
    th = np.array([0, b[1], dn(b[1]), b[3], dn(b[3]), b[5], dn(b[5]), b[7], dn(b[7]),
                   -b[1], dn(-b[1]), -b[3], dn(-b[3]), -b[5], dn(-b[5]), -b[7], dn(-b[7]), np.nan, 4.0, -4.0],
                  np.float32)
    want = [DIR_H, DIR_RL, DIR_H, DIR_V, DIR_RL, DIR_LR, DIR_V, DIR_H, DIR_LR,
            DIR_H, DIR_LR, DIR_LR, DIR_V, DIR_V, DIR_RL, DIR_RL, DIR_H, DIR_H, DIR_H, DIR_H]
    assert direction_index(th).tolist() == want


def test_canny_oracle_vs_golden_and_cv2():
    from oracle.canny import canny_birth_levels, canny_np, edges_from_depth_np, quantise_depth
    z = np.load(os.path.join(GOLDEN, "canny.npz"))
    for i in range(6):
        d = z[f"depth{i}"]
        for t in (20, 60, 120, 240):
            ref = np.unpackbits(z[f"edges{i}_t{t}"])[: d.size].reshape(d.shape) * 255
            assert np.array_equal(edges_from_depth_np(d, 0.0, 80.0, t // 2, t), ref), (i, t)
            assert np.array_equal(cv2.Canny(quantise_depth(d), t // 2, t), ref)
    _, d = scene_with_gt(96, 200, 1)
    q = quantise_depth(d)
    pairs = [(t // 2, t) for t in range(240, 19, -20)]
    lv = canny_birth_levels(q, pairs)
    for k, (lo, hi) in enumerate(pairs):
        assert np.array_equal((lv <= k) * 255, cv2.Canny(q, lo, hi))
        assert np.array_equal(canny_np(q, lo, hi), cv2.Canny(q, lo, hi))


def test_dee_oracle_vs_golden_and_cv2_sobel():
    from oracle import dee
    z = np.load(os.path.join(GOLDEN, "dee.npz"))
    for i in range(5):
        p = z[f"prob{i}"]
        sx, sy = dee.sobel5(p)
        assert np.array_equal(sx, cv2.Sobel(p, cv2.CV_64F, 1, 0, ksize=5))
        assert np.array_equal(sy, cv2.Sobel(p, cv2.CV_64F, 0, 1, ksize=5))
        nms = dee.non_max_suppression(p)
        assert np.array_equal(nms, z[f"nms{i}"])
        assert np.array_equal(dee.hysteresis(nms), z[f"hyst{i}"], equal_nan=True)
        assert np.array_equal(dee.hysteresis(p), z[f"hyst_raw{i}"], equal_nan=True)
        assert np.array_equal(dee.hysteresis(nms, 0.2, 0.5), z[f"hyst_custom{i}"], equal_nan=True)
        assert np.array_equal(dee.normals_u8(p), z[f"normals{i}"])


def test_sobel5_fused_taps():
    """Premise of the fused multiply-adds in dee_front_tma_kernel (csrc/dee.cu): a Sobel5 tap times an fp32 value is
    EXACT in fp64, and so is a power of two times any double that comes out of the row pass, so fma(tap, a, acc)
    rounds once, exactly like acc + (tap * a) with the exact product: the kernel stays bit-identical to cv2.Sobel."""
    from fractions import Fraction
    rng = np.random.default_rng(11)
    mant = rng.integers(1 << 23, 1 << 24, 4000).astype(np.float64)
    expo = rng.integers(-140, 120, 4000)
    vals = np.ldexp(mant, expo - 23).astype(np.float32)            # normals, subnormals, huge values
    vals = np.concatenate([vals, -vals, np.array([0.0, -0.0, 1.0, 3.0, np.finfo(np.float32).max], np.float32)])
    for tap in (-2.0, -1.0, 0.0, 1.0, 2.0, 4.0, 6.0):
        prod = np.float64(tap) * vals.astype(np.float64)
        for a, p in zip(vals[::7], prod[::7]):
            assert Fraction(float(p)) == Fraction(tap) * Fraction(float(a))
        assert np.array_equal(np.signbit(prod), np.signbit(np.float64(tap)) ^ np.signbit(vals))   # signed zeros too
    # column pass: doubles that are sums of such products, times 2 / 4
    acc = (vals[:-1].astype(np.float64) * 6.0 + vals[1:].astype(np.float64) * 4.0)
    for tap in (2.0, 4.0):
        for a in acc[::11]:
            assert Fraction(float(np.float64(tap) * a)) == Fraction(tap) * Fraction(float(a))


def test_matcher_oracle_vs_scipy():
    from oracle import pr_counts as opr
    for k, (shape, md) in enumerate([((60, 90), 0.0075), ((218, 1153), 0.002), ((40, 40), 0.05), ((100, 300), 0.01)]):
        pred, gt = random_boundary_maps(*shape, seed=k)
        c = opr.match_count(pred, gt, md)
        assert c == opr.match_count_scipy(pred, gt, md)
        m1, m2, _, _ = opr.correspond_pixels(pred, gt, md)
        assert (m1 > 0).sum() == c == (m2 > 0).sum()
    assert opr.match_count(np.zeros((5, 5)), np.ones((5, 5)), 0.1) == 0


def test_min_cost_assignment_with_outliers_has_maximum_cardinality():
    """Why the matcher is restated as a maximum-cardinality matching: py-bsds500's `correspond_pixels` solves
    a min-cost perfect assignment in which every pixel may instead go to an outlier node at cost `outlier_cost * max_dist`
    (100 x the largest real edge).  One more real pair replaces two outlier assignments by one edge that costs at most
    `max_dist`, so the optimum always has the maximum number of real pairs.  Checked numerically with a dense assignment
    solver on small random boundary maps: its number of real pairs equals the oracle's (and scipy's) matching size."""
    from scipy.optimize import linear_sum_assignment
    from oracle import pr_counts as opr
    for k, (shape, md) in enumerate([((24, 30), 0.05), ((30, 30), 0.08), ((16, 48), 0.04), ((20, 20), 0.15)]):
        pred, gt = random_boundary_maps(*shape, seed=40 + k)
        R = opr.match_radius(shape, md)
        P, Q = np.argwhere(pred != 0), np.argwhere(gt != 0)
        n1, n2 = len(P), len(Q)
        big, oc = 1e9, 100.0 * R
        d = np.sqrt(((P[:, None, :] - Q[None, :, :]) ** 2).sum(-1))
        cost = np.full((n1 + n2, n1 + n2), big)
        cost[:n1, :n2] = np.where(d <= R, d, big)            # real edges within the radius
        cost[:n1, n2:] = np.where(np.eye(n1) > 0, oc, big)   # predicted pixel i -> its outlier
        cost[n1:, :n2] = np.where(np.eye(n2) > 0, oc, big)   # GT pixel j <- its outlier
        cost[n1:, n2:] = 0.0                                 # outliers pair up for free
        r, c = linear_sum_assignment(cost)
        real = int(((r < n1) & (c < n2)).sum())
        assert cost[r, c].max() < big
        assert real == opr.match_count(pred, gt, md) == opr.match_count_scipy(pred, gt, md), (k, real)


def test_pr_oracle_vs_reference_golden():
    from oracle import pr_counts as opr
    z = np.load(os.path.join(GOLDEN, "pr.npz"))
    for flag in (0, 1):
        c, thr = opr.evaluate_boundaries(z["soft"], [z["soft_gt"].astype(np.float64)], 9, 0.0075, bool(flag))
        assert np.array_equal(c, z[f"soft_counts_thin{flag}"])
        assert np.array_equal(thr, z["soft_thr"])
    H, W = z["pr_shape"]
    gts = [np.unpackbits(z[f"pr_gt{i}"])[: H * W].reshape(H, W).astype(np.uint8) * 255 for i in range(3)]
    depths = [(z[f"pr_depth_u16_{i}"] / 256).astype(np.float32) for i in range(3)]
    c = opr.pr_sweep_counts(depths, gts, [int(v) for v in z["pr_range"]], tuple(int(v) for v in z["pr_crop"]))
    rec, prec, _ = opr.rec_prec_f1(c[:, 0], c[:, 1], c[:, 2], c[:, 3])
    assert np.array_equal(prec, z["pr_precision"]) and np.array_equal(rec, z["pr_recall"])
    pr = np.vstack((prec, rec)).transpose()
    assert opr.mean_recall_at_precision_range(pr) == z["pr_auc_full"]
    assert opr.mean_recall_at_precision_range(pr, 0.12, 0.65) == z["pr_auc_part"]


def test_thin_oracle_properties():
    from oracle import thin
    r = np.random.default_rng(0)
    x = r.random((60, 90)) < 0.5
    x[10:40, 20:60] = True
    t = thin.binary_thin(x)
    assert not (t & ~x).any()                       # only deletes
    assert np.array_equal(thin.binary_thin(t), t)   # idempotent
    from scipy import ndimage
    s8 = np.ones((3, 3), bool)
    assert ndimage.label(t, s8)[1] == ndimage.label(x, s8)[1]  # 8-connectivity preserved
    line = np.zeros((9, 30), bool)
    line[4, 3:27] = True
    assert np.array_equal(thin.binary_thin(line), line)  # a 1-px line is already thin
    lut1, lut2 = thin.build_luts()
    assert lut1.sum() == lut2.sum() and not lut1[0] and not lut1[255]


def _chamfer_cases():
    z = np.load(os.path.join(GOLDEN, "chamfer.npz"))
    for i in range(int(z["n"])):
        H, W = z[f"shape{i}"]
        p = np.unpackbits(z[f"pred{i}"])[: H * W].reshape(H, W).astype(np.uint8) * 255
        g = np.unpackbits(z[f"gt{i}"])[: H * W].reshape(H, W).astype(np.uint8) * 255
        yield i, z, p, g


def test_chamfer_oracle_vs_reference_golden():
    """oracle/chamfer.py against chamfer_distance of the unmodified packnet_sfm/utils/edge.py (both directions,
    incl. the empty-pred (nan) and empty-GT (scipy's virtual background sample) cases) and the 9 light metrics."""
    from oracle import chamfer as och
    for i, z, p, g in _chamfer_cases():
        for tag, (a, b) in (("pg", (p, g)), ("gp", (g, p))):
            c, pc, close, n = och.chamfer_distance(a, b)
            assert np.array_equal(c, z[f"{tag}{i}_cdist"], equal_nan=True), (i, tag)
            assert np.array_equal(pc, z[f"{tag}{i}_perc"], equal_nan=True), (i, tag)
            cond = z[f"{tag}{i}_cond"]
            assert n == int((cond >= 0).sum()) and close == int((cond == 1).sum())
    z = np.load(os.path.join(GOLDEN, "chamfer.npz"))
    H, W = z["metrics_shape"]
    d = (z["metrics_depth_u16"] / 256).astype(np.float32)
    g = np.unpackbits(z["metrics_gt"])[: H * W].reshape(H, W).astype(np.float64)
    assert np.array_equal(np.array(och.compute_edge_metrics(d, g, list(z["metrics_crop"]))), z["metrics_vals"])


ALT_CASES = load_cases("edge_loss_alt.npz")


@pytest.mark.parametrize("name", sorted(ALT_CASES))
def test_alt_loss_oracle_vs_reference_golden(name):
    """attention_loss / spatially_adaptive / +dice restatement against the unmodified GradLoss.forward."""
    from oracle.edge_loss import edge_loss_torch
    c = ALT_CASES[name]
    is_grad, is_sigmoid, thresh, weight = c["attrs"]
    t = lambda a: None if a.size == 0 else torch.from_numpy(a)
    x = torch.from_numpy(c["depth"]).requires_grad_(True)
    loss, gmap = edge_loss_torch(x, t(c["edge"]), t(c["mask"]), bool(is_grad), bool(is_sigmoid), float(thresh),
                                 t(c["normal"]), weight=float(weight), edge_loss_type=str(c["ltype"]))
    loss.backward()
    assert abs(loss.item() - float(c["loss"])) <= 1e-6 * abs(float(c["loss"]))
    assert np.array_equal(gmap.numpy(), c["grad_map"])
    assert np.abs(x.grad.numpy() - c["dgrad"]).max() <= 1e-6 * np.abs(c["dgrad"]).max()


def test_targets_oracle_vs_reference_golden():
    """oracle/targets.py against resize_depth_preserve + the /255 rule of the unmodified datasets/augmentations.py and
    the normal decode expression of datasets/gta_dataset.py:413."""
    from oracle import targets as ot
    z = np.load(os.path.join(GOLDEN, "targets.npz"))
    for i in range(int(z["n_edge"])):
        assert np.array_equal(ot.edge_target(z[f"edge_in{i}"], tuple(z[f"edge_shape{i}"])), z[f"edge_out{i}"]), i
    assert np.array_equal(ot.decode_normals(np.arange(256, dtype=np.uint8)), z["theta"])
