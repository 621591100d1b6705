"""Generate the committed golden vectors by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Writes tests/golden/*.npz.  Every expected value below is produced by reference
code imported from /root/reference (GradLoss, tools.py, edge.py,
eval_depth_edges.py) or by the exact third-party call the reference makes
(cv2.Canny / cv2.Sobel); inputs are seeded synthetic tensors.
The reference's matcher/thinner (py-bsds500) is not available, so the
eval_depth_edges goldens use the oracle stand-in for `bsds_metric.bsds`; they pin
the arithmetic around the matcher (binarise, crop, threshold grid, sums, P/R),
not the matcher itself.
"""
import glob
import os
import sys
import tempfile
import warnings

import cv2
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
warnings.filterwarnings("ignore")

from _reference_loader import load_edge, load_eval_depth_edges, load_gradloss, load_tools  # noqa: E402


def synth_scene(H, W, seed, n_rect=24, noise=0.3):
    r = np.random.default_rng(seed)
    d = np.full((H, W), 40.0, np.float32)
    for _ in range(n_rect):
        y0, x0 = r.integers(0, H), r.integers(0, W)
        h, w = r.integers(1, max(2, H // 2)), r.integers(1, max(2, W // 2))
        d[y0:y0 + h, x0:x0 + w] = r.uniform(1, 80)
    d += r.normal(0, noise, (H, W)).astype(np.float32)
    return d


def loss_inputs(B, H, W, seed, h=None, w=None):
    g = torch.Generator().manual_seed(seed)
    h, w = h or H, w or W
    depth = torch.stack([torch.from_numpy(synth_scene(h, w, seed * 10 + b, noise=0.05)) for b in range(B)])[:, None]
    depth = depth * 0.2 + torch.rand(B, 1, h, w, generator=g) * 2.0
    u = torch.rand(B, 1, H, W, generator=g)
    edge = (u < 0.05).float() * torch.clamp(torch.rand(B, 1, H, W, generator=g), min=0.3)
    k = torch.randint(0, 256, (B, 1, H, W), generator=g).float()
    normal = (360 * k / 255 - 180) * np.pi / 180
    return depth.float(), edge, normal.float()


def gen_edge_loss():
    GradLoss = load_gradloss()
    cases = {}

    def run(name, depth, edge, mask, normal, is_grad=True, is_sigmoid=True, thresh=4, weight=10.0, p2n=1.0):
        head = GradLoss("cross_entropy", True, [], weight, p2n)
        x = depth.clone().requires_grad_(True)
        loss, gmap = head(x, edge, mask, is_grad, is_sigmoid, thresh, normal)
        loss.backward()
        cases[name] = dict(
            depth=depth.numpy(), edge=edge.numpy(),
            mask=np.zeros(0, np.float32) if mask is None else mask.numpy(),
            normal=np.zeros(0, np.float32) if normal is None else normal.numpy(),
            attrs=np.array([is_grad, is_sigmoid, thresh, weight, p2n], np.float64),
            loss=np.float32(loss.item()), grad_map=gmap.numpy(), dgrad=x.grad.numpy())

    g = torch.Generator().manual_seed(123)
    d, e, n = loss_inputs(2, 48, 64, 1)
    run("normals_nomask", d, e, None, n)
    run("normals_binmask", d, e, (torch.rand(2, 1, 48, 64, generator=g) < 0.5).float(), n)
    run("normals_softmask", d, e, torch.rand(2, 1, 48, 64, generator=g), n)
    run("normals_onesmask", d, e, torch.ones(2, 1, 48, 64), n)
    run("normals_zerosmask", d, e, torch.zeros(2, 1, 48, 64), n)
    run("magnitude_nomask", d, e, None, None)
    run("p2n_weight", d, e, None, n, weight=3.0, p2n=2.5, thresh=2)
    # continuous normals with the band limits injected exactly (fp32-rounded k*pi/8)
    nc = (torch.rand(2, 1, 48, 64, generator=g) * 2 - 1) * np.pi
    lim = torch.tensor([np.float32(s * k * np.pi / 8) for k in range(1, 9, 2) for s in (-1, 1)])
    nc.view(-1)[: 8 * 40] = lim.repeat(40)
    nc.view(-1)[400:408] = torch.nextafter(lim, torch.tensor(10.0))
    nc.view(-1)[408:416] = torch.nextafter(lim, torch.tensor(-10.0))
    nc.view(-1)[416] = float("nan")
    run("normals_continuous", d, e, None, nc)
    # odd sizes (scalar path), single image
    d2, e2, n2 = loss_inputs(1, 37, 53, 2)
    run("odd_shape", d2, e2, None, n2)
    d3, e3, n3 = loss_inputs(3, 24, 40, 3)
    run("three_images_binmask", d3, e3, (torch.rand(3, 1, 24, 40, generator=g) < 0.7).float(), n3)
    # resize paths: prediction at half / 1.5x resolution of the targets
    d4, e4, n4 = loss_inputs(2, 48, 64, 4, h=24, w=32)
    run("resize_up", d4, e4, None, n4)
    d5, e5, n5 = loss_inputs(2, 32, 48, 5, h=48, w=72)
    run("resize_down", d5, e5, None, n5)
    # DEE-training mode: input already a probability (EdgeEstimationLIDARModel.py:139-144)
    prob = torch.rand(2, 1, 48, 64, generator=g) * 0.98 + 0.01
    run("dee_mode", prob, e, None, None, is_grad=False, is_sigmoid=False)
    flat = {}
    for k, c in cases.items():
        for f, v in c.items():
            flat[f"{k}/{f}"] = v
    np.savez_compressed(os.path.join(HERE, "edge_loss.npz"), **flat)
    print("edge_loss:", list(cases))


def gen_edge_loss_alt():
    """attention_loss / spatially_adaptive / +dice through the UNMODIFIED GradLoss.forward (grad_loss.py:143-156)."""
    GradLoss = load_gradloss()
    cases = {}

    def run(name, ltype, depth, edge, mask, normal, is_grad=True, is_sigmoid=True, thresh=4, weight=10.0):
        head = GradLoss(ltype, True, [], weight, 1.0)
        x = depth.clone().requires_grad_(True)
        loss, gmap = head(x, edge, mask, is_grad, is_sigmoid, thresh, normal)
        loss.backward()
        cases[name] = dict(
            depth=depth.numpy(), edge=edge.numpy(), ltype=np.array(ltype),
            mask=np.zeros(0, np.float32) if mask is None else mask.numpy(),
            normal=np.zeros(0, np.float32) if normal is None else normal.numpy(),
            attrs=np.array([is_grad, is_sigmoid, thresh, weight], np.float64),
            loss=np.float32(loss.item()), grad_map=gmap.numpy(), dgrad=x.grad.numpy())

    g = torch.Generator().manual_seed(321)
    d, e, n = loss_inputs(2, 48, 64, 11)
    e = torch.where(torch.rand(e.shape, generator=g) < 0.3, (e > 0).float(), e)   # some labels exactly 1 (num_pos)
    m = (torch.rand(2, 1, 48, 64, generator=g) < 0.6).float()
    run("attention", "attention_loss", d, e, None, n)
    run("attention_mask", "attention_loss", d, e, m, n, weight=2.5, thresh=2)
    run("spatial", "spatially_adaptive", d, e, None, n)
    run("spatial_mask", "attention_loss_spatially_adaptive", d, e, m, n)
    run("ce_dice", "cross_entropy_dice", d, e, None, n)
    run("ce_dice_mask", "cross_entropy+dice", d, e, m, n, weight=3.0)
    run("attention_dice", "attention_loss_dice", d, e, None, n)
    prob = torch.rand(2, 1, 48, 64, generator=g) * 0.98 + 0.01
    run("dee_attention", "attention_loss", prob, e, None, None, is_grad=False, is_sigmoid=False)
    run("dee_spatial_dice", "spatially_adaptive_dice", prob, e, None, None, is_grad=False, is_sigmoid=False)
    flat = {}
    for k, c in cases.items():
        for f, v in c.items():
            flat[f"{k}/{f}"] = v
    np.savez_compressed(os.path.join(HERE, "edge_loss_alt.npz"), **flat)
    print("edge_loss_alt:", {k: float(c["loss"]) for k, c in cases.items()})


def gen_edge_loss_alt2():
    """More of GradLoss.forward's alternative types through the UNMODIFIED reference: without normals (GradLayer's
    magnitude path, grad_loss.py:70-73) and with a prediction smaller / larger than the targets (F.interpolate, :127)."""
    GradLoss = load_gradloss()
    cases = {}

    def run(name, ltype, depth, edge, mask, normal, weight=10.0, thresh=4):
        head = GradLoss(ltype, True, [], weight, 1.0)
        x = depth.clone().requires_grad_(True)
        loss, gmap = head(x, edge, mask, True, True, thresh, normal)
        loss.backward()
        cases[name] = dict(
            depth=depth.numpy(), edge=edge.numpy(), ltype=np.array(ltype),
            mask=np.zeros(0, np.float32) if mask is None else mask.numpy(),
            normal=np.zeros(0, np.float32) if normal is None else normal.numpy(),
            attrs=np.array([1, 1, thresh, weight], np.float64),
            loss=np.float32(loss.item()), grad_map=gmap.numpy(), dgrad=x.grad.numpy())

    g = torch.Generator().manual_seed(654)
    d, e, n = loss_inputs(2, 48, 64, 21)
    e = torch.where(torch.rand(e.shape, generator=g) < 0.3, (e > 0).float(), e)
    m = (torch.rand(2, 1, 48, 64, generator=g) < 0.6).float()
    run("attention_mag", "attention_loss", d, e, None, None)
    run("spatial_mag_mask", "spatially_adaptive", d, e, m, None, weight=2.5, thresh=2)
    run("ce_dice_mag", "cross_entropy_dice", d, e, None, None)
    dsmall, _, _ = loss_inputs(2, 24, 32, 22)
    dbig, _, _ = loss_inputs(2, 60, 100, 23)
    run("attention_up", "attention_loss", dsmall, e, None, n)
    run("spatial_dice_down", "spatially_adaptive_dice", dbig, e, None, n, weight=3.0)
    run("attention_mag_up", "attention_loss", dsmall, e, None, None)
    flat = {}
    for k, c in cases.items():
        for f, v in c.items():
            flat[f"{k}/{f}"] = v
    np.savez_compressed(os.path.join(HERE, "edge_loss_alt2.npz"), **flat)
    print("edge_loss_alt2:", {k: float(c["loss"]) for k, c in cases.items()})


def gen_canny():
    ref_edge = load_edge()
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for i, (H, W) in enumerate([(96, 160), (64, 200), (37, 53), (5, 7), (1, 9), (128, 128)]):
            d = synth_scene(H, W, 100 + i)
            if i == 5:
                d = np.random.default_rng(7).uniform(-5, 95, (H, W)).astype(np.float32)
            path = os.path.join(tmp, f"d{i}.npy")
            np.save(path, d)
            out[f"depth{i}"] = d
            for t in (20, 60, 120, 240):
                e = ref_edge.edge_from_depth(path, None, os.path.join(tmp, "e.jpeg"),
                                             thresh_1=int(t / 2), thresh_2=int(t), is_write_edge=False)
                out[f"edges{i}_t{t}"] = np.packbits(e > 0)
        # resize branch (cv2.resize INTER_LINEAR to the GT size) -- host-side in the drop-in
        d = synth_scene(48, 80, 200)
        path = os.path.join(tmp, "dr.npy")
        np.save(path, d)
        out["depth_resize"] = d
        e = ref_edge.edge_from_depth(path, (160, 96), os.path.join(tmp, "e.jpeg"), thresh_1=20, thresh_2=40,
                                     is_write_edge=False)
        out["edges_resize"] = np.packbits(e > 0)
    np.savez_compressed(os.path.join(HERE, "canny.npz"), **out)
    print("canny:", len(out))


def synth_prob(H, W, seed):
    r = np.random.default_rng(seed)
    p = 1 / (1 + np.exp(-r.normal(-3, 2, (H, W))))
    p = cv2.blur(p, (7, 7))
    for _ in range(5):
        p[r.integers(0, H), :] += 0.6
        p[:, r.integers(0, W)] += 0.5
    return np.clip(p, 0, 1).astype(np.float32)


def gen_dee():
    tools = load_tools()
    out = {}
    for i, (H, W) in enumerate([(40, 56), (64, 96), (7, 9), (3, 3), (33, 70)]):
        p = synth_prob(H, W, 300 + i)
        if i == 1:
            p *= np.random.default_rng(5).choice([1e-12, 1e-3, 1, 1], size=p.shape).astype(np.float32)
        out[f"prob{i}"] = p
        # infer_edge_estimation.py:244-250 (the normals block, verbatim calls)
        sx = cv2.Sobel(p, cv2.CV_64F, 1, 0, ksize=5)
        sy = cv2.Sobel(p, cv2.CV_64F, 0, 1, ksize=5)
        ang = np.arctan2(-sy, sx)
        out[f"normals{i}"] = (((ang * (180 / np.pi) + 180) / 360) * 255).astype("uint8")
        nms = tools.non_max_suppression(p)
        out[f"nms{i}"] = nms
        out[f"hyst{i}"] = tools.hysteresis(nms)
        out[f"hyst_raw{i}"] = tools.hysteresis(p)
        out[f"hyst_custom{i}"] = tools.hysteresis(nms, 0.2, 0.5)
    np.savez_compressed(os.path.join(HERE, "dee.npz"), **out)
    print("dee:", len(out))


def synth_gt_and_depth(H, W, seed):
    """GT = region boundaries of a piecewise-constant scene; predicted depth =
    that scene, slightly shifted and noisy (so pred edges sit near GT edges)."""
    r = np.random.default_rng(seed)
    d = np.full((H, W), 40.0, np.float32)
    for _ in range(14):
        y0, x0 = r.integers(0, H), r.integers(0, W)
        h, w = r.integers(8, H // 2), r.integers(8, W // 2)
        d[y0:y0 + h, x0:x0 + w] = r.uniform(3, 80)
    gt = np.zeros((H, W), bool)
    gt[:, 1:] |= d[:, 1:] != d[:, :-1]
    gt[1:, :] |= d[1:, :] != d[:-1, :]
    pred = np.roll(d, (int(r.integers(-2, 3)), int(r.integers(-2, 3))), (0, 1))
    pred = pred + r.normal(0, 0.4, (H, W)).astype(np.float32)
    # 1/256 m fixed point (the KITTI u16 depth format) so the fixture stores as uint16
    pred = np.round(np.clip(pred, 0, 200) * 256).astype(np.uint16)
    return gt, (pred / 256).astype(np.float32)


def gen_pr():
    from oracle import pr_counts as opr, thin as othin
    import types
    cp = types.ModuleType("correspond_pixels")
    cp.correspond_pixels = opr.correspond_pixels
    th = types.ModuleType("thin")
    th.binary_thin = othin.binary_thin
    ede = load_eval_depth_edges(th, cp)
    out = {}
    # (a) evaluate_boundaries on a soft map, 9 thresholds, thinning off/on
    gt, depth = synth_gt_and_depth(80, 120, 400)
    r = np.random.default_rng(1)
    soft = cv2.GaussianBlur(np.roll(gt, (1, -1), (0, 1)).astype(np.float64), (5, 5), 1.0)
    soft = np.clip(soft / soft.max() + r.uniform(0, 0.15, soft.shape), 0, 1)
    out["soft"] = soft
    out["soft_gt"] = gt
    for thin_flag in (False, True):
        c_r, s_r, c_p, s_p, thr = ede.evaluate_boundaries(soft, [gt.astype(np.float64)], thresholds=9,
                                                          max_dist=0.0075, apply_thinning=thin_flag)
        out[f"soft_counts_thin{int(thin_flag)}"] = np.stack([c_r, s_r, c_p, s_p], 1).astype(np.int64)
        out["soft_thr"] = thr
    # (b) pr_evaluation end to end (files, JPEG round trip, pool) on 3 small scenes
    with tempfile.TemporaryDirectory() as tmp:
        H, W = 200, 520
        crop = [10, 510, 20, 190]  # diagonal 528 px -> match radius 1.06 px
        gts, preds = [], []
        for i in range(3):
            g, d = synth_gt_and_depth(H, W, 500 + i)
            gp = os.path.join(tmp, f"gt{i}.png")
            cv2.imwrite(gp, g.astype(np.uint8) * 255)
            dp = os.path.join(tmp, f"pred{i}.npy")
            np.save(dp, d)
            gts.append(gp)
            preds.append(dp)
            out[f"pr_gt{i}"] = np.packbits(g)
            out["pr_shape"] = np.array([H, W])
            out[f"pr_depth_u16_{i}"] = np.round(d * 256).astype(np.uint16)
        rng = [40, 100, 160, 240]
        pv, rv = ede.pr_evaluation(gts, preds, edge_thresh_range=rng, gt_crop=crop,
                                   save_folder=os.path.join(tmp, "out"), num_workers=2)
        out["pr_range"] = np.array(rng)
        out["pr_crop"] = np.array(crop)
        out["pr_precision"] = np.array(pv, np.float64)
        out["pr_recall"] = np.array(rv, np.float64)
        pr = np.vstack((pv, rv)).transpose()
        out["pr_auc_full"] = np.float64(ede.mean_recall_at_precision_range(pr))
        out["pr_auc_part"] = np.float64(ede.mean_recall_at_precision_range(pr, 0.12, 0.65))
    np.savez_compressed(os.path.join(HERE, "pr.npz"), **out)
    print("pr:", {k: v for k, v in out.items() if k.startswith("pr_") and np.size(v) < 8})


def gen_pr_mask():
    """The mask-image branch of _pred_eval (eval_depth_edges.py:182-186, 198-200, 209-210) through the UNMODIFIED
    pr_evaluation / _pred_eval: gt_crop is the path of a mask PNG with values 0 / 127 (< 0.5) / 128 (>= 0.5) / 255."""
    from oracle import pr_counts as opr, thin as othin
    import types
    cp = types.ModuleType("correspond_pixels")
    cp.correspond_pixels = opr.correspond_pixels
    th = types.ModuleType("thin")
    th.binary_thin = othin.binary_thin
    ede = load_eval_depth_edges(th, cp)
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        H, W = 120, 260
        mask = np.full((H, W), 255, np.uint8)
        mask[:, :40] = 0
        mask[:30, :] = 127
        mask[60:90, 100:200] = 128
        mask[100:, 200:] = 64
        mp = os.path.join(tmp, "mask.png")
        cv2.imwrite(mp, mask)
        out["mask"] = mask
        gts, preds = [], []
        for i in range(2):
            g, d = synth_gt_and_depth(H, W, 700 + i)
            gp = os.path.join(tmp, f"gt{i}.png")
            cv2.imwrite(gp, g.astype(np.uint8) * 255)
            dp = os.path.join(tmp, f"pred{i}.npy")
            np.save(dp, d)
            gts.append(gp)
            preds.append(dp)
            out[f"gt{i}"] = np.packbits(g)
            out[f"depth_u16_{i}"] = np.round(d * 256).astype(np.uint16)
        out["shape"] = np.array([H, W])
        rng = [40, 120, 240]
        pv, rv = ede.pr_evaluation(gts, preds, edge_thresh_range=rng, gt_crop=mp,
                                   save_folder=os.path.join(tmp, "out"), num_workers=1)
        out["range"] = np.array(rng)
        out["precision"] = np.array(pv, np.float64)
        out["recall"] = np.array(rv, np.float64)
        # _pred_eval itself on the last setting's predicted edge image of scene 0 (JPEG on disk), mask and list crop
        pe = sorted(glob.glob(os.path.join(tmp, "out", "*_pred_canny_edge.jpeg")))[0]
        out["pred_edge0"] = cv2.imread(pe)[:, :, 0]
        for tag, crop in (("mask", mp), ("crop", str([10, 250, 8, 112])), ("nocrop", "[]")):
            r = ede._pred_eval(pe, gts[0], crop)
            out[f"pe_{tag}"] = np.array([r.count_r_overall[0], r.sum_r_overall[0], r.count_p_overall[0],
                                         r.sum_p_overall[0], r.recall[0], r.precision[0]], np.float64)
    np.savez_compressed(os.path.join(HERE, "pr_mask.npz"), **out)
    print("pr_mask:", out["precision"], out["recall"], out["pe_mask"], out["pe_crop"])


def gen_targets():
    """resize_depth_preserve + the /255 rule of resize_sample from the UNMODIFIED datasets/augmentations.py, the normal
    decode expression of datasets/gta_dataset.py:413 and the float32 cast of to_tensor_sample."""
    from _reference_loader import load_augmentations
    A = load_augmentations()
    r = np.random.default_rng(11)
    out = {}
    cases = [((48, 64), (48, 64)), ((48, 64), (24, 32)), ((96, 160), (24, 40)), ((37, 53), (20, 31)), ((24, 32), (48, 64)),
             ((50, 70), (17, 70)), ((8, 8), (1, 1))]
    for i, (hw, HW) in enumerate(cases):
        e = ((r.random(hw) < 0.08) * r.integers(1, 256, hw)).astype(np.uint8)
        if i == 5:
            e = (e > 0).astype(np.uint8)          # 0/1 map: the "/255 if max > 1" rule must not fire
        if i == 6:
            e[:] = 0                               # nothing valid
        ref = A.resize_depth_preserve(e, HW)[:, :, 0]
        if np.max(ref) > 1:                        # resize_sample, augmentations.py:196-199
            ref = ref / 255
        out[f"edge_in{i}"] = e
        out[f"edge_shape{i}"] = np.array(HW)
        out[f"edge_out{i}"] = ref.astype(np.float32)   # to_tensor_sample: FloatTensor
    out["n_edge"] = np.int64(len(cases))
    v = np.arange(256, dtype=np.uint8)
    out["theta"] = ((360. * (v / 255.) - 180) * (np.pi / 180)).astype(np.float32)   # gta_dataset.py:413
    np.savez_compressed(os.path.join(HERE, "targets.npz"), **out)
    print("targets:", len(cases), "edge cases + 256 angles")


def gen_chamfer():
    """chamfer_distance of the UNMODIFIED packnet_sfm/utils/edge.py on synthetic edge maps (both directions), and
    the 9 light metrics of compute_edge_metrics restated around it (cv2.Canny + the reference chamfer_distance)."""
    import cv2
    from _reference_loader import load_utils_edge
    E = load_utils_edge()
    out = {}
    r = np.random.default_rng(5)
    cases = []
    for k, (H, W) in enumerate([(60, 90), (218, 1153), (40, 300), (7, 5)]):
        gt, depth = synth_gt_and_depth(max(H, 32), max(W, 32), 300 + k)
        gt = gt[:H, :W].astype(np.uint8) * 255
        pred = cv2.Canny((np.clip(depth[:H, :W], 0, 80) * (255.0 / 80)).astype(np.uint8), 10, 20)
        cases.append((pred, gt))
    cases.append((np.zeros((20, 30), np.uint8), cases[0][1][:20, :30]))           # no predicted pixel
    cases.append((cases[0][0][:20, :30], np.zeros((20, 30), np.uint8)))           # no GT pixel (scipy's convention)
    cases.append(((r.random((33, 47)) < 0.5).astype(np.uint8) * 255, (r.random((33, 47)) < 0.01).astype(np.uint8) * 255))
    out["n"] = np.int64(len(cases))
    for i, (p, g) in enumerate(cases):
        out[f"pred{i}"] = np.packbits(p > 127)
        out[f"gt{i}"] = np.packbits(g > 127)
        out[f"shape{i}"] = np.array(p.shape)
        for tag, (a, b) in (("pg", (p, g)), ("gp", (g, p))):
            with np.errstate(invalid="ignore", divide="ignore"):
                c_dist, perc, cond = E.chamfer_distance(a.astype(np.float64), b.astype(np.float64))
            out[f"{tag}{i}_cdist"] = np.float64(c_dist)
            out[f"{tag}{i}_perc"] = np.float64(perc)
            out[f"{tag}{i}_cond"] = cond.astype(np.int8)
    # compute_edge_metrics: model_wrapper.py cannot be imported here (yacs, MinkowskiEngine, ...): its body for a
    # depth model is restated in oracle/chamfer.py; pin the restatement's arithmetic to the reference chamfer_distance
    gt, depth = synth_gt_and_depth(160, 512, 311)
    vis = (depth * (255.0 / np.max(depth))).astype(np.uint8)
    crop = [18, 480, 60, 150]
    vals = []
    for lo, hi in ((10, 20), (20, 40), (30, 60)):
        im = cv2.Canny(vis, lo, hi)[crop[2]:crop[3], crop[0]:crop[1]]
        g = (gt.astype(np.float64) * 255)[crop[2]:crop[3], crop[0]:crop[1]]
        _, p1, _ = E.chamfer_distance(im, g)
        _, p2, _ = E.chamfer_distance(g, im)
        vals += [p1, p2, 2 * ((p1 * p2) / (p1 + p2))]
    out["metrics_depth_u16"] = np.round(depth * 256).astype(np.uint16)   # depth is on the 1/256 m grid
    out["metrics_crop"] = np.array(crop)
    out["metrics_shape"] = np.array(depth.shape)
    out["metrics_gt"] = np.packbits(gt)
    out["metrics_vals"] = np.array(vals, np.float64)
    np.savez_compressed(os.path.join(HERE, "chamfer.npz"), **out)
    print("chamfer:", out["metrics_vals"])


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "chamfer":
        gen_chamfer()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "pr_mask":
        gen_pr_mask()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "targets":
        gen_targets()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "edge_loss_alt2":
        torch.manual_seed(0)
        gen_edge_loss_alt2()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "edge_loss_alt":
        torch.manual_seed(0)
        gen_edge_loss_alt()
        sys.exit(0)
    torch.manual_seed(0)
    gen_edge_loss()
    gen_edge_loss_alt()
    gen_edge_loss_alt2()
    gen_canny()
    gen_dee()
    gen_pr()
    gen_pr_mask()
    gen_chamfer()
    gen_targets()
