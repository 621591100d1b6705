"""Oracle for the in-training "light" edge metric (TEST INFRASTRUCTURE).

Restates ``chamfer_distance`` of packnet_code/packnet_sfm/utils/edge.py:20-62
(2-D inputs, no mask): Euclidean distance transform of the non-GT pixels, mean
distance at predicted pixels and the share of predicted pixels closer than
``edge_to_edge_thresh`` to a GT pixel.
"""
from __future__ import annotations

import numpy as np
from scipy import ndimage


def chamfer_distance(im_pred, im_gt, edge_to_edge_thresh=5):
    gt = (np.asarray(im_gt) / 255) > 0.5
    pred = (np.asarray(im_pred) / 255) > 0.5
    dist = ndimage.distance_transform_edt(~gt)
    n_pred = pred.sum()
    with np.errstate(invalid="ignore", divide="ignore"):
        c_dist = float((dist * pred).sum() / n_pred)
        close = int((dist[pred] < edge_to_edge_thresh).sum())
        percentage = float(close / n_pred)
    return c_dist, percentage, close, int(n_pred)
