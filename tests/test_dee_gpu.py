"""Bit-exact parity of the DEE post-process against goldens produced by the reference
``tools.py`` loops / cv2.Sobel, and against the vectorised oracle at full size."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from synth import prob_map

pytestmark = pytest.mark.gpu


def _eq(a, b):
    return a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


def test_golden_tools():
    from mindtheedge_b200 import tools
    z = np.load(os.path.join(GOLDEN, "dee.npz"))
    for i in range(5):
        p = z[f"prob{i}"]
        nms = tools.non_max_suppression(p)
        assert _eq(nms, z[f"nms{i}"]), ("nms", i)
        assert _eq(tools.hysteresis(nms), z[f"hyst{i}"]), ("hyst", i)
        assert _eq(tools.hysteresis(p), z[f"hyst_raw{i}"]), ("hyst_raw", i)
        assert _eq(tools.hysteresis(nms, 0.2, 0.5), z[f"hyst_custom{i}"]), ("hyst_custom", i)
        assert _eq(tools.edge_normals(p), z[f"normals{i}"]), ("normals", i)


@pytest.mark.parametrize("shape", [(384, 1280), (96, 320), (48, 160), (5, 300), (130, 3)])
def test_fused_batch_vs_oracle(shape):
    """The fused tensor-level op (normals + NMS + hysteresis, as infer_edge_estimation.py:244-255
    chains them) on a batch, against the vectorised oracle."""
    from mindtheedge_b200.tools import dee_postprocess
    from oracle import dee as odee
    H, W = shape
    probs = np.stack([prob_map(H, W, 40 + k) for k in range(3)])
    probs[2] *= 0.5  # an image with no strong pixel: max(labels) comes from the border -> NaN plane (0/0)
    nrm, out = dee_postprocess(torch.from_numpy(probs).cuda())
    nrm, out = nrm.cpu().numpy(), out.cpu().numpy()
    for k in range(3):
        assert np.array_equal(nrm[k], odee.normals_u8(probs[k])), ("normals", k)
        ref = odee.hysteresis(odee.non_max_suppression(probs[k]))
        assert _eq(out[k], ref), ("edges", k, int((out[k] != ref).sum()))
    assert (out[0] > 0).any()
