"""Stand-in for ``bsds_metric.bsds`` (py-bsds500), the package the reference imports at
``eval_depth_edges.py:7`` (``from bsds_metric.bsds import thin, correspond_pixels``).

Putting this package on ``sys.modules`` under that name makes the
UNMODIFIED reference script run with the matcher and the thinner on the GPU."""
from . import correspond_pixels, thin  # noqa: F401


def install_as_bsds_metric():
    """Register this package as ``bsds_metric.bsds`` so ``eval_depth_edges.py`` imports it."""
    import sys
    import types
    pkg = types.ModuleType("bsds_metric")
    pkg.bsds = sys.modules[__name__]
    sys.modules["bsds_metric"] = pkg
    sys.modules["bsds_metric.bsds"] = sys.modules[__name__]
    sys.modules["bsds_metric.bsds.thin"] = thin
    sys.modules["bsds_metric.bsds.correspond_pixels"] = correspond_pixels
