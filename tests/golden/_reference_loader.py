"""Import the UNMODIFIED reference from /root/reference (build container only).

Used solely by tests/golden/make_golden.py to generate the committed golden
vectors; /root/reference does not exist on the GPU box, so nothing in the test
suite imports this at run time.
"""
import importlib.util
import sys
import types

REF = "/root/reference"


def load_gradloss():
    """GradLoss with the `.cuda()` -> identity shim."""
    import torch
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from packnet_code.packnet_sfm.losses.grad_loss import GradLoss
    return GradLoss


def load_tools():
    spec = importlib.util.spec_from_file_location(
        "ref_tools", REF + "/packnet_code/packnet_sfm/utils/tools.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod  # picklable for the reference's multiprocessing.Pool
    spec.loader.exec_module(mod)
    return mod


def load_edge():
    spec = importlib.util.spec_from_file_location("ref_edge", REF + "/edge.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod  # picklable for the reference's multiprocessing.Pool
    spec.loader.exec_module(mod)
    return mod


def load_eval_depth_edges(thin_mod, correspond_mod):
    """eval_depth_edges.py with a stand-in `bsds_metric.bsds` package
    (the real py-bsds500 is not vendored: SURVEY.md 8c row 3)."""
    pkg = types.ModuleType("bsds_metric")
    sub = types.ModuleType("bsds_metric.bsds")
    sub.thin = thin_mod
    sub.correspond_pixels = correspond_mod
    pkg.bsds = sub
    sys.modules["bsds_metric"] = pkg
    sys.modules["bsds_metric.bsds"] = sub
    sys.modules["bsds_metric.bsds.thin"] = thin_mod
    sys.modules["bsds_metric.bsds.correspond_pixels"] = correspond_mod
    if REF not in sys.path:
        sys.path.insert(0, REF)
    sys.modules["edge"] = load_edge()
    spec = importlib.util.spec_from_file_location("ref_eval_depth_edges", REF + "/eval_depth_edges.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod  # picklable for the reference's multiprocessing.Pool
    spec.loader.exec_module(mod)
    return mod


def load_utils_edge():
    """packnet_sfm/utils/edge.py (chamfer_distance).  Its import chain wants `yacs`, which is not installed here and
    is only used for an isinstance check in utils/types.py: a stub module stands in."""
    y = types.ModuleType("yacs")
    yc = types.ModuleType("yacs.config")

    class CfgNode(dict):
        pass
    yc.CfgNode = CfgNode
    y.config = yc
    y.CfgNode = CfgNode
    sys.modules.setdefault("yacs", y)
    sys.modules.setdefault("yacs.config", yc)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from packnet_code.packnet_sfm.utils import edge as ref_utils_edge
    return ref_utils_edge


def load_augmentations():
    """packnet_sfm/datasets/augmentations.py (resize_depth_preserve).  Needs the `yacs` stub of load_utils_edge and
    `Image.ANTIALIAS`, which Pillow >= 10 removed (the module uses it as a default argument at import time)."""
    from PIL import Image
    if not hasattr(Image, "ANTIALIAS"):
        Image.ANTIALIAS = Image.LANCZOS
    load_utils_edge()
    from packnet_code.packnet_sfm.datasets import augmentations as ref_aug
    return ref_aug
