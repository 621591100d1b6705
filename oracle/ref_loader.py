"""Import the staged UNMODIFIED reference (``oracle/_ref/``, see make_ref.py) for the CPU baseline arm.

TEST / BENCH INFRASTRUCTURE: used by ``bench.py --impl reference`` only, in a process of its own (the ``.cuda()`` ->
identity shim below is process-global and must never be active next to the product code).
"""
import importlib.util
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.exists(os.path.join(REF, "MANIFEST.json"))


def _file_module(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod   # picklable for the reference's multiprocessing.Pool
    spec.loader.exec_module(mod)
    return mod


def load_gradloss_cpu():
    """The reference GradLoss on CPU torch: ``.cuda()`` at import / construction time (attention_loss.py:13,
    grad_loss.py:51-54) is shimmed to the identity."""
    import torch
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from packnet_code.packnet_sfm.losses.grad_loss import GradLoss
    return GradLoss


def load_tools():
    return _file_module("ref_tools", "packnet_code/packnet_sfm/utils/tools.py")


def load_eval_depth_edges():
    """eval_depth_edges.py with the oracle's matcher / thinner standing in for py-bsds500 (absent: PARITY UNPINNED)."""
    from . import pr_counts as opr, thin as othin
    cp = types.ModuleType("bsds_metric.bsds.correspond_pixels")
    cp.correspond_pixels = opr.correspond_pixels
    th = types.ModuleType("bsds_metric.bsds.thin")
    th.binary_thin = othin.binary_thin
    pkg, sub = types.ModuleType("bsds_metric"), types.ModuleType("bsds_metric.bsds")
    sub.thin, sub.correspond_pixels, pkg.bsds = th, cp, sub
    sys.modules.update({"bsds_metric": pkg, "bsds_metric.bsds": sub, "bsds_metric.bsds.thin": th,
                        "bsds_metric.bsds.correspond_pixels": cp})
    sys.modules["edge"] = _file_module("edge", "edge.py")
    return _file_module("ref_eval_depth_edges", "eval_depth_edges.py")
