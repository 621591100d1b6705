"""Diagnostic: how far apart are (a) libmte, (b) the reference op chain in eager torch on CUDA,
(c) the same chain on CPU and (d) the fp64 restatement, on one config-1-like image batch."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from test_edge_loss_gpu import _inputs
from oracle.edge_loss import edge_loss_torch, edge_loss_np64
from mindtheedge_b200.losses import edge_loss
import oracle.edge_loss as oe

def errs(a, b):
    a = a.astype(np.float64); b = b.astype(np.float64)
    return "max/max=%.2e  l2=%.2e" % (np.abs(a - b).max() / np.abs(b).max(), np.linalg.norm(a - b) / np.linalg.norm(b))

for shape in [(1, 384, 1280), (4, 384, 1280)]:
    depth, edge, normal = _inputs(*shape, seed=7)
    x = depth.clone().requires_grad_(True)
    l_cpu, g_cpu = edge_loss_torch(x, edge, None, True, True, 4, normal, weight=10.0); l_cpu.backward()
    l64, g64, dx64 = edge_loss_np64(depth.numpy(), edge.numpy(), None, True, True, 4, normal.numpy(), weight=10.0)
    xg = depth.cuda().requires_grad_(True)
    l_g, g_g = edge_loss(xg, edge.cuda(), None, True, True, 4, normal.cuda(), weight=10.0); l_g.backward()
    # eager torch on CUDA (the reference's own GPU path): same op chain as the oracle but on device
    xc = depth.cuda().requires_grad_(True)
    l_e, g_e = edge_loss_torch(xc, edge.cuda(), None, True, True, 4, normal.cuda(), weight=10.0); l_e.backward()
    print(shape, "loss cpu %.8f gpu %.8f f64 %.8f" % (l_cpu.item(), l_g.item(), l64))
    print("  loss rel: mte-vs-cpu %.2e  mte-vs-f64 %.2e  cpu-vs-f64 %.2e" % (abs(l_g.item()-l_cpu.item())/abs(l_cpu.item()), abs(l_g.item()-l64)/abs(l64), abs(l_cpu.item()-l64)/abs(l64)))
    print("  grad mte vs cpu :", errs(xg.grad.cpu().numpy(), x.grad.numpy()))
    print("  grad mte vs f64 :", errs(xg.grad.cpu().numpy(), dx64))
    print("  grad cpu vs f64 :", errs(x.grad.numpy(), dx64))
    print("  grad eager-cuda vs cpu :", errs(xc.grad.cpu().numpy(), x.grad.numpy()), " loss rel %.2e" % (abs(l_e.item()-l_cpu.item())/abs(l_cpu.item())))
    print("  grad eager-cuda vs f64 :", errs(xc.grad.cpu().numpy(), dx64))
    print("  gmap mte vs cpu :", errs(g_g.cpu().numpy(), g_cpu.numpy()))
    print("  gmap mte vs f64 :", errs(g_g.cpu().numpy(), g64))
    print("  gmap cpu vs f64 :", errs(g_cpu.numpy(), g64))
