"""Tiny driver for ncu: a few fwd+bwd launches at config-3 shape (B=8, 4 scales), B=32 single scale, config 1."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import quick_loss_bench as q
cfgs = [(8, 384, 1280, 4, 2), (32, 384, 1280, 1, 2), (4, 384, 1280, 1, 2)]
if len(sys.argv) > 1: cfgs = [cfgs[int(sys.argv[1])]]
for cfg in cfgs:
    print(q.run(*cfg, iters=2))
