"""Recipe for ``oracle/_ref/``: the UNMODIFIED reference files of the hot path, staged for the CPU baseline arm.

TEST / BENCH INFRASTRUCTURE.  The reference (liortalker/MindTheEdge) is pure Python, so there is nothing to compile:
this script copies the few source files of the path VERBATIM from ``/root/reference`` (build container only) into the
git-ignored ``oracle/_ref/`` tree, keeping their package layout, and records their SHA-256 in ``MANIFEST.json``.
``oracle/_ref/`` is listed in ``.gitignore`` (no reference source ever enters the history) but not in
``.gpurunignore``, so it travels to the GPU box like a built ``.so``.  Nothing in ``mindtheedge_b200`` imports it;
``bench.py --impl reference`` (and the ``cpu_baseline`` leg, through a subprocess) is the only consumer.

    python oracle/make_ref.py            # no-op when /root/reference is absent (GPU box: uses the staged copy)
"""
import hashlib
import json
import os
import shutil
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
FILES = [
    "packnet_code/packnet_sfm/losses/grad_loss.py",       # GradLayer / GradLoss (part 1)
    "packnet_code/packnet_sfm/losses/attention_loss.py",
    "packnet_code/packnet_sfm/utils/tools.py",            # non_max_suppression / hysteresis (part 2b)
    "edge.py",                                            # edge_from_depth (part 2a)
    "eval_depth_edges.py",                                # evaluate_boundaries / _pred_eval / pr_evaluation (part 3)
]


def make() -> bool:
    if not os.path.isdir(REF):
        return os.path.exists(os.path.join(OUT, "MANIFEST.json"))
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump({"source": "liortalker/MindTheEdge (unmodified copies, see oracle/make_ref.py)", "sha256": manifest},
                  f, indent=1)
    return True


if __name__ == "__main__":
    ok = make()
    print("oracle/_ref:", "ready" if ok else "unavailable (/root/reference absent and nothing staged)")
    sys.exit(0)
