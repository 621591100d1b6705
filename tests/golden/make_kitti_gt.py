"""Pack the reference's bundled KITTI-DE ground-truth edge maps (data/kitti_de/gt/*.png, 102 images, 1-bit, 384x1280,
listed by data/kitti_de/kitti_de_annotated_edges.txt) into one small fixture.

Run in the build container only (needs /root/reference):   python tests/golden/make_kitti_gt.py
bench.py's AUC workload (BASELINE.json config 2: "bundled GT edges vs synthetic predicted depth") reads the fixture;
/root/reference does not exist on the GPU box.
"""
import glob
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
files = sorted(glob.glob("/root/reference/data/kitti_de/gt/*.png"))
maps = []
for f in files:
    im = cv2.imread(f, cv2.IMREAD_UNCHANGED)
    if im.ndim == 3:
        im = im[:, :, 0]
    maps.append(im > 0)
a = np.stack(maps)
np.savez_compressed(os.path.join(HERE, "kitti_de_gt.npz"), bits=np.packbits(a), shape=np.array(a.shape),
                    names=np.array([os.path.basename(f) for f in files]))
print(a.shape, "mean density %.4f" % a.mean())
