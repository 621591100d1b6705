"""CPU probe for the next matcher design: how does an image's matching problem split into
connected components of its proximity graph (predicted pixel -- GT pixel within max_dist * diagonal)?  Prints, for the
bench's KITTI-DE set, the number of components, the share of the largest ones, and the makespan of a greedy
longest-first packing of the components of ALL images onto 148 SMs (cost model: vertices per component)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import heapq
import numpy as np, cv2
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import connected_components
import bench
from oracle.canny import quantise_depth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 102
depths, gts = bench.kitti_like_set(n, 7000)
x0, x1, y0, y1 = 44, 1197, 153, 371
h, w = y1 - y0, x1 - x0
radius = 0.002 * np.hypot(h, w)
offs = [(dy, dx) for dy in range(-3, 4) for dx in range(-3, 4) if dy * dy + dx * dx <= radius * radius]
all_sizes, per_img = [], []
for i in range(n):
    pred = cv2.Canny(quantise_depth(depths[i]), 10, 20)[y0:y1, x0:x1] > 0      # loosest pair: every stage's pixels
    gt = gts[i][y0:y1, x0:x1] > 0
    pid = -np.ones((h, w), np.int64); pid[pred] = np.arange(pred.sum())
    qid = -np.ones((h, w), np.int64); qid[gt] = pred.sum() + np.arange(gt.sum())
    nv = int(pred.sum() + gt.sum())
    rows, cols = [], []
    py, px = np.nonzero(pred)
    for dy, dx in offs:
        yy, xx = py + dy, px + dx
        ok = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
        q = qid[yy[ok], xx[ok]]
        m = q >= 0
        rows.append(pid[py[ok][m], px[ok][m]]); cols.append(q[m])
    r, c = np.concatenate(rows), np.concatenate(cols)
    k, lab = connected_components(coo_matrix((np.ones(len(r)), (r, c)), shape=(nv, nv)), directed=False)
    sizes = np.sort(np.bincount(lab))[::-1]
    sizes = sizes[sizes > 1]                                                   # isolated vertices need no matching
    per_img.append((nv, len(sizes), sizes[:3].tolist(), float(sizes[0]) / max(nv, 1)))
    all_sizes += sizes.tolist()
per_img_nv = np.array([p[0] for p in per_img])
print("images %d: vertices per image mean %.0f max %d" % (n, per_img_nv.mean(), per_img_nv.max()))
print("components per image mean %.0f; largest component / image vertices: mean %.3f max %.3f" %
      (np.mean([p[1] for p in per_img]), np.mean([p[3] for p in per_img]), np.max([p[3] for p in per_img])))
for i in np.argsort(-per_img_nv)[:5]:
    print("  img %3d vertices %5d components %4d largest %s" % (i, per_img[i][0], per_img[i][1], per_img[i][2]))
# greedy longest-first packing of all components onto 148 SMs
loads = [0] * 148
heapq.heapify(loads)
for s in sorted(all_sizes, reverse=True):
    heapq.heappush(loads, heapq.heappop(loads) + s)
print("one CTA per image: makespan %d vertices (heaviest image), mean %.0f; components packed on 148 SMs: makespan %d vertices"
      " (largest single component %d) -> %.1fx shorter critical path under a vertices-proportional cost model" %
      (per_img_nv.max(), per_img_nv.mean(), max(loads), max(all_sizes), per_img_nv.max() / max(loads)))
