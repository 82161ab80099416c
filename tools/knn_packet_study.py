"""Algorithm-level study for the round-2 KNN kernel (CPU, numpy): how many octree leaves does a WARP-WIDE PACKET walk
visit compared with 32 independent per-thread walks?

Today every lane walks the octree on its own (ncu: 8.5 of 32 lanes active on average).  A packet walk keeps the 32 lanes
of a warp - 32 neighbouring rays (Morton-ordered 4 x 8 pixel block) at the same sample index - in lock step on ONE shared
stack: a node is pruned when its box is farther from the packet's bounding box than the WORST lane's current 8th-best
distance, and at a leaf every lane scans all (<= 32) points against its own query (broadcast loads, no divergence).
The packet visits the UNION of the leaves its lanes need; it wins when  union < (32 / 8.5) x mean per-lane leaves.

Same tree shape as csrc/grid.cu (Morton order, leaves <= 32 points, tight boxes; box bounds only - no disc bounds -
for both variants).  Usage: python tools/knn_packet_study.py
"""
import sys

import numpy as np

sys.path.insert(0, ".")
from neumesh_b200 import synth  # noqa: E402

LEAF, K, BITS = 32, 8, 10


def build(pts):
    lo, hi = pts.min(0), pts.max(0)
    q = np.minimum(((pts - lo) / (hi - lo).max() * (1 << BITS)).astype(np.int64), (1 << BITS) - 1)
    code = np.zeros(len(pts), np.int64)
    for b in range(BITS):
        for a in range(3):
            code |= ((q[:, a] >> b) & 1) << (3 * b + (2 - a))
    order = np.argsort(code, kind="stable")
    pts, code = pts[order], code[order]
    nodes = []   # (lo, hi, begin, end, children)

    def rec(b, e, level):
        i = len(nodes)
        nodes.append(None)
        kids = []
        if e - b > LEAF and level < BITS:
            shift = 3 * (BITS - level - 1)
            keys = code[b:e] >> shift
            cuts = np.searchsorted(keys, np.arange(keys[0], keys[-1] + 2))
            for c in range(len(cuts) - 1):
                if cuts[c + 1] > cuts[c]:
                    kids.append(rec(b + cuts[c], b + cuts[c + 1], level + 1))
        nodes[i] = (pts[b:e].min(0), pts[b:e].max(0), b, e, kids)
        return i

    rec(0, len(pts), 0)
    return pts, nodes


def box_d2(lo, hi, qlo, qhi):
    """squared distance between boxes [lo,hi] and [qlo,qhi] (a point when qlo == qhi)"""
    d = np.maximum(np.maximum(lo - qhi, qlo - hi), 0.0)
    return float((d * d).sum())


def walk(pts, nodes, queries):
    """lock-step walk of `queries` ([n,3]; n = 1: the per-thread walk).  -> (leaves visited, node tests)"""
    n = len(queries)
    best = np.full((n, K), np.inf)
    qlo, qhi = queries.min(0), queries.max(0)
    leaves = tests = 0
    stack = [(0.0, 0)]
    while stack:
        d, i = stack.pop()
        if d >= best[:, -1].max():
            continue
        lo, hi, b, e, kids = nodes[i]
        if not kids:
            leaves += 1
            d2 = ((queries[:, None, :] - pts[None, b:e, :]) ** 2).sum(-1)
            best = np.sort(np.concatenate([best, d2], axis=1), axis=1)[:, :K]
            continue
        cand = []
        for c in kids:
            tests += 1
            dc = box_d2(nodes[c][0], nodes[c][1], qlo, qhi)
            if dc < best[:, -1].max():
                cand.append((dc, c))
        for item in sorted(cand, reverse=True):   # nearest child on top of the stack
            stack.append(item)
    return leaves, tests


def main():
    rng = np.random.default_rng(0)
    mesh = synth.icosphere_mesh(7, seed=0)
    pts, nodes = build(np.asarray(mesh.vertices, np.float64))
    o, d = synth.frame_rays(800, 800, view=0)
    o, d = o.numpy().astype(np.float64).reshape(800, 800, 3), d.numpy().astype(np.float64).reshape(800, 800, 3)
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    rows = []
    for _ in range(120):
        y, x = rng.integers(0, 800 - 4), rng.integers(0, 800 - 8)
        oo, dd = o[y:y + 4, x:x + 8].reshape(-1, 3), d[y:y + 4, x:x + 8].reshape(-1, 3)
        mid = -(oo * dd).sum(-1)
        disc = mid ** 2 - (oo * oo).sum(-1) + 1.0
        if (disc <= 0).any():
            continue   # block misses the unit sphere: the renderer still samples it, but near = far handling differs
        near, far = mid - np.sqrt(disc), mid + np.sqrt(disc)
        s = rng.integers(0, 64)
        t = near + (far - near) * (s / 63.0)
        q = oo + dd * t[:, None]
        lu, tu = walk(pts, nodes, q)
        per = [walk(pts, nodes, q[i:i + 1]) for i in range(32)]
        l1 = np.mean([p[0] for p in per])
        t1 = np.mean([p[1] for p in per])
        dist = np.sqrt(((q[:, None, :] - pts[None, ::97, :]) ** 2).sum(-1).min(1)).mean()   # rough distance to the mesh
        rows.append((dist, l1, lu, t1, tu))
    r = np.array(rows)
    print(f"{len(r)} packets (4x8 pixel blocks at one coarse sample index), V = {len(pts)}, leaves <= {LEAF}")
    for lo_, hi_ in ((0, 0.03), (0.03, 0.1), (0.1, 0.3), (0.3, 2.0)):
        m = (r[:, 0] >= lo_) & (r[:, 0] < hi_)
        if m.sum() == 0:
            continue
        l1, lu, t1, tu = r[m, 1].mean(), r[m, 2].mean(), r[m, 3].mean(), r[m, 4].mean()
        print(f"  distance to mesh [{lo_:.2f}, {hi_:.2f}): {m.sum():3d} packets | leaves per lane {l1:6.1f}  packet union "
              f"{lu:6.1f}  (x{lu / l1:.2f}) | child-box tests per lane {t1:7.1f}  packet {tu:7.1f} (x{tu / t1:.2f}) | "
              f"break-even x{32 / 8.5:.2f}")


if __name__ == "__main__":
    main()
