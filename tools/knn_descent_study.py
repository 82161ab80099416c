"""Where do the octree walks of csrc/knn_walk.cuh spend their work?  CPU emulation (numpy) of the SAME walk - Morton
octree with leaves <= 32 points, tight box + disc bound per node, box test first and disc test only for the children
the box cannot reject, nearest-child-first, pruned by the 8th-best distance, warm-started with the previous sample's
neighbours - over the coarse samples of real rays of the bench frame.  Counts per query: internal nodes expanded, child
box tests, child disc tests, leaves scanned, points scanned; and how many of the box tests a one-float4 bounding-sphere
pre-test would already have rejected.  Round-2 planning aid, not product code.  Usage: python tools/knn_descent_study.py
"""
import sys

import numpy as np

sys.path.insert(0, ".")
from neumesh_b200 import synth  # noqa: E402

LEAF, K, BITS, DISC_MAX = 32, 8, 10, 8192


def build(pts):
    lo, hi = pts.min(0), pts.max(0)
    q = np.minimum(((pts - lo) / (hi - lo).max() * (1 << BITS)).astype(np.int64), (1 << BITS) - 1)
    code = np.zeros(len(pts), np.int64)
    for b in range(BITS):
        for a in range(3):
            code |= ((q[:, a] >> b) & 1) << (3 * b + (2 - a))
    order = np.argsort(code, kind="stable")
    pts, code = pts[order], code[order]
    nodes = []

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
        p = pts[b:e]
        c = p.mean(0)
        disc = None
        if e - b <= DISC_MAX:
            w, v = np.linalg.eigh(np.cov((p - c).T) + 1e-18 * np.eye(3)) if e - b > 2 else (None, np.eye(3))
            u = v[:, 0]
            disc = (c, u, np.abs((p - c) @ u).max(), np.linalg.norm(p - c, axis=1).max())
        bc = 0.5 * (p.min(0) + p.max(0))
        nodes[i] = dict(lo=p.min(0), hi=p.max(0), b=b, e=e, kids=kids, disc=disc, sc=bc, level=level,
                        sr=np.linalg.norm(p - bc, axis=1).max())
        return i

    rec(0, len(pts), 0)
    return pts, nodes


def box_d2(n, q):
    d = np.maximum(np.maximum(n["lo"] - q, q - n["hi"]), 0.0)
    return float(d @ d)


def disc_d2(n, q):
    if n["disc"] is None:
        return 0.0
    c, u, t, r = n["disc"]
    v = q - c
    a = float(v @ u)
    bb = np.sqrt(max(float(v @ v) - a * a, 0.0))
    h, l = max(abs(a) - t, 0.0), max(bb - r, 0.0)
    return h * h + l * l


def walk(pts, nodes, q, warm_idx, cnt):
    if warm_idx is None:
        bd, bi = np.full(K, np.inf), np.full(K, -1)
    else:
        d = ((pts[warm_idx] - q) ** 2).sum(1)
        o = np.argsort(d)
        bd, bi = d[o], warm_idx[o]
    stack = [(0.0, 0)]
    while stack:
        d, i = stack.pop()
        if d > bd[-1]:
            continue
        n = nodes[i]
        if not n["kids"]:
            cnt["leaves"] += 1
            cnt["points"] += n["e"] - n["b"]
            d2 = ((pts[n["b"]:n["e"]] - q) ** 2).sum(1)
            idx = np.arange(n["b"], n["e"])
            keep = ~np.isin(idx, bi)
            alld, alli = np.concatenate([bd, d2[keep]]), np.concatenate([bi, idx[keep]])
            o = np.argsort(alld, kind="stable")[:K]
            bd, bi = alld[o], alli[o]
            continue
        cnt["internal"] += 1
        cnt["tests_top3" if n["level"] < 3 else "tests_deep"] += len(n["kids"])
        cand = []
        for c in n["kids"]:
            ch = nodes[c]
            cnt["box"] += 1
            sph = max(np.linalg.norm(q - ch["sc"]) - ch["sr"], 0.0) ** 2
            if sph > bd[-1]:
                cnt["box_sphere_rejectable"] += 1
            b2 = box_d2(ch, q)
            if b2 <= bd[-1]:
                cnt["disc"] += 1
                b2 = max(b2, disc_d2(ch, q))
                if b2 <= bd[-1]:
                    cand.append((b2, c))
        for item in sorted(cand, reverse=True):
            stack.append(item)
    return bi


def main():
    rng = np.random.default_rng(1)
    mesh = synth.icosphere_mesh(7, seed=0)
    pts, nodes = build(np.asarray(mesh.vertices, np.float64))
    o, d = synth.frame_rays(800, 800, view=0)
    o, d = o.numpy().astype(np.float64), d.numpy().astype(np.float64)
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    keys = ("internal", "box", "box_sphere_rejectable", "disc", "leaves", "points", "tests_top3", "tests_deep")
    groups = {"hit rays (pass within 0.05 of the mesh)": [], "near-miss rays": [], "far-miss rays (> 0.3 from the mesh)": []}
    rays = rng.choice(len(o), 260, replace=False)
    for r in rays:
        mid = -(o[r] @ d[r])
        disc = mid * mid - o[r] @ o[r] + 1.0
        if disc <= 0:
            continue
        near, far = mid - np.sqrt(disc), mid + np.sqrt(disc)
        t = near + (far - near) * np.linspace(0, 1, 64)
        q = o[r] + d[r] * t[:, None]
        closest = np.sqrt(((q[:, None, :] - pts[None, ::61, :]) ** 2).sum(-1).min())
        cnt = dict.fromkeys(keys, 0)
        warm = None
        for s in range(64):
            warm = walk(pts, nodes, q[s], warm, cnt)
        g = "hit rays (pass within 0.05 of the mesh)" if closest < 0.05 else (
            "far-miss rays (> 0.3 from the mesh)" if closest > 0.3 else "near-miss rays")
        groups[g].append([cnt[k] / 64.0 for k in keys])
    print(f"per query (64 coarse samples per ray, warm-started), V = {len(pts)}, leaves <= {LEAF}, {len(nodes)} nodes")
    for g, rows in groups.items():
        if not rows:
            continue
        m = np.array(rows).mean(0)
        print(f"  {g}: {len(rows)} rays | internal nodes {m[0]:.1f} | box tests {m[1]:.1f} (sphere pre-test would reject "
              f"{100 * m[2] / max(m[1], 1e-9):.0f} %) | disc tests {m[3]:.1f} | leaves {m[4]:.1f} | points {m[5]:.0f} | "
              f"box tests under nodes of level 0-2: {m[6]:.1f}, deeper: {m[7]:.1f}")


if __name__ == "__main__":
    main()
