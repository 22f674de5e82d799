"""Feasibility: pre-split (clip) long thin triangle references before the SAH build; node visits / primitive tests per ray.
    python tools/presplit_sim.py config:c2 800 [max_refs_factor]"""
import sys, os, runpy
import numpy as np
sys.argv = [sys.argv[0]] + sys.argv[1:]
name = sys.argv[1] if len(sys.argv) > 1 else "config:c2"
NR = int(sys.argv[2]) if len(sys.argv) > 2 else 600
FACT = float(sys.argv[3]) if len(sys.argv) > 3 else 2.0
sys.argv = ["bvh_sim", name, "1"]
ns = runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), "bvh_sim.py"))
boxes, geom, kinds, prim_hit, n = ns["boxes"], ns["geom"], ns["kinds"], ns["prim_hit"], ns["n"]
area = ns["area"]

def clip_poly(poly, axis, lo, hi):
    def clip(poly, sign, c):
        out = []
        for i in range(len(poly)):
            a, b = poly[i], poly[(i + 1) % len(poly)]
            da, db = sign * (a[axis] - c), sign * (b[axis] - c)
            if da >= 0: out.append(a)
            if (da >= 0) != (db >= 0):
                t = da / (da - db); out.append(a + t * (b - a))
        return out
    poly = clip(poly, 1.0, lo)
    if poly: poly = clip(poly, -1.0, hi)
    return poly

# references: (item, box, polygon or None)
refs = []
for i in range(n):
    if kinds[i] == 2:
        refs.append((i, boxes[i].copy(), [geom[i][0:3].copy(), geom[i][3:6].copy(), geom[i][6:9].copy()]))
    else:
        refs.append((i, boxes[i].copy(), None))
budget = int(FACT * n)
import heapq
# priority: box area of triangle refs (largest first)
heap = [(-area(b), k) for k, (it, b, p) in enumerate(refs) if p is not None]
heapq.heapify(heap)
while heap and len(refs) < budget:
    negA, k = heapq.heappop(heap)
    it, b, poly = refs[k]
    ext = b[3:6] - b[0:3]
    ax = int(np.argmax(ext))
    if ext[ax] <= 0: continue
    mid = 0.5 * (b[ax] + b[3 + ax])
    parts = []
    for lo, hi in ((b[ax], mid), (mid, b[3 + ax])):
        p2 = clip_poly(poly, ax, lo, hi)
        if len(p2) >= 3:
            P = np.array(p2); parts.append((it, np.concatenate([P.min(0), P.max(0)]), p2))
    if len(parts) == 2 and area(parts[0][1]) + area(parts[1][1]) < 0.9 * area(b):
        refs[k] = parts[0]; refs.append(parts[1])
        heapq.heappush(heap, (-area(parts[0][1]), k)); heapq.heappush(heap, (-area(parts[1][1]), len(refs) - 1))
print("refs", len(refs), "of", n, "elements")
rboxes = np.array([r[1] for r in refs]); ritem = [r[0] for r in refs]

# SAH over references (reuse builder by monkeypatching globals)
def build(rb):
    cen = (rb[:, :3] + rb[:, 3:]) / 2
    nodes = []
    def ar(l, h):
        e = np.maximum(h - l, 0); return 2 * (e[:, 0] * e[:, 1] + e[:, 1] * e[:, 2] + e[:, 2] * e[:, 0])
    def rec(idx):
        if len(idx) == 1: return ~int(idx[0])
        best = (np.inf, None, None)
        for ax in range(3):
            order = idx[np.argsort(cen[idx, ax], kind="stable")]
            lo = np.minimum.accumulate(rb[order, :3], axis=0); hi = np.maximum.accumulate(rb[order, 3:], axis=0)
            rlo = np.minimum.accumulate(rb[order[::-1], :3], axis=0)[::-1]; rhi = np.maximum.accumulate(rb[order[::-1], 3:], axis=0)[::-1]
            k = np.arange(1, len(order))
            cost = ar(lo[:-1], hi[:-1]) * k + ar(rlo[1:], rhi[1:]) * (len(order) - k)
            j = int(np.argmin(cost))
            if cost[j] < best[0]: best = (cost[j], order, j + 1)
        _, order, k = best
        me = len(nodes); nodes.append(None)
        l = rec(order[:k]); r = rec(order[k:])
        nodes[me] = (np.concatenate([rb[order, :3].min(0), rb[order, 3:].max(0)]), l, r)
        return me
    sys.setrecursionlimit(100000)
    root = rec(np.arange(len(rb)))
    return nodes, root

def slab(b, o, inv, t0, t1):
    ta = (b[0:3] - o) * inv; tb = (b[3:6] - o) * inv
    return max(t0, np.nanmax(np.minimum(ta, tb))), min(t1, np.nanmin(np.maximum(ta, tb)))

def traverse(tree, rb, items, o, d, t0=1e-3):
    nodes, root = tree
    with np.errstate(divide="ignore"): inv = 1.0 / d
    best, bt, nv, npt = -1, np.inf, 0, 0
    stack = [(root, t0)]
    while stack:
        c, tn = stack.pop()
        if tn >= bt: continue
        if c < 0:
            npt += 1
            t = prim_hit(items[~c], o, d, t0, bt)
            if t is not None: best, bt = items[~c], t
            continue
        nv += 1
        kids = []
        for ch in (nodes[c][1], nodes[c][2]):
            b = rb[~ch] if ch < 0 else nodes[ch][0]
            tn_c, tf_c = slab(b, o, inv, t0, bt)
            if tn_c <= tf_c: kids.append((ch, tn_c))
        kids.sort(key=lambda x: -x[1]); stack.extend(kids)
    return best, bt, nv, npt

T0 = build(boxes); T1 = build(rboxes)
rng = np.random.default_rng(0)
cam = ns["fs"].cam[0:3].astype(float)
tgt = boxes[rng.integers(0, n, NR)]; tgt = tgt[:, :3] + rng.uniform(size=(NR, 3)) * (tgt[:, 3:] - tgt[:, :3])
d0 = tgt - cam; d0 /= np.linalg.norm(d0, axis=1, keepdims=True)
gen = (np.tile(cam, (NR, 1)), d0)
for bounce in range(4):
    o, d = gen
    if len(o) == 0: break
    s0 = [0, 0]; s1 = [0, 0]; no, nd = [], []; mism = 0
    for i in range(len(o)):
        a = traverse(T0, boxes, list(range(n)), o[i], d[i]); b = traverse(T1, rboxes, ritem, o[i], d[i])
        s0[0] += a[2]; s0[1] += a[3]; s1[0] += b[2]; s1[1] += b[3]; mism += a[0] != b[0]
        if a[0] >= 0:
            v = rng.normal(size=3); v /= np.linalg.norm(v); no.append(o[i] + a[1] * d[i]); nd.append(v)
    m = len(o)
    print("gen %d: %d rays | sah nodes %.1f prims %.1f | presplit nodes %.1f prims %.1f | mismatch %d" % (bounce, m, s0[0] / m, s0[1] / m, s1[0] / m, s1[1] / m, mism))
    if not no: break
    gen = (np.array(no), np.array(nd))
