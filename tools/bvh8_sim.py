"""Feasibility of an 8-wide (collapsed, 8-bit quantised) accelerator tree: node visits per ray vs the binary tree of traversal 3.
    python tools/bvh8_sim.py config:c2 400 8"""
import sys, os, runpy
import numpy as np
name = sys.argv[1] if len(sys.argv) > 1 else "config:c2"
NR = int(sys.argv[2]) if len(sys.argv) > 2 else 400
FACT = sys.argv[3] if len(sys.argv) > 3 else "8"
sys.argv = ["presplit_sim", name, "0", FACT]     # 0 rays: only build refs + trees
ns = runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), "presplit_sim.py"))
T1, rboxes, ritem, prim_hit, boxes, n, traverse, slab = ns["T1"], ns["rboxes"], ns["ritem"], ns["prim_hit"], ns["boxes"], ns["n"], ns["traverse"], ns["slab"]
nodes, root = T1
area = ns["area"]

def collapse(width):
    wide = []   # list of [child refs], child = ~ref (leaf) or wide index
    def boxof(c): return rboxes[~c] if c < 0 else nodes[c][0]
    def rec(c):
        me = len(wide); wide.append(None)
        kids = [nodes[c][1], nodes[c][2]]
        while len(kids) < width:
            cand = [(area(boxof(k)), i) for i, k in enumerate(kids) if k >= 0]
            if not cand: break
            _, i = max(cand)
            k = kids.pop(i); kids += [nodes[k][1], nodes[k][2]]
        out = []
        for k in kids:
            out.append((k if k < 0 else None, boxof(k).copy(), k))
        wide[me] = out
        res = []
        for leaf, b, k in out:
            res.append((leaf if leaf is not None else rec(k), b))
        wide[me] = res
        return me
    sys.setrecursionlimit(100000)
    r = rec(root)
    return wide, r

def quantise(wide):
    q = []
    for kids in wide:
        lo = np.min([b[:3] for _, b in kids], axis=0); hi = np.max([b[3:] for _, b in kids], axis=0)
        e = np.ceil(np.log2(np.maximum(hi - lo, 1e-30) / 255.0)); s = 2.0 ** e
        out = []
        for c, b in kids:
            ql = np.floor((b[:3] - lo) / s); qh = np.ceil((b[3:] - lo) / s)
            out.append((c, np.concatenate([lo + ql * s, lo + qh * s])))
        q.append(out)
    return q

def traverse_wide(wide, wroot, o, d, t0=1e-3, order="near"):
    with np.errstate(divide="ignore"): inv = 1.0 / d
    best, bt, nv, nbox, npt = -1, np.inf, 0, 0, 0
    stack = [(wroot, t0)]
    while stack:
        c, tn = stack.pop()
        if tn >= bt: continue
        if c < 0:
            npt += 1
            t = prim_hit(ritem[~c], o, d, t0, bt)
            if t is not None: best, bt = ritem[~c], t
            continue
        nv += 1
        kids = []
        for ch, b in wide[c]:
            nbox += 1
            tn_c, tf_c = slab(b, o, inv, t0, bt)
            if tn_c <= tf_c: kids.append((ch, tn_c))
        kids.sort(key=lambda x: -x[1]); stack.extend(kids)
    return best, bt, nv, nbox, npt

W8, r8 = collapse(8); W4, r4 = collapse(4)
Q8 = quantise(W8)
print("binary nodes", len(nodes), "| wide-8 nodes", len(W8), "avg children %.2f" % np.mean([len(k) for k in W8]), "| wide-4 nodes", len(W4), "avg children %.2f" % np.mean([len(k) for k in W4]))
rng = np.random.default_rng(0)

fs = ns["ns"]["fs"] if "ns" in ns else None
cam = fs.cam[0:3].astype(float)
tgt = boxes[rng.integers(0, n, NR)]; tgt = tgt[:, :3] + rng.uniform(size=(NR, 3)) * (tgt[:, 3:] - tgt[:, :3])
d0 = tgt - cam; d0 /= np.linalg.norm(d0, axis=1, keepdims=True)
gen = (np.tile(cam, (NR, 1)), d0)
for bounce in range(5):
    o, d = gen
    acc = np.zeros((4, 3)); no, nd = [], []; mism = 0
    for i in range(len(o)):
        b = traverse(T1, rboxes, ritem, o[i], d[i])
        w8 = traverse_wide(W8, r8, o[i], d[i]); w4 = traverse_wide(W4, r4, o[i], d[i]); q8 = traverse_wide(Q8, r8, o[i], d[i])
        acc[0] += (b[2], 2 * b[2], b[3]); acc[1] += (w4[2], w4[3], w4[4]); acc[2] += (w8[2], w8[3], w8[4]); acc[3] += (q8[2], q8[3], q8[4])
        mism += (b[0] != w8[0]) + (b[0] != q8[0])
        if b[0] >= 0:
            v = rng.normal(size=3); v /= np.linalg.norm(v); no.append(o[i] + b[1] * d[i]); nd.append(v)
    m = len(o); acc /= m
    print("gen %d: %4d rays | binary visits %.1f boxes %.1f prims %.1f | wide4 %.1f / %.1f / %.1f | wide8 %.1f / %.1f / %.1f | wide8 quantised %.1f / %.1f / %.1f | mismatch %d" %
          (bounce, m, *acc[0], *acc[1], *acc[2], *acc[3], mism))
    if not no: break
    gen = (np.array(no), np.array(nd))
