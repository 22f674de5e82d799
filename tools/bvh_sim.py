"""Row f4 feasibility: node visits / primitive tests per ray on the reference topology (median split, round-robin axis)
versus a SAH binary tree over the same elements, for primary and bounce rays of a config.  CPU only, double precision.
    python tools/bvh_sim.py config:c2 [n_rays]
"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracinginaweekend_b200 as rb

name = sys.argv[1] if len(sys.argv) > 1 else "config:c2"
NR = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
W = rb.World.named(name)
fs = W.export()
n = len(fs.kinds)
geom = fs.geom.astype(np.float64)
kinds = fs.kinds

def leaf_box(i):
    g = geom[i]
    if kinds[i] == 0:
        return np.concatenate([g[0:3] - g[3], g[0:3] + g[3]])
    if kinds[i] == 2:
        p = g[0:9].reshape(3, 3)
        return np.concatenate([p.min(0), p.max(0)])
    if kinds[i] == 3:
        return g[0:6].copy()
    pl = int(g[0]); ax = [(0, 1, 2), (0, 2, 1), (1, 2, 0)][pl]
    lo = np.zeros(3); hi = np.zeros(3)
    lo[ax[0]], hi[ax[0]] = g[2], g[3]; lo[ax[1]], hi[ax[1]] = g[4], g[5]; lo[ax[2]], hi[ax[2]] = g[1] - 0.01, g[1] + 0.01
    return np.concatenate([lo, hi])

boxes = np.array([leaf_box(i) for i in range(n)])

def prim_hit(i, o, d, t0, t1):
    g = geom[i]
    if kinds[i] == 0:
        oc = o - g[0:3]; a = d @ d; hb = oc @ d; c = oc @ oc - g[3] * g[3]
        disc = hb * hb - a * c
        if disc <= 0: return None
        r = np.sqrt(disc)
        for t in ((-hb - r) / a, (-hb + r) / a):
            if t0 <= t < t1: return t
        return None
    if kinds[i] == 2:
        p0, p1, p2 = g[0:3], g[3:6], g[6:9]
        nn = np.cross(p1 - p0, p2 - p0); ln = np.linalg.norm(nn)
        if ln == 0: return None
        nn /= ln
        dn = d @ nn
        if abs(dn) <= 1e-4: return None
        t = ((p0 - o) @ nn) / dn
        if not (t0 <= t < t1): return None
        p = o + t * d
        # barycentric
        v0, v1, v2 = p1 - p0, p2 - p0, p - p0
        d00, d01, d11, d20, d21 = v0 @ v0, v0 @ v1, v1 @ v1, v2 @ v0, v2 @ v1
        den = d00 * d11 - d01 * d01
        if den == 0: return None
        w1 = (d11 * d20 - d01 * d21) / den; w2 = (d00 * d21 - d01 * d20) / den
        if w1 > 0 and w2 > 0 and w1 + w2 < 1: return t
        return None
    b = boxes[i]
    with np.errstate(divide="ignore", invalid="ignore"):
        ta = (b[0:3] - o) / d; tb = (b[3:6] - o) / d
    tn = np.nanmax(np.minimum(ta, tb)); tf = np.nanmin(np.maximum(ta, tb))
    if tn <= tf and t0 <= tn < t1: return tn
    return None

# ---------------------------------------------------------------- trees: arrays of (box, left, right) with negative = ~leaf
EPS = 5.9604645e-8
leafA = np.zeros(n); leafC = np.zeros(n)
for i in range(n):
    maxabs = np.abs(boxes[i]).max()
    if kinds[i] == 0:
        leafA[i] = 8 * EPS * (maxabs + 3 * geom[i][3]); leafC[i] = 8 * 3 * EPS / geom[i][3]
    else:
        leafA[i] = 8 * EPS * maxabs
GUARD = os.environ.get("GUARD", "0") == "1"
SIGN_REF = os.environ.get("SIGN_REF", "0") == "1"
ref_axis = {}

def ref_tree():
    nodes = []
    for j in range(len(fs.nodes_lra)):
        l, r, ax = fs.nodes_lra[j]
        ref_axis[j] = int(ax)
        enc = lambda c: ~int(c & 0x7FFFFFFF) if c & 0x80000000 else int(c)
        nodes.append((fs.nodes_box[j].astype(np.float64), enc(int(l)), enc(int(r))))
    root = fs.root
    return nodes, (~int(root & 0x7FFFFFFF) if root & 0x80000000 else int(root))

def area(b):
    e = np.maximum(b[3:6] - b[0:3], 0)
    return 2 * (e[0] * e[1] + e[1] * e[2] + e[2] * e[0])

def sah_tree():
    nodes = []
    cen = (boxes[:, 0:3] + boxes[:, 3:6]) / 2
    def build(idx):
        if len(idx) == 1:
            return ~int(idx[0])
        best = (np.inf, None, None)
        for ax in range(3):
            order = idx[np.argsort(cen[idx, ax], kind="stable")]
            lo = np.minimum.accumulate(boxes[order, 0:3], axis=0); hi = np.maximum.accumulate(boxes[order, 3:6], axis=0)
            rlo = np.minimum.accumulate(boxes[order[::-1], 0:3], axis=0)[::-1]; rhi = np.maximum.accumulate(boxes[order[::-1], 3:6], axis=0)[::-1]
            def ar(l, h):
                e = np.maximum(h - l, 0); return 2 * (e[:, 0] * e[:, 1] + e[:, 1] * e[:, 2] + e[:, 2] * e[:, 0])
            k = np.arange(1, len(order))
            cost = ar(lo[:-1], hi[:-1]) * k + ar(rlo[1:], rhi[1:]) * (len(order) - k)
            j = int(np.argmin(cost))
            if cost[j] < best[0]:
                best = (cost[j], order, j + 1)
        _, order, k = best
        me = len(nodes); nodes.append(None)
        l = build(order[:k]); r = build(order[k:])
        b = np.concatenate([boxes[order, 0:3].min(0), boxes[order, 3:6].max(0)])
        nodes[me] = (b, l, r)
        return me
    sys.setrecursionlimit(100000)
    root = build(np.arange(n))
    return nodes, root

def slab(b, o, inv, t0, t1):
    ta = (b[0:3] - o) * inv; tb = (b[3:6] - o) * inv
    tn = max(t0, np.nanmax(np.minimum(ta, tb))); tf = min(t1, np.nanmin(np.maximum(ta, tb)))
    return tn, tf

_margin_cache = {}
def margins(tree):
    key = id(tree[0])
    if key in _margin_cache: return _margin_cache[key]
    nodes, root = tree
    A = [0.0] * len(nodes); C = [0.0] * len(nodes)
    def rec(c):
        if c < 0: return leafA[~c], leafC[~c]
        a1, c1 = rec(nodes[c][1]); a2, c2 = rec(nodes[c][2])
        A[c], C[c] = max(a1, a2), max(c1, c2)
        return A[c], C[c]
    rec(root)
    _margin_cache[key] = (A, C)
    return A, C

def traverse(tree, o, d, t0=1e-3, t1=np.inf, sign_order=False):
    nodes, root = tree
    mA, mC = margins(tree)
    with np.errstate(divide="ignore"):
        inv = 1.0 / d
    mscale = 2 * np.abs(inv).max()
    best, bt = -1, t1
    nv = npt = 0
    def box_of(c):
        return boxes[~c] if c < 0 else nodes[c][0]
    stack = [(root, t0)]
    while stack:
        c, tn = stack.pop()
        if tn >= bt: continue
        if c < 0:
            npt += 1
            t = prim_hit(~c, o, d, t0, bt)
            if t is not None: best, bt = ~c, t
            continue
        nv += 1
        _, l, r = nodes[c]
        kids = []
        for ch in (l, r):
            if ch < 0:
                # leaves are not box-tested in the reference; with child boxes in the parent the GPU culls them by their own box
                tn_c, tf_c = slab(boxes[~ch], o, inv, t0, bt)
            else:
                tn_c, tf_c = slab(nodes[ch][0], o, inv, t0, bt)
            mg = (mA[c] + tn_c * (1e-5 + mC[c] * tn_c)) * mscale if GUARD else 0.0
            if not (tn_c - tf_c > mg): kids.append((ch, tn_c - mg))
        if sign_order:
            if d[ref_axis[c]] > 0: kids = kids[::-1]      # left first => left must be popped last pushed
        else:
            kids.sort(key=lambda x: -x[1])
        stack.extend(kids)
    return best, bt, nv, npt

rng = np.random.default_rng(0)
# primary-like rays: from the camera position towards random points of the element boxes
cam_pos = fs.cam[0:3].astype(np.float64)
tgt = boxes[rng.integers(0, n, NR)]
tgt = tgt[:, :3] + rng.uniform(size=(NR, 3)) * (tgt[:, 3:] - tgt[:, :3])
pd = tgt - cam_pos; pd /= np.linalg.norm(pd, axis=1, keepdims=True)
po = np.tile(cam_pos, (NR, 1))
T_ref, T_sah = ref_tree(), sah_tree()
def sah_cost(tree):
    nodes, root = tree
    ra = area(nodes[root][0])
    return sum(area(b) for b, _, _ in nodes) / ra
print("SAH cost (sum node area / root area): ref %.2f  sah %.2f" % (sah_cost(T_ref), sah_cost(T_sah)))
gen = [(po, pd)]
for bounce in range(3):
    o, d = gen[-1]
    no, nd = [], []
    stats = {"ref": [0, 0], "sah": [0, 0]}
    mism = 0
    for i in range(len(o)):
        a = traverse(T_ref, o[i], d[i], sign_order=SIGN_REF); b = traverse(T_sah, o[i], d[i])
        stats["ref"][0] += a[2]; stats["ref"][1] += a[3]; stats["sah"][0] += b[2]; stats["sah"][1] += b[3]
        mism += a[0] != b[0]
        if a[0] >= 0:
            p = o[i] + a[1] * d[i]
            v = rng.normal(size=3); v /= np.linalg.norm(v)
            no.append(p); nd.append(v)
    m = len(o)
    print("gen %d: %d rays, hit %.2f | ref nodes %.1f prims %.1f | sah nodes %.1f prims %.1f | id mismatches %d" % (
        bounce, m, len(no) / m, stats["ref"][0] / m, stats["ref"][1] / m, stats["sah"][0] / m, stats["sah"][1] / m, mism))
    if not no: break
    gen.append((np.array(no), np.array(nd)))
