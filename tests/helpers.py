"""Shared test helpers: synthetic scenes, ray sets (primary / random / adversarial) and comparisons."""
from __future__ import annotations

import numpy as np

F32 = np.float32
MISS = 0xFFFFFFFF


def f2u(x):
    return np.ascontiguousarray(x, F32).view(np.uint32)


def unit_rows(v):
    v = np.asarray(v, np.float64)
    return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(F32)


def simple_camera(pos=(0.0, 1.0, 6.0), at=(0.0, 0.0, 0.0), fov=50.0, aspect=0.75, lens_radius=0.0, focus=1.0):
    """Camera fields as Camera::build()...build() computes them (float32 numpy restatement for test scenes)."""
    pos, at = np.array(pos, F32), np.array(at, F32)
    up = np.array([0, 1, 0], F32)
    fwd = at - pos
    h = F32(np.tan(F32(fov) * F32(np.pi / 180.0) * F32(0.5)))
    vw, vh = F32(2.0) * h, F32(2.0) * h * F32(aspect)

    def unit(a):
        return a * (F32(1.0) / F32(np.sqrt(np.dot(a, a).astype(F32))))
    right = unit(np.cross(fwd, up).astype(F32))
    upv = unit(np.cross(right, fwd).astype(F32))
    sfwd = fwd * (F32(focus) / F32(np.sqrt(np.dot(fwd, fwd).astype(F32))))
    ulc = F32(focus) * (right * (vw * F32(-0.5)) + upv * (vh * F32(0.5))) + sfwd
    cam = np.zeros(23, F32)
    cam[0:3], cam[3:6], cam[6:9], cam[9:12] = pos, ulc, right, upv
    cam[12:15], cam[15:18] = right * (F32(focus) * vw), upv * (F32(focus) * vh)
    cam[18] = lens_radius
    return cam


def solid_tex(rgb):
    """[kind, image, r, g, b bits, inv_frequency bits, even, odd]"""
    return [0, MISS] + list(np.array(rgb, F32).view(np.uint32)) + [0, MISS, MISS]


def checker_tex(inv_frequency, even, odd):
    return [2, MISS, 0, 0, 0, int(np.array([inv_frequency], F32).view(np.uint32)[0]), even, odd]


def marble_tex(scale, perlin):
    return [3, perlin, 0, 0, 0, int(np.array([scale], F32).view(np.uint32)[0]), MISS, MISS]


def random_perlin(rng):
    """One Perlin table row: 256 unit gradients (float bits) + three permutations of 0..255 (perlin.rs:19-36)."""
    g = rng.normal(size=(256, 3))
    g = (g / np.linalg.norm(g, axis=1, keepdims=True)).astype(F32)
    return np.concatenate([g.reshape(-1).view(np.uint32)] + [rng.permutation(256).astype(np.uint32) for _ in range(3)])


def mat(kind, tex, param=0.0):
    return [kind, tex, int(np.array([param], F32).view(np.uint32)[0])]


def random_scene(seed, n_spheres=20, n_tris=30, n_rects=4, extent=4.0, with_light=False, image_tex=None, n_boxes=0, wrap_fraction=0.0,
                 checker=False, volume_fraction=0.0, marble=False):
    """Random mixed scene in the flat element layout (kinds, item_mat, geom, mats, texs, images, cam, bg, light[, wrap, perlins]).
    n_boxes / wrap_fraction / checker exercise row f2: AxisAlignedBox primitives, run-time Transformation / Animation wrappers
    (with the reference's untransformed bounding boxes, Q2) and Checker textures; volume_fraction / marble exercise row f3:
    constant-density volumes (any boundary geometry, Isotropic and other phase functions) and Marble textures."""
    rng = np.random.default_rng(seed)
    kinds, geom = [], []
    for _ in range(n_boxes):
        g = np.zeros(24, F32)
        lo = rng.uniform(-extent, extent, 3)
        g[0:3] = lo; g[3:6] = lo + rng.uniform(0.2, 1.5, 3)
        kinds.append(3); geom.append(g)
    for _ in range(n_spheres):
        g = np.zeros(24, F32)
        g[0:3] = rng.uniform(-extent, extent, 3)
        g[3] = rng.uniform(0.1, 0.8)
        kinds.append(0); geom.append(g)
    for _ in range(n_rects):
        g = np.zeros(24, F32)
        g[0] = rng.integers(0, 3)
        g[1] = rng.uniform(-extent, extent)
        a, b = np.sort(rng.uniform(-extent, extent, 2)); c, d = np.sort(rng.uniform(-extent, extent, 2))
        g[2:6] = [a, b + 0.1, c, d + 0.1]
        kinds.append(1); geom.append(g)
    for _ in range(n_tris):
        g = np.zeros(24, F32)
        p0 = rng.uniform(-extent, extent, 3)
        g[0:3] = p0; g[3:6] = p0 + rng.uniform(-1, 1, 3); g[6:9] = p0 + rng.uniform(-1, 1, 3)
        n = np.cross(g[3:6] - g[0:3], g[6:9] - g[0:3]); n = n / (np.linalg.norm(n) + 1e-20)
        for k in range(3):
            g[9 + 3 * k: 12 + 3 * k] = n + rng.normal(0, 0.1, 3)   # non-unit, varying vertex normals (Q8)
        g[18:24] = rng.uniform(0, 1, 6)
        kinds.append(2); geom.append(g)
    order = rng.permutation(len(kinds))
    kinds = np.array(kinds, np.uint32)[order]
    geom = np.array(geom, F32)[order]
    texs = [solid_tex(rng.uniform(0.2, 0.9, 3)) for _ in range(4)] + [solid_tex([4.0, 4.0, 4.0])]
    images = []
    if image_tex is not None:
        texs.append([1, 0, 0, 0, 0, 0, MISS, MISS]); images.append(image_tex)
    mats = [mat(0, 0), mat(0, 1), mat(1, 2, 0.0), mat(1, 3, 0.3), mat(2, MISS, 1.5), mat(3, 4)]
    if image_tex is not None:
        mats.append(mat(0, 5))
    if checker:
        texs.append(checker_tex(2.5, 0, 1)); texs.append(checker_tex(0.7, len(texs) - 1, 2))   # nested checker
        mats.append(mat(0, len(texs) - 1)); mats.append(mat(1, len(texs) - 2, 0.2))
    perlins = []
    if marble:
        perlins = [random_perlin(rng), random_perlin(rng)]
        texs.append(marble_tex(4.0, 0)); texs.append(marble_tex(0.6, 1))
        mats.append(mat(0, len(texs) - 2)); mats.append(mat(1, len(texs) - 1, 0.1))
        if checker:
            texs.append(checker_tex(1.3, len(texs) - 1, 0)); mats.append(mat(0, len(texs) - 1))
    item_mat = rng.integers(0, len(mats), len(kinds)).astype(np.uint32)
    iso = []
    if volume_fraction > 0:
        iso = [len(mats), len(mats) + 1]
        mats.append(mat(4, 0)); mats.append(mat(4, len(texs) - 1))   # Isotropic over a solid and over the last texture
    cam = simple_camera(pos=(0.5, 1.0, 3.0 * extent), at=(0, 0, 0), fov=50.0, aspect=0.75)
    if wrap_fraction > 0:
        cam[19], cam[20] = 0.0, 1.0      # motion_blur(0, 1)
    bg = np.array([1, 0, 0, 0], F32)
    light = np.zeros(7, F32)
    if with_light:
        ridx = np.nonzero(kinds == 1)[0]
        if len(ridx):
            g = geom[ridx[0]]
            item_mat[ridx[0]] = 5
            light = np.array([1, g[0], g[1], g[2], g[3], g[4], g[5]], F32)
    out = dict(kinds=kinds, item_mat=item_mat, geom=geom, mats=np.array(mats, np.uint32), texs=np.array(texs, np.uint32),
               images=images, cam=cam, bg=bg, light=light)
    if marble:
        out["perlins"] = np.array(perlins, np.uint32)
    if wrap_fraction > 0 or volume_fraction > 0:
        n = len(kinds)
        wrap = np.zeros((n, 10), F32)
        wrap[:, 5] = 1.0                                   # y_cosine of the identity
        sel = rng.uniform(size=n) < wrap_fraction
        if with_light:
            sel[np.nonzero(kinds == 1)[0][:1]] = False     # the light-sampling rect must stay un-wrapped (world_builder.rs:317)
        for i in np.nonzero(sel)[0]:
            fl = rng.integers(1, 4)
            if kinds[i] == 2:
                fl = 2                                     # triangles are always fully baked: only Animation can wrap them
            ang = rng.uniform(-1.0, 1.0)
            if fl & 1:
                wrap[i, 1:4] = rng.uniform(-0.5, 0.5, 3); wrap[i, 4] = np.sin(ang); wrap[i, 5] = np.cos(ang)
            if fl & 2:
                wrap[i, 6:9] = rng.uniform(-0.8, 0.8, 3)
            wrap[i, 0] = fl
        vol = rng.uniform(size=n) < volume_fraction
        if with_light:
            vol[np.nonzero(kinds == 1)[0][:1]] = False
        for i in np.nonzero(vol)[0]:
            wrap[i, 0] += 4                                # RT_WRAP_VOLUME: the geometry is the boundary of a constant medium
            wrap[i, 9] = -1.0 / rng.uniform(0.2, 3.0)      # -1 / density
            if rng.uniform() < 0.8:
                item_mat[i] = iso[int(rng.integers(0, 2))] # mostly Isotropic; the rest keep an ordinary material as phase function
        out["wrap"] = wrap
    return out


def scene_bounds(fs):
    if len(fs.nodes_box):
        lo, hi = fs.nodes_box[0, :3].astype(np.float64), fs.nodes_box[0, 3:].astype(np.float64)
    else:
        lo, hi = np.array([-2.0, -2, -2]), np.array([2.0, 2, 2])
    return np.maximum(lo, -60.0), np.minimum(hi, 60.0)


def random_rays(fs, n, seed):
    rng = np.random.default_rng(seed)
    lo, hi = scene_bounds(fs)
    pad = 0.25 * (hi - lo)
    org = rng.uniform(lo - pad, hi + pad, size=(n, 3)).astype(F32)
    return org, unit_rows(rng.normal(size=(n, 3)))


def adversarial_rays(fs, n_per_class, seed):
    """Rays aimed at the places where acceptance rules and ties are decided: triangle vertices / edges, sphere limbs,
    rect borders and corners, exact axis-parallel directions (zero components -> inf / NaN slabs), origins lying exactly
    on node-box planes, origins on surfaces."""
    rng = np.random.default_rng(seed)
    lo, hi = scene_bounds(fs)
    pad = 0.5 * (hi - lo)
    orgs, dirs = [], []

    def aim(targets):
        targets = np.asarray(targets, np.float64)
        if len(targets) == 0:
            return
        idx = rng.integers(0, len(targets), n_per_class)
        o = rng.uniform(lo - pad, hi + pad, size=(n_per_class, 3))
        d = targets[idx] - o
        ok = np.linalg.norm(d, axis=1) > 1e-6
        orgs.append(o[ok].astype(F32)); dirs.append(unit_rows(d[ok]))

    tri = fs.geom[fs.kinds == 2]
    if len(tri):
        p = tri[:, :9].reshape(-1, 3, 3).astype(np.float64)
        aim(p.reshape(-1, 3))                                    # vertices
        aim(0.5 * (p[:, 0] + p[:, 1])); aim(0.5 * (p[:, 1] + p[:, 2]))  # edge midpoints
        w = rng.dirichlet([1, 1, 1], len(p))
        aim((p * w[:, :, None]).sum(1))                          # interior
    sph = fs.geom[fs.kinds == 0]
    if len(sph):
        c, r = sph[:, :3].astype(np.float64), sph[:, 3:4].astype(np.float64)
        u = rng.normal(size=c.shape); u /= np.linalg.norm(u, axis=1, keepdims=True)
        aim(c + r * u)                                           # surface points (limb for some origins)
        # exact tangents: origin o, tangent point t with (t-c).(t-o) = 0
        idx = rng.integers(0, len(c), n_per_class)
        o = rng.uniform(lo - pad, hi + pad, size=(n_per_class, 3))
        oc = o - c[idx]; dist = np.linalg.norm(oc, axis=1, keepdims=True)
        ok = (dist[:, 0] > r[idx, 0] * 1.01)
        w_ = oc / dist
        v = np.cross(w_, rng.normal(size=w_.shape)); v /= np.linalg.norm(v, axis=1, keepdims=True) + 1e-30
        cosb = r[idx] / dist
        tpt = c[idx] + r[idx] * (cosb * w_ + np.sqrt(np.maximum(0, 1 - cosb ** 2)) * v)
        d = tpt - o
        orgs.append(o[ok].astype(F32)); dirs.append(unit_rows(d[ok]))
    rect = fs.geom[fs.kinds == 1]
    if len(rect):
        pts = []
        for g in rect:
            plane = int(g[0]); a0, a1, n = [(0, 1, 2), (0, 2, 1), (1, 2, 0)][plane]
            for u_ in (g[2], g[3], 0.5 * (g[2] + g[3])):
                for v_ in (g[4], g[5], 0.5 * (g[4] + g[5])):
                    q = np.zeros(3); q[a0], q[a1], q[n] = u_, v_, g[1]
                    pts.append(q)
        aim(pts)
    # exact axis-parallel rays (zero direction components)
    for ax in range(3):
        for sgn in (-1.0, 1.0):
            o = rng.uniform(lo - pad, hi + pad, size=(n_per_class // 4 + 1, 3)).astype(F32)
            d = np.zeros_like(o); d[:, ax] = sgn
            orgs.append(o); dirs.append(d.astype(F32))
    # origins exactly on node-box planes with a zero direction component on that axis (0/0 = NaN slabs)
    if len(fs.nodes_box):
        m = n_per_class // 2 + 1
        idx = rng.integers(0, len(fs.nodes_box), m)
        ax = rng.integers(0, 3, m); side = rng.integers(0, 2, m)
        o = rng.uniform(lo, hi, size=(m, 3)).astype(F32)
        o[np.arange(m), ax] = fs.nodes_box[idx, ax + 3 * side]
        d = rng.normal(size=(m, 3)); d[np.arange(m), ax] = 0.0
        orgs.append(o); dirs.append(unit_rows(d + 1e-30))
        dirs[-1][np.arange(m), ax] = 0.0
    # rays sliding along slab / primitive boundaries: origin on a box plane, rect border or sphere-AABB face, with a zero or
    # tiny direction component on that axis (the hit lands within rounding distance of the boundary: "phantom" accepts)
    planes = []   # (axis, coordinate)
    if len(fs.nodes_box):
        for ax in range(3):
            planes += [(ax, v) for v in fs.nodes_box[:, ax]] + [(ax, v) for v in fs.nodes_box[:, 3 + ax]]
    for g in rect:
        plane = int(g[0]); a0, a1, n = [(0, 1, 2), (0, 2, 1), (1, 2, 0)][plane]
        planes += [(a0, g[2]), (a0, g[3]), (a1, g[4]), (a1, g[5]), (n, g[1])]
    for g in sph[:200]:
        for ax in range(3):
            planes += [(ax, g[ax] - g[3]), (ax, g[ax] + g[3])]
    if planes:
        m = n_per_class
        pick = rng.integers(0, len(planes), m)
        o = rng.uniform(lo, hi, size=(m, 3)).astype(F32)
        d = rng.normal(size=(m, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        tiny = rng.choice([0.0, 1e-12, -1e-12, 1e-9, -1e-9, 1e-7, -1e-7, 1e-5, -1e-5, 1e-3, -1e-3], m)
        for i in range(m):
            ax, v = planes[pick[i]]
            o[i, ax] = v
            d[i, ax] = 0.0
            d[i] /= max(np.linalg.norm(d[i]), 1e-30)
            d[i, ax] = tiny[i]
        orgs.append(o); dirs.append(d.astype(F32))
    return np.concatenate(orgs), np.concatenate(dirs)


def _uv_diff(a, b):
    """max |a-b| over finite entries; inf when the NaN masks differ (acos of a rounded |n.y| > 1 is NaN on both sides)."""
    if a.size == 0:
        return 0.0
    na, nb = np.isnan(a), np.isnan(b)
    if not np.array_equal(na, nb):
        return float("inf")
    return float(np.abs(a[~na] - b[~nb]).max()) if (~na).any() else 0.0


def compare_hits(got, ref, rel_tol=1e-5):
    """Hit ids bit-exact; t and normals within rel_tol relative (the north-star bound); returns a dict of counts."""
    ids_bad = int((got["hit_id"] != ref["hit_id"]).sum())
    both = (ref["hit_id"] != MISS) & (got["hit_id"] == ref["hit_id"])
    t_bits = int((f2u(got["t"][both]) != f2u(ref["t"][both])).sum())
    n_bits = int((f2u(got["normal"][both]) != f2u(ref["normal"][both])).any(axis=1).sum())
    with np.errstate(invalid="ignore", divide="ignore"):
        t_rel = np.abs(got["t"][both] - ref["t"][both]) / np.maximum(np.abs(ref["t"][both]), 1e-30)
        nscale = np.maximum(np.linalg.norm(ref["normal"][both], axis=1, keepdims=True), 1e-30)
        n_rel = np.abs(got["normal"][both] - ref["normal"][both]) / nscale
    return dict(n=len(ref["hit_id"]), hits=int(both.sum()), id_mismatch=ids_bad, t_bit_mismatch=t_bits, normal_bit_mismatch=n_bits,
                t_rel_max=float(t_rel.max()) if both.any() else 0.0, n_rel_max=float(n_rel.max()) if both.any() else 0.0,
                ff_mismatch=int((got["front_face"][both] != ref["front_face"][both]).sum()),
                uv_abs_max=_uv_diff(got["uv"][both], ref["uv"][both]))


def assert_hits_match(got, ref, rel_tol=1e-5):
    r = compare_hits(got, ref, rel_tol)
    assert r["id_mismatch"] == 0, r
    assert r["t_rel_max"] <= rel_tol and r["n_rel_max"] <= rel_tol, r
    assert r["ff_mismatch"] == 0, r
    assert r["uv_abs_max"] <= 2e-6, r   # device acosf / atan2f vs glibc: a few ulp of uv in [0, 1]
    return r


def images_equal_bits(a, b):
    a, b = np.ascontiguousarray(a, F32), np.ascontiguousarray(b, F32)
    na, nb = np.isnan(a), np.isnan(b)
    return bool(np.array_equal(na, nb) and np.array_equal(a.view(np.uint32)[~na], b.view(np.uint32)[~nb]))
