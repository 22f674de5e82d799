"""ORACLE — TEST INFRASTRUCTURE ONLY.  A SECOND, independent restatement of the functions that decide hit ids.

Written in numpy float32 from the Rust sources alone (not from oracle_core.hpp), to catch a single-author transcription slip
in the C++ oracle ("Missing" #1).  Every numpy float32 add / sub / mul / div / sqrt is one IEEE-rounded operation, in
the operand order of the Rust, so the results must agree with the C++ oracle BIT FOR BIT.  Covered, with the reference lines:

  build_bvh        BoundingVolumeHierarchy::new / _new        src/lib/hittable.rs:360-427  (stable sort by aabb.min[axis],
                                                              mid = len / 2, child axis = (axis + len(child)) % 3, pre-order ids)
  element_box      Geometry::bounding_box                     src/lib/hittable.rs:150-157, aabb.rs:20-25, rect_geometry.rs:87-98,
                                                              triangle_geometry.rs:47-49
  hit_cond         Aabb::hit_cond (per-axis independent, Q1)  src/lib/aabb.rs:65-78, math.rs:35-41
  trace            hit_recursive / hit_index_list / hit       src/lib/hittable.rs:429-473
  sphere / rect / triangle hits, GeoHitInteraction::new_from_ray
                                                              src/lib/hittable/{sphere,rect,triangle}_geometry.rs, hittable.rs:34-54

Vectorisation: the reference orders the two children of a node by `ray.direction[axis_id] > 0.0`, so all rays of one direction-sign
octant visit the tree in the same order.  Each octant's rays walk the tree together (depth-first, in that octant's child order) with a
per-ray `t_range.end`; a ray takes part in a subtree iff every ancestor passed hit_cond for it — exactly the recursion, ray by ray.
"""
from __future__ import annotations

import sys

import numpy as np

F = np.float32
LEAF = 1 << 62  # leaf reference = LEAF + item index (the Rust uses usize::MAX - index)


def element_box(kind, g):
    """Geometry::bounding_box for one element (g: 24 floats of the flat layout: sphere c,r | rect plane,dist,r0,r1 | triangle p0,p1,p2)."""
    g = np.asarray(g, F)
    if kind == 0:  # Aabb::new_radius: center -/+ (r, r, r)
        return g[0:3] - g[3], g[0:3] + g[3]
    if kind == 1:  # RectGeometry::bounding_box(0.01)
        p0, p1, n = [(0, 1, 2), (0, 2, 1), (1, 2, 0)][int(g[0])]
        mn, mx = np.zeros(3, F), np.zeros(3, F)
        mn[p0], mn[p1], mn[n] = g[2], g[4], g[1] - F(0.01)
        mx[p0], mx[p1], mx[n] = g[3], g[5], g[1] + F(0.01)
        return mn, mx
    if kind == 2:  # Aabb::new_surrounding_points
        p = g[0:9].reshape(3, 3)
        return p.min(axis=0), p.max(axis=0)
    if kind == 3:  # Geometry::AxisAlignedBox: the box itself
        return g[0:3].copy(), g[3:6].copy()
    raise ValueError(kind)


def build_bvh(mins, maxs):
    """hittable.rs:382-427.  Returns (nodes, root): nodes[i] = (min[3], max[3], axis, left, right), children = node index or LEAF + item."""
    nodes = []
    sys.setrecursionlimit(max(10000, sys.getrecursionlimit()))

    def new(h, axis):  # h: list of (ref, min, max)
        if len(h) == 1:
            return h[0]
        h.sort(key=lambda e: e[1][axis])  # slice::sort_by is stable, like list.sort
        if len(h) == 2:
            me = len(nodes)
            mn, mx = np.minimum(h[0][1], h[1][1]), np.maximum(h[0][2], h[1][2])
            nodes.append((mn, mx, axis, h[0][0], h[1][0]))
            return me, mn, mx
        mid = len(h) // 2
        left, right = h[:mid], h[mid:]
        me = len(nodes)
        nodes.append(None)  # the slot is reserved BEFORE the children are built: pre-order numbering
        l = new(left, (axis + len(left)) % 3)
        r = new(right, (axis + len(right)) % 3)
        mn, mx = np.minimum(l[1], r[1]), np.maximum(l[2], r[2])
        nodes[me] = (mn, mx, axis, l[0], r[0])
        return me, mn, mx

    h = [(LEAF + i, mins[i], maxs[i]) for i in range(len(mins))]
    root = new(h, 0)
    return nodes, root[0]


def hit_cond(mn, mx, O, D, start, end):
    """aabb.rs:65-78 for a batch of rays: each axis tested on its own against (start, end); math::minmax by `<`; f32::max / min ignore NaN."""
    ok = np.ones(len(O), bool)
    with np.errstate(all="ignore"):
        for i in range(3):
            a = (mn[i] - O[:, i]) / D[:, i]
            b = (mx[i] - O[:, i]) / D[:, i]
            lt = a < b
            t0, t1 = np.where(lt, a, b), np.where(lt, b, a)
            t_min, t_max = np.fmax(t0, start), np.fmin(t1, end)
            ok &= ~(t_max <= t_min)
    return ok


def dot(ax, ay, az, bx, by, bz):
    return ax * bx + ay * by + az * bz  # Dir3::dot: (x*x + y*y) + z*z, vec3.rs:189-191


def cross(a, b):
    return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def contains(start, end, t):
    return (start <= t) & (t < end)  # Range::contains, half open


def hit_sphere(g, O, D, start, end):
    """sphere_geometry.rs:21-43 -> (hit mask, t)."""
    with np.errstate(all="ignore"):
        ocx, ocy, ocz = O[:, 0] - g[0], O[:, 1] - g[1], O[:, 2] - g[2]
        half_b = dot(ocx, ocy, ocz, D[:, 0], D[:, 1], D[:, 2])
        c = dot(ocx, ocy, ocz, ocx, ocy, ocz) - g[3] * g[3]
        disc = half_b * half_b - c
        valid = ~(disc < F(0.0))
        sq = np.sqrt(np.where(valid, disc, F(0.0)))
        small, large = -half_b - sq, -half_b + sq
        in_s, in_l = contains(start, end, small), contains(start, end, large)
        return valid & (in_s | in_l), np.where(in_s, small, large)


def hit_rect(g, O, D, start, end):
    """rect_geometry.rs:33-42."""
    p0, p1, n = [(0, 1, 2), (0, 2, 1), (1, 2, 0)][int(g[0])]
    with np.errstate(all="ignore"):
        t = (g[1] - O[:, n]) / D[:, n]
        a, b = O[:, p0] + t * D[:, p0], O[:, p1] + t * D[:, p1]
        inside = (a >= g[2]) & (a <= g[3]) & (b >= g[4]) & (b <= g[5])
        return contains(start, end, t) & inside, t


def hit_triangle(g, O, D, start, end):
    """triangle_geometry.rs:13-31 -> (hit mask, t, w1, w2).  The per-triangle values are recomputed from the vertices, as the Rust does."""
    p0, p1, p2 = g[0:3], g[3:6], g[6:9]
    dir1, dir2 = p1 - p0, p2 - p0
    with np.errstate(all="ignore"):
        c = cross(dir1, dir2)
        ln = np.sqrt(dot(c[0], c[1], c[2], c[0], c[1], c[2]))
        s = F(1.0) / ln  # Dir3::unit = self * (1.0 / length)
        nx, ny, nz = c[0] * s, c[1] * s, c[2] * s
        denom = dot(D[:, 0], D[:, 1], D[:, 2], nx, ny, nz)
        ok = np.abs(denom) > F(0.0001)
        t = dot(p0[0] - O[:, 0], p0[1] - O[:, 1], p0[2] - O[:, 2], nx, ny, nz) / denom
        ok &= contains(start, end, t)
        qx, qy, qz = (O[:, 0] + t * D[:, 0]) - p0[0], (O[:, 1] + t * D[:, 1]) - p0[1], (O[:, 2] + t * D[:, 2]) - p0[2]
        v1 = cross((nx, ny, nz), dir2)
        w1 = dot(qx, qy, qz, v1[0], v1[1], v1[2]) / dot(dir1[0], dir1[1], dir1[2], v1[0], v1[1], v1[2])
        ok &= (w1 > F(0.0)) & (w1 < F(1.0))
        v2 = cross((nx, ny, nz), dir1)
        w2 = dot(qx, qy, qz, v2[0], v2[1], v2[2]) / dot(dir2[0], dir2[1], dir2[2], v2[0], v2[1], v2[2])
        w0 = F(1.0) - w1 - w2
        ok &= (w2 > F(0.0)) & (w0 > F(0.0))
        return ok, t, w1, w2


def trace(kinds, geom, nodes, root, org, dirs, t_min=0.001, t_max=np.inf):
    """BoundingVolumeHierarchy::hit for a batch of rays -> dict(hit_id, t, normal, front_face); hit_id 0xFFFFFFFF = miss."""
    O = np.ascontiguousarray(org, F).reshape(-1, 3)
    D = np.ascontiguousarray(dirs, F).reshape(-1, 3)
    geom = np.ascontiguousarray(geom, F).reshape(len(kinds), 24)
    n = len(O)
    start = F(t_min)
    end = np.full(n, t_max, F)
    best = np.full(n, 0xFFFFFFFF, np.uint32)
    bw1, bw2 = np.zeros(n, F), np.zeros(n, F)
    pos_sign = D > F(0.0)

    def leaf(item, idx):
        k, g = int(kinds[item]), geom[item]
        if k == 0:
            m, t = hit_sphere(g, O[idx], D[idx], start, end[idx])
        elif k == 1:
            m, t = hit_rect(g, O[idx], D[idx], start, end[idx])
        elif k == 2:
            m, t, w1, w2 = hit_triangle(g, O[idx], D[idx], start, end[idx])
            bw1[idx[m]], bw2[idx[m]] = w1[m], w2[m]
        else:
            raise NotImplementedError("AxisAlignedBox is not part of the north-star primitive set")
        h = idx[m]
        end[h] = t[m]      # hit_index_list: t_range.end = hi.t
        best[h] = item

    def visit(ref, idx, octant):
        if idx.size == 0:
            return
        if ref >= LEAF:    # a leaf is never box-tested (hittable.rs:436-437)
            leaf(ref - LEAF, idx)
            return
        mn, mx, axis, left, right = nodes[ref]
        idx = idx[hit_cond(mn, mx, O[idx], D[idx], start, end[idx])]
        first, second = (left, right) if (octant >> axis) & 1 else (right, left)  # ray.direction[axis_id] > 0.0
        visit(first, idx, octant)
        visit(second, idx, octant)

    code = pos_sign[:, 0].astype(np.int64) | (pos_sign[:, 1].astype(np.int64) << 1) | (pos_sign[:, 2].astype(np.int64) << 2)
    for octant in range(8):
        visit(root, np.nonzero(code == octant)[0], octant)

    # GeoHitInteraction of the final hit: position, surface normal, front_face flip (hittable.rs:34-54)
    nrm = np.zeros((n, 3), F)
    ff = np.zeros(n, np.uint8)
    hit = best != 0xFFFFFFFF
    with np.errstate(all="ignore"):
        for item in np.unique(best[hit]):
            idx = np.nonzero(best == item)[0]
            k, g = int(kinds[item]), geom[item]
            t = end[idx]
            P = np.stack([O[idx, i] + t * D[idx, i] for i in range(3)], axis=1)
            if k == 0:
                sn = np.stack([(P[:, i] - g[i]) / g[3] for i in range(3)], axis=1)
            elif k == 1:
                sn = np.zeros((len(idx), 3), F)
                sn[:, [(0, 1, 2), (0, 2, 1), (1, 2, 0)][int(g[0])][2]] = F(-1.0)
            else:  # math::interpolate(w0, w1, w2, normals): v0*w0 + v1*w1 + v2*w2, not re-normalised (Q8)
                w1, w2 = bw1[idx], bw2[idx]
                w0 = F(1.0) - w1 - w2
                sn = np.stack([g[9 + i] * w0 + g[12 + i] * w1 + g[15 + i] * w2 for i in range(3)], axis=1)
            front = dot(sn[:, 0], sn[:, 1], sn[:, 2], D[idx, 0], D[idx, 1], D[idx, 2]) < F(0.0)
            nrm[idx] = np.where(front[:, None], sn, -sn)
            ff[idx] = front
    t_out = np.where(hit, end, F(np.inf)).astype(F)
    return dict(hit_id=best, t=t_out, normal=nrm, front_face=ff)
