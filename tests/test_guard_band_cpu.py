"""The guard band of the tight cull (`guard_band` / `padded_cull` in rt_device.cuh), checked on the CPU against the
oracle's literal primitive tests: whenever the reference arithmetic ACCEPTS a hit for a ray that grazes an element, the ray's
carried slab interval on the element's own bounding box must not be empty by more than the band — otherwise the tight cull could
skip an element the reference hits.  One-element worlds (the root is the leaf, so the oracle applies no box test at all)."""
import math

import numpy as np
import pytest

from helpers import F32, MISS, mat, simple_camera, solid_tex

EPS = 5.9604645e-8


def bf16_up(v):
    b = np.array([v], F32).view(np.uint32)[0]
    if b & 0xFFFF:
        b = ((b | 0xFFFF) + 1) & 0xFFFFFFFF
    return np.array([b & 0xFFFF0000], np.uint32).view(F32)[0]


def band_checks(box, A, C, o, d, t_hit, start=0.001):
    """Returns (scalar-band cull?, padded cull?) for an accepted hit at t_hit, evaluated like the device does (end = +inf: the
    weakest situation, nothing has shrunk the range yet)."""
    lo, hi = box[:3].astype(np.float64), box[3:].astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / d.astype(np.float64)
        a = (lo - o) * inv; b = (hi - o) * inv
    l, h = np.minimum(a, b), np.maximum(a, b)
    tn, tf = max(np.nanmax(l), start), np.nanmin(h)
    A, C = float(bf16_up(A)), float(bf16_up(C))
    mscale = 2.0 * np.max(np.abs(inv))
    sA, sB, sC = A * mscale, 1e-5 * mscale, C * mscale
    tfp = max(tf, start)
    te = tfp + sA + tfp * (sB + sC * tfp)
    mg = sA + te * (sB + sC * te) if (sB + 2 * sC * te) <= 1.0 else math.inf
    scalar_cull = (tn - tf) > mg
    # per-axis padded test
    ia = np.abs(inv)
    te2 = math.inf
    for k in range(3):
        hp = max(h[k], start)
        tk = hp + 2 * ia[k] * (A + hp * (1e-5 + C * hp))
        if ia[k] * (1e-5 + 2 * C * tk) <= 0.5:
            te2 = min(te2, tk)
    if math.isinf(te2):
        padded_cull = False
    else:
        eps_e = A + te2 * (1e-5 + C * te2)
        padded_cull = max(np.max(l - eps_e * ia), start) > np.min(h + eps_e * ia)
    return scalar_cull, padded_cull, tn - tf, mg


def one_item_world(oracle, kind, geom):
    g = np.zeros((1, 24), F32); g[0, :len(geom)] = geom
    return oracle.OracleWorld.from_flat(np.array([kind], np.uint32), np.zeros(1, np.uint32), g, np.array([mat(0, 0)], np.uint32),
                                        np.array([solid_tex([1, 1, 1])], np.uint32), [], simple_camera(), np.array([0, 1, 1, 1], F32),
                                        np.zeros(7, F32))


def unit(v):
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


@pytest.mark.parametrize("seed", range(6))
def test_accepted_sphere_hits_are_never_culled(seed, oracle):
    rng = np.random.default_rng(seed)
    n_acc = n_outside = 0
    for _ in range(25):
        r = F32(10.0 ** rng.uniform(-2, 3))
        c = (rng.normal(size=3) * 10.0 ** rng.uniform(-1, 3)).astype(F32)
        W = one_item_world(oracle, 0, [c[0], c[1], c[2], r])
        n = 1500
        # rays that graze the sphere: aim at points at distance r (1 + delta) from the centre, |delta| from 1e-8 to 1e-3, both signs
        dist = (10.0 ** rng.uniform(math.log10(1.5), 4.5, n) * float(r)).astype(np.float64)    # 1.5 r .. 30 000 r away
        dirs = unit(rng.normal(size=(n, 3)))
        perp = unit(np.cross(dirs, rng.normal(size=(n, 3))))
        # half of the rays touch the sphere where it touches its box: perp = +-e_k, direction (nearly) perpendicular to e_k
        ax = rng.integers(0, 3, n)
        ek = np.eye(3)[ax] * rng.choice([-1, 1], n)[:, None]
        tang = rng.normal(size=(n, 3)); tang -= ek * np.sum(tang * ek, axis=1, keepdims=True)
        tang = unit(tang) + ek * (rng.normal(size=n) * 10.0 ** rng.uniform(-7, -2, n))[:, None]
        axis_case = rng.random(n) < 0.5
        dirs = np.where(axis_case[:, None], unit(tang), dirs)
        perp = np.where(axis_case[:, None], ek, perp)
        delta = rng.choice([-1, 1], n) * 10.0 ** rng.uniform(-8, -3, n)
        target = c.astype(np.float64) + perp * (float(r) * (1 + delta))[:, None]
        org = (target - dirs * dist[:, None]).astype(F32)
        d = unit(target - org.astype(np.float64)).astype(F32)
        # a third of the rays nearly parallel to an axis
        h = W.trace(org, d)
        box = np.concatenate([(c - r).astype(F32), (c + r).astype(F32)])
        maxabs = float(np.abs(box).max())
        A, C = 8 * EPS * (maxabs + 3 * float(r)), 8 * 3 * EPS / float(r)
        for i in np.nonzero(h["hit_id"] != MISS)[0]:
            sc, pc, gap, mg = band_checks(box, A, C, org[i].astype(np.float64), d[i], float(h["t"][i]))
            n_acc += 1
            n_outside += gap > 0
            assert not sc and not pc, (c, r, org[i], d[i], h["t"][i], gap, mg)
    assert n_acc > 5000 and n_outside > 20       # the set does contain accepted hits whose ray misses the exact box


@pytest.mark.parametrize("seed", range(6))
def test_accepted_triangle_hits_are_never_culled(seed, oracle):
    rng = np.random.default_rng(100 + seed)
    n_acc = n_outside = 0
    for it in range(25):
        s = 10.0 ** rng.uniform(-2, 2)
        p0 = (rng.normal(size=3) * 10.0 ** rng.uniform(-1, 2)).astype(F32)
        if it % 3 == 0:      # sliver fanned from p0 (Q4), aspect 30 .. 30 000
            asp = 10.0 ** rng.uniform(1.5, 4.5)
            e1 = rng.normal(size=3) * s * 30; e2 = e1 + rng.normal(size=3) * s * 30 / asp
            if it % 2 == 0:
                e1, e2 = (e2 - e1), e1      # ... or with p0 on the short edge
        elif it % 3 == 1:    # axis-aligned: zero-thickness box
            e1 = np.array([s, 0, 0]); e2 = np.array([0, s, 0])
        else:
            e1 = rng.normal(size=3) * s; e2 = rng.normal(size=3) * s
        p1, p2 = (p0 + e1).astype(F32), (p0 + e2).astype(F32)
        nrm = [0, 0, 1] * 3
        W = one_item_world(oracle, 2, list(p0) + list(p1) + list(p2) + nrm + [0] * 6)
        n = 1500
        # aim at points on / just outside the edges and vertices
        u = rng.uniform(0, 1, n); v = rng.uniform(0, 1, n)
        edge = rng.integers(0, 4, n)
        u = np.where(edge == 0, 0.0, u); v = np.where(edge == 1, 0.0, v)
        v = np.where(edge == 2, 1.0 - u, v)
        u = u + rng.choice([-1, 1], n) * 10.0 ** rng.uniform(-8, -3, n)
        target = p0.astype(np.float64) + u[:, None] * (p1 - p0).astype(np.float64) + v[:, None] * (p2 - p0).astype(np.float64)
        dirs = unit(rng.normal(size=(n, 3)))
        dist = rng.uniform(0.5, 300.0, n) * s
        org = (target - dirs * dist[:, None]).astype(F32)
        d = unit(target - org.astype(np.float64)).astype(F32)
        h = W.trace(org, d)
        P = np.stack([p0, p1, p2])
        box = np.concatenate([P.min(0), P.max(0)]).astype(F32)
        # pack_scene's triangle band: A = 8 eps M + eps kappa (3 M + 10 L), kappa = L^2 / |e1 x e2|  (slivers are ill-conditioned)
        E1, E2 = (p1 - p0).astype(np.float64), (p2 - p0).astype(np.float64)
        l2 = max(E1 @ E1, E2 @ E2, (E2 - E1) @ (E2 - E1))
        two_a = np.linalg.norm(np.cross(E1, E2))
        M = float(np.abs(box).max())
        A = 8 * EPS * M + EPS * (l2 / two_a) * (3 * M + 10 * math.sqrt(l2))
        for i in np.nonzero(h["hit_id"] != MISS)[0]:
            sc, pc, gap, mg = band_checks(box, A, 0.0, org[i].astype(np.float64), d[i], float(h["t"][i]))
            n_acc += 1
            n_outside += gap > 0
            assert not sc and not pc, (p0, p1, p2, org[i], d[i], h["t"][i], gap, mg)
    assert n_acc > 5000 and n_outside > 20


@pytest.mark.parametrize("seed", range(3))
def test_accepted_rect_and_box_hits_are_never_culled(seed, oracle):
    """Rects (rect_geometry.rs:33-59, box = the rect padded by 0.01 along its normal) and AxisAlignedBox primitives (aabb.rs:80-167,
    box = itself): rays aimed at their borders, edges and corners."""
    rng = np.random.default_rng(300 + seed)
    n_acc = n_outside = 0
    for it in range(30):
        s = 10.0 ** rng.uniform(-1, 3)
        lo = (rng.normal(size=3) * s).astype(F32)
        hi = (lo + np.abs(rng.normal(size=3)) * s * 10.0 ** rng.uniform(-2, 0, 3) + F32(1e-3) * F32(s)).astype(F32)
        n = 1500
        if it % 2 == 0:
            plane = int(rng.integers(0, 3))
            p0, p1, nn = [(0, 1, 2), (0, 2, 1), (1, 2, 0)][plane]
            dist = lo[nn]
            W = one_item_world(oracle, 1, [plane, dist, lo[p0], hi[p0], lo[p1], hi[p1]])
            box = np.zeros(6, F32)
            box[p0], box[3 + p0], box[p1], box[3 + p1] = lo[p0], hi[p0], lo[p1], hi[p1]
            box[nn], box[3 + nn] = F32(dist) - F32(0.01), F32(dist) + F32(0.01)
            target = np.zeros((n, 3))
            target[:, nn] = float(dist)
            for ax in (p0, p1):
                edge = rng.integers(0, 3, n)
                inside = float(lo[ax]) + rng.uniform(0, 1, n) * float(hi[ax] - lo[ax])
                border = np.where(edge == 0, float(lo[ax]), float(hi[ax])) * (1 + rng.choice([-1, 1], n) * 10.0 ** rng.uniform(-8, -4, n))
                target[:, ax] = np.where(edge == 2, inside, border)
        else:
            W = one_item_world(oracle, 3, list(lo) + list(hi))
            box = np.concatenate([lo, hi]).astype(F32)
            pick = rng.integers(0, 2, (n, 3))
            corner = np.where(pick == 0, lo.astype(np.float64), hi.astype(np.float64))
            inside = lo.astype(np.float64) + rng.uniform(0, 1, (n, 3)) * (hi - lo).astype(np.float64)
            on_face = rng.random((n, 3)) < 0.6
            target = np.where(on_face, corner * (1 + rng.choice([-1, 1], (n, 3)) * 10.0 ** rng.uniform(-8, -4, (n, 3))), inside)
        dirs = unit(rng.normal(size=(n, 3)))
        if it % 3 == 0:
            k = rng.integers(0, 3)
            dirs[:, k] *= 10.0 ** rng.uniform(-6, -1, n)
            dirs = unit(dirs)
        dist_o = rng.uniform(0.5, 200.0, n) * s
        org = (target - dirs * dist_o[:, None]).astype(F32)
        d = unit(target - org.astype(np.float64)).astype(F32)
        h = W.trace(org, d)
        A = 8 * EPS * float(np.abs(box).max())
        for i in np.nonzero(h["hit_id"] != MISS)[0]:
            sc, pc, gap, mg = band_checks(box, A, 0.0, org[i].astype(np.float64), d[i], float(h["t"][i]))
            n_acc += 1
            n_outside += gap > 0
            assert not sc and not pc, (it, box, org[i], d[i], h["t"][i], gap, mg)
    assert n_acc > 5000


def world_box_py(lbox, flags, off, sn, cs, vel, t_lo, t_hi):
    """pack_scene's world_box: the local box's corners through Transformation::apply_point, swept over the shutter interval."""
    pts = []
    for c in range(8):
        p = np.array([lbox[3 + 0] if c & 1 else lbox[0], lbox[3 + 1] if c & 2 else lbox[1], lbox[3 + 2] if c & 4 else lbox[2]], np.float64)
        if flags & 1:
            p = np.array([cs * p[0] + sn * p[2] + off[0], p[1] + off[1], -sn * p[0] + cs * p[2] + off[2]])
        for t in (t_lo, t_hi):
            pts.append(p + (np.asarray(vel, np.float64) * t if flags & 2 else 0.0))
    pts = np.array(pts)
    lo, hi = pts.min(0), pts.max(0)
    return np.concatenate([np.nextafter(lo.astype(F32), F32(-np.inf)), np.nextafter(hi.astype(F32), F32(np.inf))]).astype(F32)


@pytest.mark.parametrize("seed", range(4))
def test_accepted_hits_on_wrapped_elements_stay_within_their_world_box_band(seed, oracle):
    """Row f4 on wrapped elements: the rebuilt tree bounds a Transformation / Animation-wrapped element by its world-space box and
    adds 16 eps (coordinate magnitude) to its band for the rounding of the ray's reverse transform."""
    rng = np.random.default_rng(700 + seed)
    n_acc = 0
    for it in range(24):
        s = 10.0 ** rng.uniform(-1, 2)
        kind = [0, 3, 1][it % 3]
        if kind == 0:
            r = F32(s * rng.uniform(0.2, 2)); c = (rng.normal(size=3) * s).astype(F32)
            geom = [c[0], c[1], c[2], r]
            lbox = np.concatenate([(c - r).astype(F32), (c + r).astype(F32)])
            A_loc, C_loc = 8 * EPS * (float(np.abs(lbox).max()) + 3 * float(r)), 24 * EPS / float(r)
        else:
            lo = (rng.normal(size=3) * s).astype(F32); hi = (lo + np.abs(rng.normal(size=3)) * s + F32(1e-3 * s)).astype(F32)
            if kind == 3:
                geom = list(lo) + list(hi); lbox = np.concatenate([lo, hi]).astype(F32)
            else:
                geom = [1, lo[1], lo[0], hi[0], lo[2], hi[2]]           # XZ rect at y = lo[1]
                lbox = np.array([lo[0], lo[1] - F32(0.01), lo[2], hi[0], lo[1] + F32(0.01), hi[2]], F32)
            A_loc, C_loc = 8 * EPS * float(np.abs(lbox).max()), 0.0
        flags = int(rng.integers(1, 4))
        ang = rng.uniform(-math.pi, math.pi)
        off = (rng.normal(size=3) * s * 3).astype(F32) if flags & 1 else np.zeros(3, F32)
        sn, cs = (F32(math.sin(ang)), F32(math.cos(ang))) if flags & 1 else (F32(0), F32(1))
        vel = (rng.normal(size=3) * s).astype(F32) if flags & 2 else np.zeros(3, F32)
        time = F32(rng.uniform(0.0, 1.0))
        wrap = np.array([[flags, off[0], off[1], off[2], sn, cs, vel[0], vel[1], vel[2], 0.0]], F32)
        g = np.zeros((1, 24), F32); g[0, :len(geom)] = geom
        W = oracle.OracleWorld.from_flat(np.array([kind], np.uint32), np.zeros(1, np.uint32), g, np.array([mat(0, 0)], np.uint32),
                                         np.array([solid_tex([1, 1, 1])], np.uint32), [], simple_camera(), np.array([0, 1, 1, 1], F32),
                                         np.zeros(7, F32), wrap=wrap)
        wbox = world_box_py(lbox, flags, off.astype(np.float64), float(sn), float(cs), vel, 0.0, 1.0)
        A = A_loc + 16 * EPS * max(float(np.abs(wbox).max()), float(np.abs(lbox).max()))
        # grazing rays in WORLD space: aim at the surface of the element's world position at `time`
        n = 1500
        wb_t = world_box_py(lbox, flags, off.astype(np.float64), float(sn), float(cs), vel, float(time), float(time))
        centre = 0.5 * (wb_t[:3] + wb_t[3:]).astype(np.float64); half = 0.5 * (wb_t[3:] - wb_t[:3]).astype(np.float64)
        target = centre + rng.uniform(-1.02, 1.02, (n, 3)) * half
        dirs = unit(rng.normal(size=(n, 3)))
        dist = rng.uniform(0.5, 100.0, n) * s
        org = (target - dirs * dist[:, None]).astype(F32)
        d = unit(target - org.astype(np.float64)).astype(F32)
        h = W.trace(org, d, time=float(time))
        for i in np.nonzero(h["hit_id"] != MISS)[0]:
            sc, pc, gap, mg = band_checks(wbox, A, C_loc, org[i].astype(np.float64), d[i], float(h["t"][i]))
            n_acc += 1
            assert not sc and not pc, (it, kind, flags, wbox, org[i], d[i], h["t"][i], gap, mg)
    assert n_acc > 3000
