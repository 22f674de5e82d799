"""Row f4, the lemma behind `ref_visible` (raytracinginaweekend_b200/csrc/rt_device.cuh), checked on the CPU against the oracle's
literal `Aabb::hit_cond` (aabb.rs:65-78):

    if p = rn(o + rn(t d)) lies inside a box by more than 4e-7 (|face| + |o| + |p|) on every axis, then hit_cond(box, ray, start..e)
    is true for every e > t (start <= t) — and so it is for every LARGER box (the ancestors are nested exact unions).

The device evaluates the margin in float32 exactly as restated here (same operation order, no FMA).  Adversarial inputs: points
sitting a few ulps from a face, huge and tiny coordinates, near-axis-parallel and exactly axis-parallel rays."""
import math

import numpy as np

from helpers import F32


def ref_visible_f32(lo, hi, o, d, t):
    c = F32(4e-7)
    p = (o + t * d).astype(F32)
    s = (np.abs(o) + np.abs(p)).astype(F32)
    m_lo = (c * (np.abs(lo) + s).astype(F32)).astype(F32)
    m_hi = (c * (np.abs(hi) + s).astype(F32)).astype(F32)
    return bool(np.all((p - lo).astype(F32) > m_lo) and np.all((hi - p).astype(F32) > m_hi)), p


def test_point_inside_by_margin_implies_hit_cond(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(2026)
    checked = passed_margin = 0
    for trial in range(60000):
        scale = F32(10.0 ** rng.uniform(-3, 3))
        centre = (rng.uniform(-1, 1, 3) * scale * 10.0 ** rng.uniform(-1, 1)).astype(F32)
        half = (np.abs(rng.normal(size=3)) * scale * 10.0 ** rng.uniform(-2, 0) + F32(1e-30)).astype(F32)
        if trial % 7 == 0:
            half[rng.integers(0, 3)] *= F32(1e-4)                     # a thin box
        lo, hi = (centre - half).astype(F32), (centre + half).astype(F32)
        # hit point: mostly a few ulps to a few margins away from a face (the decisive zone), sometimes well inside
        p_target = (lo + rng.uniform(0, 1, 3) * (hi - lo)).astype(np.float64)
        k = rng.integers(0, 3)
        face = float(lo[k]) if rng.random() < 0.5 else float(hi[k])
        if trial % 3 != 0:
            p_target[k] = face + (1 if face == float(lo[k]) else -1) * abs(face if face != 0 else scale) * 10.0 ** rng.uniform(-7.5, -5.5)
        d = rng.normal(size=3)
        mode = trial % 5
        if mode == 1:
            d[rng.integers(0, 3)] *= 10.0 ** rng.uniform(-8, -2)          # nearly parallel to a slab
        if mode == 2:
            d[rng.integers(0, 3)] = 0.0                                   # exactly parallel
        d = (d / np.linalg.norm(d)).astype(F32)
        t = F32(10.0 ** rng.uniform(-2, 2)) * scale
        o = (p_target - float(t) * d.astype(np.float64)).astype(F32)
        ok, p = ref_visible_f32(lo, hi, o, d, t)
        checked += 1
        if not ok:
            continue
        passed_margin += 1
        box = np.concatenate([lo, hi]).astype(F32)
        e = np.nextafter(t, F32(np.inf))
        start = F32(0.001) * min(F32(1.0), t)
        for grow in (0.0, 1.0, 1000.0):                                  # the box itself and two nested ancestors
            g = (np.abs(rng.normal(size=6)) * grow * scale).astype(F32)
            big = box.copy(); big[:3] = (big[:3] - g[:3]).astype(F32); big[3:] = (big[3:] + g[3:]).astype(F32)
            for end in (e, F32(np.inf)):
                assert L.orc_aabb_hit_cond(oracle._p(big), oracle._p(o), oracle._p(d), start, end) == 1, (trial, lo, hi, o, d, t, p, grow, end)
    assert checked == 60000 and passed_margin > 10000      # the margin is not vacuous on the adversarial set


def test_margin_rejects_flat_boxes_and_face_points(oracle):
    """Zero-thickness boxes (invisible in the reference, aabb.rs:73) and points exactly on a face never pass the margin."""
    lo, hi = np.array([0, 0, -3], F32), np.array([1, 1, -3], F32)
    o, d = np.array([0.3, 0.3, 0], F32), np.array([0, 0, -1], F32)
    assert not ref_visible_f32(lo, hi, o, d, F32(3.0))[0]
    lo, hi = np.array([0, 0, -4], F32), np.array([1, 1, -3], F32)
    assert not ref_visible_f32(lo, hi, o, d, F32(3.0))[0]            # on the z = -3 face
    assert ref_visible_f32(lo, hi, o, d, F32(3.5))[0]
