"""The reciprocal fast path of the node test (rt_device.cuh: trav_setup / approx_slab): a slab distance is evaluated as
fma(b, inv, oinv) with inv = rn(1/d), oinv = rn(-o * inv), instead of the reference's rn(rn(b - o) / d) (aabb.rs:65-78).
Decisions taken from the approximate value are only trusted outside a band of  T.oe + RT_BAND |t|  with
T.oe = 2.4e-7 max_i |o_i / d_i| and RT_BAND = 1e-6; inside the band the exact code runs.  This test checks the error bound the band
rests on, |approx - exact| <= 4 eps |t| + eps |o/d| = 2.4e-7 |t| + 6e-8 |o/d| (+ a denormal floor), on adversarial operands,
emulating the float32 FMA exactly (the product of two float32 is exact in float64).  The kernel's bands are 4x that."""
import numpy as np

from helpers import F32


def fma32(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F32)


def test_reciprocal_slab_distance_error_bound():
    rng = np.random.default_rng(7)
    n = 2_000_000
    mag = lambda lo, hi: (10.0 ** rng.uniform(lo, hi, n) * rng.choice([-1.0, 1.0], n)).astype(F32)
    b = mag(-3, 4)                      # box plane coordinate
    o = mag(-3, 4)                      # ray origin coordinate
    d = mag(-6, 0)                      # direction component (unit vector: |d| <= 1), down to 1e-6
    near = rng.random(n) < 0.3          # planes a few ulps from the origin coordinate: catastrophic cancellation in b - o
    b = np.where(near, (o.astype(np.float64) * (1 + rng.normal(size=n) * 10.0 ** rng.uniform(-7.5, -4, n))).astype(F32), b)
    exact = ((b - o).astype(F32) / d).astype(F32)                    # the reference: two roundings
    inv = (F32(1.0) / d).astype(F32)
    oinv = ((-o).astype(F32) * inv).astype(F32)
    approx = fma32(b, inv, oinv)
    err = np.abs(approx.astype(np.float64) - exact.astype(np.float64))
    bound = 2.4e-7 * np.abs(exact.astype(np.float64)) + 6e-8 * np.abs(o.astype(np.float64) / d.astype(np.float64)) + 1e-37
    worst = np.max(err / bound)
    assert worst <= 1.0, (worst, int(np.argmax(err / bound)))
    assert worst > 0.3       # the bound is not vacuous on this operand set
    # the kernel's own band: RT_BAND |t| + T.oe
    band = 1e-6 * np.abs(approx.astype(np.float64)) + 2.4e-7 * np.abs(oinv.astype(np.float64))
    assert np.max(err / (band + 1e-37)) <= 0.5
