import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracinginaweekend_b200 as rb
from raytracinginaweekend_b200 import _cabi
for name, size, T, spp in (("config:c1", (1200, 800), 10, 10), ("config:c5", (1920, 1080), 4, 8), ("config:c2", (1920, 1080), 64, 64)):
    W = rb.World.named(name)
    rb.render(rb.Size2i(*size), 1, 1, 50, W, seed=1)
    for mode, pol in ((2, 0), (3, 0), (3, 1), (3, 2)):
        _cabi.set_option("traversal", mode); _cabi.set_option("fast_order", pol)
        _cabi.set_option("count", 1)
        _, c = rb.render(rb.Size2i(*size), 1, 1, 50, W, seed=3, return_stats=True)
        _cabi.set_option("count", 0)
        best = None
        for rep in range(2):
            _, st = rb.render(rb.Size2i(*size), T, spp, 50, W, seed=3, return_stats=True)
            if best is None or st["render_ms"] < best["render_ms"]: best = st
        print("%-10s mode %d order %d | %7.1f Mrays/s | nodes/ray %6.1f prims/ray %5.1f | redo %.5f" % (
            name, mode, pol, best["ray_segments"] / best["render_ms"] / 1e3, c["node_visits"] / c["ray_segments"], c["prim_tests"] / c["ray_segments"],
            c["redo_rays"] / c["ray_segments"]), flush=True)
    W.close()
