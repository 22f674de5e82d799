import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracinginaweekend_b200 as rb
from raytracinginaweekend_b200 import _cabi
cfgs = {"c2": ("config:c2", (1920, 1080), 64, 64), "c1": ("config:c1", (1200, 800), 10, 10), "c5": ("config:c5", (1920, 1080), 4, 8),
        "c3": ("config:c3", (1024, 1024), 32, 64), "suz": ("demo:suzanne", (1600, 1200), 16, 16)}
for key in (sys.argv[1:] or ["c2"]):
    name, size, T, spp = cfgs[key]
    for ps, extra in [tuple(int(v) for v in a.split(",")) for a in os.environ.get("PS", "100,3 200,3 400,3 400,5 800,4").split()]:
        _cabi.set_option("presplit", ps); _cabi.set_option("rebuild_depth_extra", extra)
        W = rb.World.named(name)
        t0 = time.time(); rb.render(rb.Size2i(64, 36), 1, 1, 50, W, seed=1); t_create = time.time() - t0
        _cabi.set_option("count", 1)
        _, c = rb.render(rb.Size2i(*size), 1, 1, 50, W, seed=3, return_stats=True)
        _cabi.set_option("count", 0)
        best = None
        for rep in range(2):
            _, st = rb.render(rb.Size2i(*size), T, spp, 50, W, seed=3, return_stats=True)
            if best is None or st["render_ms"] < best["render_ms"]: best = st
        print("%-12s presplit %4d extra %d | %7.1f Mrays/s | nodes/ray %6.1f prims/ray %5.1f | redo %.5f | stack %d/%d | create %.2fs" % (
            name, ps, extra, best["ray_segments"] / best["render_ms"] / 1e3, c["node_visits"] / c["ray_segments"], c["prim_tests"] / c["ray_segments"],
            c["redo_rays"] / c["ray_segments"], c["max_stack"], c["stack_capacity"], t_create), flush=True)
        W.close()
