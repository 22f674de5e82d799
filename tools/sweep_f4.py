"""Row f4 sweep: throughput / node visits / redo fraction of traversal 3 vs 2, and the depth cap of the rebuilt tree."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracinginaweekend_b200 as rb
from raytracinginaweekend_b200 import _cabi

def run(name, size, T, spp, extra, mode):
    _cabi.set_option("rebuild_depth_extra", extra)
    _cabi.set_option("traversal", mode)
    W = rb.World.named(name)
    t0 = time.time()
    rb.render(rb.Size2i(*size), 1, 1, 50, W, seed=1)
    t_first = time.time() - t0
    _cabi.set_option("count", 1)
    _, c = rb.render(rb.Size2i(*size), 1, 1, 50, W, seed=3, return_stats=True)
    _cabi.set_option("count", 0)
    best = None
    for rep in range(2):
        _, st = rb.render(rb.Size2i(*size), T, spp, 50, W, seed=3, return_stats=True)
        if best is None or st["render_ms"] < best["render_ms"]: best = st
    print("%-10s mode %d extra %2d | %7.1f Mrays/s | nodes/ray %6.1f prims/ray %5.1f | redo %.4f | first render %.1fs" % (
        name, mode, extra, best["ray_segments"] / best["render_ms"] / 1e3, c["node_visits"] / c["ray_segments"], c["prim_tests"] / c["ray_segments"],
        c["redo_rays"] / c["ray_segments"], t_first), flush=True)
    W.close()

which = sys.argv[1] if len(sys.argv) > 1 else "c2"
if which == "c2":
    for extra in (0, 2, 3, 5, 8, 12):
        run("config:c2", (1920, 1080), 32, 32, extra, 3)
    run("config:c2", (1920, 1080), 32, 32, 3, 2)
elif which == "ab":
    for T in (32, 64):
        for rep in range(2):
            for mode in (2, 3):
                run("config:c2", (1920, 1080), T, T, 3, mode)
elif which == "all":
    for name, size, T, spp in (("config:c1", (1200, 800), 10, 10), ("config:c3", (1024, 1024), 64, 64), ("config:c5", (3840, 2160), 8, 16), ("config:c4", (3840, 2160), 4, 4)):
        for mode, extra in ((2, 3), (3, 3), (3, 8)):
            run(name, size, T, spp, extra, mode)
