import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracinginaweekend_b200 as rb
from raytracinginaweekend_b200 import _cabi
t0 = time.time(); W = rb.World.named("config:c4"); print("world build %.1fs" % (time.time() - t0), flush=True)
size = rb.Size2i(3840, 2160)
t0 = time.time(); rb.render(rb.Size2i(64, 36), 1, 1, 50, W, seed=1); print("scene create + tiny render %.1fs" % (time.time() - t0), flush=True)
for mode in (3, 2):
    _cabi.set_option("traversal", mode)
    _cabi.set_option("count", 1)
    _, c = rb.render(size, 1, 1, 50, W, seed=3, return_stats=True)
    _cabi.set_option("count", 0)
    _, st = rb.render(size, 4, 4, 50, W, seed=3, return_stats=True)
    print("c4 mode %d | %7.1f Mrays/s | nodes/ray %6.1f prims/ray %5.1f | redo %.5f | %d rays %.0f ms" % (
        mode, st["ray_segments"] / st["render_ms"] / 1e3, c["node_visits"] / c["ray_segments"], c["prim_tests"] / c["ray_segments"],
        c["redo_rays"] / c["ray_segments"], st["ray_segments"], st["render_ms"]), flush=True)
