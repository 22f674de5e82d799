"""Row f4 on scenes with wrapped elements / volumes: traversal 2 (reference tree) vs 3 (rebuilt tree with true world-space bounds)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracinginaweekend_b200 as rb
from raytracinginaweekend_b200 import _cabi
for name in ("demo:cornell_box", "demo:moving_spheres", "demo:cornell_box_smoke", "demo:final_scene2", "demo:perlin_spheres", "demo:final_scene1"):
    W = rb.World.named(name)
    size = rb.Size2i(1200, 1200 * W.default_size.height // W.default_size.width)
    rb.render(rb.Size2i(64, 64), 1, 1, 50, W, seed=1)
    for mode in (2, 3):
        _cabi.set_option("traversal", mode)
        _cabi.set_option("count", 1)
        _, c = rb.render(size, 1, 1, 50, W, seed=3, return_stats=True)
        _cabi.set_option("count", 0)
        best = None
        for rep in range(2):
            _, st = rb.render(size, 16, 16, 50, W, seed=3, return_stats=True)
            if best is None or st["render_ms"] < best["render_ms"]: best = st
        print("%-24s mode %d | %7.1f Mrays/s | nodes/ray %6.1f prims/ray %5.1f | restarted %.4f | stack %d/%d" % (
            name, mode, best["ray_segments"] / best["render_ms"] / 1e3, c["node_visits"] / c["ray_segments"], c["prim_tests"] / c["ray_segments"],
            c["redo_rays"] / c["ray_segments"], c["max_stack"], c["stack_capacity"]), flush=True)
    W.close()
