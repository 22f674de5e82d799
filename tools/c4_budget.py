"""C4: pre-split budget above 150 000 elements ("presplit_large", per cent of the element count) vs scene hand-off time, node visits and Mrays/s."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracinginaweekend_b200 as rb
from raytracinginaweekend_b200 import _cabi
W = rb.World.named("config:c4")
size = rb.Size2i(3840, 2160)
for pct in [int(a) for a in sys.argv[1:]] or [200, 300, 400]:
    _cabi.set_option("presplit_large", pct)
    W.release_device()
    t0 = time.time(); sc = W.device_scene(0); t1 = time.time()
    _cabi.set_option("count", 1)
    _, c = rb.render(size, 1, 1, 50, W, seed=3, return_stats=True)
    _cabi.set_option("count", 0)
    best = 0.0
    for rep in range(2):
        _, st = rb.render(size, 16, 16, 50, W, seed=3, return_stats=True)
        best = max(best, st["ray_segments"] / st["render_ms"] / 1e3)
    print("c4 budget %4d %% | create %.3f s | %6.1f MB | %7.1f Mrays/s | nodes/ray %6.1f prims/ray %5.2f" % (
        pct, t1 - t0, sc.device_bytes() / 1e6, best, c["node_visits"] / c["ray_segments"], c["prim_tests"] / c["ray_segments"]), flush=True)
