import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracinginaweekend_b200 as rb
from raytracinginaweekend_b200 import _cabi
W = rb.World.named("config:c2")
size = rb.Size2i(1920, 1080)
rb.render(size, 1, 1, 50, W, seed=1)
for bps in (0, 9, 8, 7, 6, 5, 4):
    _cabi.set_option("blocks_per_sm", bps)
    best = 1e9
    for rep in range(2):
        _, st = rb.render(size, 64, 64, 50, W, seed=3, return_stats=True)
        best = min(best, st["render_ms"])
    print("blocks_per_sm", bps, "%.1f ms %.1f Mrays/s" % (best, st["ray_segments"] / best / 1e3), flush=True)
