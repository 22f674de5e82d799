import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracinginaweekend_b200 as rb
from raytracinginaweekend_b200 import _cabi
T = int(sys.argv[1]) if len(sys.argv) > 1 else 32
W = rb.World.named("config:c2")
rb.render(rb.Size2i(1920, 1080), 1, 1, 50, W, seed=1)
_cabi.set_option("profile", 1)
for mode in (2, 3):
    _cabi.set_option("traversal", mode)
    os.environ["RT_ITER_LOG"] = "gpurun_out/iter_T%d_m%d.csv" % (T, mode)
    for rep in range(2):
        _, st = rb.render(rb.Size2i(1920, 1080), T, T, 50, W, seed=3, return_stats=True)
    print(mode, st["render_ms"], st["extend_ms"], st["shade_ms"], st["iterations"])
