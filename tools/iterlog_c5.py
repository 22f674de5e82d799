import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracinginaweekend_b200 as rb
from raytracinginaweekend_b200 import _cabi
name = sys.argv[1] if len(sys.argv) > 1 else "config:c5"
W = rb.World.named(name)
size = rb.Size2i(1920, 1080)
rb.render(size, 1, 1, 50, W, seed=1)
_cabi.set_option("profile", 1)
for mode in (2, 3):
    _cabi.set_option("traversal", mode)
    os.environ["RT_ITER_LOG"] = "gpurun_out/iter_%s_m%d.csv" % (name.split(":")[1], mode)
    for rep in range(2):
        _, st = rb.render(size, 4, 8, 50, W, seed=3, return_stats=True)
    print(mode, st["render_ms"], st["extend_ms"], st["shade_ms"], st["iterations"])
