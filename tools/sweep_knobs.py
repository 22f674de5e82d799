"""Knob sweep for the extend kernel on C2 at the bench configuration (64 lanes x 64 spp)."""
import sys, os, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracinginaweekend_b200 as rb
from raytracinginaweekend_b200 import _cabi
W = rb.World.named("config:c2")
size = rb.Size2i(1920, 1080)
rb.render(size, 1, 1, 50, W, seed=1)
def run(**kw):
    for k, v in kw.items(): _cabi.set_option(k, v)
    best = 1e9
    for rep in range(2):
        _, st = rb.render(size, 64, 64, 50, W, seed=3, return_stats=True)
        best = min(best, st["render_ms"])
    print(kw, "%.1f ms  %.1f Mrays/s" % (best, st["ray_segments"] / best / 1e3), flush=True)
for p in (1, 2, 3, 4, 2, 1):
    run(pipes=p, leaf_weight=24, refill=28)
run(pipes=1, leaf_weight=24, refill=28, blocks_per_sm=8)
run(pipes=1, leaf_weight=24, refill=28, blocks_per_sm=0, check_every=4)
run(pipes=1, leaf_weight=24, refill=28, check_every=16)
