"""Knob sweep on the final round-2 kernels (C2 at the bench configuration unless a config is given)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracinginaweekend_b200 as rb
from raytracinginaweekend_b200 import _cabi
CFG = {"c1": (1200, 800, 10, 10), "c2": (1920, 1080, 64, 64), "c3": (1024, 1024, 256, 32), "c4": (3840, 2160, 16, 16), "c5": (3840, 2160, 16, 16)}
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
w, h, spp, T = CFG[name]
W = rb.World.named("config:" + name)
size = rb.Size2i(w, h)
rb.render(rb.Size2i(64, 36), 1, 1, 50, W, seed=1)
base = dict(pipes=0, leaf_weight=24, refill=28, tail_budget=64, check_every=8, blocks_per_sm=0, carveout=75)
def run(**kw):
    opts = dict(base); opts.update(kw)
    for k, v in opts.items(): _cabi.set_option(k, v)
    best = 1e9
    for rep in range(3):
        _, st = rb.render(size, T, spp, 50, W, seed=3, return_stats=True)
        best = min(best, st["render_ms"])
    print(name, kw, "%.2f ms  %.1f Mrays/s  iters %d" % (best, st["ray_segments"] / best / 1e3, st["iterations"]), flush=True)
if len(sys.argv) > 2 and sys.argv[2] == "carveout2":
    for c in (-1, 75, 66, 85, -1, 75):
        run(carveout=c)
    sys.exit(0)
if len(sys.argv) > 2 and sys.argv[2] == "carveout":
    for c in (-1, 100, 90, 80, 72, 60, 50, -1):
        run(carveout=c)
    for c in (100, 80, 60):
        run(carveout=c, blocks_per_sm=6)
    sys.exit(0)
run()
for p in (2, 3, 4): run(pipes=p)
for p in (2, 3): run(pipes=p, tail_budget=0)
for lw in (16, 20, 28, 32): run(leaf_weight=lw)
for rf in (24, 26, 30, 32): run(refill=rf)
for ce in (4, 16): run(check_every=ce)
for tb in (128, 256, 1024): run(tail_budget=tb)
