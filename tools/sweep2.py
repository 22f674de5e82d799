"""Scratch: timing vs lanes (thread_count) through the product API."""
import sys, os, json, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracinginaweekend_b200 as rb
from raytracinginaweekend_b200 import _cabi
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="config:c2"); ap.add_argument("--spp", type=int, default=64)
ap.add_argument("--threads", default="4,8,16"); ap.add_argument("--opt", action="append", default=[])
a = ap.parse_args()
W = rb.World.named(a.config); size = W.default_size
for o in a.opt:
    k, v = o.split("="); _cabi.set_option(k, int(v))
_cabi.set_option("profile", 1)
rb.render(size, 1, 1, 50, W)
for T in [int(x) for x in a.threads.split(",")]:
    img, st = rb.render(size, T, a.spp, 50, W, seed=3, return_stats=True)
    print(a.config, "T", T, "ms %.1f Mrays/s %.1f extend_ms %.1f shade_ms %.1f iters %d rays %d" % (st["render_ms"], st["ray_segments"] / st["render_ms"] / 1e3, st["extend_ms"], st["shade_ms"], st["iterations"], st["ray_segments"]), flush=True)
