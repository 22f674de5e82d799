"""Scratch: option sweep timing on one config."""
import sys, os, json, argparse, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracinginaweekend_b200 as rb
from raytracinginaweekend_b200 import _cabi as rt
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="config:c2"); ap.add_argument("--spp", type=int, default=8); ap.add_argument("--threads", type=int, default=4)
ap.add_argument("--opt", action="append", default=[])   # key=v1,v2,...
a = ap.parse_args()
W = rb.World.named(a.config); fs = W.export(); w, h = fs.default_size
S = rt.DeviceScene.from_flat(fs)
S.render(w, h, 1, 50, 1, 3)
keys = [o.split("=")[0] for o in a.opt]; vals = [[int(v) for v in o.split("=")[1].split(",")] for o in a.opt]
for combo in itertools.product(*vals):
    for k, v in zip(keys, combo): rt.set_option(k, v)
    img, st = S.render(w, h, a.spp, 50, a.threads, 3)
    print(a.config, dict(zip(keys, combo)), "ms %.1f Mrays/s %.1f" % (st["render_ms"], st["ray_segments"] / st["render_ms"] / 1e3), flush=True)
