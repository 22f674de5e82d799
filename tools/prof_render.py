"""Render one config through the C ABI (used for plain timing and under ncu)."""
import sys, os, json, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import raytracinginaweekend_b200 as rb
from raytracinginaweekend_b200 import _cabi as rt
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="config:c2"); ap.add_argument("--spp", type=int, default=2)
ap.add_argument("--threads", type=int, default=1); ap.add_argument("--count", type=int, default=0)
ap.add_argument("--scale", type=float, default=1.0); ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--traversal", type=int, default=3); ap.add_argument("--refill", type=int, default=28); ap.add_argument("--leaf_weight", type=int, default=24); ap.add_argument("--profile", type=int, default=0); ap.add_argument("--pipes", type=int, default=1)
a = ap.parse_args()
W = rb.World.named(a.config); fs = W.export(); w, h = fs.default_size
w = int(w * a.scale); h = int(h * a.scale)
S = rt.DeviceScene.from_flat(fs)
rt.set_option("pipes", a.pipes); rt.set_option("count", a.count); rt.set_option("traversal", a.traversal); rt.set_option("profile", a.profile); rt.set_option("refill", a.refill); rt.set_option("leaf_weight", a.leaf_weight)
for r in range(a.reps):
    img, st = S.render(w, h, a.spp, 50, a.threads, 3)
    st["Mrays_s"] = st["ray_segments"] / st["render_ms"] / 1e3
    if a.count:
        st["nodes_per_ray"] = st["node_visits"] / st["ray_segments"]; st["prims_per_ray"] = st["prim_tests"] / st["ray_segments"]
    print(a.config, w, h, a.spp, json.dumps(st), flush=True)
