"""Traversal 4 (4-wide accelerator tree) vs 3 (binary): identical films, Mrays/s, node visits / primitive tests per ray."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import raytracinginaweekend_b200 as rb
from raytracinginaweekend_b200 import _cabi
CFG = {"c1": (1200, 800, 10, 10), "c2": (1920, 1080, 64, 64), "c3": (1024, 1024, 256, 32), "c4": (3840, 2160, 16, 16), "c5": (3840, 2160, 16, 16),
       "fs2": (800, 800, 16, 16)}
which = sys.argv[1].split(",") if len(sys.argv) > 1 else ["c1", "c2", "c5", "c4"]
modes = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [3, 4]
def same(a, b):
    return np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])
for name in which:
    w, h, spp, T = CFG[name]
    W = rb.World.named("demo:final_scene2" if name == "fs2" else "config:" + name)
    films = {}
    for m in [0] + modes:
        _cabi.set_option("traversal", m)
        films[m] = rb.render(rb.Size2i(96, 54) if name in ("c4", "c5") else rb.Size2i(200, 120), 3, 6, 50, W, seed=5)
    print(json.dumps({"config": name, "films_identical_to_mode0": {m: bool(same(films[0], films[m])) for m in modes}}), flush=True)
    for m in modes:
        _cabi.set_option("traversal", m)
        _cabi.set_option("count", 1)
        _, c = rb.render(rb.Size2i(w, h), T, spp, 50, W, seed=3, return_stats=True)
        _cabi.set_option("count", 0)
        best = None
        for rep in range(3):
            _, st = rb.render(rb.Size2i(w, h), T, spp, 50, W, seed=3, return_stats=True)
            if best is None or st["render_ms"] < best["render_ms"]:
                best = st
        print(json.dumps({"config": name, "traversal": m, "Mrays_s": round(best["ray_segments"] / best["render_ms"] / 1e3, 1), "render_ms": round(best["render_ms"], 2),
                          "nodes_per_ray": round(c["node_visits"] / c["ray_segments"], 2), "prims_per_ray": round(c["prim_tests"] / c["ray_segments"], 2),
                          "redo": round(c["redo_rays"] / c["ray_segments"], 5), "max_stack": c["max_stack"], "stack_cap": c["stack_capacity"],
                          "lanes": round(c["busy_lane_rounds"] / max(1, c["warp_rounds"]), 1)}), flush=True)
    W.close()
_cabi.shutdown()
