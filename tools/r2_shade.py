"""Shade / extend device time per config (one pipeline, per-launch events): the numbers behind k_shade's launch bounds."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracinginaweekend_b200 as rb
from raytracinginaweekend_b200 import _cabi
CFG = {"c1": (1200, 800, 10, 10), "c2": (1920, 1080, 64, 64), "c3": (1024, 1024, 256, 32), "fs2": (800, 800, 16, 16), "perlin": (800, 450, 16, 16)}
for name in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["c2", "c3", "fs2", "perlin"]):
    w, h, spp, T = CFG[name]
    W = rb.World.named({"fs2": "demo:final_scene2", "perlin": "demo:perlin_spheres"}.get(name, "config:" + name))
    _cabi.set_option("pipes", 1); _cabi.set_option("profile", 1)
    best = None
    for rep in range(3):
        _, st = rb.render(rb.Size2i(w, h), T, spp, 50, W, seed=3, return_stats=True)
        if best is None or st["render_ms"] < best["render_ms"]:
            best = st
    print(json.dumps({"config": name, "render_ms": round(best["render_ms"], 2), "extend_ms": round(best["extend_ms"], 2), "shade_ms": round(best["shade_ms"], 2),
                      "shade_GBps_at_136B_per_ray": round(best["ray_segments"] * 136 / best["shade_ms"] / 1e6, 1), "rays": best["ray_segments"]}), flush=True)
    W.close()
_cabi.shutdown()
