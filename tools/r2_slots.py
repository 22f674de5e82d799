"""Sweep of the in-flight path budget ("max_slots") and pipeline count on C2 / C5 / C1."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracinginaweekend_b200 as rb
from raytracinginaweekend_b200 import _cabi
CFG = {"c1": (1200, 800, 10, 10), "c2": (1920, 1080, 64, 64), "c3": (1024, 1024, 256, 32), "c4": (3840, 2160, 16, 16), "c5": (3840, 2160, 16, 16)}
which = sys.argv[1].split(",") if len(sys.argv) > 1 else ["c2", "c5", "c1"]
for name in which:
    w, h, spp, T = CFG[name]
    W = rb.World.named("config:" + name)
    rb.render(rb.Size2i(64, 36), 1, 1, 50, W, seed=1)
    for slots_m in (8, 16, 40, 80, 160):
        for pipes in (1, 2, 4):
            _cabi.set_option("max_slots", slots_m << 20); _cabi.set_option("pipes", pipes)
            best = None
            for rep in range(2):
                _, st = rb.render(rb.Size2i(w, h), T, spp, 50, W, seed=3, return_stats=True)
                if best is None or st["render_ms"] < best["render_ms"]:
                    best = st
            print(json.dumps({"config": name, "max_slots_Mi": slots_m, "pipes": pipes, "slots": best["slots"], "Mrays_s": round(best["ray_segments"] / best["render_ms"] / 1e3, 1),
                              "render_ms": round(best["render_ms"], 2), "iters": best["iterations"]}), flush=True)
    W.close()
