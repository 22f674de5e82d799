"""Persistent path kernel ("loop" 1) vs the wavefront launches ("loop" 0): same film, Mrays/s per config and in-flight budget."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import raytracinginaweekend_b200 as rb
from raytracinginaweekend_b200 import _cabi
CFG = {"c1": (1200, 800, 10, 10), "c2": (1920, 1080, 64, 64), "c3": (1024, 1024, 256, 32), "c4": (3840, 2160, 16, 16), "c5": (3840, 2160, 16, 16)}
which = sys.argv[1].split(",") if len(sys.argv) > 1 else ["c1", "c2", "c3", "c5", "c4"]
slot_list = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [16, 40, 160]
for name in which:
    w, h, spp, T = CFG[name]
    W = rb.World.named("config:" + name)
    # bitwise A/B on a small frame first
    films = []
    for loop in (0, 1):
        _cabi.set_option("loop", loop)
        films.append(rb.render(rb.Size2i(160, 90), 4, 8, 50, W, seed=5))
    same = np.array_equal(np.isnan(films[0]), np.isnan(films[1])) and np.array_equal(films[0][~np.isnan(films[0])], films[1][~np.isnan(films[1])])
    print(json.dumps({"config": name, "small_frame_films_identical": bool(same)}), flush=True)
    for slots_m in slot_list:
        for loop in (0, 1):
            _cabi.set_option("max_slots", slots_m << 20); _cabi.set_option("loop", loop)
            best = None
            for rep in range(2):
                _, st = rb.render(rb.Size2i(w, h), T, spp, 50, W, seed=3, return_stats=True)
                if best is None or st["render_ms"] < best["render_ms"]:
                    best = st
            print(json.dumps({"config": name, "loop": loop, "max_slots_Mi": slots_m, "tracks": best["tracks"], "Mrays_s": round(best["ray_segments"] / best["render_ms"] / 1e3, 1),
                              "render_ms": round(best["render_ms"], 2), "rays": best["ray_segments"]}), flush=True)
    W.close()
