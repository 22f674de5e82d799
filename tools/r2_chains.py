"""Dynamic chain scheduling: sweep of the in-flight path budget ("max_slots") x pipelines, the reference-tree share of the
extend warps (RT_TAIL_LOG with "count"), and the measurement-only A/B without the reference-tree restarts ("debug_skip_redo")."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import raytracinginaweekend_b200 as rb
from raytracinginaweekend_b200 import _cabi
CFG = {"c1": (1200, 800, 10, 10), "c2": (1920, 1080, 64, 64), "c3": (1024, 1024, 256, 32), "c4": (3840, 2160, 16, 16), "c5": (3840, 2160, 16, 16)}
which = sys.argv[1].split(",") if len(sys.argv) > 1 else ["c2", "c5", "c1", "c3"]
sweep = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [2, 4, 8, 16, 32, 64, 160]
pipes_list = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 2, 4]
iterlog_dir = sys.argv[4] if len(sys.argv) > 4 else None
def best_of(n, *a, **k):
    best = None
    for _ in range(n):
        img, st = rb.render(*a, return_stats=True, **k)
        if best is None or st["render_ms"] < best[1]["render_ms"]:
            best = (img, st)
    return best
for name in which:
    w, h, spp, T = CFG[name]
    W = rb.World.named("config:" + name)
    rb.render(rb.Size2i(64, 36), 1, 1, 50, W, seed=1)
    ref = None
    for slots_m in sweep:
        for pipes in pipes_list:
            _cabi.set_option("max_slots", slots_m << 20); _cabi.set_option("pipes", pipes)
            img, st = best_of(2, rb.Size2i(w, h), T, spp, 50, W, seed=3)
            if ref is None: ref = img
            same = bool(np.array_equal(ref.view(np.uint32), img.view(np.uint32)))
            print(json.dumps({"config": name, "max_slots_Mi": slots_m, "pipes": pipes, "slots": st["slots"], "Mrays_s": round(st["ray_segments"] / st["render_ms"] / 1e3, 1),
                              "render_ms": round(st["render_ms"], 2), "iters": st["iterations"], "same_film": same}), flush=True)
    if iterlog_dir:  # per-launch times, one pipeline, launches back to back
        _cabi.set_option("pipes", 1); _cabi.set_option("profile", 1); _cabi.set_option("count", 1)
        for slots_m in (16, 160):
            _cabi.set_option("max_slots", slots_m << 20)
            os.environ["RT_ITER_LOG"] = os.path.join(iterlog_dir, "r2b_iter_%s_%dMi.csv" % (name, slots_m))
            rb.render(rb.Size2i(w, h), T, spp, 50, W, seed=3)
        del os.environ["RT_ITER_LOG"]
        _cabi.set_option("profile", 0); _cabi.set_option("count", 0)
    # reference-tree share + A/B without restarts (NOT a valid render; timing only)
    _cabi.set_option("max_slots", 160 << 20); _cabi.set_option("pipes", 2)
    _cabi.set_option("count", 1)
    _, st = rb.render(rb.Size2i(w, h), T, spp, 50, W, seed=3, return_stats=True)
    print(json.dumps({"config": name, "count": 1, "nodes_per_ray": round(st["node_visits"] / st["ray_segments"], 2), "prims_per_ray": round(st["prim_tests"] / st["ray_segments"], 2),
                      "redo": st["redo_rays"] / st["ray_segments"], "lanes": round(st["busy_lane_rounds"] / max(1, st["warp_rounds"]), 2), "hist": st["rounds_hist"]}), flush=True)
    _cabi.set_option("count", 0)
    for skip in (0, 1):
        _cabi.set_option("debug_skip_redo", skip)
        _, st = best_of(2, rb.Size2i(w, h), T, spp, 50, W, seed=3)
        print(json.dumps({"config": name, "debug_skip_redo": skip, "Mrays_s": round(st["ray_segments"] / st["render_ms"] / 1e3, 1), "render_ms": round(st["render_ms"], 2)}), flush=True)
    _cabi.set_option("debug_skip_redo", 0)
    W.close()
