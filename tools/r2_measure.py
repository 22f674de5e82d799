"""Render time of the BASELINE configs at the library's defaults and for a few pipeline counts, the restart (redo) rate with "count",
and optionally a per-launch log of one single-pipeline render (RT_ITER_LOG + "profile", COUNT off so the times are the product's)."""
import sys, os, json
sys.path.insert(0, os.environ.get("PKG_ROOT") or os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # PKG_ROOT: an A/B build of the package
import numpy as np
import raytracinginaweekend_b200 as rb
from raytracinginaweekend_b200 import _cabi
CFG = {"c1": (1200, 800, 10, 10), "c2": (1920, 1080, 64, 64), "c3": (1024, 1024, 256, 32), "c4": (3840, 2160, 16, 16), "c5": (3840, 2160, 16, 16)}
which = sys.argv[1].split(",") if len(sys.argv) > 1 else ["c2", "c5", "c1", "c3", "c4"]
pipes_list = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 2, 4]
iterlog_dir = sys.argv[3] if len(sys.argv) > 3 else None
extra = dict(kv.split("=") for kv in sys.argv[4].split(",")) if len(sys.argv) > 4 else {}
for k, v in extra.items(): _cabi.set_option(k, int(v))
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
    for pipes in pipes_list:
        _cabi.set_option("pipes", pipes)
        img, st = best_of(3, rb.Size2i(w, h), T, spp, 50, W, seed=3)
        print(json.dumps({"config": name, "pipes_opt": pipes, "pipes": st["pipes"], "slots": st["slots"], "Mrays_s": round(st["ray_segments"] / st["render_ms"] / 1e3, 1),
                          "render_ms": round(st["render_ms"], 2), "iters": st["iterations"], "extra": extra}), flush=True)
    _cabi.set_option("pipes", 0); _cabi.set_option("count", 1)
    if iterlog_dir: os.environ["RT_LONG_RAYS"] = os.path.join(iterlog_dir, "r2b_longrays_%s.csv" % name)
    _, st = rb.render(rb.Size2i(w, h), T, spp, 50, W, seed=3, return_stats=True)
    print(json.dumps({"config": name, "count": 1, "nodes_per_ray": round(st["node_visits"] / st["ray_segments"], 2), "prims_per_ray": round(st["prim_tests"] / st["ray_segments"], 2),
                      "redo_rays": st["redo_rays"], "redo": st["redo_rays"] / st["ray_segments"], "lanes": round(st["busy_lane_rounds"] / max(1, st["warp_rounds"]), 2), "hist": st["rounds_hist"]}), flush=True)
    _cabi.set_option("count", 0)
    os.environ.pop("RT_LONG_RAYS", None)
    if iterlog_dir:
        _cabi.set_option("pipes", 1); _cabi.set_option("profile", 1)
        os.environ["RT_ITER_LOG"] = os.path.join(iterlog_dir, "r2b_iterp1_%s%s.csv" % (name, "".join("_%s%s" % kv for kv in extra.items())))
        rb.render(rb.Size2i(w, h), T, spp, 50, W, seed=3)
        del os.environ["RT_ITER_LOG"]
        _cabi.set_option("profile", 0); _cabi.set_option("pipes", 0)
    W.close()
