"""Round-2 baseline: every BASELINE config once through the C ABI (device film), with node / primitive counts and the per-iteration log of C2."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracinginaweekend_b200 as rb
from raytracinginaweekend_b200 import _cabi
CFG = {"c1": (1200, 800, 10, 10), "c2": (1920, 1080, 64, 64), "c3": (1024, 1024, 256, 32), "c4": (3840, 2160, 16, 1), "c5": (3840, 2160, 8, 8)}
which = sys.argv[1].split(",") if len(sys.argv) > 1 else list(CFG)
for name in which:
    w, h, spp, T = CFG[name]
    t0 = time.time(); W = rb.World.named("config:" + name); t1 = time.time()
    rb.render(rb.Size2i(64, 36), 1, 1, 50, W, seed=1); t2 = time.time()
    _cabi.set_option("profile", 1)
    if name == "c2":
        os.environ["RT_ITER_LOG"] = "gpurun_out/r2_iter_c2.csv"
    _cabi.set_option("count", 1)
    _, c = rb.render(rb.Size2i(w, h), T, spp, 50, W, seed=3, return_stats=True)
    os.environ.pop("RT_ITER_LOG", None)
    _cabi.set_option("count", 0)
    best = None
    for rep in range(3):
        _, st = rb.render(rb.Size2i(w, h), T, spp, 50, W, seed=3, return_stats=True)
        if best is None or st["render_ms"] < best["render_ms"]:
            best = st
    print(json.dumps({"config": name, "world_s": t1 - t0, "create_s": t2 - t1, "Mrays_s": best["ray_segments"] / best["render_ms"] / 1e3,
                      "render_ms": best["render_ms"], "extend_ms": best["extend_ms"], "shade_ms": best["shade_ms"], "rays": best["ray_segments"],
                      "iters": best["iterations"], "nodes_per_ray": c["node_visits"] / c["ray_segments"], "prims_per_ray": c["prim_tests"] / c["ray_segments"],
                      "redo": c["redo_rays"] / c["ray_segments"]}), flush=True)
    W.close()
