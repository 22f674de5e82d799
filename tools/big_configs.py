"""Scratch: time the large configs through the product API."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracinginaweekend_b200 as rb
from raytracinginaweekend_b200 import _cabi
_cabi.set_option("profile", 1)
for name, spp, T in [("config:c4", 16, 4), ("config:c5", 32, 8), ("config:c5", 64, 16)]:
    t0 = time.time(); W = rb.World.named(name); t1 = time.time()
    size = W.default_size
    rb.render(size, 1, 1, 50, W); t2 = time.time()
    _cabi.set_option("count", 1)
    img, st = rb.render(size, 1, 1, 50, W, seed=3, return_stats=True)
    _cabi.set_option("count", 0)
    print(name, "build %.1fs first render (incl upload) %.1fs" % (t1 - t0, t2 - t1), "nodes/ray %.1f prims/ray %.1f exact/node %.5f" % (st["node_visits"] / st["ray_segments"], st["prim_tests"] / st["ray_segments"], st["exact_steps"] / max(1, st["node_visits"])), flush=True)
    img, st = rb.render(size, T, spp, 50, W, seed=3, return_stats=True)
    print(name, "spp", spp, "T", T, "ms %.1f Mrays/s %.1f extend_ms %.1f shade_ms %.1f iters %d rays %d" % (st["render_ms"], st["ray_segments"] / st["render_ms"] / 1e3, st["extend_ms"], st["shade_ms"], st["iterations"], st["ray_segments"]), flush=True)
    from PIL import Image
    Image.fromarray(rb.to_rgb8_gamma2(img)).resize((960, 540)).save("gpurun_out/%s.png" % name.split(":")[1])
    W.close()
