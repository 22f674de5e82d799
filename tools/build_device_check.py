"""Accelerator tree built on the device ("build_device" 1) vs on the host cores (0): scene hand-off time, traversal work per ray,
Mrays/s, and the film (which must not depend on the accelerator at all).   python tools/build_device_check.py config:c5 [w h spp]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import raytracinginaweekend_b200 as rb
from raytracinginaweekend_b200 import _cabi
name = sys.argv[1] if len(sys.argv) > 1 else "config:c5"
w, h, spp = (int(a) for a in sys.argv[2:5]) if len(sys.argv) > 4 else (960, 540, 8)
W = rb.World.named(name)
films = []
for dev in (0, 1, 0, 1):
    _cabi.set_option("build_device", dev)
    W.release_device()
    t0 = time.time(); sc = W.device_scene(0); t1 = time.time()
    _cabi.set_option("count", 1)
    _, c = rb.render(rb.Size2i(w, h), 1, 1, 50, W, seed=3, return_stats=True)
    _cabi.set_option("count", 0)
    best = 0.0
    for rep in range(2):
        film, st = rb.render(rb.Size2i(w, h), spp, spp, 50, W, seed=3, return_stats=True)
        best = max(best, st["ray_segments"] / st["render_ms"] / 1e3)
    films.append(np.asarray(film))
    print("%s build_device %d | create %.3f s | %6.1f MB | %7.1f Mrays/s | nodes/ray %6.2f prims/ray %5.2f | redo %d" % (
        name, dev, t1 - t0, sc.device_bytes() / 1e6, best, c["node_visits"] / c["ray_segments"], c["prim_tests"] / c["ray_segments"], c["redo_rays"]), flush=True)
same = all(np.array_equal(films[0].view(np.uint32), f.view(np.uint32)) for f in films[1:])
print("films bit-identical:", same, flush=True)
