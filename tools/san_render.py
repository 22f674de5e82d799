"""Small renders that exercise every round-2 code path (chains through few slots, forced parking, proof in k_shade, ties by visit rank,
restarts, volumes / moving elements, staged upload / download) — the command that runs under compute-sanitizer."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import raytracinginaweekend_b200 as rb
from raytracinginaweekend_b200 import _cabi
for name, w, h, spp, T in (("config:c2", 96, 54, 6, 3), ("config:c5", 64, 36, 4, 2), ("demo:cornell_box_smoke", 32, 32, 4, 2), ("demo:moving_spheres", 48, 32, 4, 2), ("config:c1", 60, 40, 4, 2)):
    W = rb.World.named(name)
    ref = None
    for slots, budget in ((160 << 20, 64), (w * h + 7, -64), (w * h // 2 + 3, -64), (160 << 20, 0)):
        _cabi.set_option("max_slots", slots); _cabi.set_option("tail_budget", budget)
        img, st = rb.render(rb.Size2i(w, h), T, spp, 50, W, seed=9, return_stats=True)
        if ref is None: ref = img
        same = bool(np.array_equal(ref.view(np.uint32), img.view(np.uint32)))
        print(name, slots, budget, "same_film", same, "iters", st["iterations"], "rays", st["ray_segments"], flush=True)
        assert same
    W.close()
print("OK")
