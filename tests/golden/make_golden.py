"""Generates the committed golden vectors from the ORACLE (the reference cannot run here: no rustc; see oracle/).

    python tests/golden/make_golden.py

For each of the scenes below it stores, in tests/golden/<scene>.npz:
  * pixel-centre primary rays (lens disabled) + random rays + adversarial rays, and the oracle's Scene::hit answers
    (hit id, t, normal, front_face, uv);
  * a REPLAY-mode render (Philox streams, so the GPU walks the same paths) at a small size.
Used by tests/test_golden_cpu.py (pins the oracle against regressions) and tests/test_gpu_parity.py (the CUDA path must
reproduce them without the oracle in the loop).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from helpers import adversarial_rays, random_rays  # noqa: E402
from oracle import binding as ob  # noqa: E402

SCENES = {"config:c1": (48, 32), "config:c2": (48, 27), "config:c3": (32, 32), "demo:defocus_blur": (40, 24), "demo:simple_plane": (40, 24),
          "demo:cornell_box": (32, 32), "demo:moving_spheres": (38, 32), "demo:cornell_box_smoke": (32, 32), "demo:perlin_spheres": (40, 30),
          "demo:final_scene2": (32, 32)}
TRACE_TIME = 0.3   # Ray::time of the traced rays (only demo:moving_spheres / demo:final_scene2 have Animation leaves)
RENDER = dict(spp=4, max_depth=50, threads=2, seed=20261018)


def main():
    only = sys.argv[1:]
    for name, (w, h) in SCENES.items():
        if only and name not in only:
            continue
        W = ob.OracleWorld.named(name)
        fs = W.export()
        po, pd = W.primary_rays(w, h)
        ro, rd = random_rays(fs, 1500, 42)
        ao, ad = adversarial_rays(fs, 120, 7)
        org = np.concatenate([po, ro, ao]); dir = np.concatenate([pd, rd, ad])
        hit = W.trace(org, dir, time=TRACE_TIME)
        img, st = W.render(w, h, RENDER["spp"], RENDER["max_depth"], RENDER["threads"], RENDER["seed"], "replay")
        out = os.path.join(HERE, name.replace(":", "_") + ".npz")
        np.savez_compressed(out, org=org, dir=dir, hit_id=hit["hit_id"], t=hit["t"], normal=hit["normal"], front_face=hit["front_face"],
                            uv=hit["uv"], image=img, image_size=np.array([w, h]), rays=np.array([st["rays"]], np.uint64),
                            n_primary=np.array([len(po)]))
        print(name, "rays", len(org), "hits", int((hit["hit_id"] != 0xFFFFFFFF).sum()), "->", os.path.relpath(out), os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
