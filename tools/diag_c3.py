import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracinginaweekend_b200 as rb
from raytracinginaweekend_b200 import _cabi
np.set_printoptions(precision=9, floatmode="unique", linewidth=200)
for name in ["config_c3", "config_c1", "config_c2", "demo_defocus_blur", "demo_simple_plane"]:
    g = np.load(f"tests/golden/{name}.npz")
    w = rb.World.named(name.replace("_", ":", 1))
    for mode in (0, 1, 2):
        _cabi.set_option("traversal", mode)
        got = w.hit(g["org"], g["dir"])
        bad = np.nonzero(got["hit_id"] != g["hit_id"])[0]
        print(name, "mode", mode, "bad", len(bad))
        for i in bad[:6]:
            print("  i", i, "org", g["org"][i], "dir", g["dir"][i], "got", got["hit_id"][i], got["t"][i], "ref", g["hit_id"][i], g["t"][i])
