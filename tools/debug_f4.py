import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import raytracinginaweekend_b200 as rb
from raytracinginaweekend_b200 import _cabi
from oracle import binding as oracle
from helpers import F32, random_scene, random_rays
sc = random_scene(61, n_spheres=6, n_tris=12, n_rects=0)
z = -3.0
tri = lambda a, b, c: np.array(list(a) + list(b) + list(c) + [0, 0, 1] * 3 + [0] * 6, F32)
tr = np.nonzero(sc["kinds"] == 2)[0]
sc["geom"][tr[0]] = tri((0, 0, z), (1, 0, z), (0, 1, z))
sc["geom"][tr[1]] = tri((0.1, 0.1, z), (1.1, 0.1, z), (0.1, 1.1, z))
sc["geom"][tr[2]] = tri((-2, -2, z - 1), (-1, -2, z - 1), (-2, -1, z - 1))
sp = np.nonzero(sc["kinds"] == 0)[0]
sc["geom"][sp[0], :4] = [3, 0, -5, 1]
sc["geom"][sp[1], :4] = [3, 0, -5, 1]
print("tr", tr[:3], "sp", sp[:2])
w, o = rb.World.from_elements(**sc), oracle.OracleWorld.from_flat(**sc)
fs = o.export()
print("nodes_lra", fs.nodes_lra.tolist())
org0, dir0 = random_rays(fs, 20000, 5)
aim = np.array([[0.3, 0.3, z], [0.5, 0.4, z], [-1.7, -1.8, z - 1], [3, 0, -5], [3.2, 0.1, -5]], F32)
org1 = np.tile(np.array([[0.2, 0.1, 4.0]], F32), (len(aim), 1)); dir1 = aim - org1
org2 = aim + np.array([0, 0, 6], F32); dir2 = np.tile(np.array([[0, 0, -1]], F32), (len(aim), 1))
org, dir = np.concatenate([org0, org1, org2]), np.concatenate([dir0, dir1, dir2])
ref = o.trace(org, dir)
for mode in (0, 2, 3):
    _cabi.set_option("traversal", mode)
    got = w.hit(org, dir)
    bad = np.nonzero(got["hit_id"] != ref["hit_id"])[0]
    print("mode", mode, "mismatch", len(bad))
    for i in bad[:10]:
        print("  ray", i, org[i], dir[i], "got", got["hit_id"][i], got["t"][i], "ref", ref["hit_id"][i], ref["t"][i], "kind", sc["kinds"][got["hit_id"][i]] if got["hit_id"][i] < 100 else None)
for it in (9, 13, 2):
    print("item", it, "kind", sc["kinds"][it], "geom", sc["geom"][it][:12])
_cabi.set_option("traversal", 1)
got = w.hit(org, dir)
print("mode 1 mismatches", np.nonzero(got["hit_id"] != ref["hit_id"])[0])
dn = dir / np.linalg.norm(dir, axis=1, keepdims=True)
refn = o.trace(org, dn.astype(F32))
for mode in (1, 2):
    _cabi.set_option("traversal", mode)
    got = w.hit(org, dn.astype(F32))
    print("unit dirs mode", mode, np.nonzero(got["hit_id"] != refn["hit_id"])[0])
print("nodes_box", fs.nodes_box.tolist())
