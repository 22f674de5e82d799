"""Oracle pin (ii): the C++ oracle against a SECOND, independent restatement of the id-deciding functions, written in numpy float32
from the Rust alone (oracle/second_restatement.py: BVH builder hittable.rs:360-427, Aabb::hit_cond aabb.rs:65-78, the ordered
recursion hittable.rs:429-473, the sphere / rect / triangle hits).  Both must agree bit for bit — tree, boxes, hit ids, t, normals,
front faces — on >= 10^6 rays of C1-C3 (pixel-centre primary rays, random rays, adversarial rays).  A transcription slip in either
restatement would have to be made twice, in two languages, to go unnoticed.  No GPU involved."""
import numpy as np
import pytest

from helpers import adversarial_rays, random_rays
from oracle import second_restatement as sr


def second_tree(fs):
    geom = np.asarray(fs.geom, np.float32).reshape(len(fs.kinds), 24)
    boxes = [sr.element_box(int(k), g) for k, g in zip(fs.kinds, geom)]
    return sr.build_bvh([b[0] for b in boxes], [b[1] for b in boxes])


@pytest.mark.parametrize("name,size,n_random", [("config:c3", (400, 400), 100000), ("config:c1", (600, 400), 150000), ("config:c2", (640, 360), 170000)])
def test_oracle_equals_second_restatement(name, size, n_random, oracle_world):
    W = oracle_world(name)
    fs = W.export()
    nodes, root = second_tree(fs)
    # the tree: same pre-order numbering, children, split axes and (bitwise) boxes as the oracle's builder
    ref = lambda c: ((c - sr.LEAF) | 0x80000000) if c >= sr.LEAF else c
    lra = np.array([[ref(l), ref(r), a] for (_, _, a, l, r) in nodes], np.uint32)
    box = np.array([np.concatenate([mn, mx]) for (mn, mx, _, _, _) in nodes], np.float32)
    assert root == fs.root and np.array_equal(lra, np.asarray(fs.nodes_lra, np.uint32).reshape(-1, 3))
    assert np.array_equal(box.view(np.uint32), np.asarray(fs.nodes_box, np.float32).reshape(-1, 6).view(np.uint32))
    # the rays
    org, d = W.primary_rays(*size)
    ro, rd = random_rays(fs, n_random, 3)
    ao, ad = adversarial_rays(fs, 300, 5)
    org, d = np.concatenate([org, ro, ao]), np.concatenate([d, rd, ad])
    got = sr.trace(fs.kinds, fs.geom, nodes, root, org, d)
    exp = W.trace(org, d)
    assert (exp["hit_id"] != 0xFFFFFFFF).sum() > 0.5 * len(org)
    assert np.array_equal(got["hit_id"], exp["hit_id"])
    assert np.array_equal(got["t"].view(np.uint32), exp["t"].view(np.uint32))
    assert np.array_equal(got["front_face"], exp["front_face"])
    nan = np.isnan(exp["normal"])
    assert np.array_equal(np.isnan(got["normal"]), nan)
    assert np.array_equal(got["normal"].view(np.uint32)[~nan], exp["normal"].view(np.uint32)[~nan])


def test_ray_count_of_the_cross_check():
    assert 400 * 400 + 600 * 400 + 640 * 360 + 100000 + 150000 + 170000 >= 10 ** 6
