"""CPU test of the two lemmas behind round 2's handling of fast-path hits, against the ORACLE (the literal
restatement of the reference traversal) and the independent numpy primitive tests of oracle/second_restatement.py.  No GPU involved.

  tie rule   : among the elements that accept exactly the smallest t, the reference returns the one its ordered DFS visits first, and
               that order is fixed by the tree and the signs of the ray direction (hittable.rs:441-445) — the per-octant visit rank.
  exact walk : if t is that minimum and t0_i <= t < t1_i (slab distances in the reference's own arithmetic, aabb.rs:65-78) on every axis
               of every ancestor box of the winner, the reference reaches it, so it returns exactly (winner, t).

The brute force evaluates EVERY element's test for every ray (no tree), so it knows the true minimum and the true tie sets.  C2 is the
scene where ties matter: the Q4 fan triangles (v0, ORIGIN, v2) of a mirror-symmetric mesh lie IN its symmetry plane and overlap."""
import numpy as np
import pytest

from oracle import second_restatement as sr

F = np.float32
LEAF = 0x80000000


def visit_ranks(fs):
    """rank[octant][item] = position of the element in the reference's ordered DFS for direction signs `octant` (bit i = d_i > 0)."""
    lra = np.asarray(fs.nodes_lra, np.int64).reshape(-1, 3)
    n_items = len(fs.kinds)
    ranks = np.zeros((8, n_items), np.int64)
    for octant in range(8):
        nxt, stack = 0, [int(fs.root)]
        while stack:
            r = stack.pop()
            if r & LEAF:
                ranks[octant, r & ~LEAF] = nxt
                nxt += 1
                continue
            left, right, axis = (int(v) for v in lra[r])
            left_first = (octant >> axis) & 1
            stack.append(right if left_first else left)   # visited second
            stack.append(left if left_first else right)
    return ranks


def parents(fs):
    lra = np.asarray(fs.nodes_lra, np.int64).reshape(-1, 3)
    parent_node = np.full(len(lra), -1, np.int64)
    parent_item = np.full(len(fs.kinds), -1, np.int64)
    for i, (left, right, _) in enumerate(lra):
        for ch in (int(left), int(right)):
            if ch & LEAF:
                parent_item[ch & ~LEAF] = i
            else:
                parent_node[ch] = i
    return parent_node, parent_item


def walk_passes(box, parent_node, first_parent, o, d, t):
    """t0_i <= t < t1_i on every axis of every ancestor box, with slab_axis' arithmetic: a = (mn - o) / d, b = (mx - o) / d, (t0, t1) = a < b ? (a, b) : (b, a)."""
    node = first_parent
    with np.errstate(all="ignore"):
        while node >= 0:
            a = (box[node, 0:3] - o) / d
            b = (box[node, 3:6] - o) / d
            lt = a < b
            t0, t1 = np.where(lt, a, b), np.where(lt, b, a)
            if not np.all((t0 <= t) & (t < t1)):
                return False
            node = parent_node[node]
    return True


def rays_for_c2(W, rng):
    o, d = W.primary_rays(160, 90)
    o, d = o.reshape(-1, 3), d.reshape(-1, 3)
    outs_o, outs_d = [o], [d]
    for cx in (-2.5, 2.5):   # rays aimed at the symmetry plane of each mesh instance (model x = 0), from both sides, all octants
        n = 20000
        target = np.stack([np.full(n, cx), rng.uniform(0.05, 1.6, n), rng.uniform(-1.3, 0.85, n)], 1).astype(F)
        origin = (target + rng.normal(size=(n, 3)) * np.array([2.0, 1.0, 2.0])).astype(F)
        dd = target - origin
        outs_o.append(origin)
        outs_d.append((dd / np.linalg.norm(dd, axis=1, keepdims=True)).astype(F))
    return np.concatenate(outs_o), np.concatenate(outs_d)


def test_reference_returns_the_first_visited_of_the_closest_tied_elements(oracle_world):
    W = oracle_world("config:c2")
    fs = W.export()
    kinds = np.asarray(fs.kinds)
    geom = np.asarray(fs.geom, F).reshape(len(kinds), 24)
    box = np.asarray(fs.nodes_box, F).reshape(-1, 6)
    ranks = visit_ranks(fs)
    parent_node, parent_item = parents(fs)
    rng = np.random.default_rng(11)
    org, dirs = rays_for_c2(W, rng)
    ref = W.trace(org, dirs)
    n_ties = n_checked = n_walk_failed = 0
    start, end = F(0.001), F(np.inf)
    for lo in range(0, len(org), 6000):
        O, D = org[lo:lo + 6000], dirs[lo:lo + 6000]
        T = np.full((len(kinds), len(O)), np.inf, F)
        for e in range(len(kinds)):
            if kinds[e] == 0:
                ok, t = sr.hit_sphere(geom[e], O, D, start, end)
            elif kinds[e] == 1:
                ok, t = sr.hit_rect(geom[e], O, D, start, end)
            else:
                ok, t, _, _ = sr.hit_triangle(geom[e], O, D, start, end)
            T[e] = np.where(ok, t, np.inf)
        tmin = T.min(axis=0)
        for j in np.nonzero(np.isfinite(tmin))[0]:
            tied = np.nonzero(T[:, j] == tmin[j])[0]
            octant = int(D[j, 0] > 0) | (int(D[j, 1] > 0) << 1) | (int(D[j, 2] > 0) << 2)
            winner = tied[np.argmin(ranks[octant, tied])]
            n_ties += len(tied) > 1
            if not walk_passes(box, parent_node, parent_item[winner], O[j], D[j], tmin[j]):
                n_walk_failed += 1      # proves nothing: the GPU path traces such a ray again on the reference tree
                continue
            n_checked += 1
            assert ref["hit_id"][lo + j] == winner, (lo + j, tied, winner, ref["hit_id"][lo + j])
            assert ref["t"][lo + j].view(np.uint32) == tmin[j].view(np.uint32)
        # rays without any accepted element: the reference misses too
        miss = ~np.isfinite(tmin)
        assert np.all(ref["hit_id"][lo:lo + len(O)][miss] == 0xFFFFFFFF)
    assert n_ties >= 40, n_ties                       # the symmetry-plane fans produce exact ties
    assert n_checked > 0.95 * (n_checked + n_walk_failed)
