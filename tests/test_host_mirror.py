"""Host-side logic of the product (no GPU): the C++ mirror of the reference API must flatten scenes exactly like the
oracle's literal restatement (element order = hit ids, baked transforms, BVH topology and boxes, tables, camera), and
the C ABI must load, export every declared symbol and fail loudly without a CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import raytracinginaweekend_b200 as rb
from raytracinginaweekend_b200 import _cabi, api, build, dist
from helpers import F32, random_scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIELDS = ["kinds", "item_mat", "geom", "nodes_box", "nodes_lra", "mats", "texs", "cam", "bg", "light", "wrap", "perlins"]


def same_flat(a, b):
    for f in FIELDS:
        x, y = getattr(a, f), getattr(b, f)
        assert x.shape == y.shape, f
        assert np.array_equal(np.ascontiguousarray(x).view(np.uint32), np.ascontiguousarray(y).view(np.uint32)), f
    assert a.root == b.root and a.depth == b.depth
    assert len(a.images) == len(b.images) and all(np.array_equal(i, j) for i, j in zip(a.images, b.images))


WORLDS = ["demo:suzanne", "demo:final_scene1", "demo:simple_plane", "demo:earth_mapped", "demo:defocus_blur", "demo:cornell_box", "demo:moving_spheres",
          "demo:cornell_box_smoke", "demo:perlin_spheres", "demo:final_scene2",
          "config:c1", "config:c2", "config:c3", "config:c5"]


@pytest.mark.parametrize("name", WORLDS)
def test_world_flatten_matches_oracle(name, oracle_world):
    w = rb.World.named(name)
    same_flat(w.export(), oracle_world(name).export())
    assert (w.default_size.width, w.default_size.height) == oracle_world(name).default_size


def test_million_triangle_world_flatten_matches_oracle(oracle):
    w = rb.World.named("config:c4")
    o = oracle.OracleWorld.named("config:c4")
    assert w.n_items == 999_945 and w.n_nodes == 999_944
    same_flat(w.export(), o.export())


@pytest.mark.parametrize("seed", range(6))
def test_bvh_build_matches_oracle_on_random_scenes(seed, oracle):
    sc = random_scene(seed, n_spheres=5 + 13 * seed, n_tris=40 * seed, n_rects=seed, with_light=seed % 2 == 1, n_boxes=3 * (seed % 3),
                      wrap_fraction=0.3 * (seed % 2), checker=seed % 3 == 0, volume_fraction=0.2 * (seed >= 3), marble=seed >= 4)
    same_flat(rb.World.from_elements(**sc).export(), oracle.OracleWorld.from_flat(**sc).export())


def test_bvh_ties_are_stable(oracle):
    """Equal sort keys: the reference uses a stable sort, so ties keep element order (affects hit ids on coincident geometry)."""
    sc = random_scene(1, n_spheres=40, n_tris=0, n_rects=0)
    sc["geom"][:, 0] = 1.0
    sc["geom"][::2, 1] = 2.0
    sc["geom"][:, 3] = 0.5
    same_flat(rb.World.from_elements(**sc).export(), oracle.OracleWorld.from_flat(**sc).export())


def test_obj_loader_matches_oracle(oracle):
    for name, n in (("suzanne.obj", 968), ("cube.obj", 12)):
        tris = rb.load_obj_mesh(os.path.join(api.ASSETS, name))
        cnt = np.zeros(1, np.uint64)
        ref = np.zeros((n, 24), F32)
        oracle.lib().orc_obj_load(os.path.join(api.ASSETS, name).encode(), oracle._p(cnt), oracle._p(ref), n)
        assert tris.shape == (n, 24) and np.array_equal(tris.view(np.uint32), ref.view(np.uint32))


def test_unknown_worlds_are_rejected():
    for name in ("demo:nope", "config:c9", ""):
        with pytest.raises(RuntimeError):
            rb.World.named(name)


def test_gamma_matches_oracle(oracle):
    rng = np.random.default_rng(0)
    v = np.concatenate([rng.uniform(-0.5, 2.0, 3000), [np.nan, 0.0, 1.0, np.inf, -np.inf, 0.25]]).astype(F32).reshape(-1, 3)
    assert np.array_equal(rb.to_rgb8_gamma2(v), oracle.to_rgb8_gamma2(v))


# ---------------------------------------------------------------- C ABI surface
def declared_symbols(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(rt_[a-z0-9_]+|rtw_[a-z0-9_]+)\s*\(", txt)))


def test_cabi_exports_every_declared_symbol():
    L = _cabi.lib()
    syms = declared_symbols("b200rt.h")
    assert set(syms) == set(_cabi.EXPORTED), (syms, _cabi.EXPORTED)
    for s in syms:
        assert getattr(L, s) is not None
    H = api.host_lib()
    hs = declared_symbols("rtiaw_host.h")
    assert len(hs) >= 10
    for s in hs:
        assert getattr(H, s) is not None


def test_struct_layouts_match_header():
    assert C.sizeof(_cabi.RtItem) == 144 and C.sizeof(_cabi.RtBvhNode) == 36
    assert C.sizeof(_cabi.RtMaterial) == 12 and C.sizeof(_cabi.RtTexture) == 32
    assert C.sizeof(_cabi.RtCamera) == 92 and C.sizeof(_cabi.RtRenderParams) == 40


def test_ctypes_structs_match_the_c_header(tmp_path):
    """Compile include/b200rt.h as plain C and compare sizeof / offsetof of every struct with the ctypes mirror in _cabi.py."""
    import subprocess
    structs = {"RtItem": ["kind", "geom", "wrap_flags", "transform", "velocity", "neg_inv_density"], "RtBvhNode": ["aabb_min", "left", "axis_id"],
               "RtMaterial": ["texture", "param"], "RtTexture": ["image", "rgb", "inv_frequency", "odd"], "RtImage": ["height", "rgb8"],
               "RtPerlin": ["perm_x", "perm_z"], "RtCamera": ["upper_left_corner", "scaled_up", "lens_radius", "time_end", "shutter_pace"],
               "RtSceneDesc": ["items", "nodes", "root", "materials", "textures", "images", "n_perlins", "perlins", "camera", "background_is_sky",
                               "background_rgb", "has_light", "light_dist", "light_r1"],
               "RtRenderParams": ["height", "samples_per_pixel", "max_depth", "thread_count", "render_mode", "seed"],
               "RtStats": ["paths", "kernel_launches", "render_ms", "shade_ms", "exact_steps", "redo_rays", "max_stack", "stack_capacity"]}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "b200rt.h"', 'int main(void) {']
    for name, fields in structs.items():
        lines.append(f'  printf("{name} %zu", sizeof({name}));')
        for f in fields:
            lines.append(f'  printf(" %zu", offsetof({name}, {f}));')
        lines.append('  printf("\\n");')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    out = subprocess.check_output([str(exe)], text=True).strip().splitlines()
    for line in out:
        parts = line.split()
        name, vals = parts[0], [int(x) for x in parts[1:]]
        ct = getattr(_cabi, name)
        assert C.sizeof(ct) == vals[0], name
        for f, off in zip(structs[name], vals[1:]):
            assert getattr(ct, f).offset == off, (name, f)


def test_native_libraries_are_in_tree():
    assert os.path.dirname(build.LIB_CUDA) == os.path.join(ROOT, "raytracinginaweekend_b200")
    assert os.path.exists(build.LIB_CUDA) and os.path.exists(build.LIB_HOST)


def test_no_cpu_fallback_without_device():
    """Without a GPU every compute entry point must fail loudly (RT_ECUDA), never compute on the CPU."""
    try:
        n = _cabi.device_count()
    except _cabi.B200RTError as e:
        assert e.code == -2
        n = 0
    if n > 0:
        pytest.skip("a CUDA device is present")
    sc = random_scene(0)
    w = rb.World.from_elements(**sc)
    with pytest.raises(RuntimeError, match="rt_scene_create"):
        rb.render(rb.Size2i(8, 8), 1, 1, 5, w)
    with pytest.raises(RuntimeError):
        w.hit(np.zeros((1, 3), F32), np.array([[0, 0, -1]], F32))
    fs = w.export()
    with pytest.raises(_cabi.B200RTError) as ei:
        _cabi.DeviceScene.from_flat(fs)
    assert ei.value.code == -2


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the package may import, include, link or load it."""
    pkg = os.path.join(ROOT, "raytracinginaweekend_b200")
    bad = re.compile(r"(^\s*(import|from)\s+oracle\b|#\s*include\s*[\"<][^\">]*oracle|liboracle|oracle/_build|oracle\.binding)", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert not bad.search(txt), f


# ---------------------------------------------------------------- lane split (rendering.rs:222-237)
def test_lane_partition():
    assert dist.n_tasks(8, 3) == 3 and dist.n_tasks(4, 64) == 4
    lanes = [list(range(*dist.lane_partition(8, 64, r, 4)[:1], dist.n_tasks(8, 64), 4)) for r in range(4)]
    assert sorted(sum(lanes, [])) == list(range(8))
    assert [dist.lane_samples(4, 10, i) for i in range(4)] == [3, 3, 2, 2]
    assert dist.lane_partition(2, 64, 3, 4)[2] == 0      # more ranks than lanes: this rank owns nothing


def test_every_option_key_is_documented_in_the_header():
    """rt_set_option keys accepted by the library (csrc/b200rt.cu) all appear, quoted, in include/b200rt.h's option list."""
    import re
    src = open(os.path.join(ROOT, "raytracinginaweekend_b200", "csrc", "b200rt.cu")).read()
    hdr = open(os.path.join(ROOT, "include", "b200rt.h")).read()
    keys = set(re.findall(r'k == "([a-z_]+)"', src))
    assert len(keys) >= 20
    missing = sorted(k for k in keys if '"%s"' % k not in hdr)
    assert not missing, missing
