"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes binding over oracle/_build/liboracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.  PARITY UNPINNED (see oracle_core.hpp).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
ASSETS = os.path.join(os.path.dirname(_HERE), "assets")

_lib = None


def build(force: bool = False) -> str:
    """Compile the C++ restatement (g++, a few seconds)."""
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "oracle_app.hpp", "oracle_core.hpp", "oracle_rng.hpp")]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_last_error.restype = C.c_char_p
        L.orc_world_create.restype = C.c_void_p
        L.orc_world_create.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_uint32, C.c_uint32, C.c_void_p]
        L.orc_world_create_flat.restype = C.c_void_p
        L.orc_world_create_flat.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p,
                                            C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        L.orc_set_trace_time.argtypes = [C.c_float]
        L.orc_world_perlin.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.orc_world_destroy.argtypes = [C.c_void_p]
        L.orc_world_info.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_world_export.argtypes = [C.c_void_p] + [C.c_void_p] * 12
        L.orc_world_image.restype = C.c_void_p
        L.orc_world_image.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_trace.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_primary_rays.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_render.restype = C.c_double
        L.orc_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_uint64, C.c_uint64, C.c_int,
                                 C.c_int, C.c_void_p, C.c_void_p]
        L.orc_render_planes.restype = C.c_double
        L.orc_render_planes.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_uint64, C.c_uint64, C.c_int,
                                        C.c_int, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        L.orc_to_rgb8_gamma2.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.orc_aabb_hit_cond.restype = C.c_int
        L.orc_aabb_hit_cond.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float]
        L.orc_obj_load.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.orc_rng_kat.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_xoroshiro_from_seed.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_probe_math.argtypes = [C.c_float, C.c_void_p]
        L.orc_probe_ln.argtypes = [C.c_float]
        L.orc_probe_ln.restype = C.c_float
        L.orc_split_work_tasks.restype = C.c_int
        L.orc_split_work_tasks.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64]
        L.orc_philox.argtypes = [C.c_uint32] * 6 + [C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


_earth_cache = None


def earth_rgb():
    """assets/earthmap.jpg decoded with PIL (libjpeg-turbo) to RGB8, row 0 = top."""
    global _earth_cache
    if _earth_cache is None:
        from PIL import Image
        im = Image.open(os.path.join(ASSETS, "earthmap.jpg")).convert("RGB")
        _earth_cache = np.ascontiguousarray(np.asarray(im, dtype=np.uint8))
    return _earth_cache


@dataclass
class FlatScene:
    """The flattened scene exactly as NodeRef::finish + BoundingVolumeHierarchy::new leave it."""
    kinds: np.ndarray          # u32 [N]   0 sphere, 1 rect, 2 triangle
    item_mat: np.ndarray       # u32 [N]
    geom: np.ndarray           # f32 [N,24]
    nodes_box: np.ndarray      # f32 [N-1,6]
    nodes_lra: np.ndarray      # u32 [N-1,3]  left,right (0x80000000|item = leaf), axis
    root: int
    depth: int
    mats: np.ndarray           # u32 [M,3]  kind, texture, param bits
    texs: np.ndarray           # u32 [T,8]  kind, image, rgb bits, inv_frequency bits, even, odd
    wrap: np.ndarray = None    # f32 [N,10] flags (1 Transformation, 2 Animation, 4 Volume), offset xyz, y_sine, y_cosine, velocity xyz, -1/density
    perlins: np.ndarray = None  # u32 [P,1536] 256 gradient vectors (float bits) + perm_x, perm_y, perm_z
    images: list = field(default_factory=list)   # list of u8 [h,w,3]
    cam: np.ndarray = None     # f32 [23]
    bg: np.ndarray = None      # f32 [4]
    light: np.ndarray = None   # f32 [7]
    default_size: tuple = (0, 0)


class OracleWorld:
    def __init__(self, handle):
        if not handle:
            raise RuntimeError("oracle: " + lib().orc_last_error().decode())
        self.h = handle
        info = np.zeros(11, dtype=np.uint64)
        lib().orc_world_info(self.h, _p(info))
        self.info = info
        self.n_items = int(info[0])
        self.n_nodes = int(info[1])
        self.default_size = (int(info[4]), int(info[5]))

    @classmethod
    def named(cls, name: str) -> "OracleWorld":
        e = earth_rgb()
        w = cls(lib().orc_world_create(name.encode(), os.path.join(ASSETS, "suzanne.obj").encode(),
                                       os.path.join(ASSETS, "cube.obj").encode(), e.shape[1], e.shape[0], _p(e)))
        w._keep = e
        return w

    @classmethod
    def from_flat(cls, kinds, item_mat, geom, mats, texs, images, cam, bg, light, wrap=None, perlins=None) -> "OracleWorld":
        kinds = np.ascontiguousarray(kinds, dtype=np.uint32)
        item_mat = np.ascontiguousarray(item_mat, dtype=np.uint32)
        geom = np.ascontiguousarray(geom, dtype=np.float32)
        mats = np.ascontiguousarray(mats, dtype=np.uint32)
        texs = np.ascontiguousarray(texs, dtype=np.uint32)
        images = [np.ascontiguousarray(im, dtype=np.uint8) for im in images]
        wh = np.array([[im.shape[1], im.shape[0]] for im in images], dtype=np.uint32).reshape(-1, 2)
        ptrs = (C.c_void_p * max(1, len(images)))(*[im.ctypes.data for im in images])
        cam = np.ascontiguousarray(cam, dtype=np.float32)
        bg = np.ascontiguousarray(bg, dtype=np.float32)
        light = np.ascontiguousarray(light, dtype=np.float32)
        wrap = None if wrap is None else np.ascontiguousarray(wrap, dtype=np.float32).reshape(len(kinds), 10)
        perlins = np.zeros((0, 1536), np.uint32) if perlins is None else np.ascontiguousarray(perlins, dtype=np.uint32).reshape(-1, 1536)
        texs = texs.reshape(-1, 8)
        return cls(lib().orc_world_create_flat(len(kinds), _p(kinds), _p(item_mat), _p(geom), len(mats), _p(mats),
                                               len(texs), _p(texs), len(images), _p(wh), ptrs, _p(cam), _p(bg),
                                               _p(light), _p(wrap), len(perlins), _p(perlins)))

    def close(self):
        if self.h:
            lib().orc_world_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def export(self) -> FlatScene:
        n, nn = self.n_items, self.n_nodes
        nm, nt, ni = int(self.info[6]), int(self.info[7]), int(self.info[8])
        fs = FlatScene(
            kinds=np.zeros(n, np.uint32), item_mat=np.zeros(n, np.uint32), geom=np.zeros((n, 24), np.float32),
            nodes_box=np.zeros((nn, 6), np.float32), nodes_lra=np.zeros((nn, 3), np.uint32),
            root=int(self.info[2]), depth=int(self.info[3]),
            mats=np.zeros((nm, 3), np.uint32), texs=np.zeros((nt, 8), np.uint32), wrap=np.zeros((n, 10), np.float32),
            perlins=np.zeros((int(self.info[10]), 1536), np.uint32),
            cam=np.zeros(23, np.float32), bg=np.zeros(4, np.float32), light=np.zeros(7, np.float32),
            default_size=self.default_size)
        wh = np.zeros((max(ni, 1), 2), np.uint32)
        lib().orc_world_export(self.h, _p(fs.kinds), _p(fs.item_mat), _p(fs.geom), _p(fs.nodes_box), _p(fs.nodes_lra),
                               _p(fs.mats), _p(fs.texs), _p(wh), _p(fs.cam), _p(fs.bg), _p(fs.light), _p(fs.wrap))
        for i in range(len(fs.perlins)):
            lib().orc_world_perlin(self.h, i, _p(fs.perlins[i]))
        for i in range(ni):
            w, h = int(wh[i, 0]), int(wh[i, 1])
            buf = (C.c_uint8 * (w * h * 3)).from_address(lib().orc_world_image(self.h, i))
            fs.images.append(np.frombuffer(buf, dtype=np.uint8).reshape(h, w, 3).copy())
        return fs

    def trace(self, org, dir, t_min=0.001, t_max=np.inf, counters=False, time=0.0):
        lib().orc_set_trace_time(time)
        org = np.ascontiguousarray(org, dtype=np.float32).reshape(-1, 3)
        dir = np.ascontiguousarray(dir, dtype=np.float32).reshape(-1, 3)
        n = org.shape[0]
        hit = np.zeros(n, np.uint32)
        t = np.zeros(n, np.float32)
        nrm = np.zeros((n, 3), np.float32)
        ff = np.zeros(n, np.uint8)
        uv = np.zeros((n, 2), np.float32)
        cnt = np.zeros(2, np.uint64)
        lib().orc_trace(self.h, n, _p(org), _p(dir), t_min, t_max, _p(hit), _p(t), _p(nrm), _p(ff), _p(uv), _p(cnt))
        out = dict(hit_id=hit, t=t, normal=nrm, front_face=ff, uv=uv)
        if counters:
            out["node_tests"] = int(cnt[0])
            out["prim_tests"] = int(cnt[1])
        return out

    def primary_rays(self, width, height):
        org = np.zeros((width * height, 3), np.float32)
        dir = np.zeros((width * height, 3), np.float32)
        lib().orc_primary_rays(self.h, width, height, _p(org), _p(dir))
        return org, dir

    def render(self, width, height, spp, max_depth=50, threads=1, seed=1, mode="faithful", render_mode="default"):
        out = np.zeros((height, width, 3), np.float32)
        stats = np.zeros(3, np.uint64)
        secs = lib().orc_render(self.h, width, height, spp, max_depth, threads, seed,
                                {"faithful": 0, "replay": 1}[mode], {"default": 0, "normals": 1}[render_mode],
                                _p(out), _p(stats))
        return out, dict(seconds=secs, rays=int(stats[0]), node_tests=int(stats[1]), prim_tests=int(stats[2]))


def _render_planes(self, width, height, spp, max_depth, threads, seed, mode, lane_first, lane_stride, render_mode="default"):
    """Un-normalised sum of this rank's lane planes (multi-GPU split of render(); see rt_render_planes_device)."""
    out = np.zeros((height, width, 3), np.float32)
    stats = np.zeros(3, np.uint64)
    secs = lib().orc_render_planes(self.h, width, height, spp, max_depth, threads, seed, {"faithful": 0, "replay": 1}[mode],
                                   {"default": 0, "normals": 1}[render_mode], lane_first, lane_stride, _p(out), _p(stats))
    return out, dict(seconds=secs, rays=int(stats[0]))


OracleWorld.render_planes = _render_planes


def to_rgb8_gamma2(rgb):
    rgb = np.ascontiguousarray(rgb, dtype=np.float32)
    out = np.zeros(rgb.shape, np.uint8)
    lib().orc_to_rgb8_gamma2(_p(rgb), rgb.size // 3, _p(out))
    return out
