"""ctypes binding over the C ABI of libb200rt.so (include/b200rt.h).

This is the only way Python reaches the render hot path; the library is hand-written sm_100a CUDA and
has no CPU fallback — loading fails loudly if the extension has not been built.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

RT_LEAF_BIT = 0x80000000
RT_MISS = 0xFFFFFFFF
RT_EINVAL_PY = -1


class RtItem(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("material", C.c_uint32), ("geom", C.c_float * 24), ("wrap_flags", C.c_uint32),
                ("transform", C.c_float * 5), ("velocity", C.c_float * 3), ("neg_inv_density", C.c_float)]


class RtBvhNode(C.Structure):
    _fields_ = [("aabb_min", C.c_float * 3), ("aabb_max", C.c_float * 3), ("left", C.c_uint32), ("right", C.c_uint32),
                ("axis_id", C.c_uint32)]


class RtMaterial(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("texture", C.c_uint32), ("param", C.c_float)]


class RtTexture(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("image", C.c_uint32), ("rgb", C.c_float * 3), ("inv_frequency", C.c_float),
                ("even", C.c_uint32), ("odd", C.c_uint32)]


class RtImage(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("rgb8", C.c_void_p)]


class RtPerlin(C.Structure):
    _fields_ = [("ranvec", C.c_void_p), ("perm_x", C.c_void_p), ("perm_y", C.c_void_p), ("perm_z", C.c_void_p)]


class RtCamera(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("upper_left_corner", C.c_float * 3), ("unit_right", C.c_float * 3),
                ("unit_up", C.c_float * 3), ("scaled_right", C.c_float * 3), ("scaled_up", C.c_float * 3),
                ("lens_radius", C.c_float), ("time_start", C.c_float), ("time_end", C.c_float),
                ("shutter_pace", C.c_float * 2)]


class RtSceneDesc(C.Structure):
    _fields_ = [("n_items", C.c_uint32), ("items", C.c_void_p), ("n_nodes", C.c_uint32), ("nodes", C.c_void_p),
                ("root", C.c_uint32), ("n_materials", C.c_uint32), ("materials", C.c_void_p),
                ("n_textures", C.c_uint32), ("textures", C.c_void_p), ("n_images", C.c_uint32), ("images", C.c_void_p),
                ("n_perlins", C.c_uint32), ("perlins", C.c_void_p), ("camera", RtCamera), ("background_is_sky", C.c_uint32), ("background_rgb", C.c_float * 3),
                ("has_light", C.c_uint32), ("light_plane", C.c_uint32), ("light_dist", C.c_float),
                ("light_r0", C.c_float * 2), ("light_r1", C.c_float * 2)]


class RtRenderParams(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("samples_per_pixel", C.c_uint64),
                ("max_depth", C.c_int32), ("thread_count", C.c_uint32), ("render_mode", C.c_uint32),
                ("seed", C.c_uint64)]


class RtStats(C.Structure):
    _fields_ = [("ray_segments", C.c_uint64), ("paths", C.c_uint64), ("iterations", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("render_ms", C.c_double), ("extend_ms", C.c_double),
                ("shade_ms", C.c_double), ("node_visits", C.c_uint64), ("prim_tests", C.c_uint64), ("exact_steps", C.c_uint64),
                ("redo_rays", C.c_uint64), ("max_stack", C.c_uint32), ("stack_capacity", C.c_uint32),
                ("slots", C.c_uint32), ("pipes", C.c_uint32), ("merge_ms", C.c_double), ("warp_rounds", C.c_uint64), ("busy_lane_rounds", C.c_uint64), ("rounds_hist", C.c_uint64 * 24)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_}
        d["rounds_hist"] = list(self.rounds_hist)
        return d


ITEM_DTYPE = np.dtype([("kind", "<u4"), ("material", "<u4"), ("geom", "<f4", (24,)), ("wrap_flags", "<u4"),
                       ("transform", "<f4", (5,)), ("velocity", "<f4", (3,)), ("neg_inv_density", "<f4")])
NODE_DTYPE = np.dtype([("aabb_min", "<f4", (3,)), ("aabb_max", "<f4", (3,)), ("left", "<u4"), ("right", "<u4"),
                       ("axis_id", "<u4")])
MATERIAL_DTYPE = np.dtype([("kind", "<u4"), ("texture", "<u4"), ("param", "<f4")])
TEXTURE_DTYPE = np.dtype([("kind", "<u4"), ("image", "<u4"), ("rgb", "<f4", (3,)), ("inv_frequency", "<f4"), ("even", "<u4"),
                          ("odd", "<u4")])
assert ITEM_DTYPE.itemsize == C.sizeof(RtItem) and NODE_DTYPE.itemsize == C.sizeof(RtBvhNode)
assert MATERIAL_DTYPE.itemsize == C.sizeof(RtMaterial) and TEXTURE_DTYPE.itemsize == C.sizeof(RtTexture)

EXPORTED = ["rt_last_error", "rt_device_count", "rt_scene_create", "rt_scene_destroy", "rt_scene_device_bytes", "rt_scene_set_ray_time",
            "rt_trace", "rt_trace_device", "rt_render", "rt_render_device", "rt_render_moments_device", "rt_render_planes_device", "rt_render_multi", "rt_render_multi_device",
            "rt_film_scale_device", "rt_to_rgb8_gamma2_device", "rt_set_option", "rt_presplit_refs", "rt_scene_accelerator_nodes", "rt_shutdown"]

_lib = None


class B200RTError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"b200rt error {code}: {msg}")
        self.code = code


def lib():
    """Load libb200rt.so (building it in-tree with nvcc if stale).  Never falls back to anything else."""
    global _lib
    if _lib is None:
        path = _build.build_cuda()
        L = C.CDLL(path)
        L.rt_last_error.restype = C.c_char_p
        L.rt_device_count.argtypes = [C.POINTER(C.c_int)]
        L.rt_scene_create.argtypes = [C.POINTER(RtSceneDesc), C.c_int, C.POINTER(C.c_void_p)]
        L.rt_scene_destroy.argtypes = [C.c_void_p]
        L.rt_scene_destroy.restype = None
        L.rt_scene_device_bytes.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.rt_scene_set_ray_time.argtypes = [C.c_void_p, C.c_float]
        L.rt_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_float, C.c_float] + [C.c_void_p] * 5
        L.rt_trace_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_float, C.c_float] + [C.c_void_p] * 6
        L.rt_render.argtypes = [C.c_void_p, C.POINTER(RtRenderParams), C.c_void_p, C.POINTER(RtStats)]
        L.rt_render_device.argtypes = [C.c_void_p, C.POINTER(RtRenderParams), C.c_void_p, C.c_void_p, C.POINTER(RtStats)]
        L.rt_render_moments_device.argtypes = [C.c_void_p, C.POINTER(RtRenderParams), C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(RtStats)]
        L.rt_render_planes_device.argtypes = [C.c_void_p, C.POINTER(RtRenderParams), C.c_uint32, C.c_uint32, C.c_void_p,
                                              C.c_void_p, C.POINTER(RtStats)]
        L.rt_render_multi.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(RtRenderParams), C.c_void_p, C.POINTER(RtStats)]
        L.rt_render_multi_device.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(RtRenderParams), C.POINTER(C.c_void_p), C.POINTER(RtStats)]
        L.rt_film_scale_device.argtypes = [C.c_void_p, C.c_uint64, C.c_float, C.c_void_p]
        L.rt_to_rgb8_gamma2_device.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        L.rt_set_option.argtypes = [C.c_char_p, C.c_int64]
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise B200RTError(rc, lib().rt_last_error().decode(errors="replace"))


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def presplit_refs(boxes6, tri9, is_tri, budget: int, device: int = -1):
    """rt_presplit_refs (test aid): the references row f4's pre-split makes of n elements; device < 0 = the host implementation."""
    boxes6 = np.ascontiguousarray(boxes6, dtype=np.float32); tri9 = np.ascontiguousarray(tri9, dtype=np.float32)
    is_tri = np.ascontiguousarray(is_tri, dtype=np.uint8)
    n = boxes6.shape[0]
    cap = max(int(budget), n)
    items = np.empty(cap, dtype=np.uint32); boxes = np.empty((cap, 6), dtype=np.float32)
    out_n = C.c_uint64(0)
    L = lib()
    L.rt_presplit_refs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
    check(L.rt_presplit_refs(boxes6.ctypes.data, tri9.ctypes.data, is_tri.ctypes.data, n, int(budget), int(device), items.ctypes.data,
                             boxes.ctypes.data, C.byref(out_n)))
    return items[:out_n.value].copy(), boxes[:out_n.value].copy()


def set_option(key: str, value: int):
    check(lib().rt_set_option(key.encode(), int(value)))


def shutdown():
    """Release the library's per-device render workspaces (rt_shutdown)."""
    if _lib is not None:
        check(_lib.rt_shutdown())


def device_count() -> int:
    n = C.c_int(0)
    check(lib().rt_device_count(C.byref(n)))
    return n.value


def textures_from_u32(texs):
    """[T, 8] u32 rows (kind, image, r, g, b bits, inv_frequency bits, even, odd) -> RtTexture records."""
    texs = np.asarray(texs, np.uint32).reshape(-1, 8)
    t = np.zeros(len(texs), TEXTURE_DTYPE)
    t["kind"], t["image"] = texs[:, 0], texs[:, 1]
    t["rgb"] = texs[:, 2:5].copy().view(np.float32)
    t["inv_frequency"] = texs[:, 5].copy().view(np.float32)
    t["even"], t["odd"] = texs[:, 6], texs[:, 7]
    return t


def items_from_flat(kinds, item_mat, geom, wrap=None):
    n = len(kinds)
    items = np.zeros(n, ITEM_DTYPE)
    items["kind"] = kinds
    items["material"] = item_mat
    items["geom"] = np.asarray(geom, np.float32).reshape(n, 24)
    if wrap is not None:
        w = np.asarray(wrap, np.float32).reshape(n, 10)
        items["wrap_flags"] = w[:, 0].astype(np.uint32)
        items["transform"] = w[:, 1:6]
        items["velocity"] = w[:, 6:9]
        items["neg_inv_density"] = w[:, 9]
    return items


def perlins_to_c(perlins):
    """[P, 1536] u32 rows (768 gradient float bits, perm_x | perm_y | perm_z) -> (RtPerlin array, keepalive)."""
    perlins = np.ascontiguousarray(np.asarray(perlins if perlins is not None else [], np.uint32).reshape(-1, 1536))
    arr = (RtPerlin * max(1, len(perlins)))()
    for i in range(len(perlins)):
        base = perlins.ctypes.data + i * 1536 * 4
        arr[i].ranvec = base
        arr[i].perm_x = base + 768 * 4
        arr[i].perm_y = base + 1024 * 4
        arr[i].perm_z = base + 1280 * 4
    return arr, perlins


def make_desc(kinds, item_mat, geom, nodes_box, nodes_lra, root, mats, texs, images, cam, bg, light, wrap=None, perlins=None):
    """Assemble an RtSceneDesc from flat numpy arrays.  Returns (desc, keepalive)."""
    n = len(kinds)
    items = items_from_flat(kinds, item_mat, geom, wrap)
    nn = len(nodes_lra)
    nodes = np.zeros(nn, NODE_DTYPE)
    if nn:
        nb = np.asarray(nodes_box, np.float32).reshape(nn, 6)
        nl = np.asarray(nodes_lra, np.uint32).reshape(nn, 3)
        nodes["aabb_min"] = nb[:, :3]
        nodes["aabb_max"] = nb[:, 3:]
        nodes["left"] = nl[:, 0]
        nodes["right"] = nl[:, 1]
        nodes["axis_id"] = nl[:, 2]
    mats = np.asarray(mats, np.uint32).reshape(-1, 3)
    m = np.zeros(len(mats), MATERIAL_DTYPE)
    m["kind"] = mats[:, 0]
    m["texture"] = mats[:, 1]
    m["param"] = mats[:, 2].copy().view(np.float32)
    t = textures_from_u32(texs)
    images = [np.ascontiguousarray(im, np.uint8) for im in images]
    imgs = (RtImage * max(1, len(images)))()
    for i, im in enumerate(images):
        imgs[i].width = im.shape[1]
        imgs[i].height = im.shape[0]
        imgs[i].rgb8 = im.ctypes.data
    d = RtSceneDesc()
    d.n_items = n
    d.items = items.ctypes.data
    d.n_nodes = nn
    d.nodes = nodes.ctypes.data if nn else None
    d.root = int(root)
    d.n_materials = len(m)
    d.materials = m.ctypes.data
    d.n_textures = len(t)
    d.textures = t.ctypes.data if len(t) else None
    d.n_images = len(images)
    d.images = C.cast(imgs, C.c_void_p)
    parr, pkeep = perlins_to_c(perlins)
    d.n_perlins = len(pkeep)
    d.perlins = C.cast(parr, C.c_void_p)
    cam = np.asarray(cam, np.float32)
    for i, name in enumerate(["position", "upper_left_corner", "unit_right", "unit_up", "scaled_right", "scaled_up"]):
        setattr(d.camera, name, (C.c_float * 3)(*cam[3 * i:3 * i + 3]))
    d.camera.lens_radius = float(cam[18])
    d.camera.time_start = float(cam[19])
    d.camera.time_end = float(cam[20])
    d.camera.shutter_pace = (C.c_float * 2)(float(cam[21]), float(cam[22]))
    bg = np.asarray(bg, np.float32)
    d.background_is_sky = 1 if bg[0] != 0 else 0
    d.background_rgb = (C.c_float * 3)(*bg[1:4])
    light = np.asarray(light, np.float32)
    d.has_light = 1 if light[0] != 0 else 0
    d.light_plane = int(light[1])
    d.light_dist = float(light[2])
    d.light_r0 = (C.c_float * 2)(float(light[3]), float(light[4]))
    d.light_r1 = (C.c_float * 2)(float(light[5]), float(light[6]))
    return d, (items, nodes, m, t, images, imgs, parr, pkeep)


class DeviceScene:
    """Owner of an RtScene* (device-resident packed scene)."""

    def __init__(self, desc: RtSceneDesc, device: int = 0, keepalive=None):
        h = C.c_void_p()
        check(lib().rt_scene_create(C.byref(desc), device, C.byref(h)))
        self.h = h
        self.device = device
        self.owned = True
        del keepalive  # rt_scene_create copies everything

    @classmethod
    def borrowed(cls, handle, device: int = 0, owner=None, generation: int = 0) -> "DeviceScene":
        """Wrap an RtScene* owned by someone else (the host mirror's World).  The wrapper keeps its owner alive and refuses to
        touch the handle once the owner has released or replaced its device copy (`owner._scene_generation` moves on)."""
        self = cls.__new__(cls)
        self._h = handle
        self.device = device
        self.owned = False
        self._owner = owner
        self._generation = generation
        return self

    @property
    def h(self):
        owner = getattr(self, "_owner", None)
        if owner is not None and owner._scene_generation != self._generation:
            raise B200RTError(RT_EINVAL_PY, "stale scene handle: the World released or re-created its device scene")
        return self._h

    @h.setter
    def h(self, value):
        self._h = value

    @classmethod
    def from_flat(cls, fs, device: int = 0) -> "DeviceScene":
        """fs: any object with the flat-scene attributes (kinds, item_mat, geom, nodes_box, nodes_lra, root, ...)."""
        desc, keep = make_desc(fs.kinds, fs.item_mat, fs.geom, fs.nodes_box, fs.nodes_lra, fs.root, fs.mats, fs.texs,
                               fs.images, fs.cam, fs.bg, fs.light, getattr(fs, "wrap", None), getattr(fs, "perlins", None))
        return cls(desc, device, keep)

    def close(self):
        if getattr(self, "_h", None):
            if getattr(self, "owned", True):
                lib().rt_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def device_bytes(self) -> int:
        b = C.c_uint64(0)
        check(lib().rt_scene_device_bytes(self.h, C.byref(b)))
        return b.value

    def accelerator_nodes(self):
        """rt_scene_accelerator_nodes (test aid): (nodes as a (n, 16) uint32 array, root reference, depth) of the accelerator tree."""
        L = lib()
        L.rt_scene_accelerator_nodes.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        n, root, depth = C.c_uint64(0), C.c_uint32(0), C.c_uint32(0)
        check(L.rt_scene_accelerator_nodes(self.h, None, 0, C.byref(n), C.byref(root), C.byref(depth)))
        nodes = np.empty((n.value, 16), dtype=np.uint32)
        if n.value:
            check(L.rt_scene_accelerator_nodes(self.h, nodes.ctypes.data, n.value, C.byref(n), C.byref(root), C.byref(depth)))
        return nodes, root.value, depth.value

    def set_ray_time(self, time: float):
        check(lib().rt_scene_set_ray_time(self.h, float(time)))

    def trace(self, org, dir, t_min=0.001, t_max=float("inf")):
        org = np.ascontiguousarray(org, np.float32).reshape(-1, 3)
        dir = np.ascontiguousarray(dir, np.float32).reshape(-1, 3)
        n = org.shape[0]
        hit = np.zeros(n, np.uint32)
        t = np.zeros(n, np.float32)
        nrm = np.zeros((n, 3), np.float32)
        ff = np.zeros(n, np.uint8)
        uv = np.zeros((n, 2), np.float32)
        check(lib().rt_trace(self.h, _p(org), _p(dir), n, t_min, t_max, _p(hit), _p(t), _p(nrm), _p(ff), _p(uv)))
        return dict(hit_id=hit, t=t, normal=nrm, front_face=ff, uv=uv)

    def render(self, width, height, spp, max_depth=50, thread_count=1, seed=1, render_mode=0):
        p = RtRenderParams(width, height, spp, max_depth, thread_count, render_mode, seed)
        out = np.zeros((height, width, 3), np.float32)
        st = RtStats()
        check(lib().rt_render(self.h, C.byref(p), _p(out), C.byref(st)))
        return out, st.as_dict()

    def render_device(self, d_out_ptr: int, width, height, spp, max_depth=50, thread_count=1, seed=1, render_mode=0,
                      stream: int = 0):
        p = RtRenderParams(width, height, spp, max_depth, thread_count, render_mode, seed)
        st = RtStats()
        check(lib().rt_render_device(self.h, C.byref(p), C.c_void_p(d_out_ptr), C.c_void_p(stream), C.byref(st)))
        return st.as_dict()

    def render_moments_device(self, d_out_ptr: int, d_sumsq_ptr: int, width, height, spp, max_depth=50, thread_count=1, seed=1,
                              render_mode=0, stream: int = 0):
        """Film + per-pixel sum over samples of (r^2 + g^2 + b^2)."""
        p = RtRenderParams(width, height, spp, max_depth, thread_count, render_mode, seed)
        st = RtStats()
        check(lib().rt_render_moments_device(self.h, C.byref(p), C.c_void_p(d_out_ptr), C.c_void_p(d_sumsq_ptr), C.c_void_p(stream),
                                             C.byref(st)))
        return st.as_dict()

    def render_planes_device(self, d_out_ptr: int, lane_first, lane_stride, width, height, spp, max_depth=50,
                             thread_count=1, seed=1, render_mode=0, stream: int = 0):
        p = RtRenderParams(width, height, spp, max_depth, thread_count, render_mode, seed)
        st = RtStats()
        check(lib().rt_render_planes_device(self.h, C.byref(p), lane_first, lane_stride, C.c_void_p(d_out_ptr),
                                            C.c_void_p(stream), C.byref(st)))
        return st.as_dict()


def render_multi(scenes, width, height, spp, max_depth=50, thread_count=1, seed=1, render_mode=0):
    """rt_render_multi: one blocking call over len(scenes) GPUs (DeviceScene objects on different devices); host film + per-GPU stats."""
    n = len(scenes)
    p = RtRenderParams(width, height, spp, max_depth, thread_count, render_mode, seed)
    hs = (C.c_void_p * n)(*[C.cast(s.h, C.c_void_p).value for s in scenes])
    out = np.zeros((height, width, 3), np.float32)
    st = (RtStats * n)()
    check(lib().rt_render_multi(hs, n, C.byref(p), _p(out), st))
    return out, [s.as_dict() for s in st]
