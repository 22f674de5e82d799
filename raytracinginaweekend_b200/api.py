"""Python face of the host mirror (librtiaw_host.so, include/rtiaw_host.h) — same names as the reference's `src/lib` API.

    world = World.named("config:c2")                       # create_world_* (src/app/main.rs:25-37)
    pixels = render(Size2i(1920, 1080), thread_count, samples_per_pixel, max_depth, world, RenderMode.Default)
                                                           # render() (src/lib/rendering.rs:121-128)
    hits = world.hit(origins, directions, (0.001, inf))    # Scene::hit (src/lib/hittable.rs:130-137), batched

Everything that touches rays runs in libb200rt.so (hand-written sm_100a CUDA); there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import enum
import os
from dataclasses import dataclass, field

import numpy as np

from . import _cabi
from . import build as _build

ASSETS = os.path.join(_build.ROOT, "assets")


class RenderMode(enum.IntEnum):  # src/lib/rendering.rs:94-98
    Default = 0
    Normals = 1


@dataclass(frozen=True)
class Size2i:  # src/lib/size2i.rs
    width: int
    height: int

    def aspect_ratio(self) -> float:
        return np.float32(self.height) / np.float32(self.width)

    def count(self) -> int:
        return self.width * self.height


@dataclass
class FlatScene:
    """The flattened scene as NodeRef::finish + BoundingVolumeHierarchy::new leave it (same fields as the oracle export)."""
    kinds: np.ndarray
    item_mat: np.ndarray
    geom: np.ndarray
    nodes_box: np.ndarray
    nodes_lra: np.ndarray
    root: int
    depth: int
    mats: np.ndarray
    texs: np.ndarray
    wrap: np.ndarray = None      # [n, 10]: RT_WRAP_* flags, offset xyz, y_sine, y_cosine, velocity xyz, -1/density
    perlins: np.ndarray = None   # [P, 1536] u32: 768 gradient float bits, perm_x | perm_y | perm_z
    images: list = field(default_factory=list)
    cam: np.ndarray = None
    bg: np.ndarray = None
    light: np.ndarray = None
    default_size: tuple = (0, 0)


_host = None
_earth = None


def host_lib():
    global _host
    if _host is None:
        _cabi.lib()  # load libb200rt.so first (fails loudly if the CUDA extension is missing)
        L = C.CDLL(_build.build_host())
        L.rtw_last_error.restype = C.c_char_p
        L.rtw_world_create.restype = C.c_void_p
        L.rtw_world_create.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_uint32, C.c_uint32, C.c_void_p]
        L.rtw_world_create_elements.restype = C.c_void_p
        L.rtw_world_create_elements.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p,
                                                C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        L.rtw_world_perlin.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        L.rtw_world_destroy.argtypes = [C.c_void_p]
        L.rtw_world_destroy.restype = None
        L.rtw_world_info.argtypes = [C.c_void_p, C.c_void_p]
        L.rtw_world_export.argtypes = [C.c_void_p] * 9
        L.rtw_world_image.restype = C.c_void_p
        L.rtw_world_image.argtypes = [C.c_void_p, C.c_uint32]
        L.rtw_world_device.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        L.rtw_world_release_device.argtypes = [C.c_void_p]
        L.rtw_world_upload_bytes.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.rtw_render.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_uint64, C.c_uint64, C.c_int32, C.c_uint32,
                                 C.c_uint64, C.c_int, C.c_void_p, C.POINTER(_cabi.RtStats)]
        L.rtw_hit.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int] + [C.c_void_p] * 5
        L.rtw_load_obj.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_uint64]
        _host = L
    return _host


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _check(rc):
    if rc != 0:
        raise RuntimeError("rtiaw host: " + host_lib().rtw_last_error().decode(errors="replace"))


def earth_rgb8() -> np.ndarray:
    """assets/earthmap.jpg decoded to RGB8 (row 0 = top); JPEG decode is caller-side in the reference too
    (image::load in src/app/worlds/world_builder.rs:19-31)."""
    global _earth
    if _earth is None:
        from PIL import Image
        _earth = np.ascontiguousarray(np.asarray(Image.open(os.path.join(ASSETS, "earthmap.jpg")).convert("RGB"), dtype=np.uint8))
    return _earth


def load_obj_mesh(path: str) -> np.ndarray:
    """load_obj_mesh (src/app/obj_loader.rs:7): [n, 24] floats (positions, normals, uvs), fan-triangulation quirk included."""
    n = C.c_uint64(0)
    _check(host_lib().rtw_load_obj(path.encode(), C.byref(n), None, 0))
    out = np.zeros((n.value, 24), np.float32)
    _check(host_lib().rtw_load_obj(path.encode(), C.byref(n), _p(out), n.value))
    return out


class World:
    """`World` (src/lib/rendering.rs:12-17): camera + scene + background + optional light-sampling provider."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("rtiaw host: " + host_lib().rtw_last_error().decode(errors="replace"))
        self.h = C.c_void_p(handle)
        self._scene_generation = 0  # bumped whenever the device scene is released / re-created: borrowed wrappers check it
        self._scene_device = None
        info = np.zeros(11, np.uint64)
        _check(host_lib().rtw_world_info(self.h, _p(info)))
        self.info = info
        self.n_items, self.n_nodes = int(info[0]), int(info[1])
        self.default_size = Size2i(int(info[4]), int(info[5]))

    @classmethod
    def named(cls, name: str) -> "World":
        need_earth = name in ("demo:earth_mapped", "demo:final_scene2", "config:c3")  # JPEG decode is caller-side (image crate in the reference)
        e = earth_rgb8() if need_earth else None
        return cls(host_lib().rtw_world_create(name.encode(), os.path.join(ASSETS, "suzanne.obj").encode(),
                                               os.path.join(ASSETS, "cube.obj").encode(),
                                               e.shape[1] if need_earth else 0, e.shape[0] if need_earth else 0, _p(e)))

    @classmethod
    def from_elements(cls, kinds, item_mat, geom, mats, texs, images, cam, bg, light, wrap=None, perlins=None) -> "World":
        """Scene::new(BoundingVolumeHierarchy::new(elements)) from SurfaceGeometry arrays (flat layouts of b200rt.h)."""
        kinds = np.ascontiguousarray(kinds, np.uint32)
        item_mat = np.ascontiguousarray(item_mat, np.uint32)
        geom = np.ascontiguousarray(geom, np.float32).reshape(len(kinds), 24)
        mats_u = np.asarray(mats, np.uint32).reshape(-1, 3)
        m = np.zeros(len(mats_u), _cabi.MATERIAL_DTYPE)
        m["kind"], m["texture"], m["param"] = mats_u[:, 0], mats_u[:, 1], mats_u[:, 2].copy().view(np.float32)
        t = _cabi.textures_from_u32(texs)
        images = [np.ascontiguousarray(im, np.uint8) for im in images]
        imgs = (_cabi.RtImage * max(1, len(images)))()
        for i, im in enumerate(images):
            imgs[i].width, imgs[i].height, imgs[i].rgb8 = im.shape[1], im.shape[0], im.ctypes.data
        cam = np.ascontiguousarray(cam, np.float32)
        bg = np.ascontiguousarray(bg, np.float32)
        light = np.ascontiguousarray(light, np.float32)
        wrap = None if wrap is None else np.ascontiguousarray(wrap, np.float32).reshape(len(kinds), 10)
        parr, pkeep = _cabi.perlins_to_c(perlins)
        return cls(host_lib().rtw_world_create_elements(len(kinds), _p(kinds), _p(item_mat), _p(geom), len(m), _p(m), len(t),
                                                        _p(t), len(images), C.cast(imgs, C.c_void_p), _p(cam), _p(bg), _p(light),
                                                        _p(wrap), len(pkeep), C.cast(parr, C.c_void_p)))

    def close(self):
        if getattr(self, "h", None):
            self._scene_generation += 1
            host_lib().rtw_world_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def export(self) -> FlatScene:
        n, nn = self.n_items, self.n_nodes
        nm, nt, ni = int(self.info[6]), int(self.info[7]), int(self.info[8])
        items = np.zeros(n, _cabi.ITEM_DTYPE)
        nodes = np.zeros(max(nn, 1), _cabi.NODE_DTYPE)
        m = np.zeros(nm, _cabi.MATERIAL_DTYPE)
        t = np.zeros(max(nt, 1), _cabi.TEXTURE_DTYPE)
        wh = np.zeros((max(ni, 1), 2), np.uint32)
        cam, bg, light = np.zeros(23, np.float32), np.zeros(4, np.float32), np.zeros(7, np.float32)
        _check(host_lib().rtw_world_export(self.h, _p(items), _p(nodes), _p(m), _p(t), _p(wh), _p(cam), _p(bg), _p(light)))
        nodes, t = nodes[:nn], t[:nt]
        fs = FlatScene(
            kinds=items["kind"].copy(), item_mat=items["material"].copy(), geom=items["geom"].copy(),
            nodes_box=np.concatenate([nodes["aabb_min"], nodes["aabb_max"]], axis=1) if nn else np.zeros((0, 6), np.float32),
            nodes_lra=np.stack([nodes["left"], nodes["right"], nodes["axis_id"]], axis=1) if nn else np.zeros((0, 3), np.uint32),
            root=int(self.info[2]), depth=int(self.info[3]),
            mats=np.stack([m["kind"], m["texture"], m["param"].copy().view(np.uint32)], axis=1),
            texs=np.concatenate([np.stack([t["kind"], t["image"]], axis=1), t["rgb"].copy().view(np.uint32),
                                 t["inv_frequency"].copy().view(np.uint32)[:, None], np.stack([t["even"], t["odd"]], axis=1)], axis=1) if nt else np.zeros((0, 8), np.uint32),
            wrap=np.concatenate([items["wrap_flags"].astype(np.float32)[:, None], items["transform"], items["velocity"],
                                 items["neg_inv_density"][:, None]], axis=1),
            cam=cam, bg=bg, light=light, default_size=(self.default_size.width, self.default_size.height))
        for i in range(ni):
            w, h = int(wh[i, 0]), int(wh[i, 1])
            buf = (C.c_uint8 * (w * h * 3)).from_address(host_lib().rtw_world_image(self.h, i))
            fs.images.append(np.frombuffer(buf, dtype=np.uint8).reshape(h, w, 3).copy())
        n_perlins = int(self.info[10])
        fs.perlins = np.zeros((n_perlins, 1536), np.uint32)
        for i in range(n_perlins):
            _check(host_lib().rtw_world_perlin(self.h, i, _p(fs.perlins[i, :768]), _p(fs.perlins[i, 768:])))
        return fs

    def device_scene(self, device: int = 0) -> "_cabi.DeviceScene":
        """RtScene* owned by this world (uploaded on first use)."""
        if self._scene_device is not None and self._scene_device != device:
            self._scene_generation += 1  # the host mirror keeps ONE device copy: asking for another device replaces it
        h = C.c_void_p()
        _check(host_lib().rtw_world_device(self.h, device, C.byref(h)))
        self._scene_device = device
        return _cabi.DeviceScene.borrowed(h, device, owner=self, generation=self._scene_generation)

    def release_device(self):
        """Drop the device copy; the next render / hit uploads the scene again."""
        self._scene_generation += 1
        self._scene_device = None
        _check(host_lib().rtw_world_release_device(self.h))

    def upload_bytes(self) -> int:
        b = C.c_uint64(0)
        _check(host_lib().rtw_world_upload_bytes(self.h, C.byref(b)))
        return b.value

    def hit(self, org, dir, t_range=(0.001, float("inf")), device: int = 0, time: float = 0.0):
        scene = self.device_scene(device)
        scene.set_ray_time(time)
        org = np.ascontiguousarray(org, np.float32).reshape(-1, 3)
        dir = np.ascontiguousarray(dir, np.float32).reshape(-1, 3)
        n = org.shape[0]
        hit, t = np.zeros(n, np.uint32), np.zeros(n, np.float32)
        nrm, ff, uv = np.zeros((n, 3), np.float32), np.zeros(n, np.uint8), np.zeros((n, 2), np.float32)
        try:
            _check(host_lib().rtw_hit(self.h, n, _p(org), _p(dir), t_range[0], t_range[1], device, _p(hit), _p(t), _p(nrm), _p(ff), _p(uv)))
        finally:
            scene.set_ray_time(0.0)  # Ray::time belongs to this batch only: do not leave it set on the shared scene
        return dict(hit_id=hit, t=t, normal=nrm, front_face=ff, uv=uv)


def render(image_size: Size2i, thread_count: int, samples_per_pixel: int, max_depth: int, world: World,
           render_mode: RenderMode = RenderMode.Default, seed: int = 0x5EED, device: int = 0, return_stats: bool = False):
    """render() of src/lib/rendering.rs:121-128: returns [height, width, 3] linear-RGB float32, row 0 = top."""
    out = np.empty((image_size.height, image_size.width, 3), np.float32)
    st = _cabi.RtStats()
    if world._scene_device is not None and world._scene_device != device:
        world._scene_generation += 1  # the World re-creates its device scene on the other GPU: earlier borrowed handles die
    world._scene_device = device
    _check(host_lib().rtw_render(world.h, image_size.width, image_size.height, thread_count, samples_per_pixel, max_depth,
                                 int(render_mode), seed, device, _p(out), C.byref(st)))
    return (out, st.as_dict()) if return_stats else out


def to_rgb8_gamma2(rgb: np.ndarray) -> np.ndarray:
    """Color::to_rgb8_gamma2 (src/lib/color.rs:43-48): u8(clamp(0, 255, sqrt(c) * 256)), NaN -> 0."""
    with np.errstate(invalid="ignore"):
        v = np.sqrt(rgb.astype(np.float32)) * np.float32(256.0)
        c = np.where(v > 255.0, np.float32(255.0), np.where(v < 0.0, np.float32(0.0), v))
        return np.where(np.isnan(c), 0, c).astype(np.uint8)
