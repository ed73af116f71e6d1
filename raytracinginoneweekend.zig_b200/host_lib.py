"""ctypes binding of host/ -> _build/librtw_host.so: the C++ twin of the reference's Zig host
(scene API, the scene builders of src/main.zig:124-293, Camera.init, flatten, PPM/PNG helpers)."""
import ctypes as C
import os

import numpy as np

from . import abi, build

_lib = None
ASSET_EARTH = os.path.join(build.ROOT, "assets", "sekaichizu.png")

# scene ids: 1..6 = the reference's `scene` constant (src/main.zig:310); 7 = config C3; 8 = config C4
SCENE_RANDOM, SCENE_TWO_SPHERES, SCENE_TWO_PERLIN, SCENE_EARTH, SCENE_SIMPLE_LIGHT, SCENE_CORNELL = 1, 2, 3, 4, 5, 6
SCENE_EARTH_GLASS_METAL, SCENE_SPHERE_FIELD = 7, 8


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(build.HOST_LIB) or os.environ.get("RTW_REBUILD") == "1":
        build.build_host()
    L = C.CDLL(build.HOST_LIB)
    dp, u32p = C.POINTER(C.c_double), C.POINTER(C.c_uint32)
    L.rtw_host_last_error.restype = C.c_char_p
    L.rtw_host_scene_create.restype = C.c_void_p
    L.rtw_host_scene_create.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_char_p]
    L.rtw_host_scene_destroy.argtypes = [C.c_void_p]
    L.rtw_host_scene_desc.argtypes = [C.c_void_p, C.POINTER(abi.SceneDesc)]
    L.rtw_host_scene_config.argtypes = [C.c_void_p, dp, u32p]
    L.rtw_host_camera_init.argtypes = [dp, dp, dp] + [C.c_double] * 6 + [C.POINTER(abi.Camera)]
    L.rtw_host_write_ppm.argtypes = [C.c_char_p, C.c_void_p, C.c_uint32, C.c_uint32]
    L.rtw_host_write_png.argtypes = [C.c_char_p, C.c_void_p, C.c_uint32, C.c_uint32]
    L.rtw_host_decode_png.argtypes = [C.c_char_p, u32p, u32p, C.c_void_p, C.c_uint64]
    L.rtw_host_random_real01.argtypes = [C.c_uint64, C.c_int, dp]
    _lib = L
    return L


def _v3(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64)).ctypes.data_as(C.POINTER(C.c_double))


def camera_init(look_from, look_at, vup, vfov, aspect, aperture, focus_dist=10.0, time0=0.0, time1=1.0):
    """Camera.init (src/main.zig:52-89)."""
    cam = abi.Camera()
    load().rtw_host_camera_init(_v3(look_from), _v3(look_at), _v3(vup), vfov, aspect, aperture, focus_dist, time0,
                                time1, C.byref(cam))
    return cam


class HostScene:
    """A scene built by the C++ host twin and flattened to the POD arrays of the C ABI."""

    def __init__(self, scene_id, grid=3, seed=42, asset=None):
        L = load()
        asset = asset or ASSET_EARTH
        h = L.rtw_host_scene_create(scene_id, grid, seed, asset.encode())
        if not h:
            raise RuntimeError(f"rtw_host_scene_create({scene_id}): {L.rtw_host_last_error().decode()}")
        self.h = C.c_void_p(h)
        self.scene_id, self.grid, self.seed = scene_id, grid, seed
        self.desc = abi.SceneDesc()
        L.rtw_host_scene_desc(self.h, C.byref(self.desc))
        f = np.zeros(12)
        u = np.zeros(4, dtype=np.uint32)
        L.rtw_host_scene_config(self.h, f.ctypes.data_as(C.POINTER(C.c_double)), u.ctypes.data_as(C.POINTER(C.c_uint32)))
        self.look_from, self.look_at = f[0:3].copy(), f[3:6].copy()
        self.vfov, self.aperture, self.aspect = float(f[6]), float(f[7]), float(f[8])
        self.background = tuple(float(x) for x in f[9:12])
        self.width, self.height, self.spp, self.max_depth = (int(x) for x in u)

    def __del__(self):
        try:
            if self.h:
                load().rtw_host_scene_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def camera(self, aspect=None):
        """The camera main() builds for this scene (src/main.zig:366-376: vup (0,1,0), focus 10, shutter 0..1)."""
        return camera_init(self.look_from, self.look_at, (0, 1, 0), self.vfov, self.aspect if aspect is None else aspect,
                           self.aperture, 10.0, 0.0, 1.0)


def write_ppm(path, rgb8):
    rgb8 = np.ascontiguousarray(rgb8, dtype=np.uint8)
    rc = load().rtw_host_write_ppm(path.encode(), rgb8.ctypes.data, rgb8.shape[1], rgb8.shape[0])
    if rc:
        raise IOError(f"cannot write {path}")


def write_png(path, rgb8):
    """8-bit RGB PNG, the format of the reference's out.png (src/main.zig:405)."""
    rgb8 = np.ascontiguousarray(rgb8, dtype=np.uint8)
    if load().rtw_host_write_png(path.encode(), rgb8.ctypes.data, rgb8.shape[1], rgb8.shape[0]):
        raise IOError(f"cannot write {path}")


def decode_png(path):
    L = load()
    w, h = C.c_uint32(0), C.c_uint32(0)
    if L.rtw_host_decode_png(path.encode(), C.byref(w), C.byref(h), None, 0):
        raise IOError(L.rtw_host_last_error().decode())
    out = np.empty((h.value, w.value, 4), dtype=np.uint8)
    if L.rtw_host_decode_png(path.encode(), C.byref(w), C.byref(h), out.ctypes.data, out.nbytes):
        raise IOError(L.rtw_host_last_error().decode())
    return out
