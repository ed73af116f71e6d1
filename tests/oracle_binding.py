"""ctypes binding of oracle/_build/liboracle.so — the CPU ORACLE (checker only).

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "_build", "liboracle.so")


def build(force=False):
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("oracle_capi.cpp", "rtw_oracle.hpp", "Makefile")]
    srcs.append(os.path.join(ROOT, "include", "rtw_cuda.h"))
    # rebuild when missing (or when asked): never implicitly under torchrun / on the GPU box
    stale = force or not os.path.exists(LIB_PATH) or (os.environ.get("RTW_REBUILD") == "1" and any(
        os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs))
    if stale:
        subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)
    return LIB_PATH


NATIVE_LIB_PATH = os.path.join(ORACLE_DIR, "_build", "liboracle_native.so")


def _cpu_stamp():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    import hashlib
                    return hashlib.sha1(line.encode()).hexdigest()[:16]
    except OSError:
        pass
    return "unknown"


def use_native_build():
    """Switch this process to the TIMING build of the oracle (-O3 -march=native, contraction allowed), compiling it on
    this machine if it was not compiled here before.  Must be called before the first lib().  Returns a description of
    what will be loaded.  Parity tests never call this: they use the portable -ffp-contract=off build."""
    global LIB_PATH
    assert _lib is None, "use_native_build() must come before the library is loaded"
    stamp_path = NATIVE_LIB_PATH + ".cpu"
    stamp = _cpu_stamp()
    try:
        have = os.path.exists(NATIVE_LIB_PATH) and open(stamp_path).read() == stamp
    except OSError:
        have = False
    if not have:
        try:
            if os.path.exists(NATIVE_LIB_PATH):
                os.remove(NATIVE_LIB_PATH)
            subprocess.run(["make", "-C", ORACLE_DIR, "native"], check=True, capture_output=True, timeout=300)
            with open(stamp_path, "w") as f:
                f.write(stamp)
            have = True
        except Exception as e:  # no compiler on this host: time the portable build and say so
            return f"built g++ -O3 -ffp-contract=off (portable parity build; native build failed: {type(e).__name__})"
    LIB_PATH = NATIVE_LIB_PATH
    return "built on this host with g++ -O3 -march=native (FMA contraction allowed)"


def _abi():
    import rtw_b200
    return rtw_b200.abi


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if LIB_PATH != NATIVE_LIB_PATH:
        build()
    abi = _abi()
    L = C.CDLL(LIB_PATH)
    dp, u32p, u64p, u8p = (C.POINTER(C.c_double), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64),
                           C.POINTER(C.c_uint8))
    L.orc_scene_builtin.restype = C.c_void_p
    L.orc_scene_builtin.argtypes = [C.c_int, C.c_int, C.c_uint64, u8p, C.c_uint32, C.c_uint32]
    L.orc_scene_from_desc.restype = C.c_void_p
    L.orc_scene_from_desc.argtypes = [C.POINTER(abi.SceneDesc)]
    L.orc_scene_free.argtypes = [C.c_void_p]
    L.orc_scene_export.argtypes = [C.c_void_p, C.POINTER(abi.SceneDesc)]
    L.orc_scene_config.argtypes = [C.c_void_p, dp, u32p]
    L.orc_scene_draws_used.restype = C.c_uint64
    L.orc_scene_draws_used.argtypes = [C.c_void_p]
    L.orc_camera_init.argtypes = [dp, dp, dp] + [C.c_double] * 6 + [C.POINTER(abi.Camera)]
    L.orc_trace_rays.argtypes = [C.c_void_p, C.c_uint32, dp, C.c_int, C.c_int, u32p, dp, dp, dp]
    L.orc_primary_hits.argtypes = [C.c_void_p, C.POINTER(abi.Camera), C.c_uint32, C.c_uint32, C.c_int,
                                   C.c_int, u32p, dp, dp]
    L.orc_render.restype = C.c_double
    L.orc_render.argtypes = [C.c_void_p, C.POINTER(abi.Camera), C.POINTER(abi.RenderParams), C.c_int,
                             C.c_int, C.c_int, dp, u8p, u64p, u64p]
    L.orc_num_threads.restype = C.c_int
    L.orc_kat_xoshiro.argtypes = [C.c_uint64, C.c_int, u64p, u64p]
    L.orc_kat_real01.argtypes = [C.c_uint64, C.c_int, dp]
    L.orc_kat_uint_less_than.argtypes = [C.c_uint64, C.c_uint64, C.c_int, u64p]
    L.orc_kat_sphere_uv.argtypes = [dp, dp]
    L.orc_kat_reflectance.restype = C.c_double
    L.orc_kat_reflectance.argtypes = [C.c_double, C.c_double]
    L.orc_kat_reflect.argtypes = [dp, dp, dp]
    L.orc_kat_refract.argtypes = [dp, dp, C.c_double, dp]
    L.orc_kat_resolve.restype = C.c_uint8
    L.orc_kat_resolve.argtypes = [C.c_double, C.c_uint32]
    L.orc_kat_texture.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, dp, dp]
    L.orc_kat_perlin_turb.restype = C.c_double
    L.orc_kat_perlin_turb.argtypes = [C.c_void_p, C.c_int, dp, C.c_int]
    L.orc_kat_bounding_box.argtypes = [C.c_void_p, C.c_uint32, dp, dp]
    L.orc_kat_aabb_hit.argtypes = [dp, dp, dp, C.c_double, C.c_double]
    L.orc_kat_hit_record.argtypes = [C.c_void_p, dp, dp]
    L.orc_kat_hit_records.argtypes = [C.c_void_p, C.c_uint32, dp, C.c_double, C.c_double, dp, u8p]
    L.orc_kat_perlin_noise.restype = C.c_double
    L.orc_kat_perlin_noise.argtypes = [C.c_void_p, C.c_int, dp]
    L.orc_kat_perlin_tables.argtypes = [C.c_void_p, C.c_int, dp, u32p]
    L.orc_kat_scatter.argtypes = [C.c_void_p, C.c_int, dp, dp, C.c_uint64, dp]
    L.orc_kat_scatter_given.argtypes = [C.c_void_p, C.c_int, dp, dp, dp, C.c_double, dp]
    L.orc_kat_get_ray.restype = C.c_uint64
    L.orc_kat_get_ray.argtypes = [C.POINTER(abi.Camera), C.c_uint64, C.c_double, C.c_double, dp]
    L.orc_kat_get_ray_given.argtypes = [C.POINTER(abi.Camera), dp, C.c_double, C.c_double, C.c_double, dp]
    L.orc_kat_ray_color.argtypes = [C.c_void_p, dp, dp, C.c_uint32, C.c_uint64, dp]
    L.orc_kat_samplers.argtypes = [C.c_uint64, C.c_int, C.c_uint32, dp]
    _lib = L
    return L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _d3(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


class OracleScene:
    """A scene held by the oracle (nested graph, f64)."""

    def __init__(self, handle, keep=None):
        if not handle:
            raise RuntimeError("oracle: scene construction failed")
        self.h = C.c_void_p(handle)
        self._keep = keep

    @classmethod
    def builtin(cls, scene_id, grid=3, seed=42, image=None):
        """Scene `scene_id` of the reference (src/main.zig:310,320-362)."""
        L = lib()
        if image is not None:
            img = np.ascontiguousarray(image, dtype=np.uint8)
            h = L.orc_scene_builtin(scene_id, grid, seed, img.ctypes.data_as(C.POINTER(C.c_uint8)),
                                    img.shape[1], img.shape[0])
        else:
            h = L.orc_scene_builtin(scene_id, grid, seed, None, 0, 0)
        return cls(h)

    @classmethod
    def from_desc(cls, desc, keep=None):
        return cls(lib().orc_scene_from_desc(C.byref(desc)), keep)

    def __del__(self):
        try:
            if self.h:
                lib().orc_scene_free(self.h)
                self.h = None
        except Exception:
            pass

    def export(self):
        d = _abi().SceneDesc()
        lib().orc_scene_export(self.h, C.byref(d))
        d._owner = self  # the pointers live as long as the oracle scene
        return d

    def config(self):
        f = np.zeros(12)
        u = np.zeros(4, dtype=np.uint32)
        lib().orc_scene_config(self.h, _dp(f), u.ctypes.data_as(C.POINTER(C.c_uint32)))
        return dict(look_from=f[0:3].copy(), look_at=f[3:6].copy(), vfov=f[6], aperture=f[7], aspect=f[8],
                    background=f[9:12].copy(), width=int(u[0]), height=int(u[1]), spp=int(u[2]),
                    max_depth=int(u[3]))

    def default_camera(self, aspect=None):
        c = self.config()
        return camera_init(c["look_from"], c["look_at"], (0, 1, 0), c["vfov"],
                           c["aspect"] if aspect is None else aspect, c["aperture"], 10.0, 0.0, 1.0)

    def trace_rays(self, rays, precision=64, use_bvh=False):
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 7)
        n = rays.shape[0]
        ids = np.zeros(n, dtype=np.uint32)
        t = np.zeros(n)
        nrm = np.zeros((n, 3))
        uv = np.zeros((n, 2))
        rc = lib().orc_trace_rays(self.h, n, _dp(rays), precision, int(use_bvh),
                                  ids.ctypes.data_as(C.POINTER(C.c_uint32)), _dp(t), _dp(nrm), _dp(uv))
        assert rc == 0
        return ids, t, nrm, uv

    def primary_hits(self, cam, width, height, precision=64, use_bvh=False):
        ids = np.zeros((height, width), dtype=np.uint32)
        t = np.zeros((height, width))
        nrm = np.zeros((height, width, 3))
        rc = lib().orc_primary_hits(self.h, C.byref(cam), width, height, precision, int(use_bvh),
                                    ids.ctypes.data_as(C.POINTER(C.c_uint32)), _dp(t), _dp(nrm))
        assert rc == 0
        return ids, t, nrm

    def render(self, cam, width, height, spp, max_depth=50, background=(0.7, 0.8, 1.0), seed=42,
               precision=64, nthreads=1, continue_stream=False, want_rgb8=True):
        """Returns dict(accum[H,W,3] f64 sums, bottom row first; rgb8[H,W,3] top row first; secs; paths; rays)."""
        abi = _abi()
        p = abi.RenderParams(width=width, height=height, spp_begin=0, spp_end=spp, spp_total=spp,
                             max_depth=max_depth, variant=0, flags=0, seed=seed)
        p.background[:] = list(background)
        accum = np.zeros((height, width, 3))
        rgb8 = np.zeros((height, width, 3), dtype=np.uint8)
        paths = C.c_uint64(0)
        rays = C.c_uint64(0)
        secs = lib().orc_render(self.h, C.byref(cam), C.byref(p), precision, nthreads, int(continue_stream),
                                _dp(accum), rgb8.ctypes.data_as(C.POINTER(C.c_uint8)) if want_rgb8 else None,
                                C.byref(paths), C.byref(rays))
        return dict(accum=accum, rgb8=rgb8, secs=secs, paths=paths.value, rays=rays.value)

    def texture_value(self, tex, u, v, p):
        out = np.zeros(3)
        lib().orc_kat_texture(self.h, tex, u, v, _dp(_d3(p)), _dp(out))
        return out

    def bounding_box(self, i):
        mn, mx = np.zeros(3), np.zeros(3)
        ok = lib().orc_kat_bounding_box(self.h, i, _dp(mn), _dp(mx))
        return (mn, mx) if ok else None

    def hit_record(self, ray7):
        out = np.zeros(12)
        ok = lib().orc_kat_hit_record(self.h, _dp(_d3(ray7)), _dp(out))
        if not ok:
            return None
        return dict(t=out[0], p=out[1:4].copy(), normal=out[4:7].copy(), u=out[7], v=out[8],
                    front_face=bool(out[9]), prim_id=int(out[10]), material=int(out[11]))


    def hit_records(self, rays, t_min=0.001, t_max=float("inf")):
        """world.hit for n rays -> (hit mask[n], out[n,12] = t, p[3], n[3], u, v, front, prim, material)"""
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 7)
        n = rays.shape[0]
        out = np.zeros((n, 12))
        mask = np.zeros(n, dtype=np.uint8)
        lib().orc_kat_hit_records(self.h, n, _dp(rays), t_min, t_max, _dp(out), mask.ctypes.data_as(C.POINTER(C.c_uint8)))
        return mask.astype(bool), out

    def perlin_noise(self, perlin, p):
        return lib().orc_kat_perlin_noise(self.h, perlin, _dp(_d3(p)))

    def perlin_turb(self, perlin, p, depth=7):
        return lib().orc_kat_perlin_turb(self.h, perlin, _dp(_d3(p)), depth)

    def perlin_tables(self, perlin):
        rv = np.zeros((256, 3))
        pm = np.zeros((3, 256), dtype=np.uint32)
        lib().orc_kat_perlin_tables(self.h, perlin, _dp(rv), pm.ctypes.data_as(C.POINTER(C.c_uint32)))
        return rv, pm

    def scatter(self, material, ray7, rec10, seed):
        """Material.scatter + emitted with Rng(seed): dict(ok, attenuation, ray, emitted, draws)"""
        out = np.zeros(14)
        ok = lib().orc_kat_scatter(self.h, material, _dp(_d3(ray7)), _dp(_d3(rec10)), seed, _dp(out))
        return dict(ok=bool(ok), attenuation=out[0:3].copy(), ray=out[3:10].copy(), emitted=out[10:13].copy(), draws=int(out[13]))

    def scatter_given(self, material, ray7, rec10, vec3, xi):
        out = np.zeros(13)
        ok = lib().orc_kat_scatter_given(self.h, material, _dp(_d3(ray7)), _dp(_d3(rec10)), _dp(_d3(vec3)), float(xi), _dp(out))
        return dict(ok=bool(ok), attenuation=out[0:3].copy(), ray=out[3:10].copy(), emitted=out[10:13].copy())

    def ray_color(self, ray7, background, depth, seed):
        out = np.zeros(5)
        lib().orc_kat_ray_color(self.h, _dp(_d3(ray7)), _dp(_d3(background)), depth, seed, _dp(out))
        return dict(color=out[0:3].copy(), rays=int(out[3]), draws=int(out[4]))


def get_ray(cam, seed, s, t):
    out = np.zeros(7)
    draws = lib().orc_kat_get_ray(C.byref(cam), seed, s, t, _dp(out))
    return out, int(draws)


def get_ray_given(cam, disk2, time_xi, s, t):
    out = np.zeros(7)
    lib().orc_kat_get_ray_given(C.byref(cam), _dp(_d3(disk2)), float(time_xi), float(s), float(t), _dp(out))
    return out


def samplers(seed, which, n):
    out = np.zeros((n, 3))
    lib().orc_kat_samplers(seed, which, n, _dp(out))
    return out


def camera_init(look_from, look_at, vup, vfov, aspect, aperture, focus_dist=10.0, time0=0.0, time1=1.0):
    cam = _abi().Camera()
    lib().orc_camera_init(_dp(_d3(look_from)), _dp(_d3(look_at)), _dp(_d3(vup)), vfov, aspect, aperture,
                          focus_dist, time0, time1, C.byref(cam))
    return cam


def num_threads():
    return lib().orc_num_threads()
