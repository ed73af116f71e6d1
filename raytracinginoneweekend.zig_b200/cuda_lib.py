"""ctypes binding of csrc/ -> _build/librtw_cuda.so (the C ABI of include/rtw_cuda.h).

There is no fallback of any kind: if the library is missing or no CUDA device is present every call
raises.  Nothing in here touches oracle/.
"""
import ctypes as C
import os

import numpy as np

from . import abi, build

_lib = None


class RtwCudaError(RuntimeError):
    pass


def lib_path():
    return build.CUDA_LIB


def load(build_if_missing=True):
    """Load librtw_cuda.so (building it in-tree first if it is absent or stale)."""
    global _lib
    if _lib is not None:
        return _lib
    # Never rebuild implicitly when the library exists: under torchrun N ranks load it at once, and on the GPU
    # box the prebuilt in-tree .so is the artefact under test.  __graft_entry__.build() does the real build.
    if build_if_missing and (not os.path.exists(build.CUDA_LIB) or os.environ.get("RTW_REBUILD") == "1"):
        build.build_cuda()
    if not os.path.exists(build.CUDA_LIB):
        raise RtwCudaError(f"{build.CUDA_LIB} is missing: run __graft_entry__.build() (no CPU fallback exists)")
    L = C.CDLL(build.CUDA_LIB)
    vp = C.c_void_p
    u32p, dp, u8p, fp = C.POINTER(C.c_uint32), C.POINTER(C.c_double), C.POINTER(C.c_uint8), C.POINTER(C.c_float)
    L.rtw_cuda_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.rtw_cuda_destroy.argtypes = [vp]
    L.rtw_cuda_destroy.restype = None
    L.rtw_cuda_last_error.argtypes = [vp]
    L.rtw_cuda_last_error.restype = C.c_char_p
    L.rtw_cuda_abi_version.restype = C.c_uint32
    L.rtw_cuda_upload_scene.argtypes = [vp, C.POINTER(abi.SceneDesc)]
    L.rtw_cuda_render.argtypes = [vp, C.POINTER(abi.Camera), C.POINTER(abi.RenderParams), vp, vp]
    L.rtw_cuda_accumulate.argtypes = [vp, C.POINTER(abi.Camera), C.POINTER(abi.RenderParams), vp, vp]
    L.rtw_cuda_resolve.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, vp, vp]
    L.rtw_cuda_resolve_multi.argtypes = [vp, C.POINTER(vp), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, vp, vp]
    L.rtw_cuda_render_multi.argtypes = [C.POINTER(vp), C.c_uint32, C.POINTER(abi.Camera), C.POINTER(abi.RenderParams), vp]
    L.rtw_cuda_trace_rays.argtypes = [vp, C.c_uint32, dp, C.c_uint32, C.c_uint32, u32p, dp, dp, dp]
    L.rtw_cuda_primary_hits.argtypes = [vp, C.POINTER(abi.Camera), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                        u32p, dp, dp]
    L.rtw_cuda_stats.argtypes = [vp, C.POINTER(abi.Stats)]
    L.rtw_cuda_measure_fp32_peak.argtypes = [vp, dp, dp]
    L.rtw_cuda_unit_camera.argtypes = [vp, C.POINTER(abi.Camera), C.POINTER(abi.RenderParams), C.c_uint32, u32p, fp]
    L.rtw_cuda_unit_samplers.argtypes = [vp, C.c_uint32, fp, fp]
    L.rtw_cuda_unit_uniforms.argtypes = [vp, C.POINTER(abi.RenderParams), C.c_uint32, u32p, fp]
    L.rtw_cuda_unit_shade.argtypes = [vp, C.POINTER(abi.RenderParams), C.c_uint32, dp, u32p, u32p, fp]
    L.rtw_cuda_create_multi.argtypes = [C.c_uint32, C.POINTER(vp)]
    L.rtw_cuda_set_option.argtypes = [vp, C.c_char_p, C.c_char_p]
    for name in abi.CUDA_SYMBOLS:
        getattr(L, name)  # AttributeError here = the .so does not export what the header declares
    _lib = L
    return L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def render_multi(contexts, cam, params, rgb8=None):
    """Single-process multi-GPU render: one Context per device, same scene uploaded to each."""
    H, W = params.height, params.width
    if rgb8 is None:
        rgb8 = np.empty((H, W, 3), dtype=np.uint8)
    arr = (C.c_void_p * len(contexts))(*[c.h for c in contexts])
    rc = contexts[0].L.rtw_cuda_render_multi(arr, len(contexts), C.byref(cam), C.byref(params), rgb8.ctypes.data)
    contexts[0]._check(rc, "rtw_cuda_render_multi")
    return rgb8


def create_multi(n_gpus):
    """rtw_cuda_create_multi: contexts on devices 0..n_gpus-1 with all-pairs peer access already enabled."""
    L = load()
    arr = (C.c_void_p * n_gpus)()
    rc = L.rtw_cuda_create_multi(n_gpus, arr)
    if rc != 0:
        raise RtwCudaError(f"rtw_cuda_create_multi({n_gpus}) -> {rc}: {L.rtw_cuda_last_error(None).decode()}")
    return [Context(device=i, handle=arr[i]) for i in range(n_gpus)]


class Context:
    """One rtw_ctx: a CUDA device + an uploaded scene."""

    def __init__(self, device=0, handle=None):
        self.L = load()
        self.h = C.c_void_p(handle) if handle else C.c_void_p()
        if not handle:
            rc = self.L.rtw_cuda_create(device, C.byref(self.h))
            if rc != 0:
                raise RtwCudaError(f"rtw_cuda_create({device}) -> {rc}: {self.L.rtw_cuda_last_error(None).decode()}")
        self.device = device
        self._keep = None

    def set_option(self, name, value):
        """Tuning knob (same names as the environment variables, e.g. RTW_BVH_BUILDER); value None = built-in default."""
        v = None if value is None else str(value).encode()
        self._check(self.L.rtw_cuda_set_option(self.h, name.encode(), v), "rtw_cuda_set_option")

    def options(self, **kw):
        """with ctx.options(RTW_BVH_BUILDER="lbvh"): ...   — set, then restore the defaults"""
        ctx = self

        class _Scope:
            def __enter__(self_inner):
                for k, v in kw.items():
                    ctx.set_option(k, v)

            def __exit__(self_inner, *exc):
                for k in kw:
                    ctx.set_option(k, None)
        return _Scope()

    def close(self):
        if getattr(self, "h", None):
            self.L.rtw_cuda_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise RtwCudaError(f"{what} -> {rc}: {self.L.rtw_cuda_last_error(self.h).decode()}")

    def upload_scene(self, desc, keep=None):
        self._check(self.L.rtw_cuda_upload_scene(self.h, C.byref(desc)), "rtw_cuda_upload_scene")
        self._keep = keep

    @staticmethod
    def params(width, height, spp_begin, spp_end, spp_total=0, max_depth=50, variant=abi.VARIANT_AUTO, flags=0,
               seed=42, background=(0.7, 0.8, 1.0)):
        p = abi.RenderParams(width=width, height=height, spp_begin=spp_begin, spp_end=spp_end, spp_total=spp_total,
                             max_depth=max_depth, variant=variant, flags=flags, seed=seed)
        p.background[:] = [float(x) for x in background]
        return p

    def render(self, cam, params, want_accum=False, rgb8=None, accum=None):
        """Host-buffer render (the drop-in call).  Returns (rgb8[H,W,3] top row first, accum[H,W,4] or None)."""
        H, W = params.height, params.width
        if rgb8 is None:
            rgb8 = np.empty((H, W, 3), dtype=np.uint8)
        if want_accum and accum is None:
            accum = np.empty((H, W, 4), dtype=np.float32)
        self._check(self.L.rtw_cuda_render(self.h, C.byref(cam), C.byref(params), rgb8.ctypes.data,
                                           accum.ctypes.data if accum is not None else None), "rtw_cuda_render")
        return rgb8, accum

    def accumulate(self, cam, params, d_accum_ptr, stream=None):
        self._check(self.L.rtw_cuda_accumulate(self.h, C.byref(cam), C.byref(params), d_accum_ptr, stream),
                    "rtw_cuda_accumulate")

    def resolve(self, d_accum_ptr, width, height, spp_total, d_rgb8_ptr, stream=None):
        self._check(self.L.rtw_cuda_resolve(self.h, d_accum_ptr, width, height, spp_total, d_rgb8_ptr, stream),
                    "rtw_cuda_resolve")

    def resolve_multi(self, d_accum_ptrs, width, height, spp_total, d_rgb8_ptr, stream=None):
        arr = (C.c_void_p * len(d_accum_ptrs))(*d_accum_ptrs)
        self._check(self.L.rtw_cuda_resolve_multi(self.h, arr, len(d_accum_ptrs), width, height, spp_total,
                                                  d_rgb8_ptr, stream), "rtw_cuda_resolve_multi")

    def trace_rays(self, rays, precision=32, variant=abi.VARIANT_AUTO):
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 7)
        n = rays.shape[0]
        ids = np.zeros(n, dtype=np.uint32)
        t, nrm, uv = np.zeros(n), np.zeros((n, 3)), np.zeros((n, 2))
        self._check(self.L.rtw_cuda_trace_rays(self.h, n, _dp(rays), precision, variant,
                                               ids.ctypes.data_as(C.POINTER(C.c_uint32)), _dp(t), _dp(nrm), _dp(uv)),
                    "rtw_cuda_trace_rays")
        return ids, t, nrm, uv

    def primary_hits(self, cam, width, height, precision=32, variant=abi.VARIANT_AUTO):
        ids = np.zeros((height, width), dtype=np.uint32)
        t, nrm = np.zeros((height, width)), np.zeros((height, width, 3))
        self._check(self.L.rtw_cuda_primary_hits(self.h, C.byref(cam), width, height, precision, variant,
                                                 ids.ctypes.data_as(C.POINTER(C.c_uint32)), _dp(t), _dp(nrm)),
                    "rtw_cuda_primary_hits")
        return ids, t, nrm

    # ---- unit probes of the stochastic device code (parity instruments) ----
    def unit_camera(self, cam, params, ijs):
        """ijs[n,3] = (i, j, sample) -> float32[n,14]: ray(7) | uniforms ju jv l1 l2 tm | disk point(2)"""
        ijs = np.ascontiguousarray(ijs, dtype=np.uint32).reshape(-1, 3)
        out = np.zeros((ijs.shape[0], 14), dtype=np.float32)
        self._check(self.L.rtw_cuda_unit_camera(self.h, C.byref(cam), C.byref(params), ijs.shape[0],
                                                ijs.ctypes.data_as(C.POINTER(C.c_uint32)), out.ctypes.data_as(C.POINTER(C.c_float))),
                    "rtw_cuda_unit_camera")
        return out

    def unit_samplers(self, u3):
        """u3[n,3] uniforms -> float32[n,8]: unit vector(3) | ball point(3) | disk point(2)"""
        u3 = np.ascontiguousarray(u3, dtype=np.float32).reshape(-1, 3)
        out = np.zeros((u3.shape[0], 8), dtype=np.float32)
        self._check(self.L.rtw_cuda_unit_samplers(self.h, u3.shape[0], u3.ctypes.data_as(C.POINTER(C.c_float)),
                                                  out.ctypes.data_as(C.POINTER(C.c_float))), "rtw_cuda_unit_samplers")
        return out

    def unit_uniforms(self, params, psb):
        """psb[n,3] = (pixel, sample, block) -> (float32[n,4] uniforms, uint32[n,4] raw Philox words)"""
        psb = np.ascontiguousarray(psb, dtype=np.uint32).reshape(-1, 3)
        out = np.zeros((psb.shape[0], 8), dtype=np.float32)
        self._check(self.L.rtw_cuda_unit_uniforms(self.h, C.byref(params), psb.shape[0], psb.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                  out.ctypes.data_as(C.POINTER(C.c_float))), "rtw_cuda_unit_uniforms")
        return out[:, :4].copy(), out[:, 4:].copy().view(np.uint32)

    def unit_shade(self, params, rays, psb):
        """rays[n,7], psb[n,3] = (pixel, sample, bounce) -> (prim_id[n], float32[n,20]); layout in include/rtw_cuda.h"""
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 7)
        psb = np.ascontiguousarray(psb, dtype=np.uint32).reshape(-1, 3)
        n = rays.shape[0]
        ids = np.zeros(n, dtype=np.uint32)
        out = np.zeros((n, 20), dtype=np.float32)
        self._check(self.L.rtw_cuda_unit_shade(self.h, C.byref(params), n, _dp(rays), psb.ctypes.data_as(C.POINTER(C.c_uint32)),
                                               ids.ctypes.data_as(C.POINTER(C.c_uint32)), out.ctypes.data_as(C.POINTER(C.c_float))),
                    "rtw_cuda_unit_shade")
        return ids, out

    def stats(self):
        s = abi.Stats()
        self._check(self.L.rtw_cuda_stats(self.h, C.byref(s)), "rtw_cuda_stats")
        return s.as_dict()

    def measure_fp32_peak(self):
        tf, mhz = C.c_double(0), C.c_double(0)
        self._check(self.L.rtw_cuda_measure_fp32_peak(self.h, C.byref(tf), C.byref(mhz)), "rtw_cuda_measure_fp32_peak")
        return tf.value, mhz.value
