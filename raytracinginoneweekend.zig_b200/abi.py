"""ctypes mirror of include/rtw_cuda.h (field for field; checked by tests/test_abi.py)."""
import ctypes as C

RTW_ABI_VERSION = 3
RTW_MISS = 0xFFFFFFFF

PRIM_SPHERE, PRIM_MOVING_SPHERE, PRIM_XY_RECT, PRIM_XZ_RECT, PRIM_YZ_RECT = range(5)
XFORM_TRANSLATE, XFORM_ROTATE_Y = range(2)
MAT_DIFFUSE, MAT_METAL, MAT_DIELECTRIC, MAT_DIFFUSE_LIGHT = range(4)
TEX_SOLID, TEX_CHECKER, TEX_NOISE, TEX_IMAGE = range(4)
VARIANT_AUTO, VARIANT_MEGA_FLAT, VARIANT_MEGA_BVH, VARIANT_WAVEFRONT = range(4)
FLAG_COUNT_EVENTS = 1
FLAG_DETERMINISTIC = 2
BVH_BUILDER_SAH, BVH_BUILDER_LBVH = range(2)


class Prim(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("material", C.c_uint32), ("xform", C.c_int32),
                ("reserved", C.c_uint32), ("v", C.c_double * 10)]


class Xform(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("outer", C.c_int32), ("v", C.c_double * 4)]


class Material(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("texture", C.c_int32), ("albedo", C.c_double * 3),
                ("param", C.c_double)]


class Texture(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("a", C.c_int32), ("b", C.c_int32), ("reserved", C.c_uint32),
                ("color", C.c_double * 3), ("scale", C.c_double)]


class Image(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("rgba8", C.POINTER(C.c_uint8))]


class Perlin(C.Structure):
    _fields_ = [("ranvec", C.POINTER(C.c_double)), ("perm_x", C.POINTER(C.c_uint32)),
                ("perm_y", C.POINTER(C.c_uint32)), ("perm_z", C.POINTER(C.c_uint32))]


class SceneDesc(C.Structure):
    _fields_ = [("n_prims", C.c_uint32), ("prims", C.POINTER(Prim)),
                ("n_xforms", C.c_uint32), ("xforms", C.POINTER(Xform)),
                ("n_materials", C.c_uint32), ("materials", C.POINTER(Material)),
                ("n_textures", C.c_uint32), ("textures", C.POINTER(Texture)),
                ("n_images", C.c_uint32), ("images", C.POINTER(Image)),
                ("n_perlins", C.c_uint32), ("perlins", C.POINTER(Perlin)),
                ("time0", C.c_double), ("time1", C.c_double)]


class Camera(C.Structure):
    _fields_ = [("origin", C.c_double * 3), ("horizontal", C.c_double * 3), ("vertical", C.c_double * 3),
                ("lower_left_corner", C.c_double * 3), ("u", C.c_double * 3), ("v", C.c_double * 3),
                ("w", C.c_double * 3), ("lens_radius", C.c_double), ("time0", C.c_double),
                ("time1", C.c_double)]


class RenderParams(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("spp_begin", C.c_uint32),
                ("spp_end", C.c_uint32), ("spp_total", C.c_uint32), ("max_depth", C.c_uint32),
                ("variant", C.c_uint32), ("flags", C.c_uint32), ("seed", C.c_uint64),
                ("background", C.c_double * 3)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "paths", "rays", "node_tests", "sphere_tests", "sphere_roots", "moving_tests", "rect_tests",
        "rect_accepts", "xform_apps", "sphere_finalise", "scatter_diffuse", "scatter_metal",
        "scatter_dielectric", "emit_hits", "tex_checker", "tex_image", "tex_noise", "nan_pixels")] + [
        ("ms_trace", C.c_double), ("ms_resolve", C.c_double), ("ms_upload", C.c_double),
        ("n_launches", C.c_uint32), ("variant_used", C.c_uint32), ("bvh_nodes", C.c_uint32),
        ("bvh_depth", C.c_uint32), ("ms_bvh_build", C.c_double), ("bvh_builder", C.c_uint32),
        ("reserved0", C.c_uint32), ("ms_wall", C.c_double)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


# every symbol include/rtw_cuda.h declares
CUDA_SYMBOLS = (
    "rtw_cuda_create", "rtw_cuda_destroy", "rtw_cuda_last_error", "rtw_cuda_abi_version",
    "rtw_cuda_upload_scene", "rtw_cuda_render", "rtw_cuda_accumulate", "rtw_cuda_resolve",
    "rtw_cuda_resolve_multi", "rtw_cuda_render_multi", "rtw_cuda_trace_rays", "rtw_cuda_primary_hits", "rtw_cuda_stats",
    "rtw_cuda_measure_fp32_peak", "rtw_cuda_unit_camera", "rtw_cuda_unit_samplers", "rtw_cuda_unit_uniforms", "rtw_cuda_unit_shade",
    "rtw_cuda_create_multi", "rtw_cuda_set_option",
)
