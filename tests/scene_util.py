"""Build rtw_scene_desc structures from Python lists (test helper)."""
import ctypes as C

import numpy as np

import rtw_b200
from rtw_b200 import abi


class DescBuilder:
    def __init__(self, time0=0.0, time1=1.0):
        self.prims, self.xforms, self.mats, self.texs = [], [], [], []
        self.images, self.perlins, self._keep = [], [], []
        self.time0, self.time1 = time0, time1

    def solid(self, rgb):
        t = abi.Texture(kind=abi.TEX_SOLID, a=-1, b=-1)
        t.color[:] = list(rgb)
        self.texs.append(t)
        return len(self.texs) - 1

    def checker(self, odd, even):
        o, e = self.solid(odd), self.solid(even)
        self.texs.append(abi.Texture(kind=abi.TEX_CHECKER, a=o, b=e))
        return len(self.texs) - 1

    def image(self, rgba):
        rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
        self._keep.append(rgba)
        self.images.append(abi.Image(width=rgba.shape[1], height=rgba.shape[0],
                                     rgba8=rgba.ctypes.data_as(C.POINTER(C.c_uint8))))
        self.texs.append(abi.Texture(kind=abi.TEX_IMAGE, a=len(self.images) - 1, b=-1))
        return len(self.texs) - 1

    def diffuse(self, tex):
        self.mats.append(abi.Material(kind=abi.MAT_DIFFUSE, texture=tex))
        return len(self.mats) - 1

    def metal(self, rgb, fuzz):
        m = abi.Material(kind=abi.MAT_METAL, texture=-1, param=fuzz)
        m.albedo[:] = list(rgb)
        self.mats.append(m)
        return len(self.mats) - 1

    def glass(self, ir):
        self.mats.append(abi.Material(kind=abi.MAT_DIELECTRIC, texture=-1, param=ir))
        return len(self.mats) - 1

    def light(self, tex):
        self.mats.append(abi.Material(kind=abi.MAT_DIFFUSE_LIGHT, texture=tex))
        return len(self.mats) - 1

    def translate(self, offset, outer=-1):
        x = abi.Xform(kind=abi.XFORM_TRANSLATE, outer=outer)
        x.v[0:3] = list(offset)
        self.xforms.append(x)
        return len(self.xforms) - 1

    def rotate_y(self, degrees, outer=-1):
        x = abi.Xform(kind=abi.XFORM_ROTATE_Y, outer=outer)
        x.v[0:2] = [float(np.sin(np.radians(degrees))), float(np.cos(np.radians(degrees)))]
        self.xforms.append(x)
        return len(self.xforms) - 1

    def _prim(self, kind, mat, xform, vals):
        p = abi.Prim(kind=kind, material=mat, xform=xform)
        p.v[0:len(vals)] = [float(v) for v in vals]
        self.prims.append(p)
        return len(self.prims) - 1

    def sphere(self, c, r, mat, xform=-1):
        return self._prim(abi.PRIM_SPHERE, mat, xform, [*c, r])

    def moving_sphere(self, c0, c1, t0, t1, r, mat, xform=-1):
        return self._prim(abi.PRIM_MOVING_SPHERE, mat, xform, [*c0, *c1, t0, t1, r])

    def rect(self, kind, a0, a1, b0, b1, k, mat, xform=-1):
        return self._prim(kind, mat, xform, [a0, a1, b0, b1, k])

    def box(self, p0, p1, mat, xform=-1):  # side order of Box.init, src/rtw/hittable.zig:437-442
        self.rect(abi.PRIM_XY_RECT, p0[0], p1[0], p0[1], p1[1], p1[2], mat, xform)
        self.rect(abi.PRIM_XY_RECT, p0[0], p1[0], p0[1], p1[1], p0[2], mat, xform)
        self.rect(abi.PRIM_XZ_RECT, p0[0], p1[0], p0[2], p1[2], p1[1], mat, xform)
        self.rect(abi.PRIM_XZ_RECT, p0[0], p1[0], p0[2], p1[2], p0[1], mat, xform)
        self.rect(abi.PRIM_YZ_RECT, p0[1], p1[1], p0[2], p1[2], p1[0], mat, xform)
        self.rect(abi.PRIM_YZ_RECT, p0[1], p1[1], p0[2], p1[2], p0[0], mat, xform)

    def build(self):
        def arr(cls, items):
            a = (cls * max(1, len(items)))(*items)
            self._keep.append(a)
            return a
        d = abi.SceneDesc(n_prims=len(self.prims), prims=arr(abi.Prim, self.prims),
                          n_xforms=len(self.xforms), xforms=arr(abi.Xform, self.xforms),
                          n_materials=len(self.mats), materials=arr(abi.Material, self.mats),
                          n_textures=len(self.texs), textures=arr(abi.Texture, self.texs),
                          n_images=len(self.images), images=arr(abi.Image, self.images),
                          n_perlins=0, perlins=None, time0=self.time0, time1=self.time1)
        d._owner = self
        return d


def random_scene(rng, n_spheres=40, n_moving=20, n_rects=20, n_boxes=3, extent=10.0, n_inst_spheres=6):
    """Mixed random scene: static + moving spheres, axis rects, instanced boxes, a big ground sphere."""
    b = DescBuilder()
    mats = [b.diffuse(b.solid((0.5, 0.5, 0.5))), b.metal((0.8, 0.8, 0.8), 0.1), b.glass(1.5),
            b.diffuse(b.checker((0.2, 0.3, 0.1), (0.9, 0.9, 0.9)))]
    b.sphere((0, -1000 - extent, 0), 1000, mats[3])
    for _ in range(n_spheres):
        b.sphere(rng.uniform(-extent, extent, 3), rng.uniform(0.1, 1.5), mats[rng.integers(0, 4)])
    for _ in range(n_moving):
        c0 = rng.uniform(-extent, extent, 3)
        b.moving_sphere(c0, c0 + rng.uniform(-1, 1, 3), 0.0, 1.0, rng.uniform(0.1, 1.0), mats[rng.integers(0, 4)])
    for _ in range(n_rects):
        a0, b0 = rng.uniform(-extent, extent, 2)
        kind = [abi.PRIM_XY_RECT, abi.PRIM_XZ_RECT, abi.PRIM_YZ_RECT][rng.integers(0, 3)]
        b.rect(kind, a0, a0 + rng.uniform(0.5, 4), b0, b0 + rng.uniform(0.5, 4), rng.uniform(-extent, extent),
               mats[rng.integers(0, 4)])
    for _ in range(n_boxes):
        t = b.translate(rng.uniform(-extent, extent, 3))
        r = b.rotate_y(rng.uniform(-90, 90), outer=t)
        b.box((0, 0, 0), rng.uniform(0.5, 3, 3), mats[rng.integers(0, 4)], xform=r)
    for k in range(n_inst_spheres):  # Translate(RotateY(sphere)) and Translate(RotateY(moving sphere))
        t = b.translate(rng.uniform(-extent, extent, 3))
        r = b.rotate_y(rng.uniform(-180, 180), outer=t)
        c = rng.uniform(-2, 2, 3)
        if k % 2 == 0:
            b.sphere(c, rng.uniform(0.3, 1.5), mats[rng.integers(0, 4)], xform=r)
        else:
            b.moving_sphere(c, c + rng.uniform(-1, 1, 3), 0.0, 1.0, rng.uniform(0.3, 1.0), mats[rng.integers(0, 4)], xform=r)
    return b.build()


def random_rays(rng, n, extent=10.0):
    rays = np.zeros((n, 7))
    rays[:, 0:3] = rng.uniform(-1.5 * extent, 1.5 * extent, (n, 3))
    target = rng.uniform(-extent, extent, (n, 3))
    rays[:, 3:6] = (target - rays[:, 0:3]) * rng.uniform(0.1, 3.0, (n, 1))  # un-normalised, like the reference
    rays[:, 6] = rng.uniform(0, 1, n)
    return rays
