"""ref_transpile — runs the reference's OWN source text, transpiled to Python at test time.

TEST INFRASTRUCTURE.  The reference (nsfisis/RayTracingInOneWeekend.zig) ships no tests or golden vectors and there is
no Zig toolchain in the image, so its behaviour cannot be observed by building it.  Its hot path, however, is 1.7 k
lines of straight-line f64 arithmetic in a small subset of Zig.  `zig2py` parses that subset and emits Python that
evaluates the same statements in the same order on IEEE binary64 — i.e. this module EXECUTES
/root/reference/src/{main,rc}.zig and src/rtw/*.zig, it does not restate them.  What it pins:

  * every pure function of the path (Vec3, Ray.at, Aabb, all nine Hittable.hit bodies and boudingBox rules, the four
    materials, reflect/refract/reflectance, the four textures, Perlin noise/turb/permute, Camera.init/getRay,
    rayColor, the six scene builders, main()'s render loop and quantisation) against oracle/ — bit for bit where the
    oracle keeps the reference's statement order, which is everywhere it claims Real=double semantics;
  * committed fixtures (tests/golden/ref_golden.json, made by tests/golden/make_ref_golden.py from this module) carry
    the same pin to machines without /root/reference.

Not pinned by the reference's text (imports from outside its tree, restated in zig2py/runtime.py): Zig std's
DefaultPrng / Random.float / uintLessThan, libm, zigimg's PNG decoder (Pillow here; PNG is lossless).
"""
import os
import re

from zig2py import parse, transpile, runtime

REF_ROOT = os.environ.get("RTW_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "src", "main.zig"))


class _Module:
    def __init__(self, g, path):
        object.__setattr__(self, "_g", g)
        object.__setattr__(self, "_path", path)

    def __getattr__(self, n):
        try:
            return self._g[n]
        except KeyError:
            raise AttributeError(f"{self._path}: no top-level name {n!r} (yet)")


class Loader:
    """One transpiled instance of the reference.  `patches`: {relative path: [(regex, replacement), ...]} applied to the
    source text before parsing; every pattern must match exactly once (a changed reference fails loudly)."""

    def __init__(self, patches=None, root=None):
        self.root = os.path.join(root or REF_ROOT, "src")
        self.patches = patches or {}
        self.modules = {}
        self.sources = {}
        self.pending = []
        self.resolving = False
        runtime.Image.asset_root = root or REF_ROOT

    def module(self, rel):
        rel = os.path.normpath(rel)
        if rel in self.modules:
            return self.modules[rel]
        path = os.path.join(self.root, rel)
        with open(path) as f:
            src = f.read()
        for pat, repl in self.patches.get(rel, []):
            src, n = re.subn(pat, repl, src)
            if n != 1:
                raise RuntimeError(f"patch {pat!r} matched {n} times in {rel} (expected exactly 1)")
        py = transpile(parse(src, rel))
        self.sources[rel] = py
        here = os.path.dirname(rel)
        g = {"_rt": runtime}

        def _imp(name, _here=here):
            if name == "std":
                return runtime.std
            if name == "zigimg":
                return runtime.zigimg
            return self.module(os.path.join(_here, name))

        def _lazy(name, thunk, _g=g):
            self.pending.append((_g, name, thunk, rel))

        g["_imp"] = _imp
        g["_lazy"] = _lazy
        mod = _Module(g, rel)
        self.modules[rel] = mod
        exec(compile(py, f"<zig2py {rel}>", "exec"), g)
        self.resolve()
        return mod

    def resolve(self):
        if self.resolving:
            return
        self.resolving = True
        try:
            progress = True
            while self.pending and progress:
                progress = False
                for item in list(self.pending):
                    g, name, thunk, rel = item
                    try:
                        val = thunk()
                    except (AttributeError, NameError):
                        continue
                    g[name] = val
                    self.pending.remove(item)
                    progress = True
        finally:
            self.resolving = False
        if self.pending:
            raise RuntimeError("unresolved top-level names: " + ", ".join(f"{rel}:{name}" for _, name, _, rel in self.pending))


_default = None


def ref():
    """The unpatched reference library modules: ref().vec, .ray, .aabb, .hittable, .material, .texture, .perlin, .rand"""
    global _default
    if _default is None:
        ld = Loader()

        class R:
            loader = ld
            rtw = ld.module("rtw.zig")
            vec = ld.module("rtw/vec.zig")
            ray = ld.module("rtw/ray.zig")
            aabb = ld.module("rtw/aabb.zig")
            hit_record = ld.module("rtw/hit_record.zig")
            hittable = ld.module("rtw/hittable.zig")
            material = ld.module("rtw/material.zig")
            texture = ld.module("rtw/texture.zig")
            perlin = ld.module("rtw/perlin.zig")
            rand = ld.module("rtw/rand.zig")
            rc = ld.module("rc.zig")
            main = ld.module("main.zig")

        _default = R
    return _default


def run_main(scene, width, height, spp, max_depth=50, seed=42):
    """Runs the reference's main() (src/main.zig:295-406) with its source constants replaced: scene id, image size,
    samples per pixel, depth, seed.  Returns (rgb rows top-first as a list of (r,g,b), draws consumed)."""
    aspect = f"{float(width)} / {float(height)}"
    patches = {"main.zig": [
        (r"const scene = 6;", f"const scene = {scene};"),
        (r"DefaultPrng\.init\(42\)", f"DefaultPrng.init({seed})"),
        (r"var aspect_ratio: f64 = 3\.0 / 2\.0;", f"var aspect_ratio: f64 = {aspect};"),
        (r"var image_width: u32 = 600;", f"var image_width: u32 = {width};"),
        (r"var image_height: u32 = @as\(u32, @intFromFloat\(@divTrunc\(@as\(f64, @floatFromInt\(image_width\)\), aspect_ratio\)\)\);",
         f"var image_height: u32 = {height};"),
        (r"const max_depth = 50;", f"const max_depth = {max_depth};"),
        (r"var samples_per_pixel: u32 = 50;", f"var samples_per_pixel: u32 = {spp};"),
        (r"samples_per_pixel = 400;", f"samples_per_pixel = {spp};"),
        (r"aspect_ratio = 1\.0;\s*image_width = 600;\s*image_height = 600;\s*samples_per_pixel = 200;",
         f"aspect_ratio = {aspect}; image_width = {width}; image_height = {height}; samples_per_pixel = {spp};"),
    ]}
    ld = Loader(patches)
    runtime.Image.created.clear()
    m = ld.module("main.zig")
    m.main()
    img = runtime.Image.created[-1]
    return [(p.r, p.g, p.b) for p in img.pixels.rgb24]


# ---- reference-side scene graphs from the flattened ABI description -------------------------------------------------
def _v3(a):
    return runtime.Anon(x=float(a[0]), y=float(a[1]), z=float(a[2]))


def scene_from_desc(desc):
    """Builds the reference's own `Hittable` graph (a top-level list) + materials from an rtw_scene_desc, using the
    reference's types and constructors only.  Returns (world, [material pointers])."""
    R = ref()
    rt_, A = runtime, runtime.Anon
    Texture, Material, Hittable = R.texture.Texture, R.material.Material, R.hittable.Hittable
    alloc = rt_.Allocator()

    def texture(i):
        t = desc.textures[i]
        if t.kind == 0:
            return Texture.makeSolid(_v3(t.color))
        if t.kind == 1:
            odd, even = rt_.Ptr(Texture, texture(t.a)), rt_.Ptr(Texture, texture(t.b))
            return rt_.co(Texture, A(checker=A(allocator=alloc, odd=odd, even=even)))
        if t.kind == 2:
            p = desc.perlins[t.a]
            pn = rt_.undefined(R.perlin.Perlin)
            pn.randomVec = rt_.std.ArrayList(R.vec.Vec3).init(alloc)
            pn.permX, pn.permY, pn.permZ = (rt_.std.ArrayList(rt_.usize).init(alloc) for _ in range(3))
            for k in range(256):
                pn.randomVec.append(_v3([p.ranvec[3 * k], p.ranvec[3 * k + 1], p.ranvec[3 * k + 2]]))
                pn.permX.append(int(p.perm_x[k])); pn.permY.append(int(p.perm_y[k])); pn.permZ.append(int(p.perm_z[k]))
            return rt_.co(Texture, A(noise=A(perlin=pn, scale=float(t.scale))))
        im = desc.images[t.a]
        raw = bytes(bytearray(im.rgba8[0:im.width * im.height * 4]))
        img = rt_.Image(im.width, im.height, rt_._Pixels(raw=raw))
        return rt_.co(Texture, A(image=A(image=img)))

    mats = []
    for i in range(desc.n_materials):
        m = desc.materials[i]
        if m.kind == 0:
            v = A(diffuse=A(albedo=texture(m.texture)))
        elif m.kind == 1:
            v = A(metal=A(albedo=_v3(m.albedo), fuzz=float(m.param)))
        elif m.kind == 2:
            v = A(dielectric=A(ir=float(m.param)))
        else:
            v = A(diffuse_light=A(emit=texture(m.texture)))
        rc = R.rc.Rc(Material).init(alloc)
        rt_.store(rc.get_mut(), v)
        mats.append(rc)

    objects = rt_.std.ArrayList(Hittable).init(alloc)
    for i in range(desc.n_prims):
        p = desc.prims[i]
        v, mat = p.v, mats[p.material]
        if p.kind == 0:
            h = A(sphere=A(center=_v3(v[0:3]), radius=float(v[3]), material=mat))
        elif p.kind == 1:
            h = A(movingSphere=A(center0=_v3(v[0:3]), center1=_v3(v[3:6]), time0=float(v[6]), time1=float(v[7]),
                                 radius=float(v[8]), material=mat))
        elif p.kind == 2:
            h = A(xyRect=A(x0=float(v[0]), x1=float(v[1]), y0=float(v[2]), y1=float(v[3]), k=float(v[4]), material=mat))
        elif p.kind == 3:
            h = A(xzRect=A(x0=float(v[0]), x1=float(v[1]), z0=float(v[2]), z1=float(v[3]), k=float(v[4]), material=mat))
        else:
            h = A(yzRect=A(y0=float(v[0]), y1=float(v[1]), z0=float(v[2]), z1=float(v[3]), k=float(v[4]), material=mat))
        h = rt_.co(Hittable, h)
        x = p.xform
        while x >= 0:  # innermost -> outermost
            xf = desc.xforms[x]
            cell = R.rc.Rc(Hittable).init(alloc)
            rt_.store(cell.get_mut(), h)
            if xf.kind == 0:
                h = Hittable.makeTranslate(cell, _v3(xf.v[0:3]))
            else:
                # RotateY.init takes an angle and calls std.math.sin/cos (hittable.zig:513-515); the ABI carries sin and cos.
                # The reference's own init runs (it also caches the rotated bbox, :516-556) with the std.math shim
                # returning exactly the ABI's values for this one call.
                keep = (rt_.std.math.sin, rt_.std.math.cos)
                rt_.std.math.sin, rt_.std.math.cos = (lambda t, v=float(xf.v[0]): v), (lambda t, v=float(xf.v[1]): v)
                try:
                    h = Hittable.makeRotateY(cell, 0.0)
                finally:
                    rt_.std.math.sin, rt_.std.math.cos = keep
            x = xf.outer
        objects.append(h)
    world = rt_.co(Hittable, A(list=A(objects=objects)))
    return world, mats


def builtin_world(sid, seed=42):
    """The reference's own scene builder for `scene = sid` (main.zig:124-293, selected at :320-362), run on a fresh
    DefaultPrng(seed) exactly as main() does.  Returns (world, generator) — the generator to continue the stream."""
    R = ref()
    m, alloc = R.main, runtime.Allocator()
    gen = runtime.Xoshiro256(seed)
    rng = gen.random()
    world = {1: lambda: m.generateRandomScene(rng, alloc), 2: lambda: m.generateTwoSpheres(rng, alloc),
             3: lambda: m.generateTwoPerlinSpheres(rng, alloc), 4: lambda: m.generateEarthScene(alloc),
             5: lambda: m.generateSimpleLightScene(rng, alloc), 6: lambda: m.generateCornellBox(alloc)}[sid]()
    return world, gen


def hit_record(world, ray7, t_min=0.001, t_max=float("inf")):
    """world.hit(r, t_min, t_max, &rec) -> None | dict like OracleScene.hit_record (without ids)"""
    R = ref()
    r = runtime.co(R.ray.Ray, runtime.Anon(origin=_v3(ray7[0:3]), dir=_v3(ray7[3:6]), time=float(ray7[6])))
    rec = runtime.undefined(R.hit_record.HitRecord)
    if not world.hit(r, t_min, t_max, rec):
        return None
    return dict(t=rec.t, p=(rec.p.x, rec.p.y, rec.p.z), normal=(rec.normal.x, rec.normal.y, rec.normal.z), u=rec.u, v=rec.v,
                front_face=rec.front_face, material=rec.material)
