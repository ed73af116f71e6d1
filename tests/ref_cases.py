"""ref_cases — the case families that pin the oracle to the reference, evaluated two ways.

TEST INFRASTRUCTURE.  Each family has seeded inputs, `ref_*` evaluates them by EXECUTING the transpiled reference
(tests/ref_transpile.py; needs /root/reference), `orc_*` evaluates them with the CPU oracle (oracle/).  Outputs are plain
lists of Python floats / ints / bools / None so they compare with == (bit-exact: the same f64 statements in the same
order on both sides) and serialise to JSON without loss (tests/golden/ref_golden.json, made by
tests/golden/make_ref_golden.py, carries the reference's outputs to machines that do not have /root/reference).

  used by  tests/test_ref_pin.py     reference (live) == oracle, large N           (-m "not gpu", needs /root/reference)
           tests/test_ref_golden.py  committed reference outputs == oracle, small N (-m "not gpu", runs anywhere)
"""
import math

import numpy as np

import oracle_binding as ob
import scene_util
import rtw_b200
from rtw_b200 import abi

MAIN_CASES = [(6, 12, 12, 3), (1, 15, 10, 2), (2, 9, 6, 3), (3, 9, 6, 3), (4, 9, 6, 3), (5, 24, 16, 4)]  # scene, W, H, spp


def earth():
    return rtw_b200.host_lib.decode_png(rtw_b200.host_lib.ASSET_EARTH)


def fl(x):
    return [float(v) for v in x]


# ---- scenes available to both sides -----------------------------------------------------------------------------
_scene_cache = {}


def scene(key):
    """key: 'builtin:<id>' (the reference's generate* with seed 42) or 'random:<seed>' (scene_util.random_scene with fewer
    primitives, so a ray costs ~1 ms on the transpiled side).  Returns dict(desc, osc)."""
    if key in _scene_cache:
        return _scene_cache[key]
    kind, arg = key.split(":")
    if kind == "builtin":
        sid = int(arg)
        osc = ob.OracleScene.builtin(sid, 3, 42, earth() if sid == 4 else None)
        desc = osc.export()
    else:
        rng = np.random.default_rng(int(arg))
        desc = scene_util.random_scene(rng, n_spheres=10, n_moving=8, n_rects=8, n_boxes=2, n_inst_spheres=4)
        osc = ob.OracleScene.from_desc(desc)
    _scene_cache[key] = dict(desc=desc, osc=osc)
    return _scene_cache[key]


_ref_world = {}


def ref_world(key):
    import ref_transpile as rt
    if key not in _ref_world:
        if key.startswith("builtin:"):  # the reference's own builder: the real nested graph (Box inside RotateY inside Translate)
            _ref_world[key] = (rt.builtin_world(int(key.split(":")[1]))[0], None)
        else:
            _ref_world[key] = rt.scene_from_desc(scene(key)["desc"])
    return _ref_world[key]


def rays_for(key, n, seed):
    rng = np.random.default_rng(seed)
    if key == "builtin:6":  # Cornell: rays from inside the room and from the camera side
        rays = np.zeros((n, 7))
        rays[:, 0:3] = rng.uniform(-50, 605, (n, 3))
        rays[: n // 3, 0:3] = (278, 278, -800)
        target = rng.uniform(0, 555, (n, 3))
        rays[:, 3:6] = (target - rays[:, 0:3]) * rng.uniform(0.2, 2.0, (n, 1))
        rays[:, 6] = rng.uniform(0, 1, n)
        return rays
    if key.startswith("builtin"):
        rays = scene_util.random_rays(rng, n, extent=6.0)
        rays[: n // 3, 0:3] = (13, 2, 3)
        return rays
    return scene_util.random_rays(rng, n)


# ---- family: whole program (src/main.zig:295-406) ---------------------------------------------------------------------
def ref_main(sid, W, H, spp):
    import ref_transpile as rt
    return [list(p) for p in rt.run_main(sid, W, H, spp)]


def orc_main(sid, W, H, spp):
    osc = ob.OracleScene.builtin(sid, 3, 42, earth() if sid == 4 else None)
    cfg = osc.config()
    cam = osc.default_camera(aspect=W / H)
    r = osc.render(cam, W, H, spp, 50, cfg["background"], seed=42, precision=64, nthreads=1, continue_stream=True)
    return [[int(c) for c in px] for px in r["rgb8"].reshape(-1, 3)]


# ---- family: world.hit (src/rtw/hittable.zig:47-59 and every hit body below it) --------------------------------------------
def _rec_list(t, p, n, u, v, front):
    return [float(t), *fl(p), *fl(n), None if u is None else float(u), None if v is None else float(v), bool(front)]


def ref_hits(key, rays, t_min=0.001, t_max=math.inf):
    import ref_transpile as rt
    world, _ = ref_world(key)
    out = []
    for r in rays:
        h = rt.hit_record(world, r, t_min, t_max)
        out.append(None if h is None else _rec_list(h["t"], h["p"], h["normal"], h["u"], h["v"], h["front_face"]))
    return out


def orc_hits(key, rays, t_min=0.001, t_max=math.inf):
    mask, o = scene(key)["osc"].hit_records(rays, t_min, t_max)
    return [(_rec_list(o[i, 0], o[i, 1:4], o[i, 4:7], o[i, 7], o[i, 8], o[i, 9] != 0) if mask[i] else None) for i in range(len(rays))]


def same_hits(a, b):
    """equal, except that u,v of a MovingSphere hit are undefined in the reference (None) and 0 in the oracle"""
    if len(a) != len(b):
        return False
    for x, y in zip(a, b):
        if (x is None) != (y is None):
            return False
        if x is None:
            continue
        for k, (p, q) in enumerate(zip(x, y)):
            if k in (7, 8) and (p is None or q is None):
                continue
            if p != q and not (p != p and q != q):
                return False
    return True


# ---- family: boudingBox rules (hittable.zig:61-73 and the per-variant bodies) ---------------------------------------------
def ref_boxes(key):
    import ref_transpile as rt
    R = rt.ref()
    world, _ = ref_world(key)
    d = scene(key)["desc"]
    out = []
    for obj in world.payload.objects.items:
        bb = rt.runtime.undefined(R.aabb.Aabb)
        ok = obj.boudingBox(d.time0, d.time1, bb)
        out.append([bb.min.x, bb.min.y, bb.min.z, bb.max.x, bb.max.y, bb.max.z] if ok else None)
    return out


def orc_boxes(key):
    osc, d = scene(key)["osc"], scene(key)["desc"]
    out = []
    for i in range(d.n_prims):  # top-level objects (fewer than leaves when the scene has boxes: the tail is None)
        bb = osc.bounding_box(i)
        if bb is None:
            break
        out.append(fl(bb[0]) + fl(bb[1]))
    return out


# ---- family: Aabb.hit (aabb.zig:8-45) ---------------------------------------------------------------------------------
def aabb_inputs(n, seed):
    rng = np.random.default_rng(seed)
    lo = rng.uniform(-5, 5, (n, 3))
    hi = lo + rng.uniform(0.01, 4, (n, 3))
    rays = scene_util.random_rays(rng, n, extent=6.0)
    aim = lo + (hi - lo) * rng.uniform(-0.3, 1.3, (n, 3))  # most rays aim at (or just past) their box
    rays[:, 3:6] = (aim - rays[:, 0:3]) * rng.uniform(0.3, 3.0, (n, 1))
    rays[::7, 3] = 0.0  # axis-parallel rays: division by zero in the reference's statements
    rays[::11, 4] = 0.0
    tr = np.sort(rng.uniform(0, 4, (n, 2)), axis=1)
    tr[::2] = (0.001, math.inf)
    return lo, hi, rays, tr


def ref_aabb(lo, hi, rays, tr):
    import ref_transpile as rt
    R, A, co = rt.ref(), rt.runtime.Anon, rt.runtime.co
    out = []
    for i in range(len(lo)):
        bb = co(R.aabb.Aabb, A(min=rt._v3(lo[i]), max=rt._v3(hi[i])))
        r = co(R.ray.Ray, A(origin=rt._v3(rays[i, 0:3]), dir=rt._v3(rays[i, 3:6]), time=float(rays[i, 6])))
        out.append(bool(bb.hit(r, float(tr[i, 0]), float(tr[i, 1]))))
    return out


def orc_aabb(lo, hi, rays, tr):
    L, dp = ob.lib(), ob._dp
    return [bool(L.orc_kat_aabb_hit(dp(np.ascontiguousarray(lo[i])), dp(np.ascontiguousarray(hi[i])), dp(np.ascontiguousarray(rays[i])),
                                    float(tr[i, 0]), float(tr[i, 1]))) for i in range(len(lo))]


# ---- family: reflect / refract / reflectance / getSphereUv (material.zig:87-91,112-121, hittable.zig:145-150) --------------
def helper_inputs(n, seed):
    rng = np.random.default_rng(seed)
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    nrm = rng.normal(size=(n, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    eta = np.where(rng.uniform(size=n) < 0.5, 1.5, 1.0 / 1.5) * rng.uniform(0.8, 1.2, n)
    cosv = rng.uniform(0, 1, n)
    return v, nrm, eta, cosv


def ref_helpers(v, nrm, eta, cosv):
    import ref_transpile as rt
    R, FP = rt.ref(), rt.runtime.FieldPtr
    M = R.material
    out = []

    class Cell:
        u = None
        v = None

    for i in range(len(v)):
        a, b = rt.runtime.co(R.vec.Vec3, rt._v3(v[i])), rt.runtime.co(R.vec.Vec3, rt._v3(nrm[i]))
        rf, rr = M.reflect(a, b), M.refract(a, b, float(eta[i]))
        c = Cell()
        R.hittable.Sphere.getSphereUv(a, FP(c, "u"), FP(c, "v"))
        out.append([rf.x, rf.y, rf.z, rr.x, rr.y, rr.z, M.DielectricMaterial.reflectance(float(cosv[i]), float(eta[i])), c.u, c.v])
    return out


def orc_helpers(v, nrm, eta, cosv):
    L, dp = ob.lib(), ob._dp
    out = []
    for i in range(len(v)):
        a, b = np.ascontiguousarray(v[i]), np.ascontiguousarray(nrm[i])
        rf, rr, uv = np.zeros(3), np.zeros(3), np.zeros(2)
        L.orc_kat_reflect(dp(a), dp(b), dp(rf))
        L.orc_kat_refract(dp(a), dp(b), float(eta[i]), dp(rr))
        L.orc_kat_sphere_uv(dp(a), dp(uv))
        out.append([*fl(rf), *fl(rr), float(L.orc_kat_reflectance(float(cosv[i]), float(eta[i]))), float(uv[0]), float(uv[1])])
    return out


# ---- family: Texture.value (texture.zig:36-145) and Perlin noise / turb (perlin.zig:47-91,103-124) -----------------------
def texture_scene():
    """one desc holding a solid, a checker, a nested checker, the earth image and the noise texture of builtin scene 3"""
    if "tex" in _scene_cache:
        return _scene_cache["tex"]
    b = scene_util.DescBuilder()
    s3 = scene("builtin:3")["desc"]
    t_solid = b.solid((0.4, 0.2, 0.1))
    t_check = b.checker((0.2, 0.3, 0.1), (0.9, 0.9, 0.9))
    b.texs.append(abi.Texture(kind=abi.TEX_CHECKER, a=t_check, b=t_solid))
    t_nested = len(b.texs) - 1
    t_img = b.image(earth())
    b.texs.append(abi.Texture(kind=abi.TEX_NOISE, a=0, b=-1, scale=4.0))
    t_noise = len(b.texs) - 1
    for t in (t_solid, t_check, t_nested, t_img, t_noise):
        b.sphere((0, 0, 0), 1.0, b.diffuse(t))
    d = b.build()
    d.n_perlins, d.perlins = s3.n_perlins, s3.perlins
    d._perlin_owner = s3
    _scene_cache["tex"] = dict(desc=d, osc=ob.OracleScene.from_desc(d), textures=[t_solid, t_check, t_nested, t_img, t_noise])
    return _scene_cache["tex"]


def texture_inputs(n, seed):
    rng = np.random.default_rng(seed)
    uvp = np.zeros((n, 5))
    uvp[:, 0:2] = rng.uniform(-0.1, 1.1, (n, 2))  # out-of-range u,v exercise the clamps (texture.zig:124-125)
    uvp[::50, 0] = 1.0
    uvp[::70, 1] = 1.0
    uvp[:, 2:5] = rng.uniform(-8, 8, (n, 3))
    uvp[: n // 10, 2:5] *= 100.0
    # v == 0 exactly is the reference's out-of-bounds row (texture.zig:130 clamps the ROW with width-1):
    # keep v away from the last row so both sides read inside the image
    uvp[:, 1] = np.maximum(uvp[:, 1], 0.01)
    return uvp


def ref_textures(uvp):
    import ref_transpile as rt
    ts = texture_scene()
    _, mats = ref_world("tex")
    out = []
    for k in range(5):
        tex = mats[k].get().payload.albedo  # the reference's Texture value inside DiffuseMaterial
        for row in uvp:
            c = tex.value(float(row[0]), float(row[1]), rt.runtime.co(rt.ref().vec.Vec3, rt._v3(row[2:5])))
            out.append([c.x, c.y, c.z])
    return out


def orc_textures(uvp):
    ts = texture_scene()
    out = []
    for t in ts["textures"]:
        for row in uvp:
            out.append(fl(ts["osc"].texture_value(t, float(row[0]), float(row[1]), row[2:5])))
    return out


def perlin_inputs(n, seed):
    rng = np.random.default_rng(seed)
    p = rng.uniform(-40, 40, (n, 3))
    p[: n // 8] = np.round(p[: n // 8])          # lattice points
    p[n // 8: n // 4, 0] = np.round(p[n // 8: n // 4, 0]) + 0.5
    return p


def ref_perlin(points):
    import ref_transpile as rt
    texture_scene()
    _, mats = ref_world("tex")
    pn = mats[4].get().payload.albedo.payload.perlin
    V = rt.ref().vec.Vec3
    out = []
    for p in points:
        q = rt.runtime.co(V, rt._v3(p))
        out.append([pn.noise(q), pn.turb(q, 7)])
    return out


def orc_perlin(points):
    osc = texture_scene()["osc"]
    return [[float(osc.perlin_noise(0, p)), float(osc.perlin_turb(0, p, 7))] for p in points]


def book_perlin_noise(tables, p):
    """Shirley's 'The Next Week' perlin_interp — raw (u-i) in the weight vector.  NOT the reference (perlin.zig:77 passes
    the smoothed values); kept here only to show that the pin tells the two apart (round 1 shipped this one)."""
    rv, pm = tables
    f = np.floor(p)
    u, v, w = p - f
    uu, vv, ww = (x * x * (3 - 2 * x) for x in (u, v, w))
    i, j, k = (int(x) for x in f)
    acc = 0.0
    for di in range(2):
        for dj in range(2):
            for dk in range(2):
                c = rv[pm[0][(i + di) & 255] ^ pm[1][(j + dj) & 255] ^ pm[2][(k + dk) & 255]]
                acc += ((di * uu + (1 - di) * (1 - uu)) * (dj * vv + (1 - dj) * (1 - vv)) * (dk * ww + (1 - dk) * (1 - ww))
                        * (c[0] * (u - di) + c[1] * (v - dj) + c[2] * (w - dk)))
    return acc


# ---- family: Material.scatter / emitted (material.zig:16-121) ------------------------------------------------------------
def scatter_scene():
    if "scatter" in _scene_cache:
        return _scene_cache["scatter"]
    b = scene_util.DescBuilder()
    mats = [b.diffuse(b.solid((0.4, 0.2, 0.1))), b.diffuse(b.checker((0.2, 0.3, 0.1), (0.9, 0.9, 0.9))), b.metal((0.7, 0.6, 0.5), 0.0),
            b.metal((0.8, 0.8, 0.9), 0.35), b.glass(1.5), b.glass(1.0 / 1.5), b.light(b.solid((15, 15, 15)))]
    for m in mats:
        b.sphere((0, 0, 0), 1.0, m)
    d = b.build()
    _scene_cache["scatter"] = dict(desc=d, osc=ob.OracleScene.from_desc(d), n_mats=len(mats))
    return _scene_cache["scatter"]


def scatter_inputs(n, seed):
    rng = np.random.default_rng(seed)
    n_mats = scatter_scene()["n_mats"]
    rows = []
    for i in range(n):
        nrm = rng.normal(size=3)
        nrm /= np.linalg.norm(nrm)
        d = rng.normal(size=3) * rng.uniform(0.05, 3.0)
        front = bool(np.dot(d, nrm) < 0)
        if not front:
            nrm = -nrm  # HitRecord.normal always faces the ray
        if i % 9 == 0:  # grazing incidence
            d = d - np.dot(d, nrm) * nrm * 0.999999
        ray = [*rng.uniform(-5, 5, 3), *d, rng.uniform()]
        rec = [*rng.uniform(-5, 5, 3), *nrm, rng.uniform(), rng.uniform(), 1.0 if front else 0.0, rng.uniform(0.01, 9)]
        rows.append((int(i % n_mats), ray, rec, int(rng.integers(1, 1 << 40))))
    return rows


def ref_scatter(rows):
    import ref_transpile as rt
    R, A, co, rt_ = rt.ref(), rt.runtime.Anon, rt.runtime.co, rt.runtime
    scatter_scene()
    _, mats = ref_world("scatter")
    out = []
    for mi, ray, rec10, seed in rows:
        gen = rt_.Xoshiro256(seed)
        rng = gen.random()
        r_in = co(R.ray.Ray, A(origin=rt._v3(ray[0:3]), dir=rt._v3(ray[3:6]), time=float(ray[6])))
        rec = rt_.undefined(R.hit_record.HitRecord)
        rec.p, rec.normal = co(R.vec.Vec3, rt._v3(rec10[0:3])), co(R.vec.Vec3, rt._v3(rec10[3:6]))
        rec.u, rec.v, rec.front_face, rec.t = float(rec10[6]), float(rec10[7]), rec10[8] != 0.0, float(rec10[9])
        mat = mats[mi].get()
        rec.material = mat
        att, sc = rt_.undefined(R.vec.Vec3), rt_.undefined(R.ray.Ray)
        em = mat.emitted(rec.u, rec.v, rec.p)  # main.zig:116 evaluates emitted before scatter
        ok = mat.scatter(r_in, rec, att, sc, rng)
        o = [bool(ok), gen.draws, em.x, em.y, em.z]
        if ok or mat.deref().tag == "metal":  # an absorbed metal bounce still wrote attenuation and the ray
            o += [att.x, att.y, att.z, sc.origin.x, sc.origin.y, sc.origin.z, sc.dir.x, sc.dir.y, sc.dir.z, sc.time]
        out.append(o)
    return out


def orc_scatter(rows):
    sc = scatter_scene()
    kinds = [sc["desc"].materials[i].kind for i in range(sc["n_mats"])]
    out = []
    for mi, ray, rec10, seed in rows:
        r = sc["osc"].scatter(mi, ray, rec10, seed)
        o = [r["ok"], r["draws"], *fl(r["emitted"])]
        if r["ok"] or kinds[mi] == abi.MAT_METAL:  # an absorbed metal bounce still wrote attenuation and the ray
            o += [*fl(r["attenuation"]), *fl(r["ray"])]
        out.append(o)
    return out


# ---- family: Camera.init / getRay (main.zig:52-100) -------------------------------------------------------------------------
def camera_inputs(n, seed):
    rng = np.random.default_rng(seed)
    rows = []
    for i in range(n):
        rows.append(dict(look_from=fl(rng.uniform(-20, 20, 3)), look_at=fl(rng.uniform(-2, 2, 3)), vup=[0.0, 1.0, 0.0],
                         vfov=float(rng.uniform(10, 90)), aspect=float(rng.uniform(0.5, 2.5)),
                         aperture=float(0.0 if i % 3 == 0 else rng.uniform(0, 2)), focus=float(rng.uniform(1, 20)),
                         t0=0.0, t1=float(rng.uniform(0.1, 2)), s=float(rng.uniform(0, 1.01)), t=float(rng.uniform(0, 1.01)),
                         seed=int(rng.integers(1, 1 << 40))))
    return rows


def ref_camera(rows):
    import ref_transpile as rt
    R, rt_ = rt.ref(), rt.runtime
    out = []
    for c in rows:
        cam = R.main.Camera.init(rt._v3(c["look_from"]), rt._v3(c["look_at"]), rt._v3(c["vup"]), c["vfov"], c["aspect"], c["aperture"],
                                 c["focus"], c["t0"], c["t1"])
        gen = rt_.Xoshiro256(c["seed"])
        r = cam.getRay(gen.random(), c["s"], c["t"])
        o = []
        for f in ("origin", "horizontal", "vertical", "lower_left_corner", "u", "v", "w"):
            vv = getattr(cam, f)
            o += [vv.x, vv.y, vv.z]
        o += [cam.lens_radius, cam.time0, cam.time1, r.origin.x, r.origin.y, r.origin.z, r.dir.x, r.dir.y, r.dir.z, r.time, gen.draws]
        out.append(o)
    return out


def orc_camera(rows):
    out = []
    for c in rows:
        cam = ob.camera_init(c["look_from"], c["look_at"], c["vup"], c["vfov"], c["aspect"], c["aperture"], c["focus"], c["t0"], c["t1"])
        ray, draws = ob.get_ray(cam, c["seed"], c["s"], c["t"])
        o = []
        for f in ("origin", "horizontal", "vertical", "lower_left_corner", "u", "v", "w"):
            o += fl(getattr(cam, f))
        o += [cam.lens_radius, cam.time0, cam.time1, *fl(ray), draws]
        out.append(o)
    return out


# ---- family: rayColor (main.zig:103-122) ---------------------------------------------------------------------------------------
def ref_ray_color(key, rays, bg, depth, seeds):
    import ref_transpile as rt
    R, A, co, rt_ = rt.ref(), rt.runtime.Anon, rt.runtime.co, rt.runtime
    world, _ = ref_world(key)
    out = []
    for r7, seed in zip(rays, seeds):
        gen = rt_.Xoshiro256(int(seed))
        r = co(R.ray.Ray, A(origin=rt._v3(r7[0:3]), dir=rt._v3(r7[3:6]), time=float(r7[6])))
        c = R.main.rayColor(r, rt._v3(bg), world, gen.random(), depth)
        out.append([c.x, c.y, c.z, gen.draws])
    return out


def orc_ray_color(key, rays, bg, depth, seeds):
    osc = scene(key)["osc"]
    out = []
    for r7, seed in zip(rays, seeds):
        r = osc.ray_color(r7, bg, depth, int(seed))
        out.append([*fl(r["color"]), r["draws"]])
    return out


# ---- family: rand.zig:22-40 -------------------------------------------------------------------------------------------------
def ref_samplers(seed, n):
    import ref_transpile as rt
    R, rt_ = rt.ref(), rt.runtime
    out = []
    for which, fn in enumerate((R.rand.randomPointInUnitSphere, R.rand.randomPointInUnitDisk, R.rand.randomUnitVector)):
        rng = rt_.Xoshiro256(seed + which).random()
        for _ in range(n):
            v = fn(rng)
            out.append([v.x, v.y, v.z])
    return out


def orc_samplers(seed, n):
    out = []
    for which in range(3):
        out += [fl(v) for v in ob.samplers(seed + which, which, n)]
    return out


# ---- family: Perlin.init (perlin.zig:18-38, permute :93-101) ---------------------------------------------------------------------
def ref_perlin_tables(seed=42):
    import ref_transpile as rt
    R, rt_ = rt.ref(), rt.runtime
    pn = R.perlin.Perlin.init(rt_.Allocator(), rt_.Xoshiro256(seed).random())
    return [[v.x, v.y, v.z] for v in pn.randomVec.items], [list(pn.permX.items), list(pn.permY.items), list(pn.permZ.items)]


def orc_perlin_tables():
    rv, pm = scene("builtin:3")["osc"].perlin_tables(0)  # scene 3 draws nothing before Texture.makeNoise (main.zig:144)
    return [fl(v) for v in rv], [[int(x) for x in row] for row in pm]
