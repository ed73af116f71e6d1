"""Known-answer tests that pin the CPU oracle.

The reference ships no tests or golden vectors (parity unpinned); these values are derived from the
reference's own formulas (file:line in each test) and, for the RNG, from the published Xoshiro256++ /
SplitMix64 algorithms that Zig's std.Random.DefaultPrng implements.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def test_xoshiro_sequence_and_splitmix_state(oracle):
    # Zig std Xoshiro256 "sequence" test vector for seed 0 (std.Random.DefaultPrng, src/main.zig:300)
    out = np.zeros(6, dtype=np.uint64)
    st = np.zeros(4, dtype=np.uint64)
    oracle.lib().orc_kat_xoshiro(0, 6, out.ctypes.data_as(C.POINTER(C.c_uint64)), st.ctypes.data_as(C.POINTER(C.c_uint64)))
    assert [int(x) for x in out] == [0x53175d61490b23df, 0x61da6f3dc380d507, 0x5c0fdf91ec9a7bfc,
                                     0x02eebf8c3bbe5e1a, 0x7eca04ebaf4a5eea, 0x0543c37757f08d9a]
    assert [int(x) for x in st] == [0xe220a8397b1dcdaf, 0x6e789e6aa1b965f4, 0x06c45d188009454f, 0xf88bb8a8724c81ec]


def test_real01_seed42(oracle):
    # src/rtw/rand.zig:13-15 over Random.float(f64) (52 mantissa bits, geometric exponent)
    r = np.zeros(4)
    oracle.lib().orc_kat_real01(42, 4, _dp(r))
    assert r.tolist() == [0.6969372117194047, 0.47274502314109507, 0.5152564274971367, 0.9257049799795629]
    big = np.zeros(100000)
    oracle.lib().orc_kat_real01(7, big.size, _dp(big))
    assert big.min() >= 0.0 and big.max() < 1.0 and abs(big.mean() - 0.5) < 0.005


def test_uint_less_than_is_a_bounded_uniform(oracle):
    out = np.zeros(20000, dtype=np.uint64)
    oracle.lib().orc_kat_uint_less_than(3, 7, out.size, out.ctypes.data_as(C.POINTER(C.c_uint64)))
    assert out.max() == 6
    counts = np.bincount(out.astype(np.int64), minlength=7)
    assert np.all(np.abs(counts - out.size / 7) < 5 * np.sqrt(out.size / 7))


def test_camera_scene1(oracle):
    # Camera.init src/main.zig:52-89 with the scene-1 settings of main.zig:320-326
    cam = oracle.camera_init((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 1.5, 0.1)
    np.testing.assert_allclose(list(cam.w), [0.963624111659, 0.148249863332, 0.222374794998], atol=1e-11)
    np.testing.assert_allclose(list(cam.u), [0.224859506699, 0, -0.974391195695], atol=1e-11)
    np.testing.assert_allclose(list(cam.v), [-0.144453361594, 0.988949937066, -0.033335391137], atol=1e-11)
    np.testing.assert_allclose(list(cam.horizontal), [1.189463936994, 0, -5.154343726972], atol=1e-11)
    np.testing.assert_allclose(list(cam.vertical), [-0.509420502061, 3.487571129492, -0.117558577399], atol=1e-11)
    np.testing.assert_allclose(list(cam.lower_left_corner), [3.023737165939, -1.226284198068, 3.412203202202], atol=1e-11)
    assert cam.lens_radius == 0.05 and cam.time0 == 0.0 and cam.time1 == 1.0


def test_camera_cornell(oracle):
    # main.zig:352-361: the image is mirrored in x (u = (-1,0,0)), as in the book
    cam = oracle.camera_init((278, 278, -800), (278, 278, 0), (0, 1, 0), 40.0, 1.0, 0.0)
    np.testing.assert_allclose(list(cam.w), [0, 0, -1], atol=1e-15)
    np.testing.assert_allclose(list(cam.u), [-1, 0, 0], atol=1e-15)
    np.testing.assert_allclose(list(cam.horizontal), [-7.279404685324, 0, 0], atol=1e-11)
    np.testing.assert_allclose(list(cam.lower_left_corner), [281.639702342662, 274.360297657338, -790], atol=1e-9)


@pytest.mark.parametrize("p,uv", [((1, 0, 0), (0.5, 0.5)), ((0, 1, 0), (0.5, 1.0)), ((0, 0, 1), (0.25, 0.5)),
                                  ((0, -1, 0), (0.5, 0.0)), ((0, 0, -1), (0.75, 0.5))])
def test_sphere_uv(oracle, p, uv):
    # getSphereUv src/rtw/hittable.zig:145-150
    out = np.zeros(2)
    oracle.lib().orc_kat_sphere_uv(_dp(np.array(p, dtype=np.float64)), _dp(out))
    np.testing.assert_allclose(out, uv, atol=1e-15)


def test_sphere_uv_seam_depends_on_sign_of_zero(oracle):
    out = np.zeros(2)
    oracle.lib().orc_kat_sphere_uv(_dp(np.array([-1.0, 0.0, 0.0])), _dp(out))   # -z = -0.0 -> atan2(-0, -1) = -pi
    assert out[0] == 0.0 and out[1] == 0.5
    oracle.lib().orc_kat_sphere_uv(_dp(np.array([-1.0, 0.0, -0.0])), _dp(out))  # -z = +0.0 -> atan2(+0, -1) = +pi
    assert out[0] == 1.0


def test_schlick_reflect_refract(oracle):
    L = oracle.lib()
    # reflectance src/rtw/material.zig:87-91
    assert abs(L.orc_kat_reflectance(1.0, 1.5) - 0.04) < 1e-15
    assert abs(L.orc_kat_reflectance(0.0, 1.5) - 1.0) < 1e-15
    assert abs(L.orc_kat_reflectance(0.5, 1.5) - 0.07) < 1e-15
    o = np.zeros(3)
    L.orc_kat_reflect(_dp(np.array([1.0, -1.0, 0.0])), _dp(np.array([0.0, 1.0, 0.0])), _dp(o))  # material.zig:112-114
    assert o.tolist() == [1.0, 1.0, 0.0]
    d = np.array([0.6, -0.8, 0.0])
    L.orc_kat_refract(_dp(d), _dp(np.array([0.0, 1.0, 0.0])), 1.0, _dp(o))  # material.zig:116-121: eta=1 -> unchanged
    np.testing.assert_allclose(o, d, atol=1e-15)


@pytest.mark.parametrize("avg,u8", [(0.0, 0), (0.25, 128), (0.5, 181), (0.7, 214), (0.8, 228), (1.0, 255), (4.0, 255)])
def test_resolve_quantisation(oracle, avg, u8):
    # src/main.zig:395-400: 256 * clamp(sqrt(sum/spp), 0, 0.999) truncated
    assert oracle.lib().orc_kat_resolve(avg * 8, 8) == u8


def test_sphere_hit_records(oracle):
    # Sphere.hit src/rtw/hittable.zig:95-131 on a one-sphere scene
    import rtw_b200
    abi = rtw_b200.abi
    prim = abi.Prim(kind=abi.PRIM_SPHERE, material=0, xform=-1)
    prim.v[0:4] = [0.0, 0.0, -1.0, 0.5]
    mat = abi.Material(kind=abi.MAT_DIELECTRIC, texture=-1, param=1.5)
    d = abi.SceneDesc(n_prims=1, prims=C.pointer(prim), n_materials=1, materials=C.pointer(mat), time0=0, time1=1)
    s = oracle.OracleScene.from_desc(d, keep=(prim, mat))
    h = s.hit_record([0, 0, 0, 0, 0, -1, 0])
    assert h["t"] == 0.5 and h["front_face"] and h["prim_id"] == 0
    np.testing.assert_allclose(h["p"], [0, 0, -0.5]); np.testing.assert_allclose(h["normal"], [0, 0, 1])
    assert (h["u"], h["v"]) == (0.25, 0.5)
    h = s.hit_record([0, 0, -1, 0, 0, -1, 0])  # from the centre: far root, back face, flipped normal
    assert h["t"] == 0.5 and not h["front_face"]
    np.testing.assert_allclose(h["normal"], [0, 0, 1])
    assert s.hit_record([0, 0, 0, 0, 0, 1, 0]) is None


def test_checker_and_image_texture(oracle, earth_rgba):
    s1 = oracle.OracleScene.builtin(1)
    d = s1.export()
    checker = next(i for i in range(d.n_textures) if d.textures[i].kind == 1)
    # texture.zig:79-82 with the colours of main.zig:165: sin(1)^3 > 0 -> even, one negative factor -> odd
    assert s1.texture_value(checker, 0, 0, (0.1, 0.1, 0.1)).tolist() == [0.9, 0.9, 0.9]
    assert s1.texture_value(checker, 0, 0, (0.1, 0.1, -0.1)).tolist() == [0.2, 0.3, 0.1]
    s4 = oracle.OracleScene.builtin(4, image=earth_rgba)
    h, w = earth_rgba.shape[:2]
    assert (w, h) == (500, 282)

    def texel(i, j):
        px = earth_rgba[j, i]
        return [0, 0, 1.0] if px[3] == 0 else [px[0] / 255.0, px[1] / 255.0, px[2] / 255.0]
    # texture.zig:121-144: (u,v)=(0,1) -> texel (0,0); (1,0) -> i=499, j=281 (row clamp fixed to height-1)
    np.testing.assert_allclose(s4.texture_value(0, 0.0, 1.0, (0, 0, 0)), texel(0, 0))
    np.testing.assert_allclose(s4.texture_value(0, 1.0, 0.0, (0, 0, 0)), texel(499, 281))
    np.testing.assert_allclose(s4.texture_value(0, 0.5, 0.5, (0, 0, 0)), texel(250, 141))
    a0 = np.argwhere(earth_rgba[..., 3] == 0)[0]
    u, v = (a0[1] + 0.5) / w, 1.0 - (a0[0] + 0.5) / h
    assert s4.texture_value(0, u, v, (0, 0, 0)).tolist() == [0, 0, 1.0]
    frac0 = (earth_rgba[..., 3] == 0).mean()
    assert 0.66 < frac0 < 0.68  # 66.8 % of texels are "ocean"


def test_leaf_bounding_boxes(oracle):
    # boudingBox rules src/rtw/hittable.zig:133-143, 203-217, 305-316, 491-498, 516-556
    s1 = oracle.OracleScene.builtin(1)
    mn, mx = s1.bounding_box(0)
    assert mn.tolist() == [-1000, -2000, -1000] and mx.tolist() == [1000, 0, 1000]
    d = s1.export()
    k = next(i for i in range(d.n_prims) if d.prims[i].kind == 1)
    p = d.prims[k]
    mn, mx = s1.bounding_box(k)
    np.testing.assert_allclose(mn, [p.v[0] - 0.2, p.v[1] - 0.2, p.v[2] - 0.2])
    np.testing.assert_allclose(mx, [p.v[3] + 0.2, p.v[4] + 0.2, p.v[5] + 0.2])
    s6 = oracle.OracleScene.builtin(6)
    mn, mx = s6.bounding_box(2)  # light: xzRect k=554 padded +-1e-4 on y
    np.testing.assert_allclose(mn, [213, 553.9999, 227]); np.testing.assert_allclose(mx, [343, 554.0001, 332])
    mn, mx = s6.bounding_box(6)  # Translate(265,0,295) o RotateY(15 deg) o Box(165,330,165)
    c, s = np.cos(np.radians(15)), np.sin(np.radians(15))
    xs = [c * x + s * z for x in (0, 165) for z in (0, 165)]
    zs = [-s * x + c * z for x in (0, 165) for z in (0, 165)]
    np.testing.assert_allclose(mn, [min(xs) + 265, 0, min(zs) + 295]); np.testing.assert_allclose(mx, [max(xs) + 265, 330, max(zs) + 295])


def test_aabb_slab_test(oracle):
    # Aabb.hit src/rtw/aabb.zig:8-45
    L = oracle.lib()
    mn, mx = np.array([-1.0, -1, -1]), np.array([1.0, 1, 1])
    assert L.orc_kat_aabb_hit(_dp(mn), _dp(mx), _dp(np.array([0, 0, -5, 0.1, 0.1, 1, 0.0])), 0.001, 1e30) == 1
    assert L.orc_kat_aabb_hit(_dp(mn), _dp(mx), _dp(np.array([0, 0, -5, 0.1, 0.1, -1, 0.0])), 0.001, 1e30) == 0
    assert L.orc_kat_aabb_hit(_dp(mn), _dp(mx), _dp(np.array([0, 0, -5, 0.1, 0.1, 1, 0.0])), 0.001, 3.9) == 0


def test_scene_inventory(oracle, earth_rgba):
    # object counts / order of the six builders, src/main.zig:124-293
    s1 = oracle.OracleScene.builtin(1).export()
    kinds = [s1.prims[i].kind for i in range(s1.n_prims)]
    assert kinds[:4] == [0, 0, 0, 0] and 4 < s1.n_prims <= 40
    assert s1.prims[0].v[3] == 1000 and s1.prims[0].v[1] == -1000
    s6 = oracle.OracleScene.builtin(6).export()
    assert s6.n_prims == 18 and s6.n_xforms == 4 and s6.n_materials == 4
    assert [s6.prims[i].kind for i in range(6)] == [4, 4, 3, 3, 3, 2]
    assert [s6.prims[i].kind for i in range(6, 12)] == [2, 2, 3, 3, 4, 4]   # Box side order hittable.zig:437-442
    assert s6.xforms[s6.prims[6].xform].kind == 1 and s6.xforms[s6.xforms[s6.prims[6].xform].outer].kind == 0
    assert oracle.OracleScene.builtin(2).export().n_prims == 2
    assert oracle.OracleScene.builtin(3).export().n_perlins == 1
    assert oracle.OracleScene.builtin(4, image=earth_rgba).export().n_images == 1
    assert oracle.OracleScene.builtin(5).export().n_prims == 3
    cfg = oracle.OracleScene.builtin(6).config()
    assert (cfg["width"], cfg["height"], cfg["spp"], cfg["max_depth"]) == (600, 600, 200, 50)
    cfg = oracle.OracleScene.builtin(1).config()
    assert (cfg["width"], cfg["height"], cfg["spp"]) == (600, 400, 50)


def test_nested_graph_equals_flattened_graph(oracle):
    """Translate(RotateY(Box)) evaluated as nested lists (reference) == per-rect instance chains (the ABI)."""
    s6 = oracle.OracleScene.builtin(6)
    flat = oracle.OracleScene.from_desc(s6.export(), keep=s6)
    cam = s6.default_camera()
    a = s6.primary_hits(cam, 200, 200, 64)
    b = flat.primary_hits(cam, 200, 200, 64)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])


def test_cpu_bvh_equals_linear_scan(oracle):
    s = oracle.OracleScene.builtin(1, grid=11)
    rng = np.random.default_rng(5)
    n = 20000
    rays = np.zeros((n, 7))
    rays[:, 0:3] = rng.uniform(-12, 12, (n, 3)) * [1, 0.2, 1] + [0, 3, 0]
    rays[:, 3:6] = rng.normal(size=(n, 3))
    rays[:, 6] = rng.uniform(0, 1, n)
    for prec in (64, 32):
        a = s.trace_rays(rays, prec, use_bvh=False)
        b = s.trace_rays(rays, prec, use_bvh=True)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        assert (a[0] != 0xFFFFFFFF).mean() > 0.3


def test_white_furnace_and_energy_bound(oracle):
    """rayColor src/main.zig:103-122: a diffuse sphere of albedo a under a white background converges to known
    values: a camera ray that hits once sees a*1 + ... ; radiance never exceeds the background."""
    import rtw_b200
    abi = rtw_b200.abi
    prim = abi.Prim(kind=abi.PRIM_SPHERE, material=0, xform=-1)
    prim.v[0:4] = [0.0, 0.0, 0.0, 1.0]
    tex = abi.Texture(kind=abi.TEX_SOLID, a=-1, b=-1)
    tex.color[:] = [0.5, 0.5, 0.5]
    mat = abi.Material(kind=abi.MAT_DIFFUSE, texture=0)
    d = abi.SceneDesc(n_prims=1, prims=C.pointer(prim), n_materials=1, materials=C.pointer(mat), n_textures=1,
                      textures=C.pointer(tex), time0=0, time1=1)
    s = oracle.OracleScene.from_desc(d, keep=(prim, mat, tex))
    cam = oracle.camera_init((0, 0, 4), (0, 0, 0), (0, 1, 0), 20.0, 1.0, 0.0)
    r = s.render(cam, 32, 32, 64, 50, (1.0, 1.0, 1.0), seed=3)
    img = r["accum"] / 64
    assert img.max() <= 1.0 + 1e-12
    centre = img[12:20, 12:20].mean()
    assert abs(centre - 0.5) < 0.02  # convex body: every scattered ray escapes -> exactly albedo
    r0 = s.render(cam, 8, 8, 4, 50, (0.0, 0.0, 0.0), seed=3)
    assert r0["accum"].max() == 0.0  # no emitters, black background


def test_golden_fixtures(oracle, earth_rgba):
    """Fixtures written by tests/golden/make_golden.py from this oracle (the reference has none to offer):
    they freeze the restatement so that a later edit cannot silently change it."""
    with open(os.path.join(GOLDEN, "oracle_golden.json")) as f:
        g = json.load(f)
    for key, want in g["primary_hits"].items():
        sid, w, h, prec = (int(x) for x in key.split("_"))
        s = oracle.OracleScene.builtin(sid, image=earth_rgba if sid == 4 else None)
        ids, t, n = s.primary_hits(s.default_camera(), w, h, prec)
        assert ids.astype(np.int64).sum() == want["id_sum"], key
        assert int((ids == 0xFFFFFFFF).sum()) == want["miss"], key
        assert abs(t.sum() - want["t_sum"]) <= 1e-9 * abs(want["t_sum"]), key
    for key, want in g["render_1t"].items():
        sid, w, h, spp = (int(x) for x in key.split("_"))
        s = oracle.OracleScene.builtin(sid, image=earth_rgba if sid == 4 else None)
        cfg = s.config()
        r = s.render(s.default_camera(), w, h, spp, 50, cfg["background"], precision=64, nthreads=1, continue_stream=True)
        assert r["rays"] == want["rays"], key
        np.testing.assert_allclose(r["accum"].sum(axis=(0, 1)), want["sum"], rtol=1e-9)
        assert int(r["rgb8"].astype(np.int64).sum()) == want["rgb8_sum"], key


def test_perlin_known_answers(oracle):
    """Perlin.noise hand-evaluated from perlin.zig:47-77,103-124 (NOT from the oracle).  Tables: identity permutations,
    gradient (+1,0,0) where the table index is odd and (-1,0,0) where it is even, so in the cell [0,1)^3 the corner
    (di,dj,dk) has gradient s(di)s(dj)s(dk)(1,0,0), s(0) = -1, s(1) = +1, and with the Hermite-smoothed U,V,W that
    perlin.zig:77 hands to perlinInterp (whose weight vector, :114, is therefore (U-i, V-j, W-k)) the sum factorises:

        noise = [-(1-U)U + U(U-1)] * (2V-1) * (2W-1) = -2U(1-U)(2V-1)(2W-1).

    The book's perlin_interp (raw u in the weight vector) gives [-(1-U)u + U(u-1)] for the first factor: different
    unless u = U.  Every number below is a dyadic rational: the expected values are exact in f64."""
    import ctypes as C
    import scene_util
    from rtw_b200 import abi
    b = scene_util.DescBuilder()
    b.texs.append(abi.Texture(kind=abi.TEX_NOISE, a=0, b=-1, scale=4.0))
    b.sphere((0, 0, 0), 1.0, b.diffuse(0))
    d = b.build()
    rv = np.zeros((256, 3))
    rv[:, 0] = np.where(np.arange(256) & 1, 1.0, -1.0)
    perm = np.arange(256, dtype=np.uint32)
    pl = abi.Perlin(ranvec=rv.ctypes.data_as(C.POINTER(C.c_double)), perm_x=perm.ctypes.data_as(C.POINTER(C.c_uint32)),
                    perm_y=perm.ctypes.data_as(C.POINTER(C.c_uint32)), perm_z=perm.ctypes.data_as(C.POINTER(C.c_uint32)))
    arr = (abi.Perlin * 1)(pl)
    d.n_perlins, d.perlins = 1, arr
    s = oracle.OracleScene.from_desc(d)

    def smooth(x):
        return x * x * (3 - 2 * x)

    def expect(u, v, w):
        U, V, W = smooth(u), smooth(v), smooth(w)
        return -2 * U * (1 - U) * (2 * V - 1) * (2 * W - 1)

    assert expect(0.25, 0.25, 0.75) == 0.12462615966796875  # = 2 * (5/32)(27/32) * (11/16)^2, by hand
    for p in ((0.25, 0.25, 0.75), (0.5, 0.25, 0.75), (0.75, 0.5, 0.125), (0.125, 0.875, 0.375)):
        assert s.perlin_noise(0, p) == expect(*p), p
    # the book's formula would give -[(1-U)u + U(1-u)](2V-1)(2W-1) = 0.1550903... at the first point
    book = -((1 - smooth(0.25)) * 0.25 + smooth(0.25) * 0.75) * (2 * smooth(0.25) - 1) * (2 * smooth(0.75) - 1)
    assert abs(book - 0.15509033203125) < 1e-15 and abs(s.perlin_noise(0, (0.25, 0.25, 0.75)) - book) > 0.03
    # cells with negative coordinates wrap with & 255 (perlin.zig:67-69): p = (-0.75, 0.25, 0.75) has i = -1, so the x
    # indices are 255 and 0 -> the parity pattern flips sign along x
    assert s.perlin_noise(0, (-0.75, 0.25, 0.75)) == -expect(0.25, 0.25, 0.75)
    # turb = |sum of 7 octaves| (perlin.zig:79-91)
    acc, q, wgt = 0.0, np.array([0.25, 0.25, 0.75]), 1.0
    for _ in range(7):
        acc += wgt * s.perlin_noise(0, q)
        wgt *= 0.5
        q = q * 2.0
    assert s.perlin_turb(0, (0.25, 0.25, 0.75), 7) == abs(acc)
