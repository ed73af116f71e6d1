"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Tolerances (north_star: "hit ids and hit/miss masks bit-exact, a stated tolerance only on t/normal; converged
images within a stated RMSE tolerance because the RNG streams differ"):
  * reference-order probes (precision 32 / 64): prim ids, hit/miss mask, t AND normal bit-identical to the
    oracle evaluated at the same precision (both sides: no FMA contraction, IEEE div/sqrt);
  * production arithmetic (precision 0: robust fp32 sphere form, FMA on) against the f64 oracle: id mismatches
    <= 5e-4 of pixels (silhouettes / edges only), |dt| <= 2e-3 * max(1,|t|), |dn|_inf <= 2e-3 on agreeing pixels;
  * images: per-channel mean within 0.5 % of a 16x-spp oracle render (bias gate), and
    RMSE(GPU_N, ref) <= 1.25 * RMSE(oracle_N, ref) (noise gate).
"""
import ctypes as C
import os

import numpy as np
import pytest

import scene_util

pytestmark = pytest.mark.gpu
MISS = 0xFFFFFFFF


def _scene(rtw, oracle, sid, grid=3):
    hs = rtw.HostScene(sid, grid=grid)
    return hs, oracle.OracleScene.from_desc(hs.desc, keep=hs)


# ---------------------------------------------------------------------------------------------------------
# primary hits
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("sid,grid,W,H", [(1, 3, 600, 400), (1, 11, 480, 270), (2, 3, 300, 200), (3, 3, 150, 100),
                                          (4, 3, 300, 200), (5, 3, 300, 200), (6, 3, 600, 600), (7, 3, 480, 270)])
@pytest.mark.parametrize("precision", [32, 64])
def test_primary_hits_bit_exact(rtw, oracle, ctx, sid, grid, W, H, precision):
    hs, osc = _scene(rtw, oracle, sid, grid)
    cam = hs.camera(aspect=W / H)
    ctx.upload_scene(hs.desc, keep=hs)
    oid, ot, on = osc.primary_hits(cam, W, H, precision)
    assert 0 < (oid == MISS).sum() < oid.size or sid == 6
    for variant in (rtw.abi.VARIANT_MEGA_FLAT, rtw.abi.VARIANT_MEGA_BVH):
        gid, gt, gn = ctx.primary_hits(cam, W, H, precision, variant)
        assert np.array_equal(gid, oid), f"{(gid != oid).sum()} id mismatches"
        assert np.array_equal(gt, ot), "t differs in bits"
        assert np.array_equal(gn, on), "normal differs in bits"


@pytest.mark.parametrize("sid,grid,W,H", [(1, 3, 600, 400), (1, 11, 600, 400), (2, 3, 300, 200), (4, 3, 300, 200),
                                          (5, 3, 300, 200), (6, 3, 600, 600), (7, 3, 480, 270)])
def test_production_primary_hits_vs_f64_oracle(rtw, oracle, ctx, sid, grid, W, H):
    hs, osc = _scene(rtw, oracle, sid, grid)
    cam = hs.camera(aspect=W / H)
    ctx.upload_scene(hs.desc, keep=hs)
    oid, ot, on = osc.primary_hits(cam, W, H, 64)
    for variant in (rtw.abi.VARIANT_MEGA_FLAT, rtw.abi.VARIANT_MEGA_BVH):
        gid, gt, gn = ctx.primary_hits(cam, W, H, 0, variant)
        mism = gid != oid
        assert mism.mean() <= 5e-4, f"{mism.sum()} of {oid.size} ids differ from the f64 oracle"
        # ... and every one of them sits on a silhouette or a seam: 8-adjacent to a pixel where the ORACLE's id map changes, and
        # what the device saw there is either (a) the surface one of the oracle's neighbouring pixels shows — fp32 moved a
        # silhouette by less than a pixel — or (b) a surface that MEETS the oracle's at that point (|dt| within the t
        # tolerance): the shared edge of two faces of a box or of two walls, where the reference's own tie rule decides by
        # the last bit of t.  fp32 neither invents nor loses a surface.
        pad = np.pad(oid, 1, mode="edge")
        neigh = np.stack([pad[1 + dy:1 + dy + H, 1 + dx:1 + dx + W] for dy in (-1, 0, 1) for dx in (-1, 0, 1)])
        on_edge = (neigh != oid[None]).any(axis=0)
        assert on_edge[mism].all(), f"{(mism & ~on_edge).sum()} mismatching pixels are not on an id edge of the oracle map"
        silhouette = (neigh == gid[None]).any(axis=0)
        seam = (gid != MISS) & (oid != MISS) & (np.abs(gt - ot) <= 2e-3 * np.maximum(1.0, np.abs(ot)))
        bad = mism & ~(silhouette | seam)
        detail = [(int(y), int(x), int(gid[y, x]), int(oid[y, x]), float(gt[y, x]), float(ot[y, x]), neigh[:, y, x].tolist())
                  for y, x in list(zip(*np.nonzero(bad)))[:8]]
        assert not bad.any(), f"{bad.sum()} mismatching pixels are neither a shifted silhouette nor a seam: {detail}"
        ok = ~mism & (oid != MISS)
        assert (np.abs(gt - ot)[ok] <= 2e-3 * np.maximum(1.0, np.abs(ot[ok]))).all()
        assert np.abs(gn - on)[ok].max() <= 2e-3


def test_primary_hits_at_baseline_resolution(rtw, oracle, ctx):
    """BASELINE.json configs[1] geometry (1920x1080): fp32 reference-order ids bit-exact over all 2M pixels."""
    hs, osc = _scene(rtw, oracle, 1)
    cam = hs.camera(aspect=16 / 9)
    ctx.upload_scene(hs.desc, keep=hs)
    oid, ot, _ = osc.primary_hits(cam, 1920, 1080, 32)
    gid, gt, _ = ctx.primary_hits(cam, 1920, 1080, 32, rtw.abi.VARIANT_MEGA_BVH)
    assert np.array_equal(gid, oid) and np.array_equal(gt, ot)


# ---------------------------------------------------------------------------------------------------------
# arbitrary rays on random scenes (moving spheres, rects, instanced boxes), BVH == linear scan
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_trace_rays_random_scene(rtw, oracle, ctx, seed):
    rng = np.random.default_rng(seed)
    desc = scene_util.random_scene(rng)
    osc = oracle.OracleScene.from_desc(desc, keep=desc)
    ctx.upload_scene(desc, keep=desc)
    rays = scene_util.random_rays(rng, 50000)
    for precision in (32, 64):
        oid, ot, on, ouv = osc.trace_rays(rays, precision)
        assert 0.2 < (oid != MISS).mean() < 1.0
        for variant in (rtw.abi.VARIANT_MEGA_FLAT, rtw.abi.VARIANT_MEGA_BVH):
            gid, gt, gn, guv = ctx.trace_rays(rays, precision, variant)
            assert np.array_equal(gid, oid) and np.array_equal(gt, ot) and np.array_equal(gn, on)
            # u,v come from atan2/acos (libm vs CUDA libm): tolerance, not bits (seam u=0 == u=1)
            # acos(-y) is NaN when rounding leaves |y| a hair above 1 (same in the reference): same NaN mask
            assert np.array_equal(np.isnan(guv), np.isnan(ouv))
            du = np.abs(guv[:, 0] - ouv[:, 0])
            du = np.minimum(du, 1.0 - du)
            assert np.nanmax(du) <= (1e-5 if precision == 32 else 1e-12)
            assert np.nanmax(np.abs(guv[:, 1] - ouv[:, 1])) <= (2e-4 if precision == 32 else 1e-7)
    # production arithmetic: flat and BVH must agree with each other exactly, and with f64 almost everywhere
    a = ctx.trace_rays(rays, 0, rtw.abi.VARIANT_MEGA_FLAT)
    b = ctx.trace_rays(rays, 0, rtw.abi.VARIANT_MEGA_BVH)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    oid = osc.trace_rays(rays, 64)[0]
    assert (a[0] != oid).mean() <= 1e-3


@pytest.mark.parametrize("builder", ["sah", "lbvh"])
def test_bvh_equals_linear_scan_many_prims(rtw, oracle, ctx, knobs, builder):
    """10^4-sphere field (config C4's generator at grid 50): device BVH == device linear scan == oracle, for the
    host (binned SAH) and the device (Morton / radix tree) builder."""
    knobs.setenv("RTW_BVH_BUILDER", builder)
    hs = rtw.HostScene(rtw.host_lib.SCENE_SPHERE_FIELD, grid=50)
    osc = oracle.OracleScene.from_desc(hs.desc, keep=hs)
    ctx.upload_scene(hs.desc, keep=hs)
    rng = np.random.default_rng(11)
    n = 20000
    rays = np.zeros((n, 7))
    rays[:, 0:3] = rng.uniform(-40, 40, (n, 3)) * [1, 0.05, 1] + [0, 5, 0]
    rays[:, 3:6] = rng.normal(size=(n, 3)) * [1, 0.3, 1]
    rays[:, 6] = rng.uniform(0, 1, n)
    oid, ot, _, _ = osc.trace_rays(rays, 32, use_bvh=True)
    flat = ctx.trace_rays(rays, 32, rtw.abi.VARIANT_MEGA_FLAT)
    bvh = ctx.trace_rays(rays, 32, rtw.abi.VARIANT_MEGA_BVH)
    assert np.array_equal(flat[0], bvh[0]) and np.array_equal(flat[1], bvh[1])
    assert np.array_equal(bvh[0], oid) and np.array_equal(bvh[1], ot)
    st = ctx.stats()
    assert st["bvh_depth"] <= 64
    assert st["bvh_builder"] == (rtw.abi.BVH_BUILDER_LBVH if builder == "lbvh" else rtw.abi.BVH_BUILDER_SAH)


@pytest.mark.parametrize("seed", [5, 6])
def test_device_built_bvh_mixed_scene(rtw, oracle, ctx, knobs, seed):
    """Device builder on a mixed scene (rects, instanced boxes, moving and instanced spheres, a ground sphere that
    dwarfs the rest and is grafted next to the root): reference-order probe through the BVH == linear scan == oracle."""
    knobs.setenv("RTW_BVH_BUILDER", "lbvh")
    rng = np.random.default_rng(seed)
    desc = scene_util.random_scene(rng, n_spheres=400, n_moving=200, n_rects=60, n_boxes=8, n_inst_spheres=20)
    osc = oracle.OracleScene.from_desc(desc, keep=desc)
    ctx.upload_scene(desc, keep=desc)
    st = ctx.stats()
    assert st["bvh_builder"] == rtw.abi.BVH_BUILDER_LBVH and 2 <= st["bvh_depth"] <= 64 and st["bvh_nodes"] >= 4
    rays = scene_util.random_rays(rng, 40000)
    for precision in (32, 64):
        oid, ot, on, _ = osc.trace_rays(rays, precision)
        gid, gt, gn, _ = ctx.trace_rays(rays, precision, rtw.abi.VARIANT_MEGA_BVH)
        assert np.array_equal(gid, oid) and np.array_equal(gt, ot) and np.array_equal(gn, on)
    a = ctx.trace_rays(rays, 0, rtw.abi.VARIANT_MEGA_FLAT)
    b = ctx.trace_rays(rays, 0, rtw.abi.VARIANT_MEGA_BVH)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    # the same paths whichever builder made the tree (the BVH only culls)
    hs_cam = rtw.camera_init((26, 6, 8), (0, 0, 0), (0, 1, 0), 40.0, 1.5, 0.0, 10.0, 0.0, 1.0)
    p = ctx.params(96, 64, 0, 8, 8, 20, rtw.abi.VARIANT_MEGA_BVH, rtw.abi.FLAG_DETERMINISTIC, 7, (0.7, 0.8, 1.0))
    _, acc_l = ctx.render(hs_cam, p, want_accum=True)
    knobs.setenv("RTW_BVH_BUILDER", "sah")
    ctx.upload_scene(desc, keep=desc)
    assert ctx.stats()["bvh_builder"] == rtw.abi.BVH_BUILDER_SAH
    _, acc_s = ctx.render(hs_cam, p, want_accum=True)
    assert np.array_equal(acc_l, acc_s)


def test_device_built_bvh_picks_the_morton_grid_by_surface_area_cost(rtw, oracle, ctx):
    """The device builder fits one radix tree per candidate Morton grid (cells of a thin axis 1x, (1/thin)x, (1/thin)^2x,
    (1/thin)^4x as long as the cubic grid's) and emits the one with the smallest surface-area cost.  On a sheet of spheres
    (config C4's generator: 9*10^4 spheres on the curved ground) that is not the cubic grid, and the rays need fewer node
    tests; the hits are the same whichever grid made the tree, and equal to the oracle's."""
    hs = rtw.HostScene(rtw.host_lib.SCENE_SPHERE_FIELD, grid=150)
    osc = oracle.OracleScene.from_desc(hs.desc, keep=hs)
    cam = hs.camera(aspect=1.5)
    rng = np.random.default_rng(23)
    n = 20000
    rays = np.zeros((n, 7))
    rays[:, 0:3] = rng.uniform(-120, 120, (n, 3)) * [1, 0.05, 1] + [0, 8, 0]
    rays[:, 3:6] = rng.normal(size=(n, 3)) * [1, 0.3, 1]
    rays[:, 6] = rng.uniform(0, 1, n)
    oid, ot, _, _ = osc.trace_rays(rays, 32, use_bvh=True)
    p = ctx.params(192, 128, 0, 4, 4, 50, rtw.abi.VARIANT_MEGA_BVH, rtw.abi.FLAG_COUNT_EVENTS, 42, hs.background)
    tests_per_ray = {}
    for label, opts in (("auto", {}), ("cubic", {"RTW_LBVH_POW": "0"}), ("p2", {"RTW_LBVH_POW": "2"})):
        with ctx.options(RTW_BVH_BUILDER="lbvh", **opts):
            ctx.upload_scene(hs.desc, keep=hs)
            assert ctx.stats()["bvh_builder"] == rtw.abi.BVH_BUILDER_LBVH and ctx.stats()["bvh_depth"] <= 64
            gid, gt, _, _ = ctx.trace_rays(rays, 32, rtw.abi.VARIANT_MEGA_BVH)
            assert np.array_equal(gid, oid) and np.array_equal(gt, ot)
            ctx.render(cam, p)
            st = ctx.stats()
            tests_per_ray[label] = st["node_tests"] / st["rays"]
    # 9.76 against 10.00 here (the sheet is only 22 units high over 300 x 300); 46.9 against 51.7 on the 10^6-sphere config
    assert tests_per_ray["auto"] < 0.985 * tests_per_ray["cubic"], tests_per_ray
    assert tests_per_ray["auto"] <= 1.02 * tests_per_ray["p2"], tests_per_ray
    # a scene that fills its bounds: the extents are alike, the cubic grid is the only candidate (same tree as RTW_LBVH_POW=0)
    b = scene_util.DescBuilder()
    m = b.diffuse(b.solid((0.5, 0.5, 0.5)))
    for c in rng.uniform(-20, 20, (3000, 3)):
        b.sphere(tuple(c), 0.3, m)
    desc = b.build()
    rays2 = scene_util.random_rays(rng, 8000, extent=25.0)
    got = []
    for opts in ({}, {"RTW_LBVH_POW": "0"}):
        with ctx.options(RTW_BVH_BUILDER="lbvh", **opts):
            ctx.upload_scene(desc, keep=desc)
            got.append((ctx.stats()["bvh_nodes"], ctx.trace_rays(rays2, 0, rtw.abi.VARIANT_MEGA_BVH)))
    assert got[0][0] == got[1][0]
    assert np.array_equal(got[0][1][0], got[1][1][0]) and np.array_equal(got[0][1][1], got[1][1][1])


def test_device_built_bvh_coincident_centroids(rtw, oracle, ctx, knobs):
    """Equal Morton keys (stacks of concentric spheres): the radix tree splits ties by position and stays shallow."""
    knobs.setenv("RTW_BVH_BUILDER", "lbvh")
    b = scene_util.DescBuilder()
    m = b.diffuse(b.solid((0.5, 0.5, 0.5)))
    for k in range(300):
        b.sphere((-50, 0, 0), 0.5 + 0.001 * k, m)
        b.sphere((50, 0, 0), 0.5 + 0.001 * (k % 7), m)
    desc = b.build()
    osc = oracle.OracleScene.from_desc(desc, keep=desc)
    ctx.upload_scene(desc, keep=desc)
    st = ctx.stats()
    assert st["bvh_builder"] == rtw.abi.BVH_BUILDER_LBVH and st["bvh_depth"] <= 16
    rng = np.random.default_rng(3)
    n = 4000
    rays = np.zeros((n, 7))
    rays[:, 0:3] = rng.uniform(-60, 60, (n, 3)) * [1, 0.005, 0.005]
    rays[:, 3:6] = rng.normal(size=(n, 3)) * [1, 0.002, 0.002]
    oid, ot, _, _ = osc.trace_rays(rays, 32)
    gid, gt, _, _ = ctx.trace_rays(rays, 32, rtw.abi.VARIANT_MEGA_BVH)
    assert (oid != MISS).mean() > 0.3
    assert np.array_equal(gid, oid) and np.array_equal(gt, ot)


def test_ties_later_element_wins(rtw, oracle, ctx):
    """Two coincident rects: the reference's scan keeps the later one (hittable.zig:235-242, t_max inclusive)."""
    b = scene_util.DescBuilder()
    m = b.diffuse(b.solid((0.5, 0.5, 0.5)))
    for _ in range(3):
        b.rect(rtw.abi.PRIM_XY_RECT, -1, 1, -1, 1, 0.0, m)
    b.sphere((5, 0, 0), 1.0, m)
    b.sphere((5, 0, 0), 1.0, m)
    desc = b.build()
    ctx.upload_scene(desc, keep=desc)
    osc = oracle.OracleScene.from_desc(desc, keep=desc)
    rays = np.array([[0.1, 0.2, 3, 0, 0, -1, 0], [5, 0, 4, 0, 0, -1, 0.5]], dtype=np.float64)
    for precision in (32, 64, 0):
        for variant in (rtw.abi.VARIANT_MEGA_FLAT, rtw.abi.VARIANT_MEGA_BVH):
            ids = ctx.trace_rays(rays, precision, variant)[0]
            assert ids.tolist() == [2, 4], (precision, variant, ids)
    assert osc.trace_rays(rays, 64)[0].tolist() == [2, 4]


def test_box_edges_and_corners(rtw, oracle, ctx):
    """Rays aimed EXACTLY at the edges and corners of boxes (dyadic coordinates: every t is exact in fp32 and f64): all the
    faces meeting there are hit at the same t and the reference's scan keeps the LAST one in Box.init order z1, z0, y1, y0,
    x1, x0 (hittable.zig:437-442, 235-242).  The flat scan handles a box as three slabs (box records) and must reproduce that
    — and a room with a missing face (five walls, the Cornell layout), entered through the opening, from inside and on its
    seams.  Also checked: a box is watertight (a ray at an edge never slips between two faces)."""
    b = scene_util.DescBuilder()
    m = b.diffuse(b.solid((0.5, 0.5, 0.5)))
    b.box((0, 0, 0), (1, 2, 4), m)                      # prims 0..5: z1 z0 y1 y0 x1 x0
    t = b.translate((8, 0, 0))
    r = b.rotate_y(0.0, outer=t)                        # an instance chain with an exact (identity) rotation
    b.box((0, 0, 0), (2, 2, 2), m, xform=r)             # prims 6..11
    # room [16,32]^3 without its z0 wall, walls in the Cornell order: yz@x1, yz@x0, xz@y0, xz@y1, xy@z1   (prims 12..16)
    b.rect(rtw.abi.PRIM_YZ_RECT, 16, 32, 16, 32, 32, m)
    b.rect(rtw.abi.PRIM_YZ_RECT, 16, 32, 16, 32, 16, m)
    b.rect(rtw.abi.PRIM_XZ_RECT, 16, 32, 16, 32, 16, m)
    b.rect(rtw.abi.PRIM_XZ_RECT, 16, 32, 16, 32, 32, m)
    b.rect(rtw.abi.PRIM_XY_RECT, 16, 32, 16, 32, 32, m)
    desc = b.build()
    ctx.upload_scene(desc, keep=desc)
    osc = oracle.OracleScene.from_desc(desc, keep=desc)
    rays = np.array([
        [-1, -1, 1.5, 1, 1, 0, 0],      # edge x0/y0 of box 1 from outside: faces y0 (3), x0 (5) at t = 1       -> 5
        [-1, 1, -1, 1, 0, 1, 0],        # edge x0/z0: faces z0 (1), x0 (5)                                       -> 5
        [0.5, -1, -1, 0, 1, 1, 0],      # edge y0/z0: faces z0 (1), y0 (3)                                       -> 3
        [-1, -1, -1, 1, 1, 1, 0],       # corner (0,0,0): z0 (1), y0 (3), x0 (5)                                 -> 5
        [0.5, 1, 2, 0.5, 1, 2, 0],      # from inside to the corner (1,2,4): z1 (0), y1 (2), x1 (4) at t = 1     -> 4
        [0.5, 1, 1.5, 0.5, 1, 0, 0],    # from inside to the edge x1/y1: y1 (2), x1 (4)                          -> 4
        [7, -1, 1, 1, 1, 0, 0],         # instanced box: edge x0/y0 -> y0 (9), x0 (11)                            -> 11
        [7, -1, -1, 1, 1, 1, 0],        # instanced box: corner                                                  -> 11
        [24, 24, 0, 0, 0, 1, 0],        # room: in through the missing z0 wall, straight to the back wall        -> 16
        [24, 24, 24, 8, -8, 0, 0],      # room from inside to the seam x1/y0: yz@x1 (12), xz@y0 (14) at t = 1     -> 14
        [24, 24, 24, 8, 8, 8, 0],       # room corner (32,32,32): yz@x1 (12), xz@y1 (15), xy@z1 (16)              -> 16
        [24, 24, 24, -8, -8, 0, 0],     # seam x0/y0: yz@x0 (13), xz@y0 (14)                                      -> 14
        [24, 24, 24, 0, 0, -1, 0],      # out through the opening: nothing                                       -> miss
        [40, 24, 24, -1, 0, 0, 0],      # from outside through the x1 wall (a rect has two faces)                 -> 12
    ], dtype=np.float64)
    want = [5, 5, 3, 5, 4, 4, 11, 11, 16, 14, 16, 14, MISS, 12]
    assert osc.trace_rays(rays, 64)[0].tolist() == want
    for precision in (32, 64, 0):
        for variant in (rtw.abi.VARIANT_MEGA_FLAT, rtw.abi.VARIANT_MEGA_BVH):
            ids, t, _, _ = ctx.trace_rays(rays, precision, variant)
            assert ids.tolist() == want, (precision, variant, ids.tolist())
    # box records on and off trace the same surfaces (rays through random interior points of the three solids; rays aimed
    # exactly AT edges graze by construction and are covered by the exact cases above)
    rng = np.random.default_rng(8)
    n = 40000
    more = np.zeros((n, 7))
    more[:, 0:3] = rng.uniform(-6, 40, (n, 3))
    lo = np.array([[0, 0, 0], [8, 0, 0], [16, 16, 16]], dtype=np.float64)
    hi = np.array([[1, 2, 4], [10, 2, 2], [32, 32, 32]], dtype=np.float64)
    which = rng.integers(0, 3, n)
    more[:, 3:6] = lo[which] + (hi[which] - lo[which]) * rng.uniform(0, 1, (n, 3)) - more[:, 0:3]
    oid = osc.trace_rays(more, 64)[0]
    with_boxes = ctx.trace_rays(more, 0, rtw.abi.VARIANT_MEGA_FLAT)
    with ctx.options(RTW_BOX_PRIMS=0):
        ctx.upload_scene(desc, keep=desc)
        plain = ctx.trace_rays(more, 0, rtw.abi.VARIANT_MEGA_FLAT)
    assert (with_boxes[0] != plain[0]).mean() < 1e-3 and (with_boxes[0] != oid).mean() < 1e-3 and (oid != MISS).mean() > 0.5
    same = with_boxes[0] == plain[0]
    assert np.abs(with_boxes[1] - plain[1])[same].max() < 1e-4


def test_edge_cases(rtw, oracle, ctx):
    # empty scene: everything misses, the image is the background (resolve KAT: (.7,.8,1) -> 214,228,255)
    b = scene_util.DescBuilder()
    b.diffuse(b.solid((0.5, 0.5, 0.5)))
    desc = b.build()
    ctx.upload_scene(desc, keep=desc)
    cam = rtw.camera_init((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 1.5, 0.1)
    ids = ctx.primary_hits(cam, 33, 17, 32)[0]
    assert (ids == MISS).all()
    for variant in (rtw.abi.VARIANT_MEGA_FLAT, rtw.abi.VARIANT_MEGA_BVH):
        rgb, acc = ctx.render(cam, ctx.params(33, 17, 0, 5, 5, 50, variant), want_accum=True)
        assert (rgb == np.array([214, 228, 255], dtype=np.uint8)).all()
        assert (acc[..., 3] == 5).all()
    # ragged image sizes (not multiples of the 8x4 tile), 1x1, zero samples
    hs = rtw.HostScene(1)
    ctx.upload_scene(hs.desc, keep=hs)
    for W, H in ((1, 1), (7, 3), (9, 5), (64, 1), (1, 33)):
        rgb, acc = ctx.render(hs.camera(), ctx.params(W, H, 0, 3, 3), want_accum=True)
        assert rgb.shape == (H, W, 3) and (acc[..., 3] == 3).all() and np.isfinite(acc).all()
    rgb, acc = ctx.render(hs.camera(), ctx.params(8, 8, 4, 4, 4), want_accum=True)  # empty spp range
    assert (acc == 0).all() and (rgb == 0).all()
    # depth limit: max_depth = 1 -> only background / emitted light survives one bounce deep paths (main.zig:105-108)
    rgb1, acc1 = ctx.render(hs.camera(), ctx.params(60, 40, 0, 8, 8, 1), want_accum=True)
    assert acc1[..., :3].max() <= 8 * 1.0 + 1e-3
    assert (acc1[20, 30, :3] == 0).all()  # centre pixel: every sample hits a sphere, scatters once, then depth runs out
    # rays parallel to a rect plane / starting on it never hit it (t = inf or nan compares false)
    b = scene_util.DescBuilder()
    m = b.diffuse(b.solid((0.5, 0.5, 0.5)))
    b.rect(rtw.abi.PRIM_XZ_RECT, -1, 1, -1, 1, 0.0, m)
    desc = b.build()
    ctx.upload_scene(desc, keep=desc)
    osc = oracle.OracleScene.from_desc(desc, keep=desc)
    rays = np.array([[0, 0, -3, 0, 0, 1, 0], [0, 1, -3, 0, 0, 1, 0], [0, 1, 0, 0, -1, 0, 0], [0, 0, 0, 0, 1, 0, 0]], dtype=np.float64)
    want = osc.trace_rays(rays, 64)[0]
    for precision in (32, 64, 0):
        for variant in (rtw.abi.VARIANT_MEGA_FLAT, rtw.abi.VARIANT_MEGA_BVH):
            assert np.array_equal(ctx.trace_rays(rays, precision, variant)[0], want), (precision, variant)


def test_validation_errors(rtw, ctx):
    b = scene_util.DescBuilder()
    m = b.diffuse(b.solid((0.5, 0.5, 0.5)))
    b.sphere((0, 0, 0), 1, m)
    desc = b.build()
    desc.prims[0].material = 7
    with pytest.raises(rtw.RtwCudaError, match="material 7 out of range"):
        ctx.upload_scene(desc)
    desc.prims[0].material = 0
    desc.prims[0].kind = 9
    with pytest.raises(rtw.RtwCudaError, match="bad kind"):
        ctx.upload_scene(desc)
    desc.prims[0].kind = 0
    ctx.upload_scene(desc, keep=desc)
    cam = rtw.camera_init((0, 0, 5), (0, 0, 0), (0, 1, 0), 20.0, 1.0, 0.0)
    with pytest.raises(rtw.RtwCudaError, match="unknown variant"):
        ctx.render(cam, ctx.params(8, 8, 0, 1, 1, 50, 17))
    with pytest.raises(rtw.RtwCudaError, match="spp_end < spp_begin"):
        ctx.render(cam, ctx.params(8, 8, 3, 1, 1))
    c2 = rtw.Context(0)
    with pytest.raises(rtw.RtwCudaError, match="no scene uploaded"):
        c2.render(cam, ctx.params(8, 8, 0, 1, 1))
    c2.close()
    # the wavefront packs (sample, bounce) into 26 + 6 bits: parameters beyond that are refused, not mis-rendered
    with pytest.raises(rtw.RtwCudaError, match="max_depth <= 63"):
        ctx.render(cam, ctx.params(8, 8, 0, 1, 1, 64, rtw.abi.VARIANT_WAVEFRONT))
    ctx.render(cam, ctx.params(8, 8, 0, 1, 1, 63, rtw.abi.VARIANT_WAVEFRONT))
    with pytest.raises(rtw.RtwCudaError, match="unknown option"):
        ctx.set_option("RTW_NO_SUCH_KNOB", "1")
    # tables with a count but no pointer
    for field in ("materials", "textures"):
        b2 = scene_util.DescBuilder()
        b2.sphere((0, 0, 0), 1, b2.diffuse(b2.solid((0.5, 0.5, 0.5))))
        d2 = b2.build()
        setattr(d2, field, None)
        with pytest.raises(rtw.RtwCudaError, match=f"{field} is null"):
            ctx.upload_scene(d2)
    # checker graphs: a cycle, and nesting deeper than the device follows
    b3 = scene_util.DescBuilder()
    t0 = b3.checker((0, 0, 0), (1, 1, 1))
    b3.sphere((0, 0, 0), 1, b3.diffuse(t0))
    d3 = b3.build()
    d3.textures[t0].a = t0
    with pytest.raises(rtw.RtwCudaError, match="checker nesting"):
        ctx.upload_scene(d3)
    b4 = scene_util.DescBuilder()
    t = b4.solid((0.1, 0.2, 0.3))
    for _ in range(9):
        b4.texs.append(rtw.abi.Texture(kind=rtw.abi.TEX_CHECKER, a=t, b=t))
        t = len(b4.texs) - 1
    b4.sphere((0, 0, 0), 1, b4.diffuse(t))
    with pytest.raises(rtw.RtwCudaError, match="checker nesting"):
        ctx.upload_scene(b4.build())


def test_depth_zero_is_black(rtw, oracle, ctx):
    """rayColor returns (0,0,0) before intersecting anything when depth == 0 (main.zig:105-108): the frame is black and
    every sample is counted — all variants, including the sky pixels (no background either)."""
    hs = rtw.HostScene(1)
    ctx.upload_scene(hs.desc, keep=hs)
    osc = oracle.OracleScene.from_desc(hs.desc, keep=hs)
    assert (osc.ray_color([13, 2, 3, -13, -2, -3, 0.5], hs.background, 0, 1)["color"] == 0).all()
    for variant in (1, 2, 3):
        for flags in (0, rtw.abi.FLAG_DETERMINISTIC):
            rgb, acc = ctx.render(hs.camera(), ctx.params(40, 24, 0, 6, 6, 0, variant, flags, 42, hs.background), want_accum=True)
            assert (rgb == 0).all() and (acc[..., :3] == 0).all() and (acc[..., 3] == 6).all()


def test_negative_radius_sphere(rtw, oracle, ctx):
    """The hollow-glass idiom: a sphere of radius -0.9 inside one of radius 1 (Sphere.hit uses r*r and divides by r,
    hittable.zig:99,120: the inner normal flips).  Its boudingBox would be inverted; the library boxes |r|."""
    b = scene_util.DescBuilder()
    glass = b.glass(1.5)
    grey = b.diffuse(b.solid((0.5, 0.5, 0.5)))
    b.sphere((0, -1001, 0), 1000.0, grey)
    b.sphere((0, 0, 0), 1.0, glass)
    b.sphere((0, 0, 0), -0.9, glass)
    b.moving_sphere((2.5, 0, 0), (2.5, 0.3, 0), 0.0, 1.0, -0.5, glass)
    for k in range(70):  # enough primitives for a real tree
        b.sphere((np.cos(k) * 4, -0.8, np.sin(k) * 4), 0.2, grey)
    desc = b.build()
    ctx.upload_scene(desc, keep=desc)
    osc = oracle.OracleScene.from_desc(desc, keep=desc)
    rng = np.random.default_rng(4)
    rays = scene_util.random_rays(rng, 30000, extent=1.5)
    for precision in (32, 64):
        oid, ot, on, _ = osc.trace_rays(rays, precision)
        assert (oid == 2).sum() > 500 and (oid == 3).sum() > 100
        for variant in (1, 2):
            gid, gt, gn, _ = ctx.trace_rays(rays, precision, variant)
            assert np.array_equal(gid, oid) and np.array_equal(gt, ot) and np.array_equal(gn, on)
    oid, ot, on, _ = osc.trace_rays(rays, 64)
    for variant in (1, 2):
        gid, gt, gn, _ = ctx.trace_rays(rays, 0, variant)
        ok = gid == oid
        assert ok.mean() > 0.999
        hit = ok & (oid != MISS)
        assert np.abs(gn - on)[hit].max() < 2e-3  # the flipped inner normal included


def test_accumulate_calls_on_two_streams_do_not_race(rtw, ctx):
    """One work queue per context: two rtw_cuda_accumulate calls of one context issued on different streams are ordered
    on the device by the library.  The sum of both equals the two sample ranges rendered one after the other."""
    import torch
    hs = rtw.HostScene(1)
    ctx.upload_scene(hs.desc, keep=hs)
    W, H = 256, 144
    cam = hs.camera(aspect=W / H)
    det = rtw.abi.FLAG_DETERMINISTIC
    pa = ctx.params(W, H, 0, 24, 48, 50, 0, det, 42, hs.background)
    pb = ctx.params(W, H, 24, 48, 48, 50, 0, det, 42, hs.background)
    seq_a, seq_b = torch.zeros(H, W, 4, device="cuda"), torch.zeros(H, W, 4, device="cuda")
    ctx.accumulate(cam, pa, seq_a.data_ptr(), None)
    torch.cuda.synchronize()
    ctx.accumulate(cam, pb, seq_b.data_ptr(), None)
    torch.cuda.synchronize()
    par_a, par_b = torch.zeros_like(seq_a), torch.zeros_like(seq_b)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    for _ in range(3):
        par_a.zero_(); par_b.zero_()
        torch.cuda.synchronize()
        ctx.accumulate(cam, pa, par_a.data_ptr(), s1.cuda_stream)
        ctx.accumulate(cam, pb, par_b.data_ptr(), s2.cuda_stream)
        torch.cuda.synchronize()
        assert torch.equal(par_a, seq_a) and torch.equal(par_b, seq_b)


# ---------------------------------------------------------------------------------------------------------
# resolve (K4) and the multi-buffer reduce+resolve
# ---------------------------------------------------------------------------------------------------------
def test_resolve_kat_and_row_flip(rtw, oracle, ctx):
    import torch
    W, H, spp = 8, 5, 8
    avgs = [0.0, 0.25, 0.5, 0.7, 0.8, 1.0, 4.0, float("nan")]
    want = [0, 128, 181, 214, 228, 255, 255, 0]   # SURVEY App. C; NaN -> 0 (App. B Q16)
    acc = torch.zeros(H, W, 4, device="cuda")
    for i, a in enumerate(avgs):
        acc[:, i, 0] = a * spp
        acc[:, i, 1] = a * spp
    acc[:, :, 2] = torch.arange(H, device="cuda").float()[:, None] * spp / 16.0   # row marker in blue
    out = torch.zeros(H, W, 3, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.resolve(acc.data_ptr(), W, H, spp, out.data_ptr())
    torch.cuda.synchronize()
    o = out.cpu().numpy()
    assert o[0, :, 0].tolist() == want and o[3, :, 1].tolist() == want
    for j in range(H):  # accumulation row j lands in image row H-1-j (main.zig:396)
        assert o[H - 1 - j, 0, 2] == oracle.lib().orc_kat_resolve(j * spp / 16.0, spp)
    # every representable average: GPU quantisation == oracle quantisation
    vals = torch.linspace(0, 1.2, 4096, device="cuda")
    acc = torch.zeros(1, 4096, 4, device="cuda")
    acc[0, :, 0] = vals * 3
    out = torch.zeros(1, 4096, 3, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.resolve(acc.data_ptr(), 4096, 1, 3, out.data_ptr())
    torch.cuda.synchronize()
    got = out.cpu().numpy()[0, :, 0]
    sums = acc[0, :, 0].cpu().numpy().astype(np.float64)
    ref = np.array([oracle.lib().orc_kat_resolve(float(s), 3) for s in sums])
    assert (np.abs(got.astype(int) - ref.astype(int)) <= 1).all() and (got != ref).mean() < 0.01  # fp32 vs f64 sqrt at bin edges


def test_resolve_multi_sums_buffers(rtw, ctx):
    import torch
    W, H = 64, 32
    g = torch.Generator(device="cuda").manual_seed(1)
    bufs = [torch.rand(H, W, 4, device="cuda", generator=g) * 10 for _ in range(4)]
    one = torch.zeros(H, W, 3, dtype=torch.uint8, device="cuda")
    multi = torch.zeros_like(one)
    total = bufs[0] + bufs[1] + bufs[2] + bufs[3]
    torch.cuda.synchronize()
    ctx.resolve(total.data_ptr(), W, H, 40, one.data_ptr())
    ctx.resolve_multi([b.data_ptr() for b in bufs], W, H, 40, multi.data_ptr())
    torch.cuda.synchronize()
    assert torch.equal(one, multi)


# ---------------------------------------------------------------------------------------------------------
# rendered images: statistical parity with the f64 oracle
# ---------------------------------------------------------------------------------------------------------
def _image_gates(rtw, oracle, ctx, hs, osc, W, H, spp, variant, ref_mult=16):
    cam = hs.camera(aspect=W / H)
    nt = oracle.num_threads()
    ref = osc.render(cam, W, H, spp * ref_mult, hs.max_depth, hs.background, seed=1001, precision=64, nthreads=max(2, nt))
    cpu = osc.render(cam, W, H, spp, hs.max_depth, hs.background, seed=2002, precision=64, nthreads=max(2, nt))
    ref_img = ref["accum"] / (spp * ref_mult)
    cpu_img = cpu["accum"] / spp
    p = ctx.params(W, H, 0, spp, spp, hs.max_depth, variant, rtw.abi.FLAG_COUNT_EVENTS, 42, hs.background)
    rgb, acc = ctx.render(cam, p, want_accum=True)
    st = ctx.stats()
    gpu_img = acc[..., :3].astype(np.float64) / spp
    assert np.isfinite(gpu_img).all() and st["nan_pixels"] == 0
    # gate 1: bias
    gm, rm = gpu_img.mean(axis=(0, 1)), ref_img.mean(axis=(0, 1))
    rel = np.abs(gm - rm) / rm
    # gate 2: noise
    rmse_gpu = np.sqrt(((gpu_img - ref_img) ** 2).mean())
    rmse_cpu = np.sqrt(((cpu_img - ref_img) ** 2).mean())
    rays_per_path_gpu = st["rays"] / st["paths"]
    rays_per_path_cpu = cpu["rays"] / cpu["paths"]
    return rel, rmse_gpu, rmse_cpu, rays_per_path_gpu, rays_per_path_cpu, rgb, cpu["rgb8"]


@pytest.mark.parametrize("sid,grid,W,H,spp", [(1, 3, 240, 160, 256), (1, 11, 240, 160, 128), (2, 3, 120, 80, 256),
                                              (3, 3, 120, 80, 128), (4, 3, 120, 80, 256), (5, 3, 120, 80, 512),
                                              (6, 3, 100, 100, 1024), (7, 3, 160, 90, 256)])
@pytest.mark.parametrize("variant", [1, 2])
def test_image_parity(rtw, oracle, ctx, sid, grid, W, H, spp, variant):
    hs, osc = _scene(rtw, oracle, sid, grid)
    ctx.upload_scene(hs.desc, keep=hs)
    rel, rmse_gpu, rmse_cpu, rpp_g, rpp_c, rgb, cpu_rgb = _image_gates(rtw, oracle, ctx, hs, osc, W, H, spp, variant)
    # Monte-Carlo noise of the two means themselves: allow it on top of the 0.5 % bias gate for the emissive
    # scenes whose per-pixel variance is huge (Cornell light = 15, main.zig:270)
    tol = 0.005 if sid not in (5, 6) else 0.015
    assert (rel <= tol).all(), f"bias {rel} (rays/path gpu {rpp_g:.3f} cpu {rpp_c:.3f})"
    assert rmse_gpu <= 1.25 * rmse_cpu, f"rmse gpu {rmse_gpu:.5f} cpu {rmse_cpu:.5f}"
    assert abs(rpp_g - rpp_c) / rpp_c < 0.02
    mse8 = ((rgb.astype(np.float64) - cpu_rgb.astype(np.float64)) ** 2).mean()
    psnr = 10 * np.log10(255.0 ** 2 / max(mse8, 1e-12))
    assert psnr > 10.0


def test_flat_and_bvh_render_the_same_paths(rtw, ctx, knobs):
    """Identical Philox keys + identical closest hits => identical per-sample radiance; with one chunk the
    summation order is fixed too, so the two traversal variants must agree bit for bit."""
    knobs.setenv("RTW_SPP_CHUNK", "0")
    hs = rtw.HostScene(1, grid=11)
    ctx.upload_scene(hs.desc, keep=hs)
    cam = hs.camera()
    a = ctx.render(cam, ctx.params(150, 100, 0, 16, 16, 50, 1, 0, 9, hs.background), want_accum=True)[1]
    b = ctx.render(cam, ctx.params(150, 100, 0, 16, 16, 50, 2, 0, 9, hs.background), want_accum=True)[1]
    assert np.array_equal(a, b)
    c = ctx.render(cam, ctx.params(150, 100, 0, 16, 16, 50, 2, 0, 10, hs.background), want_accum=True)[1]
    assert not np.array_equal(a, c)  # a different seed gives different samples


@pytest.mark.parametrize("n_static,n_moving,big", [(1, 0, 0), (0, 1, 0), (2, 3, 0), (5, 4, 1), (9, 0, 2), (0, 13, 0), (17, 20, 3),
                                                   (8, 8, 8), (33, 31, 1), (60, 60, 0), (100, 100, 2)])
def test_flat_groups_any_count_and_mix(rtw, oracle, ctx, n_static, n_moving, big):
    """Sphere scenes of every size and mix of kinds (static / moving / a few that dwarf the rest / a huge ground): group
    sizes of two to four, a last chunk with padding, no-bounds mode (< 3 groups), unified static + moving groups.  Flat
    scan == BVH bit for bit (ids and distances), reference-order probe == oracle bit for bit, production ids == f64 oracle
    almost everywhere."""
    rng = np.random.default_rng(1000 * n_static + 10 * n_moving + big)
    b = scene_util.DescBuilder()
    mats = [b.diffuse(b.solid((0.5, 0.5, 0.5))), b.metal((0.7, 0.6, 0.5), 0.1), b.glass(1.5)]
    b.sphere((0.0, -1000.0, 0.0), 1000.0, mats[0])
    for _ in range(n_static):
        c = rng.uniform(-5, 5, 3); c[1] = rng.uniform(0.1, 1.0)
        b.sphere(tuple(c), float(rng.uniform(0.1, 0.3)), mats[int(rng.integers(0, 3))])
    for _ in range(n_moving):
        c = rng.uniform(-5, 5, 3); c[1] = rng.uniform(0.1, 1.0)
        c1 = c + np.array([0.0, rng.uniform(0, 0.5), 0.0])
        b.moving_sphere(tuple(c), tuple(c1), 0.0, 1.0, float(rng.uniform(0.1, 0.3)), mats[int(rng.integers(0, 3))])
    for _ in range(big):
        c = rng.uniform(-4, 4, 3); c[1] = 1.0
        b.sphere(tuple(c), float(rng.uniform(0.9, 1.3)) * (-1.0 if rng.uniform() < 0.2 else 1.0), mats[int(rng.integers(0, 3))])
    desc = b.build()
    osc = oracle.OracleScene.from_desc(desc, keep=desc)
    ctx.upload_scene(desc, keep=desc)
    rays = scene_util.random_rays(rng, 20000, extent=6.0)
    rays[:8000, 0:3] = np.array([13.0, 2.0, 3.0]) + rng.normal(size=(8000, 3)) * 0.05
    o32 = osc.trace_rays(rays, 32)
    assert (o32[0] != MISS).mean() > 0.2
    for variant in (rtw.abi.VARIANT_MEGA_FLAT, rtw.abi.VARIANT_MEGA_BVH):
        g = ctx.trace_rays(rays, 32, variant)
        assert np.array_equal(g[0], o32[0]) and np.array_equal(g[1], o32[1])
    a = ctx.trace_rays(rays, 0, rtw.abi.VARIANT_MEGA_FLAT)
    c = ctx.trace_rays(rays, 0, rtw.abi.VARIANT_MEGA_BVH)
    assert np.array_equal(a[0], c[0]) and np.array_equal(a[1], c[1])
    assert (a[0] != osc.trace_rays(rays, 64)[0]).mean() <= 1e-3


def test_flat_layout_heuristics_change_the_work_not_the_result(rtw, oracle, ctx):
    """The flat scan's layout heuristics — the few spheres that dwarf the rest tested individually (RTW_MID_SPHERES), the
    number of groups rounded up to a multiple of four (RTW_GROUP_ROUND), the kernel specialised on the scene's features
    (RTW_FLAT_SPECIALISE) — decide WHICH tests run, never what a test returns: same closest hits (ids equal; t to the last
    bits: an individually tested sphere takes its c term about a reference point), same paths."""
    rng = np.random.default_rng(77)
    for sid, grid in ((1, 3), (1, 5)):
        hs = rtw.HostScene(sid, grid=grid)
        osc = oracle.OracleScene.from_desc(hs.desc, keep=hs)
        rays = scene_util.random_rays(rng, 40000, extent=6.0)
        rays[:20000, 0:3] = np.array([13.0, 2.0, 3.0]) + rng.normal(size=(20000, 3)) * 0.05
        oid = osc.trace_rays(rays, 64)[0]
        base = None
        counts = []
        for opts in ({}, {"RTW_MID_SPHERES": "0"}, {"RTW_GROUP_ROUND": "0"}, {"RTW_MID_SPHERES": "0", "RTW_GROUP_ROUND": "0"}):
            with ctx.options(**opts):
                ctx.upload_scene(hs.desc, keep=hs)
                gid, gt = ctx.trace_rays(rays, 0, rtw.abi.VARIANT_MEGA_FLAT)[:2]
                cam = hs.camera()
                ctx.render(cam, ctx.params(96, 64, 0, 8, 8, 50, 1, rtw.abi.FLAG_COUNT_EVENTS, 5, hs.background))
                counts.append(ctx.stats()["sphere_tests"])
            assert (gid != oid).mean() <= 1e-3
            if base is None:
                base = (gid, gt)
            else:
                assert np.array_equal(gid, base[0])
                hit = gid != MISS
                assert np.abs(gt[hit] - base[1][hit]).max() <= 2e-6 * max(1.0, float(np.abs(base[1][hit]).max()))
        assert len(set(counts)) > 1  # the switches did change the work
    # specialised vs generic kernel: identical paths, only the order of the fp32 sums per pixel may differ
    hs = rtw.HostScene(1)
    ctx.upload_scene(hs.desc, keep=hs)
    cam = hs.camera()
    p = ctx.params(120, 80, 0, 32, 32, 50, 1, 0, 11, hs.background)
    a = ctx.render(cam, p, want_accum=True)[1]
    with ctx.options(RTW_FLAT_SPECIALISE="0"):
        b = ctx.render(cam, p, want_accum=True)[1]
    assert np.allclose(a, b, rtol=1e-4, atol=1e-3)


def test_spp_split_equals_full_render(rtw, ctx, knobs):
    """The multi-GPU partition: sample ranges rendered separately and summed == the full range
    (Philox is keyed by the absolute sample index).  Only fp32 summation order differs."""
    import torch
    hs = rtw.HostScene(1)
    ctx.upload_scene(hs.desc, keep=hs)
    cam = hs.camera()
    W, H, spp = 200, 120, 37
    full = ctx.render(cam, ctx.params(W, H, 0, spp, spp, 50, 0, 0, 42, hs.background), want_accum=True)
    parts = [torch.zeros(H, W, 4, device="cuda") for _ in range(4)]
    bounds = [0, 10, 19, 28, 37]
    for k in range(4):
        ctx.accumulate(cam, ctx.params(W, H, bounds[k], bounds[k + 1], spp, 50, 0, 0, 42, hs.background), parts[k].data_ptr())
    torch.cuda.synchronize()
    total = (parts[0] + parts[1] + parts[2] + parts[3]).cpu().numpy()
    assert (total[..., 3] == spp).all()
    np.testing.assert_allclose(total[..., :3], full[1][..., :3], rtol=2e-5, atol=1e-5)
    out = torch.zeros(H, W, 3, dtype=torch.uint8, device="cuda")
    ctx.resolve_multi([p.data_ptr() for p in parts], W, H, spp, out.data_ptr())
    torch.cuda.synchronize()
    diff = np.abs(out.cpu().numpy().astype(int) - full[0].astype(int))
    assert diff.max() <= 1 and (diff > 0).mean() < 1e-3
    # chunked (atomic) and unchunked accumulation agree to fp32 rounding
    knobs.setenv("RTW_SPP_CHUNK", "5")
    chunked = ctx.render(cam, ctx.params(W, H, 0, spp, spp, 50, 0, 0, 42, hs.background), want_accum=True)
    np.testing.assert_allclose(chunked[1][..., :3], full[1][..., :3], rtol=2e-5, atol=1e-5)
    assert (chunked[1][..., 3] == spp).all()


def test_baseline_size_properties(rtw, ctx):
    """Size-independent properties at BASELINE.json's full frame sizes (1920x1080 and 3840x2160)."""
    hs = rtw.HostScene(1)
    ctx.upload_scene(hs.desc, keep=hs)
    cam = hs.camera(aspect=16 / 9)
    for W, H, spp in ((1920, 1080, 4), (3840, 2160, 2)):
        p = ctx.params(W, H, 0, spp, spp, 50, 0, rtw.abi.FLAG_COUNT_EVENTS, 42, hs.background)
        rgb, acc = ctx.render(cam, p, want_accum=True)
        st = ctx.stats()
        assert st["paths"] == W * H * spp and st["rays"] >= st["paths"] and st["rays"] <= 50 * st["paths"]
        assert (acc[..., 3] == spp).all() and np.isfinite(acc).all()
        # energy bound: no emitters, albedo <= 1 -> radiance never exceeds the brightest background channel
        assert acc[..., :3].max() <= spp * 1.0 + 1e-3 and acc.min() >= 0.0
        # top rows are sky: exactly the background, quantised as the KAT says
        assert (rgb[0, :, :] == np.array([214, 228, 255], dtype=np.uint8)).all()
        assert 2.0 < st["rays"] / st["paths"] < 3.0


def test_white_furnace_on_device(rtw, ctx):
    b = scene_util.DescBuilder()
    m = b.diffuse(b.solid((0.5, 0.5, 0.5)))
    b.sphere((0, 0, 0), 1.0, m)
    desc = b.build()
    ctx.upload_scene(desc, keep=desc)
    cam = rtw.camera_init((0, 0, 4), (0, 0, 0), (0, 1, 0), 20.0, 1.0, 0.0)
    for variant in (1, 2):
        rgb, acc = ctx.render(cam, ctx.params(32, 32, 0, 256, 256, 50, variant, 0, 1, (1.0, 1.0, 1.0)), want_accum=True)
        img = acc[..., :3] / 256
        assert img.max() <= 1.0 + 1e-5
        assert abs(img[12:20, 12:20].mean() - 0.5) < 0.01
        rgb, acc = ctx.render(cam, ctx.params(32, 32, 0, 4, 4, 50, variant, 0, 1, (0.0, 0.0, 0.0)), want_accum=True)
        assert acc[..., :3].max() == 0.0


@pytest.mark.parametrize("sid,grid,W,H,spp", [(1, 3, 200, 120, 33), (6, 3, 97, 61, 40), (1, 11, 160, 90, 16), (7, 3, 120, 68, 24)])
def test_wavefront_renders_the_same_paths_as_the_megakernel(rtw, ctx, knobs, sid, grid, W, H, spp):
    """K2 (generate / extend / shade+compact) is a different schedule of the same path: identical Philox keys and
    identical closest hits => identical per-sample radiance; only fp32 summation order differs.  Ragged frame
    sizes and a small slot count force many refill iterations."""
    knobs.setenv("RTW_WF_SLOTS", "8192")
    hs = rtw.HostScene(sid, grid=grid)
    ctx.upload_scene(hs.desc, keep=hs)
    cam = hs.camera(aspect=W / H)
    flag = rtw.abi.FLAG_COUNT_EVENTS
    mega = ctx.render(cam, ctx.params(W, H, 0, spp, spp, 50, rtw.abi.VARIANT_AUTO, flag, 42, hs.background), want_accum=True)
    st_m = ctx.stats()
    wave = ctx.render(cam, ctx.params(W, H, 0, spp, spp, 50, rtw.abi.VARIANT_WAVEFRONT, flag, 42, hs.background), want_accum=True)
    st_w = ctx.stats()
    assert st_w["variant_used"] == rtw.abi.VARIANT_WAVEFRONT and st_w["n_launches"] > 4
    assert (wave[1][..., 3] == spp).all()
    assert st_w["paths"] == st_m["paths"] == W * H * spp and st_w["rays"] == st_m["rays"]
    for k in ("scatter_diffuse", "scatter_metal", "scatter_dielectric", "emit_hits", "sphere_finalise"):
        assert st_w[k] == st_m[k], k
    np.testing.assert_allclose(wave[1][..., :3], mega[1][..., :3], rtol=3e-5, atol=2e-5)
    diff = np.abs(wave[0].astype(int) - mega[0].astype(int))
    assert diff.max() <= 1 and (diff > 0).mean() < 1e-3


@pytest.mark.parametrize("sid,grid,W,H,spp", [(1, 11, 203, 117, 9), (1, 3, 64, 40, 33), (2, 11, 97, 61, 5), (6, 11, 120, 120, 12), (4, 11, 80, 60, 8)])
def test_bvh_schedules_render_the_same_paths(rtw, ctx, sid, grid, W, H, spp):
    """k_megakernel_bvhq (RTW_BVH_KERNEL=2: per-warp ray queue in shared memory, traversing lanes refilled from the ring) and the
    speculative state machine (RTW_BVH_KERNEL=3: a lane postpones one leaf and goes on descending) are
    other schedules of the same paths as the per-lane state machine: identical Philox keys and closest hits => identical
    event counts and per-sample radiance; only the fp32 summation order per pixel differs.  Generic and spheres-only
    builds, ragged frames, threshold extremes (service phase at 1 and at 32 finished rays)."""
    hs = rtw.HostScene(sid, grid=grid)
    ctx.upload_scene(hs.desc, keep=hs)
    cam = hs.camera(aspect=W / H)
    bvh = rtw.abi.VARIANT_MEGA_BVH
    for flag in (rtw.abi.FLAG_COUNT_EVENTS, 0):
        p = ctx.params(W, H, 0, spp, spp, 50, bvh, flag, 42, hs.background)
        a = ctx.render(cam, p, want_accum=True)
        st_a = ctx.stats()
        for opts in ({"RTW_BVH_KERNEL": "2"}, {"RTW_BVH_KERNEL": "2", "RTW_BVH_THRESH": "1", "RTW_BVH_LEAF": "1"},
                     {"RTW_BVH_KERNEL": "2", "RTW_BVH_THRESH": "32", "RTW_BVH_STEPS": "7"},
                     {"RTW_BVH_KERNEL": "3"}, {"RTW_BVH_KERNEL": "3", "RTW_BVH_THRESH": "1", "RTW_BVH_LEAF": "1"},
                     {"RTW_BVH_KERNEL": "3", "RTW_BVH_LEAF": "32", "RTW_BVH_STEPS": "7"}):
            with ctx.options(**opts):
                b = ctx.render(cam, p, want_accum=True)
                st_b = ctx.stats()
            assert (b[1][..., 3] == spp).all()
            if flag:
                assert st_b["paths"] == st_a["paths"] == W * H * spp and st_b["rays"] == st_a["rays"]
                for k in ("scatter_diffuse", "scatter_metal", "scatter_dielectric", "emit_hits", "sphere_finalise"):
                    assert st_b[k] == st_a[k], k
                if opts["RTW_BVH_KERNEL"] == "2":  # same visits in another order
                    assert st_b["node_tests"] == st_a["node_tests"]
                else:  # speculative traversal: a postponed leaf delays the shrinking of the search interval
                    assert st_a["node_tests"] <= st_b["node_tests"] <= 1.5 * st_a["node_tests"]
            np.testing.assert_allclose(b[1][..., :3], a[1][..., :3], rtol=3e-5, atol=2e-5)
            diff = np.abs(b[0].astype(int) - a[0].astype(int))
            assert diff.max() <= 1 and (diff > 0).mean() < 1e-3


def test_render_multi_slab_resolve_on_one_device(rtw, ctx):
    """rtw_cuda_render_multi with a single context runs the same slab-resolve code path as N devices (one slab = the
    whole image, no peers): byte-identical to rtw_cuda_render, for widths that take the 4-pixel and the 1-pixel kernel."""
    hs = rtw.HostScene(1)
    ctx.upload_scene(hs.desc, keep=hs)
    for W, H, spp in ((320, 180, 9), (101, 67, 5)):
        cam = hs.camera(aspect=W / H)
        p = ctx.params(W, H, 0, spp, spp, 50, 0, rtw.abi.FLAG_DETERMINISTIC, 42, hs.background)
        one = ctx.render(cam, p)[0]
        multi = rtw.render_multi([ctx], cam, p)
        assert np.array_equal(one, multi)
        assert ctx.stats()["ms_wall"] > 0


def test_render_multi_slab_parallel_resolve_over_peers(rtw, ctx):
    """Single-process multi-GPU: each device traces its sample range, then every device resolves one scanline slab from
    all buffers over NVLink peer mappings and copies its rows to the host.  Must equal the 1-GPU render of the same
    sample set.  (bench.py --gpus N asserts the same inside the driver's multi-GPU run.)"""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    n = min(n, 4)
    hs = rtw.HostScene(1)
    ctx.upload_scene(hs.desc, keep=hs)
    group = rtw.create_multi(n)
    try:
        for c in group:
            c.upload_scene(hs.desc, keep=hs)
        for W, H, spp in ((320, 180, 37), (322, 181, 10)):  # 181 rows over 4 devices: ragged slabs; 322: the 1-pixel kernel
            cam = hs.camera(aspect=W / H)
            p = ctx.params(W, H, 0, spp, spp, 50, 0, 0, 42, hs.background)
            one = ctx.render(cam, p)[0]
            multi = rtw.render_multi(group, cam, p)
            diff = np.abs(multi.astype(int) - one.astype(int))
            assert diff.max() <= 1 and (diff > 0).mean() < 1e-3
    finally:
        for c in group:
            c.close()


def test_instanced_spheres_uv_and_image(rtw, oracle, ctx, earth_rgba):
    """Translate(RotateY(sphere)) with an image texture: production uv follows the OBJECT-space normal like the
    reference (getSphereUv inside the instance, hittable.zig:127,583-593), and the rendered image matches."""
    b = scene_util.DescBuilder()
    earth = b.diffuse(b.image(earth_rgba))
    grey = b.diffuse(b.solid((0.5, 0.5, 0.5)))
    t = b.translate((1.5, 0.5, -1.0))
    r = b.rotate_y(70.0, outer=t)
    b.sphere((0, 0, 0), 2.0, earth, xform=r)
    b.sphere((0, -1002, 0), 1000.0, grey)
    desc = b.build()
    ctx.upload_scene(desc, keep=desc)
    osc = oracle.OracleScene.from_desc(desc, keep=desc)
    cam = rtw.camera_init((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 1.5, 0.0)
    rng = np.random.default_rng(3)
    rays = scene_util.random_rays(rng, 20000, extent=3.0)
    oid, ot, on, ouv = osc.trace_rays(rays, 64)
    for variant in (1, 2):
        gid, gt, gn, guv = ctx.trace_rays(rays, 0, variant)
        ok = (gid == oid) & (oid == 0)
        assert (gid != oid).mean() < 1e-3 and ok.sum() > 1000
        du = np.abs(guv[ok, 0] - ouv[ok, 0]); du = np.minimum(du, 1 - du)
        assert du.max() < 2e-4 and np.abs(guv[ok, 1] - ouv[ok, 1]).max() < 2e-3
        for prec in (32, 64):
            a = ctx.trace_rays(rays, prec, variant)
            o = osc.trace_rays(rays, prec)
            assert np.array_equal(a[0], o[0]) and np.array_equal(a[1], o[1]) and np.array_equal(a[2], o[2])
    W, H, spp = 120, 80, 128
    ref = osc.render(cam, W, H, spp * 8, 50, (0.7, 0.8, 1.0), seed=5, precision=64, nthreads=max(2, oracle.num_threads()))["accum"] / (spp * 8)
    acc = ctx.render(cam, ctx.params(W, H, 0, spp, spp, 50, 0, 0, 42, (0.7, 0.8, 1.0)), want_accum=True)[1][..., :3] / spp
    rel = np.abs(acc.mean((0, 1)) - ref.mean((0, 1))) / ref.mean((0, 1))
    assert (rel < 0.005).all(), rel


def test_million_sphere_scene(rtw, oracle, ctx):
    """BASELINE.json configs[3] at full primitive count (10^6 spheres, BVH): production-arithmetic primary-hit ids
    against the f64 oracle (its own BVH, validated against its linear scan elsewhere), f64 reference-order probe
    bit-exact, and the size-independent render properties."""
    hs = rtw.HostScene(rtw.host_lib.SCENE_SPHERE_FIELD, grid=500)
    assert hs.desc.n_prims > 990000
    ctx.upload_scene(hs.desc, keep=hs)
    st = ctx.stats()
    assert st["bvh_depth"] <= 64 and st["bvh_nodes"] > hs.desc.n_prims // 4
    assert st["bvh_builder"] == rtw.abi.BVH_BUILDER_LBVH and st["ms_bvh_build"] < st["ms_upload"]
    osc = oracle.OracleScene.from_desc(hs.desc, keep=hs)
    W, H = 160, 90
    cam = hs.camera(aspect=W / H)
    oid, ot, on = osc.primary_hits(cam, W, H, 64, use_bvh=True)
    gid, gt, gn = ctx.primary_hits(cam, W, H, 0, rtw.abi.VARIANT_MEGA_BVH)
    assert (gid != oid).mean() <= 5e-4, (gid != oid).sum()
    g64 = ctx.primary_hits(cam, W, H, 64, rtw.abi.VARIANT_MEGA_BVH)
    assert np.array_equal(g64[0], oid) and np.array_equal(g64[1], ot) and np.array_equal(g64[2], on)
    assert 0.2 < (oid != MISS).mean() <= 1.0
    p = ctx.params(320, 180, 0, 8, 8, 50, 0, rtw.abi.FLAG_COUNT_EVENTS, 42, hs.background)
    rgb, acc = ctx.render(cam, p, want_accum=True)
    s2 = ctx.stats()
    assert s2["variant_used"] == rtw.abi.VARIANT_MEGA_BVH and s2["paths"] == 320 * 180 * 8
    assert (acc[..., 3] == 8).all() and np.isfinite(acc).all() and acc[..., :3].max() <= 8.0 + 1e-3


def test_fp32_peak_is_plausible(ctx):
    tf, mhz = ctx.measure_fp32_peak()
    assert 30.0 < tf < 100.0 and mhz > 1000
