"""Host-side logic (no GPU): the C++ twin of the Zig host builds the same scenes as the oracle's independent
restatement of src/main.zig:124-293, flattens them to the ABI with ids in depth-first order, decodes the PNG."""
import ctypes as C

import numpy as np
import pytest


def _prim_tuple(p):
    return (p.kind, p.xform, tuple(p.v))


def _mat_expanded(d, mi):
    """A material with its texture tree resolved (table indices are a flatten convention, not semantics)."""
    m = d.materials[mi]

    def tex(ti):
        if ti < 0:
            return None
        t = d.textures[ti]
        if t.kind == 0:
            return ("solid", tuple(t.color))
        if t.kind == 1:
            return ("checker", tex(t.a), tex(t.b))
        if t.kind == 2:
            pl = d.perlins[t.a]
            return ("noise", t.scale, tuple(pl.ranvec[i] for i in range(768)), tuple(pl.perm_x[i] for i in range(256)),
                    tuple(pl.perm_y[i] for i in range(256)), tuple(pl.perm_z[i] for i in range(256)))
        im = d.images[t.a]
        return ("image", im.width, im.height, bytes(C.string_at(im.rgba8, im.width * im.height * 4)))
    return (m.kind, tex(m.texture) if m.kind in (0, 3) else None, tuple(m.albedo) if m.kind == 1 else None,
            m.param if m.kind in (1, 2) else None)


@pytest.mark.parametrize("sid,grid", [(1, 3), (1, 11), (2, 3), (3, 3), (4, 3), (5, 3), (6, 3)])
def test_host_scene_equals_oracle_scene(rtw, oracle, earth_rgba, sid, grid):
    hs = rtw.HostScene(sid, grid=grid)
    osc = oracle.OracleScene.builtin(sid, grid=grid, image=earth_rgba if sid == 4 else None)
    a, b = hs.desc, osc.export()
    assert a.n_prims == b.n_prims
    for i in range(a.n_prims):
        assert _prim_tuple(a.prims[i]) == _prim_tuple(b.prims[i]), i
        assert _mat_expanded(a, a.prims[i].material) == _mat_expanded(b, b.prims[i].material), i
    assert a.n_xforms == b.n_xforms
    for i in range(a.n_xforms):
        assert (a.xforms[i].kind, a.xforms[i].outer, tuple(a.xforms[i].v)[:3]) == \
               (b.xforms[i].kind, b.xforms[i].outer, tuple(b.xforms[i].v)[:3])
    cfg = osc.config()
    assert (hs.width, hs.height, hs.spp, hs.max_depth) == (cfg["width"], cfg["height"], cfg["spp"], cfg["max_depth"])
    assert hs.background == tuple(cfg["background"]) and hs.vfov == cfg["vfov"] and hs.aperture == cfg["aperture"]
    ca, cb = hs.camera(), osc.default_camera()
    for f, _ in rtw.abi.Camera._fields_:
        va, vb = getattr(ca, f), getattr(cb, f)
        assert (list(va) == list(vb)) if hasattr(va, "__len__") else (va == vb), f


def test_material_identity_is_deduplicated(rtw):
    """Cornell: 18 leaves share 4 Rc(Material) cells (src/main.zig:262-270, Box.init clones hittable.zig:437-442)."""
    hs = rtw.HostScene(6)
    d = hs.desc
    assert d.n_materials == 4 and d.n_prims == 18
    white = d.prims[3].material
    assert {d.prims[i].material for i in range(6, 18)} == {white}
    assert d.prims[4].material == white and d.prims[5].material == white
    # scene 1: one material per object (main.zig:192)
    s1 = rtw.HostScene(1).desc
    assert s1.n_materials == s1.n_prims


def test_host_random_is_zig_default_prng(rtw):
    out = np.zeros(4)
    rtw.host_lib.load().rtw_host_random_real01(42, 4, out.ctypes.data_as(C.POINTER(C.c_double)))
    assert out.tolist() == [0.6969372117194047, 0.47274502314109507, 0.5152564274971367, 0.9257049799795629]


def test_png_decoder_matches_pillow(rtw):
    from PIL import Image
    ref = np.array(Image.open(rtw.host_lib.ASSET_EARTH).convert("RGBA"))
    got = rtw.host_lib.decode_png(rtw.host_lib.ASSET_EARTH)
    assert got.shape == (282, 500, 4) and np.array_equal(got, ref)


def test_synthetic_config_scenes(rtw):
    c3 = rtw.HostScene(rtw.host_lib.SCENE_EARTH_GLASS_METAL)
    assert c3.desc.n_prims == 4 and c3.desc.n_images == 1 and (c3.width, c3.height, c3.spp) == (1920, 1080, 1000)
    c4 = rtw.HostScene(rtw.host_lib.SCENE_SPHERE_FIELD, grid=20)
    n = c4.desc.n_prims
    assert 0.99 * (40 * 40 + 4) < n <= 40 * 40 + 4
    # small spheres rest on the curved ground: |c - ground centre| = 1000.2
    for i in range(4, n, 97):
        p = c4.desc.prims[i]
        assert abs(np.sqrt(p.v[0] ** 2 + (p.v[1] + 1000) ** 2 + p.v[2] ** 2) - 1000.2) < 1e-9


def test_ppm_writer(rtw, tmp_path):
    img = (np.arange(5 * 7 * 3) % 256).astype(np.uint8).reshape(5, 7, 3)
    path = str(tmp_path / "x.ppm")
    rtw.host_lib.write_ppm(path, img)
    raw = open(path, "rb").read()
    assert raw.startswith(b"P6\n7 5\n255\n") and raw[len(b"P6\n7 5\n255\n"):] == img.tobytes()


def test_png_writer_round_trips(rtw, tmp_path):
    """out.png as the reference writes it (src/main.zig:405): Pillow and our own decoder read back the same pixels."""
    from PIL import Image
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    path = str(tmp_path / "out.png")
    rtw.host_lib.write_png(path, img)
    assert np.array_equal(np.array(Image.open(path).convert("RGB")), img)
    back = rtw.host_lib.decode_png(path)
    assert np.array_equal(back[..., :3], img) and (back[..., 3] == 255).all()
