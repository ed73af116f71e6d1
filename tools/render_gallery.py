"""Render small PNGs of the reference's scenes on the GPU (docs/gallery/), through the public C-ABI call."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rtw_b200
out = os.path.join(ROOT, "docs", "gallery")
os.makedirs(out, exist_ok=True)
ctx = rtw_b200.Context(0)
for sid, name, w, spp in ((1, "scene1_random_spheres", 480, 400), (2, "scene2_two_spheres", 360, 200), (3, "scene3_perlin", 360, 200),
                          (4, "scene4_earth", 360, 200), (5, "scene5_simple_light", 360, 1500), (6, "scene6_cornell", 400, 2000),
                          (7, "c3_earth_glass_metal", 480, 400)):
    hs = rtw_b200.HostScene(sid)
    h = int(w / hs.aspect)
    ctx.upload_scene(hs.desc, keep=hs)
    rgb, _ = ctx.render(hs.camera(), ctx.params(w, h, 0, spp, spp, 50, 0, 0, 42, hs.background))
    rtw_b200.host_lib.write_png(os.path.join(out, name + ".png"), rgb)
    print(name, w, h, spp, f"{ctx.stats()['ms_trace']:.1f} ms")
hs = rtw_b200.HostScene(1, grid=11)
ctx.upload_scene(hs.desc, keep=hs)
rgb, _ = ctx.render(hs.camera(), ctx.params(480, 320, 0, 400, 400, 50, 0, 0, 42, hs.background))
rtw_b200.host_lib.write_png(os.path.join(out, "scene1_book_grid11.png"), rgb)
