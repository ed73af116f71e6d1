import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rtw_b200
W, H, spp = 1920, 1080, int(sys.argv[1]) if len(sys.argv) > 1 else 16
var = int(sys.argv[2]) if len(sys.argv) > 2 else 3
grid = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ctx = rtw_b200.Context(0)
hs = rtw_b200.HostScene(1, grid=grid)
ctx.upload_scene(hs.desc, keep=hs)
rgb, _ = ctx.render(hs.camera(aspect=W / H), ctx.params(W, H, 0, spp, spp, 50, var, 0, 42, hs.background))
print(ctx.stats()["ms_trace"], ctx.stats()["n_launches"])
