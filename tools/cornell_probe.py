import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rtw_b200
ctx = rtw_b200.Context(0)
hs = rtw_b200.HostScene(6)
ctx.upload_scene(hs.desc, keep=hs)
rgb, _ = ctx.render(hs.camera(), ctx.params(600, 600, 0, 200, 200, 50, 0, 0, 42, hs.background))
print(ctx.stats()["ms_trace"])
