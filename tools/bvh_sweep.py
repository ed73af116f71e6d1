"""Development helper: sweep the BVH state machine's knobs (service threshold, interior steps per round, leaf threshold)
on the two BVH configs and print ms per frame sample.   python tools/bvh_sweep.py [quick]"""
import itertools, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import rtw_b200
from rtw_b200 import abi

ctx = rtw_b200.Context(0)
W, H = 1920, 1080
cases = [("c4", 8, 500, 24), ("c2p", 1, 11, 48)]
grid = list(itertools.product((16, 20, 24, 28), (1, 2, 3), (4, 8, 12)))
if len(sys.argv) > 1 and sys.argv[1] == "quick":
    grid = [(20, 2, 8), (24, 2, 8), (16, 2, 8), (20, 3, 8), (20, 2, 12)]
for name, sid, g, spp in cases:
    hs = rtw_b200.HostScene(sid, grid=g)
    ctx.upload_scene(hs.desc, keep=hs)
    cam = hs.camera(aspect=W / H)
    accum = torch.zeros(H, W, 4, device="cuda")
    p = ctx.params(W, H, 0, spp, spp, 50, abi.VARIANT_MEGA_BVH, 0, 42, hs.background)
    for th, st, lf in grid:
        ctx.set_option("RTW_BVH_THRESH", th); ctx.set_option("RTW_BVH_STEPS", st); ctx.set_option("RTW_BVH_LEAF", lf)
        ctx.accumulate(cam, p, accum.data_ptr(), None)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2):
            ctx.accumulate(cam, p, accum.data_ptr(), None)
        e1.record(); torch.cuda.synchronize()
        print(json.dumps({"case": name, "thresh": th, "steps": st, "leaf": lf, "ms": round(e0.elapsed_time(e1) / 2, 2)}), flush=True)
    for k in ("RTW_BVH_THRESH", "RTW_BVH_STEPS", "RTW_BVH_LEAF"):
        ctx.set_option(k, None)
    del accum
