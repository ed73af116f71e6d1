"""Development helper: time the headline workload for a few variants and print event counts per ray."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import rtw_b200
from rtw_b200 import abi
W, H = 1920, 1080
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 200
scene = int(sys.argv[2]) if len(sys.argv) > 2 else 1
grid = int(sys.argv[3]) if len(sys.argv) > 3 else 3
variants = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [1, 2]
ctx = rtw_b200.Context(0)
hs = rtw_b200.HostScene(scene, grid=grid)
cam = hs.camera(aspect=W / H)
ctx.upload_scene(hs.desc, keep=hs)
print("prims", hs.desc.n_prims, "upload ms", ctx.stats()["ms_upload"])
accum = torch.zeros(H, W, 4, device="cuda")
for var in variants:
    p = ctx.params(W, H, 0, spp, spp, 50, var, 0, 42, hs.background)
    for _ in range(1 if var == 3 else 2):
        ctx.accumulate(cam, p, accum.data_ptr(), None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ctx.accumulate(cam, p, accum.data_ptr(), None)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    ps = ctx.params(W, H, 0, min(spp, 32), min(spp, 32), 50, var, abi.FLAG_COUNT_EVENTS, 42, hs.background)
    ctx.render(cam, ps)
    st = ctx.stats()
    rays = st["rays"]
    print(f"variant {var}: {ms:.2f} ms  {W*H*spp/ms/1e3:.0f} Mpaths/s  {rays/st['paths']*W*H*spp/ms/1e3:.0f} Mrays/s | per ray: "
          f"sphere_tests {st['sphere_tests']/rays:.1f} roots {st['sphere_roots']/rays:.2f} node_tests {st['node_tests']/rays:.1f} rect_tests {st['rect_tests']/rays:.1f}")
