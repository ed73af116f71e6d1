import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import rtw_b200, oracle_binding as ob, scene_util
from rtw_b200 import abi
ctx = rtw_b200.Context(0)
for seed in (1, 2, 3):
    rng = np.random.default_rng(seed)
    desc = scene_util.random_scene(rng)
    ctx.upload_scene(desc, keep=desc)
    rays = scene_util.random_rays(rng, 50000)
    for prec in (32, 64, 0):
        for var in (1, 2):
            ctx.trace_rays(rays, prec, var)
hs = rtw_b200.HostScene(8, grid=50)
osc = ob.OracleScene.from_desc(hs.desc, keep=hs)
ctx.upload_scene(hs.desc, keep=hs)
print(ctx.stats()["bvh_nodes"], ctx.stats()["bvh_depth"])
rng = np.random.default_rng(11)
n = 20000
rays = np.zeros((n, 7))
rays[:, 0:3] = rng.uniform(-40, 40, (n, 3)) * [1, 0.05, 1] + [0, 5, 0]
rays[:, 3:6] = rng.normal(size=(n, 3)) * [1, 0.3, 1]
rays[:, 6] = rng.uniform(0, 1, n)
olin = osc.trace_rays(rays, 32, use_bvh=False)
for rep in range(3):
    flat = ctx.trace_rays(rays, 32, abi.VARIANT_MEGA_FLAT)
    bvh = ctx.trace_rays(rays, 32, abi.VARIANT_MEGA_BVH)
    b64 = ctx.trace_rays(rays, 64, abi.VARIANT_MEGA_BVH)
    bp = ctx.trace_rays(rays, 0, abi.VARIANT_MEGA_BVH)
    d = np.nonzero(flat[0] != bvh[0])[0]
    print("rep", rep, "flat!=bvh", len(d), "flat!=olin", (flat[0] != olin[0]).sum(), "bvh!=olin", (bvh[0] != olin[0]).sum(), "prod bvh != olin", (bp[0] != olin[0]).sum())
    for k in d[:6]:
        print("  ray", k, "flat", flat[0][k], flat[1][k], "bvh", bvh[0][k], bvh[1][k], "olin", olin[0][k], olin[1][k])
    if len(d): print("  first/last differing index", d[0], d[-1])
