import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import rtw_b200, oracle_binding as ob
from rtw_b200 import abi
ctx = rtw_b200.Context(0)
hs = rtw_b200.HostScene(8, grid=50)
osc = ob.OracleScene.from_desc(hs.desc, keep=hs)
ctx.upload_scene(hs.desc, keep=hs)
rng = np.random.default_rng(11)
n = 20000
rays = np.zeros((n, 7))
rays[:, 0:3] = rng.uniform(-40, 40, (n, 3)) * [1, 0.05, 1] + [0, 5, 0]
rays[:, 3:6] = rng.normal(size=(n, 3)) * [1, 0.3, 1]
rays[:, 6] = rng.uniform(0, 1, n)
for rep in range(3):
    oid, ot, _, _ = osc.trace_rays(rays, 32, use_bvh=True)
    olin = osc.trace_rays(rays, 32, use_bvh=False)
    flat = ctx.trace_rays(rays, 32, abi.VARIANT_MEGA_FLAT)
    bvh = ctx.trace_rays(rays, 32, abi.VARIANT_MEGA_BVH)
    d = np.nonzero(flat[0] != bvh[0])[0]
    print("rep", rep, "flat!=bvh", len(d), "bvh!=oracle_bvh", (bvh[0] != oid).sum(), "flat!=oracle_lin", (flat[0] != olin[0]).sum(), "oracle bvh!=lin", (oid != olin[0]).sum())
    for k in d[:8]:
        print("  ray", k, "flat", flat[0][k], flat[1][k], "bvh", bvh[0][k], bvh[1][k], "olin", olin[0][k], olin[1][k], "obvh", oid[k], ot[k])
