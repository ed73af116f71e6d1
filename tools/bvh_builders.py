"""Development helper: host (binned SAH) vs device (Morton / radix tree) BVH builder on one scene —
upload time, BVH share of it, tree size, render time and node tests per ray.
usage: bvh_builders.py [scene grid spp]   (default: scene 8 = config C4, grid 500, 64 spp)"""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import rtw_b200
from rtw_b200 import abi

scene = int(sys.argv[1]) if len(sys.argv) > 1 else 8
grid = int(sys.argv[2]) if len(sys.argv) > 2 else 500
spp = int(sys.argv[3]) if len(sys.argv) > 3 else 64
W, H = 1920, 1080
ctx = rtw_b200.Context(0)
hs = rtw_b200.HostScene(scene, grid=grid)
cam = hs.camera(aspect=W / H)
accum = torch.zeros(H, W, 4, device="cuda")
ctx.set_option("RTW_UPLOAD_TRACE", "1")
for builder, leaf_max in (("sah", 4), ("lbvh", 4), ("lbvh", 2), ("lbvh", 1)):
    ctx.set_option("RTW_BVH_BUILDER", builder)
    ctx.set_option("RTW_BVH_LEAF_MAX", leaf_max)
    ctx.upload_scene(hs.desc, keep=hs)
    up = ctx.stats()
    p = ctx.params(W, H, 0, spp, spp, 50, abi.VARIANT_MEGA_BVH, 0, 42, hs.background)
    ctx.accumulate(cam, p, accum.data_ptr(), None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ctx.accumulate(cam, p, accum.data_ptr(), None)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    ps = ctx.params(W, H, 0, 8, 8, 50, abi.VARIANT_MEGA_BVH, abi.FLAG_COUNT_EVENTS, 42, hs.background)
    ctx.render(cam, ps)
    st = ctx.stats()
    print(json.dumps({"builder": builder, "leaf_max": leaf_max, "prims": hs.desc.n_prims, "upload_ms": round(up["ms_upload"], 2),
                      "bvh_build_ms": round(up["ms_bvh_build"], 2), "bvh_nodes": up["bvh_nodes"], "bvh_depth": up["bvh_depth"],
                      "builder_used": up["bvh_builder"], "spp": spp, "render_ms": round(ms, 2),
                      "node_tests_per_ray": round(st["node_tests"] / st["rays"], 1),
                      "sphere_tests_per_ray": round(st["sphere_tests"] / st["rays"], 1)}), flush=True)
