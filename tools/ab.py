"""Development helper: A/B timing of kernel variants / tuning options on one GPU.

    python tools/ab.py '<label>:<scene>:<grid>:<W>x<H>x<spp>:<variant>[:OPT=VAL,...]' ...

Each case: upload, 2 warm-up launches, 3 timed launches (CUDA events), then an instrumented replay for events per ray.
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import rtw_b200
from rtw_b200 import abi
if os.environ.get("RTW_AB_LIB"):  # A/B of two builds inside one GPU call: point the loader at a saved copy of librtw_cuda.so
    rtw_b200.build.CUDA_LIB = os.path.abspath(os.environ["RTW_AB_LIB"])

ctx = rtw_b200.Context(0)
scenes = {}
for spec in sys.argv[1:]:
    f = spec.split(":")
    label, sid, grid = f[0], int(f[1]), int(f[2])
    W, H, spp = (int(x) for x in f[3].split("x"))
    variant = int(f[4])
    opts = dict(kv.split("=") for kv in f[5].split(",")) if len(f) > 5 and f[5] else {}
    key = (sid, grid)
    if key not in scenes:
        scenes[key] = rtw_b200.HostScene(sid, grid=grid)
    hs = scenes[key]
    with ctx.options(**opts):
        ctx.upload_scene(hs.desc, keep=hs)
        cam = hs.camera(aspect=W / H)
        accum = torch.zeros(H, W, 4, device="cuda")
        p = ctx.params(W, H, 0, spp, spp, 50, variant, 0, 42, hs.background)
        for _ in range(2):
            ctx.accumulate(cam, p, accum.data_ptr(), None)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            ctx.accumulate(cam, p, accum.data_ptr(), None)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        sspp = min(spp, 16)
        ctx.render(cam, ctx.params(W, H, 0, sspp, sspp, 50, variant, abi.FLAG_COUNT_EVENTS, 42, hs.background))
        st = ctx.stats()
    rays = max(1, st["rays"])
    print(json.dumps({"label": label, "scene": sid, "grid": grid, "prims": hs.desc.n_prims, "size": f"{W}x{H}x{spp}", "variant_used": st["variant_used"],
                      "opts": opts, "ms": round(ms, 3), "mpaths_s": round(W * H * spp / ms / 1e3, 1), "rays_per_path": round(rays / st["paths"], 3),
                      "node_tests": round(st["node_tests"] / rays, 2), "sphere_tests": round(st["sphere_tests"] / rays, 2),
                      "rect_tests": round(st["rect_tests"] / rays, 2)}), flush=True)
    del accum
