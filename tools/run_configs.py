"""Measure every BASELINE.json config on one GPU and print one JSON line per row.

    python tools/run_configs.py [--quick] [--no-cpu]

Per config: Mpaths/s, Mrays/s, rays/path, counted flops/ray, FP32 roofline fraction (against the FFMA peak measured
in this run), the oracle port's Mrays/s on all host threads (bounded sample), the bias / noise gates at reduced size,
and primary-hit id mismatches (fp32 reference-order probe vs oracle<float>; production arithmetic vs oracle<double>).
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import rtw_b200
from rtw_b200 import abi
import oracle_binding as ob
import bench as B

ap = argparse.ArgumentParser()
ap.add_argument("--quick", action="store_true")
ap.add_argument("--no-cpu", action="store_true")
ap.add_argument("--only", default="")
args = ap.parse_args()

ctx = rtw_b200.Context(0)
peak_tf, _ = ctx.measure_fp32_peak()
nt = ob.num_threads()
CONFIGS = [  # name, scene, grid, W, H, spp, variant, gate size (W,H,spp)
    ("C1a cornell 600x600x200", 6, 3, 600, 600, 200, 0, (100, 100, 512)),
    ("C1b scene1 600x400x50", 1, 3, 600, 400, 50, 0, (240, 160, 128)),
    ("C2 scene1 1920x1080x500 flat", 1, 3, 1920, 1080, 500, 1, (240, 135, 128)),
    ("C2 scene1 1920x1080x500 bvh", 1, 3, 1920, 1080, 500, 2, (240, 135, 128)),
    ("C2' scene1 grid11 1920x1080x500 bvh", 1, 11, 1920, 1080, 500, 0, (240, 135, 128)),
    ("C3 earth+glass+metal 1920x1080x1000", 7, 3, 1920, 1080, 1000, 0, (240, 135, 128)),
    ("C4 1M spheres 1920x1080x256", 8, 500, 1920, 1080, 256, 0, None),
    ("C5 scene1 3840x2160x1000", 1, 3, 3840, 2160, 1000, 0, None),
]
for name, sid, grid, W, H, spp, variant, gate in CONFIGS:
    if args.only and args.only not in name:
        continue
    if args.quick:
        spp = max(8, spp // 10)
    t0 = time.time()
    hs = rtw_b200.HostScene(sid, grid=grid)
    t_host = time.time() - t0
    cam = hs.camera(aspect=W / H)
    ctx.upload_scene(hs.desc, keep=hs)
    up = ctx.stats()
    accum = torch.zeros(H, W, 4, device="cuda")
    p = ctx.params(W, H, 0, spp, spp, 50, variant, 0, 42, hs.background)
    ctx.accumulate(cam, ctx.params(W, H, 0, max(1, spp // 8), spp, 50, variant, 0, 42, hs.background), accum.data_ptr(), None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3 if W * H * spp < 3e9 else 1
    e0.record()
    for _ in range(reps):
        ctx.accumulate(cam, p, accum.data_ptr(), None)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    sspp = min(spp, 16)
    ctx.render(cam, ctx.params(W, H, 0, sspp, sspp, 50, variant, abi.FLAG_COUNT_EVENTS, 42, hs.background))
    st = ctx.stats()
    rpp = st["rays"] / st["paths"]
    paths = W * H * spp
    flops_ray = B.counted_flops(st) / st["rays"]
    row = {"config": name, "prims": hs.desc.n_prims, "variant_used": st["variant_used"], "ms": ms, "mpaths_s": paths / ms / 1e3,
           "mrays_s": paths * rpp / ms / 1e3, "rays_per_path": rpp, "flops_per_ray": flops_ray,
           "fp32_frac": paths * rpp * flops_ray / (ms * 1e-3) / 1e12 / peak_tf, "peak_tflops": peak_tf,
           "upload_ms": up["ms_upload"], "bvh_build_ms": up["ms_bvh_build"],
           "bvh_builder": "lbvh" if up["bvh_builder"] == abi.BVH_BUILDER_LBVH else "sah", "bvh_nodes": up["bvh_nodes"], "bvh_depth": up["bvh_depth"], "host_scene_s": t_host}
    if not args.no_cpu:
        osc = ob.OracleScene.from_desc(hs.desc, keep=hs)
        big = hs.desc.n_prims > 2000
        # primary-hit ids at a reduced frame (oracle cost), fp32 ref-order bit-exact + production vs f64
        pw, ph = (480, 270) if not big else (240, 135)
        pc = hs.camera(aspect=pw / ph)
        o32 = osc.primary_hits(pc, pw, ph, 32, use_bvh=big)[0]
        o64 = osc.primary_hits(pc, pw, ph, 64, use_bvh=big)[0]
        g32 = ctx.primary_hits(pc, pw, ph, 32, abi.VARIANT_MEGA_BVH)[0]
        gp = ctx.primary_hits(pc, pw, ph, 0, variant)[0]
        row["id_mismatch_fp32_reforder_vs_oracle_float"] = int((g32 != o32).sum())
        row["id_mismatch_production_vs_oracle_double"] = int((gp != o64).sum())
        row["id_pixels"] = pw * ph
        if not big:
            cw, chh, cs = (480, 270, 4)
            cc = hs.camera(aspect=cw / chh)
            r = osc.render(cc, cw, chh, cs, 50, hs.background, seed=3, precision=64, nthreads=nt, want_rgb8=False)
            row["cpu_mrays_s_all_threads"] = r["rays"] / r["secs"] / 1e6
            row["cpu_threads"] = nt
            r1 = osc.render(cc, cw // 2, chh // 2, 2, 50, hs.background, seed=3, precision=64, nthreads=1, want_rgb8=False)
            row["cpu_mrays_s_1_thread"] = r1["rays"] / r1["secs"] / 1e6
        if gate:
            gw, gh, gs = gate
            gc = hs.camera(aspect=gw / gh)
            ref = osc.render(gc, gw, gh, gs * 16, 50, hs.background, seed=1001, precision=64, nthreads=nt, want_rgb8=False)["accum"] / (gs * 16)
            cpu = osc.render(gc, gw, gh, gs, 50, hs.background, seed=2002, precision=64, nthreads=nt, want_rgb8=False)["accum"] / gs
            acc = ctx.render(gc, ctx.params(gw, gh, 0, gs, gs, 50, variant, 0, 42, hs.background), want_accum=True)[1]
            gpu = acc[..., :3].astype(np.float64) / gs
            row["bias_rel"] = [float(x) for x in np.abs(gpu.mean((0, 1)) - ref.mean((0, 1))) / ref.mean((0, 1))]
            row["rmse_gpu_over_cpu"] = float(np.sqrt(((gpu - ref) ** 2).mean()) / np.sqrt(((cpu - ref) ** 2).mean()))
    print(json.dumps(row), flush=True)
    del accum
