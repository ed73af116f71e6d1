"""Ad-hoc GPU check used during development (not a test): parity probes + quick timings."""
import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import rtw_b200
from rtw_b200 import abi
import oracle_binding as ob

out = {}
ctx = rtw_b200.Context(0)
print("fp32 peak", ctx.measure_fp32_peak(), flush=True)
for sid, W, H in ((1, 600, 400), (6, 600, 600), (2, 300, 200), (5, 300, 200)):
    hs = rtw_b200.HostScene(sid)
    cam = hs.camera()
    ctx.upload_scene(hs.desc, keep=hs)
    osc = ob.OracleScene.from_desc(hs.desc, keep=hs)
    for prec in (32, 64):
        oid, ot, on = osc.primary_hits(cam, W, H, prec)
        for var in (abi.VARIANT_MEGA_FLAT, abi.VARIANT_MEGA_BVH):
            gid, gt, gn = ctx.primary_hits(cam, W, H, prec, var)
            print(f"scene {sid} prec {prec} var {var}: id mismatches {(gid != oid).sum()} / {oid.size}; "
                  f"t bit-equal {np.array_equal(gt, ot)} max|dt| {np.abs(gt - ot).max():.3e}; max|dn| {np.abs(gn - on).max():.3e}", flush=True)
    oid, ot, on = osc.primary_hits(cam, W, H, 64)
    for var in (abi.VARIANT_MEGA_FLAT, abi.VARIANT_MEGA_BVH):
        gid, gt, gn = ctx.primary_hits(cam, W, H, 0, var)
        mm = gid != oid
        same = ~mm & (oid != 0xFFFFFFFF)
        print(f"scene {sid} PRODUCTION var {var}: id mismatches vs f64 oracle {mm.sum()} / {oid.size}; "
              f"max rel|dt| {(np.abs(gt - ot)[same] / np.maximum(1, np.abs(ot[same]))).max():.3e}; max|dn| {np.abs(gn - on)[same].max():.3e}", flush=True)

# render timings
for sid, W, H, spp in ((1, 600, 400, 50), (6, 600, 600, 200), (1, 1920, 1080, 100)):
    hs = rtw_b200.HostScene(sid)
    aspect = W / H
    cam = hs.camera(aspect=aspect)
    ctx.upload_scene(hs.desc, keep=hs)
    for var in (abi.VARIANT_MEGA_FLAT, abi.VARIANT_MEGA_BVH):
        p = ctx.params(W, H, 0, spp, spp, 50, var, abi.FLAG_COUNT_EVENTS, 42, hs.background)
        rgb, acc = ctx.render(cam, p, want_accum=True)
        st = ctx.stats()
        p = ctx.params(W, H, 0, spp, spp, 50, var, 0, 42, hs.background)
        for _ in range(2):
            t0 = time.time(); rgb, acc = ctx.render(cam, p, want_accum=True); t1 = time.time()
        st2 = ctx.stats()
        print(f"scene {sid} {W}x{H}x{spp} var {var}: trace {st2['ms_trace']:.2f} ms wall {1e3*(t1-t0):.1f} ms; "
              f"{st['paths']/st2['ms_trace']/1e3:.1f} Mpaths/s {st['rays']/st2['ms_trace']/1e3:.1f} Mrays/s rays/path {st['rays']/max(1,st['paths']):.2f} "
              f"mean {acc[..., :3].mean(axis=(0, 1)) / spp} nan {st2['nan_pixels']}", flush=True)
        print("   stats", {k: v for k, v in st.items() if isinstance(v, int) and v}, flush=True)
        rtw_b200.host_lib.write_ppm(os.path.join(ROOT, "gpurun_out", f"scene{sid}_{W}x{H}_v{var}.ppm"), rgb)
    if W <= 600:
        osc = ob.OracleScene.from_desc(hs.desc, keep=hs)
        r = osc.render(cam, W, H, min(spp, 32), 50, hs.background, seed=7, precision=64, nthreads=ob.num_threads())
        print(f"   oracle f64 mean {r['accum'].mean(axis=(0, 1)) / min(spp, 32)} ({r['rays']/r['secs']/1e6:.2f} Mrays/s on {ob.num_threads()} threads; rays/path {r['rays']/r['paths']:.2f})", flush=True)
