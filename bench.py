#!/usr/bin/env python
"""bench.py — headline benchmark of the path-tracing hot path (BASELINE.json metric: Mpaths/s & Mrays/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" = one full frame: every pixel of the 1920x1080 random-spheres scene (the reference's scene 1,
src/main.zig:157-221, camera main.zig:320-326 at 16:9), 500 samples per pixel PER GPU, depth 50
(BASELINE.json configs[1]).  With N > 1 ranks (one process per GPU, torchrun) rank r traces sample indices
[500 r, 500 (r+1)) of every pixel into its own fp32 buffer, one NCCL reduce sums the buffers onto rank 0, rank 0
resolves: weak scaling (per-GPU work fixed), the spp-split data path of SURVEY §8(e).

`value`   device-timed whole-job Mpaths/s, inputs resident in HBM (CUDA events on the launching stream, max over
          ranks, L2 flushed between iterations).
`e2e`     the same metric through the host-buffer C-ABI call (rtw_cuda_upload_scene + rtw_cuda_render with HOST
          buffers: scene tables and camera go host->device, the u8 image comes device->host, every step).
`roofline` FP32-pipe roofline of the path-tracing kernel: counted algorithmic flops (event counters of an
          instrumented replay x the per-event constants of SURVEY §8d) / kernel time, against the FFMA peak measured
          on this device in this run.
`cpu_baseline` the oracle port (f64 restatement of the reference, linear scan) on the box's host cores, on a
          bounded sample of the same frame.

--impl reference: the reference's own CPU algorithm (the oracle port; the Zig build cannot be produced in this
image) on all host threads, same config/metric.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH, HEIGHT, SPP, DEPTH, SEED, GRID = 1920, 1080, 500, 50, 42, 3

# algorithmic flops per event (FMA = 2)
FLOPS = dict(paths=45, rays=3 + 9, node_tests=25, sphere_tests=23, sphere_roots=6, moving_tests=12, sphere_finalise=26,
             rect_tests=2, rect_accepts=18, xform_apps=30, scatter_diffuse=40, scatter_metal=55, scatter_dielectric=60,
             tex_checker=8, tex_image=8)


def counted_flops(st):
    return float(sum(st[k] * v for k, v in FLOPS.items()))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s, p in zip(sm, power) if p >= 0.5 * max(power)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


REF_SAMPLE_SPP = 4  # the reference arm renders this many of the 500 spp per step (cost is exactly linear in spp)


def oracle_for_timing():
    """(oracle_binding module, its scene-1 OracleScene, camera, background, build description) for the CPU legs.

    Timing uses the oracle built for THIS machine (-O3 -march=native, FMA contraction allowed) — compiled here on first
    use because -march=native code must not travel between hosts; the parity build (-ffp-contract=off, portable) is only
    the fallback if that compile fails.  The scene comes from the oracle's own restatement of generateRandomScene
    (main.zig:157-221), so no library of the product (librtw_host.so / librtw_cuda.so) is mapped into this process."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob
    build = ob.use_native_build()
    osc = ob.OracleScene.builtin(1, GRID, SEED)
    cfg = osc.config()
    cam = osc.default_camera(aspect=WIDTH / HEIGHT)
    return ob, osc, cam, tuple(float(x) for x in cfg["background"]), build


def run_reference(args, rank):
    """The reference's CPU implementation of the path (oracle port), all host threads, same config."""
    if rank != 0:
        return
    ob, osc, cam, background, build = oracle_for_timing()
    nt = host_threads()  # torchrun exports OMP_NUM_THREADS=1; the reference arm may use every host core
    spp_step = REF_SAMPLE_SPP
    for _ in range(args.warmup):
        osc.render(cam, WIDTH, HEIGHT, 1, DEPTH, background, seed=1, precision=64, nthreads=nt, want_rgb8=False)
    secs, paths, rays = 0.0, 0, 0
    for k in range(args.steps):
        r = osc.render(cam, WIDTH, HEIGHT, spp_step, DEPTH, background, seed=100 + k, precision=64, nthreads=nt)
        secs += r["secs"]; paths += r["paths"]; rays += r["rays"]
    val = paths / secs / 1e6
    sample = (f"{WIDTH}x{HEIGHT} x {spp_step} spp per step (of the {SPP}-spp frame: ms_per_step is the time of this sample, the metric is "
              f"per path); f64 oracle port, linear scan, OpenMP over scanlines; {build}")
    cfg = config_dict(args.gpus)
    cfg["sample_spp"] = spp_step
    print(json.dumps({
        "impl": "reference", "metric": "Mpaths/s", "value": val, "unit": "Mpaths/s", "mrays_per_s": rays / secs / 1e6,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg, "gpu_launches": 0,
        "cpu_baseline": {"value": val, "unit": "Mpaths/s", "cores": nt, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def config_dict(n_gpus, spp=SPP):
    return {"workload": f"reference scene 1 (random spheres, grid half-extent {GRID}, <=40 prims, checker ground, moving "
                        f"diffuse spheres) {WIDTH}x{HEIGHT}, {spp} spp per GPU, depth {DEPTH}"
                        + (" = BASELINE.json configs[1]" if spp == SPP else " (NOT the baseline spp)"),
            "width": WIDTH, "height": HEIGHT, "spp_per_gpu": spp, "spp_total": spp * n_gpus, "max_depth": DEPTH,
            "seed": SEED, "partition": f"spp split x{n_gpus} + NCCL reduce to rank 0" if n_gpus > 1 else "single GPU",
            "l2": "flushed between timed iterations (256 MiB write)"}


S_WIDTH, S_HEIGHT, S_SPP = 3840, 2160, 1000  # BASELINE.json configs[4] / the north_star target


def strong_scaling(ctx, hs, rank, local_rank, world, sync_all, host_barrier, reps=3):
    """STRONG scaling on the record (the north_star target): ONE 3840x2160 frame of scene 1 at 1000 spp, depth 50, its
    samples split over the N GPUs, timed by the host's wall clock around whole frames (barrier + device
    synchronize on both sides, max over ranks), two ways:
      nccl  one process per GPU (this torchrun job): every rank traces 1000/N spp into its own fp32 buffer, ONE NCCL
            reduce sums the buffers onto rank 0, rank 0 resolves;
      peer  rank 0 alone drives all N devices through the library's own multi-GPU call (rtw_cuda_create_multi +
            rtw_cuda_render_multi): every device traces its share, then EVERY device resolves one scanline slab from all N
            buffers over NVLink peer mappings and copies its rows to the host (reduction fused into the resolve, reduce-
            scatter shaped, D2H included in the time) — what a Zig/C++ host gets without NCCL.
    Also asserts, inside this driver-visible run, that the N-device frame equals the 1-device frame (+-1 LSB)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import rtw_b200
    W, H, SP = S_WIDTH, S_HEIGHT, S_SPP
    cam = hs.camera(aspect=W / H)
    lo, hi = rtw_b200.dist.spp_range(rank, world, SP)
    accum = torch.zeros(H, W, 4, dtype=torch.float32, device="cuda")
    rgb8 = torch.zeros(H, W, 3, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    params = ctx.params(W, H, lo, hi, SP, DEPTH, 0, 0, SEED, hs.background)

    def frame():
        accum.zero_()
        rtw_b200.dist.render_distributed(lambda a, b: ctx.accumulate(cam, params, accum.data_ptr(), stream),
                                         lambda buf: ctx.resolve(buf.data_ptr(), W, H, SP, rgb8.data_ptr(), stream),
                                         accum, SP, rank, world)
    frame()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(reps):
        frame()
    sync_all()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    nccl_ms = float(t[0]) * 1e3 / reps
    paths = W * H * SP
    out = {"workload": f"scene 1 {W}x{H}, {SP} spp TOTAL split over {world} GPU(s), depth {DEPTH} = BASELINE.json configs[4]; host wall clock per frame",
           "n_gpus": world, "scaling": "strong", "reps": reps,
           "nccl": {"ms_per_frame": nccl_ms, "value": paths / (nccl_ms * 1e-3) / 1e6, "unit": "Mpaths/s",
                    "exchange": f"one ncclReduce of {W * H * 16 / 1e6:.1f} MB per rank onto rank 0, then resolve on rank 0" if world > 1 else "none"}}
    del accum
    peer = None
    # The other ranks must leave their GPUs IDLE while rank 0 drives them: they wait on the host (gloo), not in an NCCL
    # barrier — an NCCL kernel spinning on GPU g for rank 0 while rank 0's own kernels are queued on GPU g from another
    # process is two contexts time-slicing one device around a kernel that cannot finish (B200_PROFILING.md warns of it).
    torch.cuda.synchronize()
    host_barrier()
    if rank == 0:
        try:
            group = rtw_b200.create_multi(world) if world > 1 else [ctx]
            for c in group:
                if c is not ctx:
                    c.upload_scene(hs.desc, keep=hs)
            host = np.empty((H, W, 3), dtype=np.uint8)
            pm = ctx.params(W, H, 0, SP, SP, DEPTH, 0, 0, SEED, hs.background)
            rtw_b200.render_multi(group, cam, pm, rgb8=host)
            t0 = time.perf_counter()
            walls = []
            for _ in range(reps):
                rtw_b200.render_multi(group, cam, pm, rgb8=host)
                walls.append(group[0].stats()["ms_wall"])
            peer_ms = (time.perf_counter() - t0) * 1e3 / reps
            st = group[0].stats()
            # N devices == 1 device on the same sample set (small frame; +-1 LSB from the fp32 summation order)
            cw, ch, csp = 480, 270, 8 * world + 3
            ccam = hs.camera(aspect=cw / ch)
            cp = ctx.params(cw, ch, 0, csp, csp, DEPTH, 0, 0, SEED, hs.background)
            multi = rtw_b200.render_multi(group, ccam, cp)
            single = ctx.render(ccam, cp)[0]
            diff = np.abs(multi.astype(int) - single.astype(int))
            same = bool(diff.max() <= 1 and (diff > 0).mean() < 1e-3)
            assert same, f"render_multi over {world} devices differs from the 1-device frame (max {diff.max()}, {(diff > 0).mean():.2e} of channels)"
            peer = {"ms_per_frame": peer_ms, "value": paths / (peer_ms * 1e-3) / 1e6, "unit": "Mpaths/s",
                    "ms_wall_in_library": sum(walls) / len(walls), "ms_trace_dev0": st["ms_trace"], "ms_resolve_dev0": st["ms_resolve"],
                    "exchange": (f"every GPU reads its {H // world}-row slab from the other {world - 1} buffers over NVLink peer mappings inside the "
                                 f"resolve kernel ({(world - 1) * W * H * 16 / world / 1e6:.1f} MB ingress per GPU), parallel D2H of the slabs") if world > 1 else "none",
                    "equals_single_device_frame": same, "check": f"{cw}x{ch}x{csp} spp, max |diff| {int(diff.max())} LSB, {float((diff > 0).mean()):.1e} of channels differ"}
            for c in group:
                if c is not ctx:
                    c.close()
        except rtw_b200.RtwCudaError as e:  # e.g. no peer access between the devices of this box
            peer = {"unavailable": str(e)}
    out["peer"] = peer
    host_barrier()
    sync_all()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--spp", type=int, default=SPP)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling block (C5: 3840x2160, 1000 spp split over the GPUs)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3

    import numpy as np
    import torch
    import torch.distributed as dist
    import rtw_b200
    from rtw_b200 import abi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    host_group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        host_group = dist.new_group(backend="gloo")  # host-side barriers (no kernel on any GPU)

    def host_barrier():
        if world > 1:
            dist.barrier(group=host_group)
    spp = args.spp
    ctx = rtw_b200.Context(local_rank)
    hs = rtw_b200.HostScene(1, grid=GRID, seed=SEED)
    cam = hs.camera(aspect=WIDTH / HEIGHT)
    ctx.upload_scene(hs.desc, keep=hs)
    lo, hi = rtw_b200.dist.spp_range(rank, world, spp * world)
    spp_total = spp * world

    accum = torch.zeros(HEIGHT, WIDTH, 4, dtype=torch.float32, device="cuda")
    rgb8 = torch.zeros(HEIGHT, WIDTH, 3, dtype=torch.uint8, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    params = ctx.params(WIDTH, HEIGHT, lo, hi, spp_total, DEPTH, args.variant, 0, SEED, hs.background)

    def trace(lo_, hi_):  # this rank's sample indices of every pixel, added into its fp32 buffer
        assert (lo_, hi_) == (lo, hi)
        ctx.accumulate(cam, params, accum.data_ptr(), stream)

    def resolve(buf):
        ctx.resolve(buf.data_ptr(), WIDTH, HEIGHT, spp_total, rgb8.data_ptr(), stream)

    def step(after_trace=None):
        accum.zero_()
        # spp split -> (NCCL reduce onto rank 0) -> resolve on rank 0: the control flow the gloo tests cover on CPU
        rtw_b200.dist.render_distributed(trace, resolve, accum, spp_total, rank, world, after_accumulate=after_trace)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sync_all()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    sync_all()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)  # L2 flush, outside the timed events
        ev[k][0].record()
        step(after_trace=ev[k][2].record)
        ev[k][1].record()
    sync_all()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = sum(a.elapsed_time(b) for a, b, _ in ev)
    ms_kernel = sum(a.elapsed_time(c) for a, _, c in ev)  # zero + path-tracing kernel
    t = torch.tensor([ms_total, ms_kernel], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_kernel = float(t[0]), float(t[1])
    paths_step = WIDTH * HEIGHT * spp * world
    value = paths_step * args.steps / (ms_total * 1e-3) / 1e6

    # ---- e2e: the host-buffer C-ABI call, H2D + D2H inside the timed region ----------------------------------
    host_img = torch.empty(HEIGHT, WIDTH, 3, dtype=torch.uint8).pin_memory()
    host_np = host_img.numpy()
    scene_bytes = (hs.desc.n_prims * ctypes.sizeof(abi.Prim) + hs.desc.n_materials * ctypes.sizeof(abi.Material)
                   + hs.desc.n_textures * ctypes.sizeof(abi.Texture) + hs.desc.n_xforms * ctypes.sizeof(abi.Xform))
    h2d = scene_bytes + ctypes.sizeof(abi.Camera) + ctypes.sizeof(abi.RenderParams)
    d2h = HEIGHT * WIDTH * 3
    e2e_steps = max(2, min(args.steps, 3))

    def e2e_step():
        ctx.upload_scene(hs.desc, keep=hs)
        if world == 1:
            ctx.render(cam, params, rgb8=host_np)  # blocking; copies the image into pinned host memory
        else:
            step()
            if rank == 0:
                host_img.copy_(rgb8, non_blocking=True)
            torch.cuda.synchronize()
    e2e_step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    sync_all()
    e2e_secs = time.perf_counter() - t0
    te = torch.tensor([e2e_secs], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = paths_step * e2e_steps / float(te[0]) / 1e6

    # ---- roofline: replay one step with event counters on (same Philox keys => same paths) -------------------
    pstat = ctx.params(WIDTH, HEIGHT, lo, hi, spp_total, DEPTH, args.variant, abi.FLAG_COUNT_EVENTS, SEED, hs.background)
    ctx.render(cam, pstat, rgb8=host_np)
    st = ctx.stats()
    sums = torch.tensor([float(st["rays"]), float(st["paths"]), counted_flops(st)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    rays_step, flops_step = float(sums[0]), float(sums[2])

    strong = None if args.no_strong else strong_scaling(ctx, hs, rank, local_rank, world, sync_all, host_barrier)

    if rank == 0:
        peak_tf, peak_mhz = ctx.measure_fp32_peak()
        kernel_ms = ms_kernel / args.steps
        achieved = counted_flops(st) / (kernel_ms * 1e-3) / 1e12  # rank 0's kernel, rank 0's flops
        # DRAM bytes per launch from an ncu --set full capture of THIS kernel variant on THIS workload (profiles/traffic.json,
        # keyed "<variant_used>:<width>x<height>x<spp>"); null when no matching capture is committed
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get(f"{st['variant_used']}:{WIDTH}x{HEIGHT}x{spp}", {}).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "mrays_per_s": rays_step * args.steps / (ms_total * 1e-3) / 1e6,
            "rays_per_path": rays_step / paths_step, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config_dict(world, spp),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mpaths/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "call": "rtw_cuda_upload_scene + rtw_cuda_render (host buffers)" if world == 1
                    else "upload + accumulate + NCCL reduce + resolve + D2H image"},
            "gpu_launches": 2 * args.steps,
            "kernel": {"name": "k_megakernel_flat<STATS=0, MINB=8, FEAT=FF_SPHERES>" if st["variant_used"] == 1 else "k_megakernel_bvh<0>",
                       "variant": st["variant_used"], "ms_per_launch": kernel_ms,
                       "flops_per_ray_counted": counted_flops(st) / max(1, st["rays"])},
            "strong": strong,
            "roofline": {"bound": "fp32", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                         "traffic": traffic, "peak_source": f"FFMA microbenchmark measured in this run on this device ({peak_mhz:.0f} MHz max clock)",
                         "note": "FP32-pipe roofline (north_star: not a dense contraction, no tensor cores; HBM traffic is one 33 MB accumulation buffer per frame)"},
        }
        if not args.no_cpu_baseline and world == 1:
            ob, osc, ocam, obg, obuild = oracle_for_timing()
            nt = host_threads()
            cspp = 12
            r = osc.render(ocam, WIDTH, HEIGHT, cspp, DEPTH, obg, seed=5, precision=64, nthreads=nt, want_rgb8=False)
            r1 = osc.render(ocam, WIDTH // 4, HEIGHT // 4, 8, DEPTH, obg, seed=5, precision=64, nthreads=1,
                            continue_stream=False, want_rgb8=False)
            line["cpu_baseline"] = {
                "value": r["paths"] / r["secs"] / 1e6, "unit": "Mpaths/s", "cores": nt, "kind": "port",
                "mrays_per_s": r["rays"] / r["secs"] / 1e6,
                "sample": f"{WIDTH}x{HEIGHT} x {cspp} spp of the {SPP}-spp frame ({r['secs']:.1f} s wall on {nt} threads); f64 oracle port, linear scan; {obuild}",
                "value_1_thread": r1["paths"] / r1["secs"] / 1e6,
                "sample_1_thread": f"{WIDTH // 4}x{HEIGHT // 4} x 8 spp, single sequential stream as the reference ({r1['secs']:.1f} s)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
