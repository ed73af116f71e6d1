"""Build recipes for the native parts (run by __graft_entry__.build(), tests and bench.py).

  csrc/  -> librtw_cuda.so   nvcc, -gencode arch=compute_100a,code=sm_100a (cross-compiles without a GPU)
  host/  -> librtw_host.so   g++   (the C++ twin of the Zig host: scene API, builders, flatten, PPM)
            rtw_render       g++   (CLI mirroring src/main.zig, links both)

Everything is built IN-TREE under _build/ so the .so files travel to the GPU box with the snapshot.
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
OUT = os.path.join(PKG, "_build")
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")
CUDA_LIB = os.path.join(OUT, "librtw_cuda.so")
HOST_LIB = os.path.join(OUT, "librtw_host.so")
RENDER_BIN = os.path.join(OUT, "rtw_render")

NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-ccbin", CXX] + ARCH


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("build step failed: " + " ".join(cmd))
    if verbose and (r.stdout or r.stderr):
        print(r.stdout + r.stderr, flush=True)


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _deps(d, exts):
    return [os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith(exts)]


class _BuildLock:
    """Serialises in-tree builds across processes (torchrun starts N ranks at once)."""

    def __enter__(self):
        import fcntl
        os.makedirs(OUT, exist_ok=True)
        self.f = open(os.path.join(OUT, ".build.lock"), "w")
        fcntl.flock(self.f, fcntl.LOCK_EX)
        return self

    def __exit__(self, *exc):
        import fcntl
        fcntl.flock(self.f, fcntl.LOCK_UN)
        self.f.close()


def build_cuda(force=False, verbose=False, extra=()):
    with _BuildLock():
        return _build_cuda(force, verbose, extra)


def build_host(force=False, verbose=False):
    with _BuildLock():
        return _build_host(force, verbose)


def _build_cuda(force=False, verbose=False, extra=()):
    os.makedirs(OUT, exist_ok=True)
    hdrs = _deps(CSRC, (".h", ".cuh")) + [os.path.join(ROOT, "include", "rtw_cuda.h"), os.path.abspath(__file__)]
    units = [
        ("rtw_kernels.cu", ["-Xptxas", "-v"]),          # production arithmetic: FMA contraction on
        ("rtw_wavefront.cu", []),                       # K2 wavefront schedule (same device functions)
        ("rtw_unit.cu", []),                            # unit probes of the stochastic code (same flags as rtw_kernels.cu)
        ("rtw_probe.cu", ["-fmad=false"]),              # reference-order probe: no contraction, IEEE div/sqrt
        ("rtw_lbvh.cu", []),                            # device-side BVH build for large scenes
        ("rtw_api.cpp", []),
        ("rtw_bvh.cpp", []),
    ]
    objs = []
    rebuilt = False
    for src, flags in units:
        s = os.path.join(CSRC, src)
        o = os.path.join(OUT, src.rsplit(".", 1)[0] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            _run([NVCC] + NVCC_COMMON + list(flags) + list(extra) + ["-c", s, "-o", o], verbose)
            rebuilt = True
    if rebuilt or not os.path.exists(CUDA_LIB):
        _run([NVCC, "-shared", "-ccbin", CXX] + ARCH + ["-o", CUDA_LIB] + objs, verbose)
    return CUDA_LIB


def _build_host(force=False, verbose=False):
    os.makedirs(OUT, exist_ok=True)
    srcs = _deps(HOST, (".cpp",))
    hdrs = _deps(HOST, (".hpp", ".h")) + [os.path.join(ROOT, "include", "rtw_cuda.h"), os.path.abspath(__file__)]
    lib_srcs = [s for s in srcs if not s.endswith("main.cpp")]
    flags = ["-O2", "-std=c++17", "-fPIC", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include")]
    if force or _stale(HOST_LIB, lib_srcs + hdrs):
        _run([CXX] + flags + ["-shared", "-o", HOST_LIB] + lib_srcs + ["-lz"], verbose)
    main = os.path.join(HOST, "main.cpp")
    if os.path.exists(main) and (force or _stale(RENDER_BIN, [main, HOST_LIB] + hdrs)):
        _run([CXX] + flags + ["-o", RENDER_BIN, main, "-L", OUT, "-lrtw_host", "-ldl",
                              "-Wl,-rpath,$ORIGIN"], verbose)
    return HOST_LIB


def build_all(force=False, verbose=False):
    return build_cuda(force, verbose), build_host(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
