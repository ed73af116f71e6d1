"""rtw-b200: the B200-native path-tracing loop of nsfisis/RayTracingInOneWeekend.zig.

Python here is harness only (ctypes over the C ABI of include/rtw_cuda.h and over the C++ twin of
the Zig host); the product is csrc/ (CUDA, sm_100a) and host/ (C++).  Import as `rtw_b200`
(see rtw_b200.py at the repo root: the directory name contains a dot).
"""
from . import abi, build, cuda_lib, dist, host_lib  # noqa: F401
from .cuda_lib import Context, RtwCudaError, render_multi, create_multi  # noqa: F401
from .host_lib import HostScene, camera_init  # noqa: F401
