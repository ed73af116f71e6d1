"""Small-frame run of every kernel family, meant to be run under compute-sanitizer --tool memcheck."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import rtw_b200
from rtw_b200 import abi
os.environ["RTW_WF_SLOTS"] = "4096"
ctx = rtw_b200.Context(0)
for sid, grid in ((1, 3), (6, 3), (1, 11), (7, 3), (5, 3)):
    hs = rtw_b200.HostScene(sid, grid=grid)
    ctx.upload_scene(hs.desc, keep=hs)
    cam = hs.camera(aspect=97 / 61)
    for variant in (1, 2, 3):
        for flags in (0, abi.FLAG_COUNT_EVENTS, abi.FLAG_DETERMINISTIC):
            if variant == 3 and flags == abi.FLAG_DETERMINISTIC:
                continue
            rgb, acc = ctx.render(cam, ctx.params(97, 61, 0, 5, 5, 50, variant, flags, 42, hs.background), want_accum=True)
            assert (acc[..., 3] == 5).all()
    for prec in (0, 32, 64):
        for variant in (1, 2):
            ctx.primary_hits(cam, 33, 17, prec, variant)
print("sanitize run ok")
