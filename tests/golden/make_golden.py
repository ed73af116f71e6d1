"""Writes tests/golden/oracle_golden.json from the CPU oracle (oracle/).

The reference has no golden vectors and cannot be run here (no zig toolchain), so these fixtures pin the
oracle to itself: a checksum of primary-hit maps and of small single-thread renders that continue the
seed-42 stream after scene generation exactly as the reference's main() does (src/main.zig:300-301, 382-394).
Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import numpy as np  # noqa: E402
import rtw_b200  # noqa: E402
import oracle_binding as ob  # noqa: E402

earth = rtw_b200.host_lib.decode_png(rtw_b200.host_lib.ASSET_EARTH)
out = {"primary_hits": {}, "render_1t": {}}
for sid in (1, 2, 3, 4, 5, 6):
    s = ob.OracleScene.builtin(sid, image=earth if sid == 4 else None)
    for prec in (64, 32):
        w, h = 96, 64
        ids, t, n = s.primary_hits(s.default_camera(), w, h, prec)
        out["primary_hits"][f"{sid}_{w}_{h}_{prec}"] = dict(id_sum=int(ids.astype(np.int64).sum()),
                                                             miss=int((ids == 0xFFFFFFFF).sum()), t_sum=float(t.sum()))
    w, h, spp = 24, 16, 4
    cfg = s.config()
    r = s.render(s.default_camera(), w, h, spp, 50, cfg["background"], precision=64, nthreads=1, continue_stream=True)
    out["render_1t"][f"{sid}_{w}_{h}_{spp}"] = dict(rays=int(r["rays"]), sum=[float(x) for x in r["accum"].sum(axis=(0, 1))],
                                                    rgb8_sum=int(r["rgb8"].astype(np.int64).sum()))
with open(os.path.join(HERE, "oracle_golden.json"), "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)
print("wrote", len(out["primary_hits"]), "+", len(out["render_1t"]), "fixtures")
