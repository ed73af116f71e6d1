"""Development helper: per-stage upload times (RTW_UPLOAD_TRACE), several uploads in a row.

    python tools/upload_trace.py [scene=8] [grid=500] [repeats=6]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["RTW_UPLOAD_TRACE"] = "1"
import rtw_b200
ctx = rtw_b200.Context(0)
hs = rtw_b200.HostScene(int(sys.argv[1]) if len(sys.argv) > 1 else 8, grid=int(sys.argv[2]) if len(sys.argv) > 2 else 500)
for k in range(int(sys.argv[3]) if len(sys.argv) > 3 else 6):
    ctx.upload_scene(hs.desc, keep=hs)
    st = ctx.stats()
    print(f"upload {k}: {st['ms_upload']:.2f} ms, bvh {st['ms_bvh_build']:.2f} ms", flush=True)
