"""Development helper: per-stage upload times (RTW_UPLOAD_TRACE) of the 1M-sphere scene, several uploads in a row."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["RTW_UPLOAD_TRACE"] = "1"
import rtw_b200
ctx = rtw_b200.Context(0)
hs = rtw_b200.HostScene(8, grid=int(sys.argv[1]) if len(sys.argv) > 1 else 500)
for k in range(int(sys.argv[2]) if len(sys.argv) > 2 else 6):
    ctx.upload_scene(hs.desc, keep=hs)
    st = ctx.stats()
    print(f"upload {k}: {st['ms_upload']:.1f} ms, bvh {st['ms_bvh_build']:.1f} ms", flush=True)
