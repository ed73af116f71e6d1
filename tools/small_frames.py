import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, rtw_b200
ctx = rtw_b200.Context(0)
for sid, W, H, spp in ((6, 600, 600, 200), (1, 600, 400, 50), (1, 1920, 1080, 500)):
    hs = rtw_b200.HostScene(sid)
    ctx.upload_scene(hs.desc, keep=hs)
    cam = hs.camera(aspect=W / H)
    acc = torch.zeros(H, W, 4, device="cuda")
    p = ctx.params(W, H, 0, spp, spp, 50, 0, 0, 42, hs.background)
    for _ in range(3): ctx.accumulate(cam, p, acc.data_ptr(), None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): ctx.accumulate(cam, p, acc.data_ptr(), None)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"scene {sid} {W}x{H}x{spp}: {ms:.3f} ms  {W*H*spp/ms/1e3:.0f} Mpaths/s (RTW_BATCH_SPP={os.environ.get('RTW_BATCH_SPP','auto')})")
