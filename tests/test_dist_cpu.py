"""The N>1 control flow on CPU: world_size-2 (and 3) gloo groups run the spp split + reduce + resolve with the
ORACLE standing in for the kernels (checker-side only), and must reproduce the single-rank result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_spp_range_is_a_disjoint_cover(rtw):
    for spp in (0, 1, 7, 8, 500, 1000, 1001):
        for world in (1, 2, 3, 4, 8):
            got = [rtw.dist.spp_range(r, world, spp) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == spp
            assert all(got[i][1] == got[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in got]
            assert max(sizes) - min(sizes) <= 1
    assert rtw.dist.spp_range(1, 2, 10, spp_begin=100) == (105, 110)
    with pytest.raises(ValueError):
        rtw.dist.spp_range(2, 2, 10)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, spp, W, H, out_path):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    sys.path.insert(0, os.path.dirname(here))
    import rtw_b200
    import oracle_binding as ob
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    osc = ob.OracleScene.builtin(6)
    cam = osc.default_camera()
    accum = torch.zeros(H, W, 4, dtype=torch.float32)

    def accumulate(lo, hi):
        # per-sample determinism stand-in: sample s is rendered with seed 1000+s, whichever rank owns it
        for s in range(lo, hi):
            r = osc.render(cam, W, H, 1, 50, (0, 0, 0), seed=1000 + s, precision=64, nthreads=2, want_rgb8=False)
            accum[..., :3] += torch.from_numpy(r["accum"].astype(np.float32))
            accum[..., 3] += 1

    def resolve(a):
        return a.clone()
    res = rtw_b200.dist.render_distributed(accumulate, resolve, accum, spp, rank, world)
    if rank == 0:
        np.save(out_path, res.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_spp_split_matches_single_rank(rtw, oracle, tmp_path, world):
    spp, W, H = 7, 24, 24
    out = str(tmp_path / f"w{world}.npy")
    mp.spawn(_worker, args=(world, _free_port(), spp, W, H, out), nprocs=world, join=True)
    got = np.load(out)
    osc = oracle.OracleScene.builtin(6)
    cam = osc.default_camera()
    want = np.zeros((H, W, 3), dtype=np.float32)
    for s in range(spp):
        want += osc.render(cam, W, H, 1, 50, (0, 0, 0), seed=1000 + s, precision=64, nthreads=2, want_rgb8=False)["accum"].astype(np.float32)
    assert (got[..., 3] == spp).all()
    np.testing.assert_allclose(got[..., :3], want, rtol=1e-5, atol=1e-6)
