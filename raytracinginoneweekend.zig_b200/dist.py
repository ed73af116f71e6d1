"""Multi-GPU plumbing: samples-per-pixel split + one sum onto rank 0.

One process per GPU (torch.distributed, NCCL over NVLink / NVSwitch).  Rank r of R traces the sample indices
`spp_range(r, R, spp)` of EVERY pixel into its own fp32 W*H*4 buffer; because Philox is keyed by the absolute
sample index the union over ranks is exactly the 1-GPU sample set.  The only exchange step of the path is the
sum of those buffers onto rank 0 (`dist.reduce`), after which rank 0 runs the resolve kernel.  The reference
has no analogue (it is single-threaded); this replaces nothing and adds the partition north_star asks for.
"""


def spp_range(rank, world, spp_total, spp_begin=0):
    """Disjoint cover of [spp_begin, spp_begin+spp_total): the first (spp_total mod world) ranks take one extra."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(spp_total, world)
    lo = spp_begin + rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def render_distributed(accumulate, resolve, accum, spp_total, rank, world, group=None, reduce_op=None,
                       after_accumulate=None):
    """accumulate(lo, hi) adds this rank's samples into `accum` (a torch tensor, H x W x 4 fp32);
    then the buffers are summed onto rank 0, which calls resolve(accum) and returns its result.
    `accumulate`/`resolve` are callables so the same control flow runs on CPU tensors under gloo in tests and on
    device pointers under NCCL in bench.py; `after_accumulate` (optional) is called between the two phases (bench.py
    records a CUDA event there to time the path-tracing kernel on its own)."""
    import torch.distributed as dist
    lo, hi = spp_range(rank, world, spp_total)
    accumulate(lo, hi)
    if after_accumulate is not None:
        after_accumulate()
    if world > 1:
        dist.reduce(accum, dst=0, op=reduce_op or dist.ReduceOp.SUM, group=group)
    if rank == 0:
        return resolve(accum)
    return None
