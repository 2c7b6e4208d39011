"""Multi-GPU plumbing: one process per GPU (torchrun), samples sharded by sample index, one reduce of the float4
accumulation image over NCCL / NVLink.  (The reference is single-GPU, src/main.cpp:222; sharding follows
BASELINE.json's north star.)  torch.distributed is plumbing only: the reduce runs on the context's own buffer."""
import numpy as np


def sample_range(rank, world, spp, first_sample=0):
    """Contiguous block of sample indices for `rank`: the blocks of all ranks tile [first, first+spp) exactly."""
    if not (0 <= rank < world) or spp < 0:
        raise ValueError("bad rank/world/spp")
    b = first_sample + spp * rank // world
    e = first_sample + spp * (rank + 1) // world
    return b, e - b


class _DevArray:
    def __init__(self, ptr, n_floats):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 3}


def accum_tensor(ctx):
    """zero-copy torch view (float32, 4 per pixel) of a Context's accumulation buffer in HBM"""
    import torch
    ptr, nbytes = ctx.accum_device_ptr()
    return torch.as_tensor(_DevArray(ptr, nbytes // 4), device=torch.device("cuda", ctx.device))


def reduce_image(t, dst=0, group=None):
    """sum the per-rank accumulation images into rank `dst` (NCCL on GPUs; gloo in the CPU tests)"""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return t


def render_sharded(ctx, spp, max_depth, seed, rank, world, stream=None, first_sample=0):
    """render this rank's share of `spp` samples, then combine on rank 0; returns (first, count) rendered here.

    `stream` (a torch.cuda.Stream): the render AND the reduce are issued on it -- the context is switched to that stream
    first, so the reduce is ordered after the last kernel that adds to the image.  Without a stream the context's own
    stream is synchronised before the reduce."""
    import torch
    b, n = sample_range(rank, world, spp, first_sample)
    if stream is not None:
        ctx.set_stream(stream.cuda_stream)  # no-op if the caller did it already; otherwise the reduce could overtake k_bounce
    ctx.clear()
    if n:
        ctx.render(b, n, max_depth, seed)
    if world > 1:
        t = accum_tensor(ctx)
        if stream is not None:
            with torch.cuda.stream(stream):
                reduce_image(t)
        else:
            ctx.sync()
            reduce_image(t)
    return b, n
