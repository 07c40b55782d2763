"""torch.distributed plumbing for the multi-process launch (one process per GPU, torchrun): the path's only exchange
step -- ONE gather of the RGB8 framebuffer to rank 0 and ONE reduce of the ray counters (NCCL over NVLink on GPUs,
gloo in the CPU tests).  Nothing here computes pixels; the de-interleave of the gathered row tiles runs in the CUDA
library (r1_deinterleave_rows)."""
import ctypes as C

import torch
import torch.distributed as dist

import rays1bench_b200 as r1


def max_local_rows(height, row_tile, world):
    return max(r1.local_rows(height, row_tile, r, world) for r in range(world))


def gather_framebuffer(local_rgb, height, row_tile, rank, world, dst=0):
    """local_rgb: uint8 [max_local_rows, width, 3] (rows past this rank's own count are padding).
    Returns on ``dst`` a uint8 tensor [world, max_local_rows, width, 3]; None elsewhere.  One collective."""
    assert local_rgb.dtype == torch.uint8 and local_rgb.is_contiguous()
    if world == 1:
        return local_rgb.unsqueeze(0)
    if rank == dst:
        gathered = torch.empty((world,) + tuple(local_rgb.shape), dtype=torch.uint8, device=local_rgb.device)
        dist.gather(local_rgb, list(gathered.unbind(0)), dst=dst)
        return gathered
    dist.gather(local_rgb, None, dst=dst)
    return None


def reduce_ray_count(count, world, dst=0):
    """count: int64 tensor [1] (the device-side uint64 counter viewed as int64).  Sum lands on ``dst``.  One collective."""
    if world > 1:
        dist.reduce(count, dst=dst, op=dist.ReduceOp.SUM)
    return count


def deinterleave(gathered, width, height, row_tile, world):
    """[world, max_local_rows, width, 3] on a CUDA device -> [height, width, 3] via the library's kernel."""
    assert gathered.is_cuda and gathered.is_contiguous()
    out = torch.empty((height, width, 3), dtype=torch.uint8, device=gathered.device)
    stride = gathered.stride(0)
    stream = torch.cuda.current_stream(gathered.device).cuda_stream
    r1._check(r1.lib.r1_deinterleave_rows(gathered.device.index, C.c_void_p(gathered.data_ptr()), stride, C.c_void_p(out.data_ptr()),
                                          width, height, row_tile, world, C.c_void_p(stream)), "r1_deinterleave_rows")
    return out
