"""Multi-GPU plumbing for the one-process-per-GPU launch (torchrun): sample-range sharding and the single
reduce of the per-GPU accumulators.  torch.distributed is plumbing only; the reduce is one
NCCL `reduce(sum, f32)` over NVLink on GPUs and the same call over gloo in the CPU tests."""
from __future__ import annotations


def weak_sample_range(rank: int, spp_per_rank: int) -> tuple[int, int]:
    """Weak scaling: every rank renders spp_per_rank samples; rank g owns global samples [g*spp, (g+1)*spp)."""
    return rank * spp_per_rank, spp_per_rank


def strong_sample_range(rank: int, world: int, total_spp: int) -> tuple[int, int]:
    """Strong scaling: total_spp split evenly (the same split rdr_create_multi uses, rdr_multi.cpp)."""
    b = total_spp * rank // world
    e = total_spp * (rank + 1) // world
    return b, e - b


STRIPE_ROWS = 16


def stripe_rows_owned(rank: int, world: int, height: int, stripe_rows: int = STRIPE_ROWS) -> list[int]:
    """Image rows rank `rank` renders under row-stripe sharding: stripe s = rows [s*stripe_rows, (s+1)*stripe_rows)
    belongs to rank s % world (the rule of rdr_set_row_stripes / stripe_pixel in rdr_layout.h).  The stripes of the
    ranks are disjoint and cover the image, so reducing the accumulators adds zeros: bit-identical to one GPU."""
    if world <= 1 or stripe_rows <= 0:
        return list(range(height))
    return [y for y in range(height) if (y // stripe_rows) % world == rank]


def reduce_accum(accum, dst: int = 0):
    """Sums the per-rank float RGBA accumulators onto rank `dst` (in place on dst)."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM)
    return accum


def device_accum_tensor(renderer, n_pixels: int, device_index: int):
    """Wraps the renderer's device accumulator (rdr_accum_device_ptr) as a torch tensor without copying."""
    import torch
    ptr, nbytes = renderer.accum_device_ptr()
    assert nbytes == n_pixels * 16

    class _Arr:
        __cuda_array_interface__ = {"shape": (n_pixels * 4,), "typestr": "<f4", "data": (ptr, False), "version": 2}

    return torch.as_tensor(_Arr(), device=f"cuda:{device_index}")


def peer_pixel_slice(rank: int, world: int, n_pixels: int) -> tuple[int, int]:
    """The pixels rank `rank` combines in the fused reduce + resolve (peer_combine_kernel): (first, count) of the contiguous
    slice [n*rank/world, n*(rank+1)/world) -- the rule of rdr_peer_combine (rdr_api.cpp) and of the multi-GPU handle
    (rdr_multi.cpp).  The slices of the ranks are disjoint and cover the image."""
    first = n_pixels * rank // world
    return first, n_pixels * (rank + 1) // world - first
