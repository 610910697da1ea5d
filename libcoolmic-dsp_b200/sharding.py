"""Stream sharding across GPUs: the partition rule, and the helpers of the CPU (gloo) test.

Streams are independent, so rank r owns the contiguous range
[r*N/R, (r+1)*N/R) and the data path needs no exchange. The one collective -- per-stream meter
results to rank 0 once per reporting interval -- is done by the C library itself over NCCL
(cmgpu_comm_*, cmgpu_gather_results in csrc/cmgpu_comm.cu; binding.Comm). What is left here:
`stream_range`, and `gather_rows` / `decode_rows`, which tests/test_sharding.py uses to push the same
raw rows through a world-size-2 gloo gather on CPU and decode them with the C ABI
(cmgpu_meter_decode / cmgpu_finalise) -- the host-side half of the N > 1 path, without GPUs.
"""
from __future__ import annotations

import ctypes as C

import numpy as np


def stream_range(total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced (sizes differ by at most one), covering [0, total) exactly once."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad world/rank")
    return total * rank // world, total * (rank + 1) // world


def wrap_device_rows(ptr: int, n_u64: int):
    """A torch CUDA tensor aliasing the engine's device meter table (no copy)."""
    import torch

    class _Dev:
        __cuda_array_interface__ = {"shape": (n_u64,), "typestr": "<i8", "data": (ptr, False), "version": 2}

    return torch.as_tensor(_Dev(), device="cuda")


def gather_rows(mine, dist, rank: int, world: int, counts: list[int] | None = None):
    """Gather each rank's int64 row tensor to rank 0. `counts` = elements per rank when shards are
    uneven (padded to the maximum for the collective). Returns the list of tensors on rank 0."""
    import torch

    if counts is None:
        counts = [mine.numel()] * world
    width = max(counts)
    if mine.numel() < width:
        pad = torch.zeros(width, dtype=mine.dtype, device=mine.device)
        pad[: mine.numel()] = mine
        mine = pad
    out = [torch.empty(width, dtype=mine.dtype, device=mine.device) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, out, dst=0)
    if rank != 0:
        return None
    return [t[:n] for t, n in zip(out, counts)]


def decode_rows(lib, rows: np.ndarray, channels: int):
    """rows: uint64/int64 [n_streams * (2C+2)] on the host -> (MeterState * n_streams)."""
    from .binding import MeterState

    rows = np.ascontiguousarray(rows).view(np.uint64)
    row = 2 * channels + 2
    n = rows.size // row
    out = (MeterState * n)()
    rc = lib.cmgpu_meter_decode(rows.ctypes.data_as(C.POINTER(C.c_uint64)), n, channels, out)
    if rc != 0:
        raise RuntimeError(f"cmgpu_meter_decode failed: {rc}")
    return out
