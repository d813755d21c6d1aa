"""Multi-GPU plumbing: one process per GPU (torchrun), paths sharded by global trajectory index.

The data path has no collective: each rank simulates its own contiguous block of trajectories with
disjoint Philox counters. Only the partial sums [sum, sumsq, n] (and, for LSM, the per-date regression
moments) are sum-allreduced — NCCL on GPU boxes, gloo in the CPU tests.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi as abi


def _cuda_device(device=None):
    """The CUDA device a rank's collectives run on: the engine's device (LOCAL_RANK), NOT torch's current device — under
    plain torchrun a user who never called torch.cuda.set_device would otherwise put every rank on cuda:0."""
    import os
    import torch
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    return torch.device("cuda", int(device))


def allreduce_sum_f64(v: np.ndarray, group=None, device=None) -> np.ndarray:
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64).copy())
    if dist.get_backend(group) == "nccl":
        t = t.to(_cuda_device(device))
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy().reshape(np.shape(v))


def allreduce_max_int(x: int, group=None, device=None) -> int:
    """Max over ranks of a small integer (return codes: every rank learns whether ANY rank failed)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([int(x)], dtype=torch.int64)
    if dist.get_backend(group) == "nccl":
        t = t.to(_cuda_device(device))
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return int(t.cpu().item())


class _DevArray:
    """Wrap a raw device pointer for torch.as_tensor (zero copy) via __cuda_array_interface__."""

    def __init__(self, ptr: int, count: int, stream: int):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 3,
                                         "strides": None, "stream": stream or 1}


def make_comm(shard, group=None, device=None):
    """hh_comm whose callback sum-allreduces a device buffer in place with torch.distributed (NCCL). `device`: the
    engine's CUDA device (default LOCAL_RANK)."""
    import torch
    import torch.distributed as dist
    rank, world = shard

    def _cb(user, dev_ptr, count, stream):
        try:
            if dist.get_backend(group) == "nccl":
                dev = _cuda_device(device)
                with torch.cuda.device(dev):
                    ext = torch.cuda.ExternalStream(stream, device=dev) if stream else torch.cuda.current_stream(dev)
                    with torch.cuda.stream(ext):
                        # stream-ordered: ProcessGroupNCCL makes `ext` wait for its collective, so the library's next
                        # kernel (the fit) sees the reduced moments without a host round trip
                        t = torch.as_tensor(_DevArray(dev_ptr, count, stream), device=dev)
                        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            else:  # gloo (CPU tests with an injected engine): dev_ptr is a host pointer
                buf = (C.c_double * count).from_address(dev_ptr)
                t = torch.frombuffer(buf, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            return 0
        except Exception as e:  # never let an exception cross the C boundary
            import sys
            print(f"[hedgehog.jl_b200] allreduce callback failed: {e!r}", file=sys.stderr)
            return 1

    cb = abi.hh_allreduce_fn(_cb)
    comm = abi.hh_comm(cb, None, rank, world)
    return comm, cb


def connect_peers(engine, group=None):
    """Map every rank's LSM mailbox into every other rank's process (CUDA IPC handles exchanged with all_gather_object).
    After this, solve(::PricingProblem{American}, ::LSM) exchanges the per-date regression moments inside the pass
    kernel's tail over NVLink instead of calling NCCL 49 times (include/hedgehog_mc.h, hh_peer_*)."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world == 1:
        return None
    mine = engine.peer_export()
    handles = [None] * world
    dist.all_gather_object(handles, mine, group=group)
    engine.peer_connect(rank, world, handles)
    dist.barrier(group)  # nobody posts into a mailbox before everybody has mapped it
    return rank, world


def peer_comm(engine):
    """hh_comm selecting the in-kernel peer exchange (no callback)."""
    rank, world = engine.peers
    return abi.hh_comm(abi.hh_allreduce_fn(), None, rank, world)
