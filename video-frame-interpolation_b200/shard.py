"""Multi-GPU plumbing: one process per GPU, independent frame pairs per rank.

Inference has no data-path collective: frame pair ``i`` belongs to rank ``i % world`` (or a contiguous chunk), every
rank holds the full 5.7 MB of weights, the host gathers outputs by index.  Training adds exactly one exchange: a
flat-bucket gradient all-reduce after backward (between /root/reference/train.py:125 and :128), averaged before
gradient clipping so clipping sees the same values as a single-GPU run.

Everything here works with the ``gloo`` backend on CPU tensors too, which is how tests/test_shard.py covers the
N > 1 path without GPUs.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Iterable, List, Sequence

import torch
import torch.distributed as dist


@dataclass(frozen=True)
class Topology:
    rank: int
    world: int
    local_rank: int

    @property
    def is_root(self) -> bool:
        return self.rank == 0


def topology_from_env() -> Topology:
    return Topology(int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)),
                    int(os.environ.get("LOCAL_RANK", 0)))


def init_distributed(backend: str | None = None) -> Topology:
    """Initialise torch.distributed from the torchrun environment (no-op for a single process)."""
    topo = topology_from_env()
    if topo.world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(topo.local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", topo.local_rank))
        else:
            dist.init_process_group(backend)
    return topo


def shard_pairs(num_pairs: int, rank: int, world: int, mode: str = "interleave") -> List[int]:
    """Indices of the frame pairs ``(f_i, f_{i+1})`` this rank processes.

    ``interleave``: i % world == rank (balanced to within one pair, streams in decode order);
    ``contiguous``: equal chunks, remainder spread over the first ranks (keeps each rank's reads sequential).
    Every index in ``range(num_pairs)`` is owned by exactly one rank; empty shards are allowed.
    """
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    if num_pairs < 0:
        raise ValueError("num_pairs must be >= 0")
    if mode == "interleave":
        return list(range(rank, num_pairs, world))
    if mode == "contiguous":
        base, rem = divmod(num_pairs, world)
        start = rank * base + min(rank, rem)
        return list(range(start, start + base + (1 if rank < rem else 0)))
    raise ValueError(f"unknown mode {mode!r}")


def batches(indices: Sequence[int], batch: int) -> Iterable[List[int]]:
    for i in range(0, len(indices), batch):
        yield list(indices[i:i + batch])


def gather_by_index(local: dict, world: int) -> dict:
    """Host-side merge of per-rank ``{pair_index: result}`` dicts (results must be picklable / CPU tensors)."""
    if world == 1 or not dist.is_initialized():
        return dict(local)
    parts = [None] * world
    dist.all_gather_object(parts, local)
    merged = {}
    for p in parts:
        merged.update(p)
    return merged


class GradBucket:
    """One flat fp32 bucket holding every parameter gradient (1,430,045 params = 5.72 MB for EMA_VFI).

    ``attach`` re-points each ``param.grad`` into the bucket, so kernels that accumulate into ``.grad`` (including
    vfi_dcn_bwd_weight's atomics into grad_weight / grad_bias) write straight into it and the all-reduce needs no
    gather copy.  ``allreduce_mean`` is the single collective of the training path.
    """

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.views = []
        off = 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    def attach(self) -> None:
        for p, v in zip(self.params, self.views):
            if p.grad is not None and p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
            p.grad = v

    def zero(self) -> None:
        self.flat.zero_()

    def allreduce_mean(self, async_op: bool = False):
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return None
        world = dist.get_world_size()
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=async_op)
        if async_op:
            return work, world
        self.flat.div_(world)
        return None

    def finish(self, handle) -> None:
        if handle is None:
            return
        work, world = handle
        work.wait()
        self.flat.div_(world)
