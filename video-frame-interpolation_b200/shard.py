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
from typing import Iterable, List, Optional, Sequence, Tuple

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


def bind_to_gpu(local_rank: int) -> dict:
    """Pin this process to the CPU cores (and thereby the NUMA node) the driver reports as local to GPU ``local_rank`` --
    BEFORE pinned host buffers are allocated, so that they land in that node's memory and the H2D / D2H DMA of the ranks of
    one box do not all cross the same memory controller.  Uses NVML's ideal CPU affinity; a box that attaches every GPU to
    one node (seen on this pool: CPU affinity 0-31 for all eight GPUs) gets the cores of that node split evenly between the
    ranks instead, which at least keeps the staging threads off each other's cores.  Returns what was done (for the bench line).
    Never raises: affinity is an optimisation."""
    info = {"bound": False}
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1 and 64 * w + b < ncpu]
        try:
            bus = pynvml.nvmlDeviceGetPciInfo(h).busId
            bus = bus.decode() if isinstance(bus, bytes) else bus
            with open(f"/sys/bus/pci/devices/{bus[-12:].lower()}/numa_node") as f:
                info["numa_node"] = int(f.read().strip())
        except Exception:
            info["numa_node"] = None
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0))) or sorted(os.sched_getaffinity(0))
        n_local = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))
        same = [i for i in range(pynvml.nvmlDeviceGetCount())
                if list(pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(i), (ncpu + 63) // 64)) == list(words)]
        if len(same) > 1 and n_local > 1 and len(allowed) >= len(same):
            k = same.index(local_rank) if local_rank in same else 0
            per = len(allowed) // len(same)
            allowed = allowed[k * per:(k + 1) * per]
        os.sched_setaffinity(0, allowed)
        info.update(bound=True, cpus=len(allowed), first_cpu=allowed[0], gpus_sharing_affinity=len(same))
    except Exception as exc:  # noqa: BLE001
        info["error"] = str(exc)[:100]
    return info


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
    gather copy.  ``allreduce_mean`` is the single collective of the training path (/root/reference/train.py:125-128: it sits
    between ``loss.backward()`` and ``clip_grad_norm_``).

    Overlap: with ``groups`` -- lists of parameters in the order their gradients become FINAL during backward (the last
    block first) -- the bucket is laid out group by group, and ``enable_overlap()`` registers post-accumulate hooks that
    launch the all-reduce of a group's segment the moment its last gradient has been accumulated, while autograd is still
    working on the earlier layers.  Only the segment of the group that finishes last is exposed; ``finish_overlap()`` waits for
    all segments, divides by the world size and returns how long the wait after the end of backward was.
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], groups: Optional[Sequence[Sequence[torch.nn.Parameter]]] = None):
        params = [p for p in params if p.requires_grad]
        if not params:
            raise ValueError("no trainable parameters")
        if groups is not None:
            ordered = [p for g in groups for p in g if p.requires_grad]
            if sorted(map(id, ordered)) != sorted(map(id, params)):
                raise ValueError("groups must partition the trainable parameters")
            params = ordered
        self.params = params
        dev = self.params[0].device
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.views = []
        off = 0
        self._offset = {}
        self._index = {id(p): i for i, p in enumerate(self.params)}
        for p in self.params:
            self._offset[id(p)] = off
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self.segments: List[Tuple[int, int]] = []          # (start, end) of each group in the flat bucket
        self._group_of = {}
        if groups is not None:
            for gi, g in enumerate(groups):
                ps = [p for p in g if p.requires_grad]
                lo = self._offset[id(ps[0])]
                self.segments.append((lo, lo + sum(p.numel() for p in ps)))
                for p in ps:
                    self._group_of[id(p)] = gi
        self._pending: List[int] = []
        self._works: list = []
        self._hooks: list = []
        self.launched: List[int] = []                      # order in which the group all-reduces were launched (last step)

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

    # ---------------------------------------------------------------------------------------------- overlapped exchange
    def enable_overlap(self) -> None:
        """Register the hooks (once).  Requires ``groups`` and attached gradients (``attach()`` after the first backward, or
        ``zero(); attach()`` before it), so that autograd accumulates in place into the bucket."""
        if not self.segments:
            raise ValueError("enable_overlap needs a GradBucket built with groups=")
        if self._hooks:
            return
        for p in self.params:
            self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    def disable_overlap(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []

    def begin_step(self) -> None:
        """Call before every backward: resets the per-group counters."""
        counts = [0] * len(self.segments)
        for p in self.params:
            counts[self._group_of[id(p)]] += 1
        self._pending = counts
        self._works = []
        self.launched = []

    def _on_grad(self, p: torch.nn.Parameter) -> None:
        gi = self._group_of[id(p)]
        v = self.views[self._index[id(p)]]
        if p.grad is not None and p.grad.data_ptr() != v.data_ptr():
            # autograd replaced .grad (first backward before attach): fold it into the bucket and re-point
            v.copy_(p.grad)
            p.grad = v
        self._pending[gi] -= 1
        if self._pending[gi] == 0:
            self.launched.append(gi)
            if dist.is_initialized() and dist.get_world_size() > 1:
                lo, hi = self.segments[gi]
                # NCCL orders the collective after everything queued on the current stream so far (this group's kernels)
                self._works.append(dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, async_op=True))

    def finish_overlap(self, measure: bool = False, average: bool = True) -> float:
        """Wait for the group all-reduces and (``average``) turn sums into means.  With ``measure`` (CUDA only) returns the exposed time in
        milliseconds: from the end of backward on the compute stream to the completion of the last all-reduce."""
        if any(self._pending):
            raise RuntimeError(f"finish_overlap: gradients still missing for groups {[i for i, c in enumerate(self._pending) if c]}")
        world = dist.get_world_size() if dist.is_initialized() else 1
        cuda = self.flat.is_cuda and measure
        if cuda:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        for w in self._works:
            w.wait()                                        # makes the current stream wait for the collective
        if cuda:
            e1.record()
        if world > 1 and average:
            self.flat.div_(world)
        self._works = []
        if cuda:
            e1.synchronize()
            return float(e0.elapsed_time(e1))
        return 0.0
