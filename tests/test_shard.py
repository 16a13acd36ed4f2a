"""Multi-GPU plumbing on CPU: shard arithmetic, and the N > 1 exchange over gloo with world_size 2."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vfi_b200 import shard


@pytest.mark.parametrize("mode", ["interleave", "contiguous"])
@pytest.mark.parametrize("n,world", [(0, 1), (1, 4), (7, 2), (256, 8), (257, 8), (5, 8)])
def test_every_pair_has_exactly_one_owner(n, world, mode):
    owned = [shard.shard_pairs(n, r, world, mode) for r in range(world)]
    flat = sorted(i for s in owned for i in s)
    assert flat == list(range(n))
    assert max(len(s) for s in owned) - min(len(s) for s in owned) <= 1
    if mode == "contiguous":
        for s in owned:
            assert s == list(range(s[0], s[0] + len(s))) if s else True


def test_batches_and_bad_arguments():
    assert list(shard.batches([0, 1, 2, 3, 4], 2)) == [[0, 1], [2, 3], [4]]
    with pytest.raises(ValueError):
        shard.shard_pairs(4, 2, 2)
    with pytest.raises(ValueError):
        shard.shard_pairs(4, 0, 2, mode="zigzag")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    topo = shard.init_distributed("gloo")
    assert (topo.rank, topo.world) == (rank, world)
    torch.manual_seed(0)                      # identical replicas
    model = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3, padding=1), torch.nn.Conv2d(4, 2, 3, padding=1))
    bucket = shard.GradBucket(model.parameters())
    # each rank trains on its own shard of 6 "frame pairs"
    data = torch.arange(6 * 3 * 5 * 5, dtype=torch.float32).reshape(6, 3, 5, 5) / 100.0
    mine = shard.shard_pairs(6, rank, world)
    bucket.zero()
    bucket.attach()
    model(data[mine]).square().sum().backward()
    bucket.attach()                           # grads produced by autograd are folded into the flat bucket
    local = bucket.flat.clone()
    h = bucket.allreduce_mean(async_op=True)
    bucket.finish(h)
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    expect = sum(gathered) / world
    ok = torch.allclose(bucket.flat, expect, rtol=1e-6, atol=1e-7)
    same_ptr = all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(bucket.params, bucket.views))
    merged = shard.gather_by_index({i: float(data[i].sum()) for i in mine}, world)
    if rank == 0:
        out.put((ok, same_ptr, sorted(merged), bucket.numel))
    dist.destroy_process_group()


def test_world_size_2_gradient_exchange_and_gather_over_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, same_ptr, keys, numel = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok and same_ptr and keys == list(range(6))
    assert numel == 3 * 4 * 9 + 4 + 4 * 2 * 9 + 2


def _overlap_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    shard.init_distributed("gloo")
    torch.manual_seed(0)                      # identical replicas
    blocks = torch.nn.ModuleList([torch.nn.Conv2d(3, 3, 3, padding=1) for _ in range(3)])
    groups = [list(b.parameters()) for b in reversed(blocks)]          # the last block's gradients are final first
    bucket = shard.GradBucket(blocks.parameters(), groups=groups)
    bucket.zero()
    bucket.attach()
    bucket.enable_overlap()
    data = torch.arange(8 * 3 * 6 * 6, dtype=torch.float32).reshape(8, 3, 6, 6) / 300.0
    mine = shard.shard_pairs(8, rank, world, "contiguous")
    results = []
    for step in range(2):                     # twice: the counters must reset
        bucket.zero()
        bucket.begin_step()
        x = data[mine]
        for b in blocks:
            x = b(x)
        x.square().sum().backward()
        exposed = bucket.finish_overlap()
        results.append((bucket.flat.clone(), list(bucket.launched), exposed))
    # reference: every rank's local gradient, averaged
    ref_blocks = torch.nn.ModuleList([torch.nn.Conv2d(3, 3, 3, padding=1) for _ in range(3)])
    ref_blocks.load_state_dict(blocks.state_dict())
    x = data[mine]
    for b in ref_blocks:
        x = b(x)
    x.square().sum().backward()
    local = torch.cat([p.grad.flatten() for g in reversed(ref_blocks) for p in g.parameters()])
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    expect = sum(gathered) / world
    ok = all(torch.allclose(r[0], expect, rtol=1e-5, atol=1e-6) for r in results)
    same_ptr = all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(bucket.params, bucket.views))
    if rank == 0:
        out.put((ok, same_ptr, [r[1] for r in results], bucket.segments))
    bucket.disable_overlap()
    dist.destroy_process_group()


def test_world_size_2_overlapped_group_exchange_over_gloo():
    """The all-reduce of a group is launched from the autograd hook of its last gradient -- block 3 first, block 1 last -- and
    the bucket holds the mean of the ranks' gradients afterwards, step after step."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_overlap_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, same_ptr, launched, segments = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok and same_ptr
    assert launched == [[0, 1, 2], [0, 1, 2]]           # group 0 = the last block
    assert segments == [(0, 84), (84, 168), (168, 252)]


def test_grad_bucket_groups_must_partition_the_parameters():
    m = torch.nn.ModuleList([torch.nn.Linear(2, 2), torch.nn.Linear(2, 2)])
    with pytest.raises(ValueError, match="partition"):
        shard.GradBucket(m.parameters(), groups=[list(m[0].parameters())])
    b = shard.GradBucket(m.parameters())
    with pytest.raises(ValueError, match="groups"):
        b.enable_overlap()
