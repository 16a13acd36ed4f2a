"""Multi-GPU behaviour on real devices (needs >= 2 GPUs; skipped otherwise): one process per GPU over NCCL.

* inference: frame pairs sharded across ranks (no collective) give bit-identical results to one GPU doing them all;
* training: per-rank backward + ONE flat-bucket NCCL all-reduce reproduces the single-GPU full-batch gradient.
"""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist

    import vfi_b200
    from oracle import torch_ref
    from vfi_b200 import shard

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    topo = shard.init_distributed("nccl")
    dev = torch.device("cuda", rank)
    torch.backends.cudnn.allow_tf32 = False               # the stock offset_conv must not add TF32 noise to the comparison
    g = torch.Generator().manual_seed(0)
    N = 8                                                    # frame pairs / samples
    frames = torch.randn(N, 3, 48, 64, generator=g)
    flows = 3 * torch.randn(N, 2, 48, 64, generator=g)
    xs = torch.randn(N, 67, 24, 32, generator=g)

    # ---- inference sharding: no collective, host gathers by index
    mine = shard.shard_pairs(N, rank, world)
    local = {i: vfi_b200.warp(frames[i:i + 1].to(dev), flows[i:i + 1].to(dev)).cpu() for i in mine}
    merged = shard.gather_by_index(local, world)

    # ---- training: identical replicas, each on its shard; one flat-bucket all-reduce
    torch.manual_seed(1)
    blk = torch_ref.FusionBlock(67).to(dev)
    with torch.no_grad():
        blk.offset_conv.weight.normal_(0, 0.02)
        blk.offset_conv.bias.normal_(0, 0.5)
    vfi_b200.install()
    bucket = shard.GradBucket(blk.parameters())
    bucket.zero()
    bucket.attach()
    # loss = mean over the GLOBAL batch -> each rank contributes sum over its shard / N, summed (not averaged) below
    (blk(xs[mine].to(dev)).square().sum() / N).backward()
    bucket.attach()
    dist.all_reduce(bucket.flat, op=dist.ReduceOp.SUM)
    torch.cuda.synchronize(dev)

    if rank == 0:
        single = torch.cat([vfi_b200.warp(frames[i:i + 1].to(dev), flows[i:i + 1].to(dev)).cpu() for i in range(N)])
        same_inference = all(torch.equal(merged[i], single[i:i + 1]) for i in range(N)) and sorted(merged) == list(range(N))
        blk.zero_grad(set_to_none=True)
        (blk(xs.to(dev)).square().sum() / N).backward()
        ref = torch.cat([p.grad.flatten() for p in blk.parameters()])
        err = float((bucket.flat - ref).abs().max() / ref.abs().max())
        q.put((same_inference, err))
    vfi_b200.uninstall()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_sharded_inference_and_gradient_allreduce():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    same_inference, err = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert same_inference
    assert err <= 2e-4     # fp32 atomics + a different batch split: summation order differs, nothing else
