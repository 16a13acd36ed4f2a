"""Pinned-memory copy bandwidth of this box: H2D alone, D2H alone, both at once (the ceiling of bench.py's e2e leg)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    # under torchrun every rank probes its own GPU AT THE SAME TIME (a barrier lines the ranks up): what the host fabric gives
    # when all the GPUs of the box copy at once is the ceiling of the multi-GPU e2e leg
    from vfi_b200 import shard

    topo = shard.init_distributed()
    dev = torch.device("cuda", topo.local_rank)
    torch.cuda.set_device(dev)
    aff = shard.bind_to_gpu(topo.local_rank)
    n = 1 << 30
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device=dev)
    d_out = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    res = {}

    def run(h2d, d2h, reps=4, piece=n):
        if topo.world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
        for _ in range(reps):
            for o in range(0, n, piece):
                if h2d:
                    with torch.cuda.stream(s1):
                        d_in[o:o + piece].copy_(h_in[o:o + piece], non_blocking=True)
                if d2h:
                    with torch.cuda.stream(s2):
                        h_out[o:o + piece].copy_(d_out[o:o + piece], non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
        b.record(); torch.cuda.synchronize()
        return reps * n / (a.elapsed_time(b) * 1e-3) / 1e9

    run(True, True, 1)
    res["h2d_alone_GBps"] = run(True, False)
    res["d2h_alone_GBps"] = run(False, True)
    res["both_each_GBps"] = run(True, True)
    res["h2d_alone_8MB_pieces_GBps"] = run(True, False, piece=8 << 20)
    res["both_each_8MB_pieces_GBps"] = run(True, True, piece=8 << 20)
    res.update(rank=topo.rank, world=topo.world, cpu_affinity=aff)
    if topo.world > 1:
        allr = [None] * topo.world
        torch.distributed.all_gather_object(allr, res)
        if topo.is_root:
            for r in allr:
                print(json.dumps(r))
            print(json.dumps({"world": topo.world, "sum_h2d_alone_GBps": sum(r["h2d_alone_GBps"] for r in allr),
                              "sum_both_each_GBps": sum(r["both_each_GBps"] for r in allr)}))
        torch.distributed.destroy_process_group()
    else:
        print(json.dumps(res))


if __name__ == "__main__":
    main()
