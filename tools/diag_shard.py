#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/diag_shard.py : per-rank sample counts 2^17 / 2^18 of the diff-drive T=100 solve,
sample-sharded over the N ranks (peer exchange and NCCL) and, beside it, every rank solving the same K unsharded: the
difference is what the exchange (and the ranks waiting for each other) costs.  Diagnostic only."""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import bench
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    D, D1 = bench.Dist(world), bench.Dist(1)
    args = types.SimpleNamespace()
    name = "diff_drive_K1M_T100"
    for e in (17, 18):
        K = 1 << e
        row = {"K_per_gpu": K, "world": world}
        for label, shard, xc in (("unsharded", "none", "p2p"), ("p2p", "samples", "p2p"), ("nccl", "samples", "nccl")):
            DD = D1 if shard == "none" else D
            ctl, res, _ = bench.measure(name, K, shard, args, DD, rank, local, 50, 5, None, exchange=xc, kernel_iters=3)
            ctl.close()
            D.barrier()
            row[label + "_ms"] = round(D.max(res["ms_per_step"]), 5)
            row[label + "_e2e_ms"] = round(D.max(res["e2e"]["ms_per_step"]), 5)
        if rank == 0:
            print(json.dumps(row), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
