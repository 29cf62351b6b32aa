"""world_size-2 test of the sample-sharded solve's host logic on CPU (gloo): every rank owns half of the samples of
ONE solve, builds its partial record {c_min, sum w, sum w^2, -, sum w*u} (here from the FP64 oracle's costs and
clamped controls -- on the GPU box the kernels produce it), the records are all-gathered (the one exchange step of
the path, ncclAllGather on the B200s) and merged with mppi_merge_partials.  Both ranks must obtain the unsharded
oracle's control sequence."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle
    from ccv_mppi_path_tracker_b200 import merge_partials
    from common import make_case
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    K, T = 2048, 30
    case = make_case("steering", K, T, seed=17)  # every rank builds the same global inputs
    U, lam = case["U"], case["sp"]["lambda_"]
    full = oracle.solve("steering", case["sp"], K, T, case["state"], case["dt"], case["path"], case["eps"], case["u0"],
                        want=("cost", "controls", "stats"))
    Kg = K // world
    sl = slice(rank * Kg, (rank + 1) * Kg)
    # this rank's shard only: its own slice of the noise tensor (disjoint Philox sub-stream on the GPU)
    mine = oracle.solve("steering", case["sp"], Kg, T, case["state"], case["dt"], case["path"], case["eps"][:, sl],
                        case["u0"], want=("cost", "controls"))
    assert np.array_equal(mine["cost"], full["cost"][sl])
    m = mine["cost"].min()
    w = np.exp(-(mine["cost"] - m) / lam)
    rec = np.concatenate([[m, w.sum(), (w ** 2).sum(), 0.0], (w[:, None] * mine["controls"].reshape(Kg, -1)).sum(0)])
    rec = torch.tensor(rec, dtype=torch.float32)
    gathered = [torch.zeros_like(rec) for _ in range(world)]
    dist.all_gather(gathered, rec)  # the single exchange step
    u, st = merge_partials(torch.stack(gathered).numpy(), lam)
    rng_ = np.array(case["sp"]["u_max"][:U]) - np.array(case["sp"]["u_min"][:U])
    err = np.abs(u.reshape(T - 1, U) - full["u_new"]) / rng_
    np.save(os.path.join(out_dir, f"u_{rank}.npy"), u)
    assert err.max() < 1e-5, err.max()
    assert abs(st["c_min"] - full["stats"][0]) < 1e-4 and abs(st["ess"] - full["stats"][2]) < 1e-3 * full["stats"][2]
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sample_shards_merge_to_the_unsharded_solve(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    u0, u1 = np.load(tmp_path / "u_0.npy"), np.load(tmp_path / "u_1.npy")
    assert np.array_equal(u0, u1)  # rank-ordered merge: every rank computes the same bits
