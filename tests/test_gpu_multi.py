"""Two-GPU test of the sample-sharded solve (one process per GPU): the NVLink peer exchange and the NCCL all-gather
give bit-identical controls on both ranks, equal (to FP32 summation order) to the unsharded solve of the same
global sample set.  Needs >= 2 CUDA devices (run with `gpurun --gpus 2`); skipped otherwise."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from ccv_mppi_path_tracker_b200 import CONTROLLERS, _capi, comm_unique_id
    from common import make_case
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    K, T = 4096, 40
    case = make_case("steering", K, T, seed=31)
    res = {}
    for mode in ("p2p", "p2p_graph", "p2p_fused", "nccl", "nccl_fused"):
        ctl = CONTROLLERS["steering"](launch=True, device=rank, horizon=T, num_samples=K // world)
        ctl.set_path(case["path"])
        ctl.set_seed(77, 0)
        ctl.set_shard(rank * (K // world), K, 0)
        ctl.use_graph(mode == "p2p_graph")  # the peer exchange replays from a CUDA graph; NCCL keeps stream launches
        ctl.set_option(_capi.OPT_FUSE_CONTROLS, 1 if mode.endswith("fused") else 0)
        if mode.startswith("p2p"):
            t = torch.frombuffer(bytearray(ctl.comm_export(world)), dtype=torch.uint8).cuda()
            allh = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allh, t)
            ctl.comm_connect(b"".join(bytes(x.cpu().numpy().tobytes()) for x in allh), rank, world)
        else:
            idt = torch.zeros(_capi.COMM_ID_BYTES, dtype=torch.uint8, device="cuda")
            if rank == 0:
                idt.copy_(torch.frombuffer(bytearray(comm_unique_id()), dtype=torch.uint8))
            dist.broadcast(idt, 0)
            ctl.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)
        dist.barrier()
        us = [ctl.solve(case["state"], case["dt"]).copy() for _ in range(4)]
        res[mode] = np.stack(us)
        dist.barrier()
        ctl.close()
    assert np.array_equal(res["p2p"], res["nccl"])
    assert np.array_equal(res["p2p"], res["p2p_graph"])
    assert np.array_equal(res["p2p_fused"], res["nccl_fused"])
    rng_ = np.array(case["sp"]["u_max"][:3]) - np.array(case["sp"]["u_min"][:3])
    assert (np.abs(res["p2p"] - res["p2p_fused"]) / rng_).max() < 5e-4  # per-CTA vs per-chunk summation order
    # the general form of the exchange (several robots per handle; a record of more than 2 x 256 columns): every word
    # polled in batches over the ranks, merge from the buffer -- against the NCCL all-gather + merge kernel, bit for bit
    for model, Kg, Tg, Rg in (("steering", 2048, 40, 3), ("full_body", 1024, 120, 1)):
        cg = make_case(model, Kg, Tg, seed=33)
        states = np.tile(cg["state"], (Rg, 1))
        states[:, 0] += 0.04 * np.arange(Rg)
        got = {}
        for mode in ("p2p", "nccl"):
            ctl = CONTROLLERS[model](launch=True, n_robots=Rg, device=rank, horizon=Tg, num_samples=Kg // world,
                                     **cg["overrides"])
            for r in range(Rg):
                ctl.set_path(cg["path"], robot=r)
            ctl.set_seed(78, 0)
            ctl.set_shard(rank * (Kg // world), Kg, 0)
            if mode == "p2p":
                t = torch.frombuffer(bytearray(ctl.comm_export(world)), dtype=torch.uint8).cuda()
                allh = [torch.zeros_like(t) for _ in range(world)]
                dist.all_gather(allh, t)
                ctl.comm_connect(b"".join(bytes(x.cpu().numpy().tobytes()) for x in allh), rank, world)
            else:
                idt = torch.zeros(_capi.COMM_ID_BYTES, dtype=torch.uint8, device="cuda")
                if rank == 0:
                    idt.copy_(torch.frombuffer(bytearray(comm_unique_id()), dtype=torch.uint8))
                dist.broadcast(idt, 0)
                ctl.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)
            dist.barrier()
            got[mode] = np.stack([np.array(ctl.solve(states, cg["dt"]), copy=True) for _ in range(3)])
            dist.barrier()
            ctl.close()
        assert np.array_equal(got["p2p"], got["nccl"]), (model, Rg)
        mine = torch.from_numpy(got["p2p"].copy()).cuda()
        both = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(both, mine)
        assert torch.equal(both[0], both[1]), (model, Rg)  # every rank holds the same controls
    # a peer that stops solving: the other rank's solve fails after MPPI_OPT_EXCHANGE_TIMEOUT_MS instead of merging
    # garbage, and its controls / warm start keep their previous values
    ctl = CONTROLLERS["steering"](launch=True, device=rank, horizon=T, num_samples=K // world)
    ctl.set_path(case["path"])
    ctl.set_seed(77, 0)
    ctl.set_shard(rank * (K // world), K, 0)
    ctl.set_option(_capi.OPT_EXCHANGE_TIMEOUT_MS, 50)
    t = torch.frombuffer(bytearray(ctl.comm_export(world)), dtype=torch.uint8).cuda()
    allh = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allh, t)
    ctl.comm_connect(b"".join(bytes(x.cpu().numpy().tobytes()) for x in allh), rank, world)
    dist.barrier()
    u_ok = ctl.solve(case["state"], case["dt"]).copy()
    dist.barrier()
    if rank == 0:
        try:
            ctl.solve(case["state"], case["dt"])
            timed_out = False
        except _capi.MppiError as e:
            timed_out = e.code == _capi.MPPI_ERR_NCCL and "timed out" in str(e)
        assert timed_out, "rank 0 solved alone and did not report the missing peer"
        assert np.array_equal(ctl.optimal_solution[0], u_ok)  # controls untouched by the failed solve
    dist.barrier()
    ctl.close()
    np.save(os.path.join(out_dir, f"u_{rank}.npy"), res["p2p"])
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_peer_exchange_equals_nccl_and_unsharded(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from ccv_mppi_path_tracker_b200 import CONTROLLERS
    from common import make_case
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    u0, u1 = np.load(tmp_path / "u_0.npy"), np.load(tmp_path / "u_1.npy")
    assert np.array_equal(u0, u1)  # rank-ordered merge: same bits on every rank
    K, T = 4096, 40
    case = make_case("steering", K, T, seed=31)
    ctl = CONTROLLERS["steering"](launch=True, device=0, horizon=T, num_samples=K)
    ctl.set_path(case["path"])
    ctl.set_seed(77, 0)
    full = np.stack([ctl.solve(case["state"], case["dt"]).copy() for _ in range(4)])
    ctl.close()
    rng_ = np.array(case["sp"]["u_max"][:3]) - np.array(case["sp"]["u_min"][:3])
    assert (np.abs(u0 - full) / rng_).max() < 1e-4
