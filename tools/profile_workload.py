#!/usr/bin/env python
"""Drive one bench.py workload for a profiler: closed-loop warm-up, then N back-to-back device-resident solves (the
timed region of bench.py, plain stream launches).  Run it under ncu:

  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
      python tools/profile_workload.py --workload diff_drive_K1M_T100
  ncu --set full --clock-control none --import-source on -k regex:rollout_cost -s 30 -c 2 -o gpurun_out/prof \
      python tools/profile_workload.py --workload diff_drive_K1M_T100

Every solve launches K2 (rollout_cost_*) exactly once, so `-s 30` skips the warm-up (25 closed-loop + 5 resident)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    from ccv_mppi_path_tracker_b200 import CONTROLLERS, _capi
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="diff_drive_K1M_T100", choices=sorted(bench.WORKLOADS))
    ap.add_argument("--K", type=int, default=0, help="override the workload's sample count (strong-scaling shards)")
    ap.add_argument("--solves", type=int, default=8)
    ap.add_argument("--sync", action="store_true", help="synchronous mppi_solve() calls instead of back-to-back enqueues")
    args = ap.parse_args()
    model, K, T, R = bench.WORKLOADS[args.workload]
    K = args.K or K
    U = bench.NUM_CONTROLS[model]
    torch.cuda.set_device(0)
    ov = {"roll_off": False} if model == "full_body" else {}
    ctl = CONTROLLERS[model](launch=True, n_robots=R, device=0, horizon=T, num_samples=K, **ov)
    paths_, states = bench.synthetic_inputs(model, R, bench.CLOSED_LOOP_WARMUP + args.solves + 60)
    for r in range(R):
        ctl.set_path(paths_[r], robot=r)
    ctl.set_seed(0x5EED0000 + 4, 0)
    stream = torch.cuda.Stream(priority=-1)
    torch.cuda.set_stream(stream)
    ctl.set_stream(stream.cuda_stream)
    for _ in range(bench.CLOSED_LOOP_WARMUP):
        bench.plant_step(model, states, ctl.solve(states, 0.1).reshape(R, T - 1, U), 0.1)
    if args.sync:
        for _ in range(5 + args.solves):
            bench.plant_step(model, states, ctl.solve(states, 0.1).reshape(R, T - 1, U), 0.1)
    else:
        ctl.set_option(_capi.OPT_FEEDBACK_WARM_START, 0)
        ctl.upload(states, 0.1, with_nominal=True)
        for _ in range(5 + args.solves):
            ctl.enqueue()
        ctl.synchronize()
    print("fused_controls", int(ctl.get_option(_capi.INFO_FUSED_CONTROLS)), "launches/solve", ctl.launch_count())
    ctl.close()


if __name__ == "__main__":
    main()
