#!/usr/bin/env python
"""Where a synchronous solve spends its time on the host clock: mppi_upload / mppi_enqueue + mppi_synchronize /
mppi_download timed separately (medians), beside the one-call mppi_solve().  Diagnostic only."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from ccv_mppi_path_tracker_b200 import CONTROLLERS, _capi
    import bench
    torch.cuda.set_device(0)
    cases = [("diff_drive", 1024, 50, 1024), ("diff_drive", 1 << 20, 100, 1), ("diff_drive", 1 << 17, 100, 1)]
    only = os.environ.get("DIAG_CASES")
    if only:
        cases = [cases[int(i)] for i in only.split(",")]
    opts = [a.split("=") for a in sys.argv[1:]]
    for model, K, T, R in cases:
        ov = {"roll_off": False} if model == "full_body" else {}
        ctl = CONTROLLERS[model](launch=True, n_robots=R, device=0, horizon=T, num_samples=K, **ov)
        for k, v in opts:
            ctl.set_option(getattr(_capi, k), float(v))
        paths_, states = bench.synthetic_inputs(model, R, 200)
        for r in range(R):
            ctl.set_path(paths_[r], robot=r)
        if R >= 8:
            ctl.set_option(_capi.OPT_UPLOAD_WARM_START, 0)
        U = bench.NUM_CONTROLS[model]
        stream = torch.cuda.Stream(priority=-1)
        torch.cuda.set_stream(stream)
        ctl.set_stream(stream.cuda_stream)
        for _ in range(10):
            bench.plant_step(model, states, ctl.solve(states, 0.1).reshape(R, T - 1, U), 0.1)
        t_solve, t_up, t_run, t_down = [], [], [], []
        for _ in range(30):
            t0 = time.perf_counter()
            u = ctl.solve(states, 0.1)
            t_solve.append(time.perf_counter() - t0)
            bench.plant_step(model, states, u.reshape(R, T - 1, U), 0.1)
        for _ in range(30):
            t0 = time.perf_counter()
            ctl.upload(states, 0.1, with_nominal=R < 8)
            t1 = time.perf_counter()
            ctl.enqueue()
            ctl.synchronize()
            t2 = time.perf_counter()
            u = ctl.download()
            t3 = time.perf_counter()
            t_up.append(t1 - t0)
            t_run.append(t2 - t1)
            t_down.append(t3 - t2)
            bench.plant_step(model, states, u.reshape(R, T - 1, U), 0.1)
        med = lambda a: round(float(np.median(a)) * 1e6, 1)
        km = {k: round(v * 1e3, 1) for k, v in ctl.time_kernels(5).items()}
        print(json.dumps({"model": model, "K": K, "T": T, "R": R, "solve_us": med(t_solve), "upload_us": med(t_up),
                          "enqueue_sync_us": med(t_run), "download_us": med(t_down), "io_bytes": ctl.io_bytes(),
                          "kernels_us": km}), flush=True)
        ctl.close()


if __name__ == "__main__":
    main()
