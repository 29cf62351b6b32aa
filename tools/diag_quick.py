#!/usr/bin/env python
"""Back-to-back device-resident solve time (bench.py's stationary timed region, default options) for a few shapes.
Diagnostic only."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    from ccv_mppi_path_tracker_b200 import CONTROLLERS, _capi
    torch.cuda.set_device(0)
    cases = [("diff_drive", 1 << 20, 100, 1), ("diff_drive", 1 << 19, 100, 1), ("diff_drive", 1 << 17, 100, 1),
             ("diff_drive", 1 << 16, 100, 1), ("full_body", 16384, 100, 1), ("steering", 4096, 50, 1),
             ("diff_drive", 1024, 50, 1024)]
    only = os.environ.get("DIAG_CASES")  # e.g. "0,2,6": indices into the list above
    if only:
        cases = [cases[int(i)] for i in only.split(",")]
    opts = [a.split("=") for a in sys.argv[1:]]
    for model, K, T, R in cases:
        U = bench.NUM_CONTROLS[model]
        ov = {"roll_off": False} if model == "full_body" else {}
        ctl = CONTROLLERS[model](launch=True, n_robots=R, device=0, horizon=T, num_samples=K, **ov)
        for k, v in opts:
            ctl.set_option(getattr(_capi, k), float(v))
        paths_, states = bench.synthetic_inputs(model, R, 200)
        for r in range(R):
            ctl.set_path(paths_[r], robot=r)
        stream = torch.cuda.Stream(priority=-1)
        torch.cuda.set_stream(stream)
        ctl.set_stream(stream.cuda_stream)
        for _ in range(25):
            bench.plant_step(model, states, ctl.solve(states, 0.1).reshape(R, T - 1, U), 0.1)
        ctl.set_option(_capi.OPT_FEEDBACK_WARM_START, 0)
        ctl.upload(states, 0.1, with_nominal=True)
        best = 1e9
        for rep in range(3):
            for _ in range(5):
                ctl.enqueue()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(30):
                ctl.enqueue()
            e1.record(stream)
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 30)
        km = {k: round(v * 1e3, 1) for k, v in ctl.time_kernels(5).items()}
        print(json.dumps({"model": model, "K": K, "T": T, "R": R, "ms": round(best, 5),
                          "steps_per_s": f"{K * R * (T - 1) / best * 1e3:.4g}", "kernels_us": km}), flush=True)
        ctl.close()


if __name__ == "__main__":
    main()
