#!/usr/bin/env python
"""Back-to-back solve time with the internal generator (noise prefetch) and with external noise (no generator at all):
the difference is what the concurrently running generator costs the solve.  Diagnostic only."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    from ccv_mppi_path_tracker_b200 import CONTROLLERS, _capi
    torch.cuda.set_device(0)
    for K in (1 << 17, 1 << 16, 1 << 18):
        model, T, R, U = "diff_drive", 100, 1, 2
        for ext in (False, True):
            ctl = CONTROLLERS[model](launch=True, n_robots=R, device=0, horizon=T, num_samples=K)
            paths_, states = bench.synthetic_inputs(model, R, 200)
            ctl.set_path(paths_[0], robot=0)
            stream = torch.cuda.Stream(priority=-1)
            torch.cuda.set_stream(stream)
            ctl.set_stream(stream.cuda_stream)
            for _ in range(25):
                bench.plant_step(model, states, ctl.solve(states, 0.1).reshape(R, T - 1, U), 0.1)
            if ext:
                ctl.set_noise(np.random.default_rng(1).standard_normal((R, T - 1, K, U), dtype=np.float32))
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
            print(json.dumps({"K": K, "external_noise": ext, "ms": round(best, 5), "launches": ctl.launch_count()}), flush=True)
            ctl.close()


if __name__ == "__main__":
    main()
