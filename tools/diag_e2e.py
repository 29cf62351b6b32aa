#!/usr/bin/env python
"""Synchronous mppi_solve() time (host buffers in and out) with plain stream launches and with the CUDA graph, for a
few shapes and option settings.  Diagnostic only."""
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
    cases = [("diff_drive", 1 << 20, 100, 1), ("diff_drive", 1 << 17, 100, 1), ("full_body", 16384, 100, 1),
             ("steering", 4096, 50, 1), ("diff_drive", 1024, 50, 1024)]
    for model, K, T, R in cases:
        for fuse, pre in ((-1, -1), (0, 0), (0, 1), (1, 0), (1, 1)):
            row = {"model": model, "K": K, "T": T, "R": R, "fuse": fuse, "prefetch": pre}
            for graph in (False, True):
                ov = {"roll_off": False} if model == "full_body" else {}
                ctl = CONTROLLERS[model](launch=True, n_robots=R, device=0, horizon=T, num_samples=K, **ov)
                paths_, states = bench.synthetic_inputs(model, R, 200)
                for r in range(R):
                    ctl.set_path(paths_[r], robot=r)
                ctl.set_option(_capi.OPT_FUSE_CONTROLS, fuse)
                ctl.set_option(_capi.OPT_NOISE_PREFETCH, pre)
                if R >= 8:
                    ctl.set_option(_capi.OPT_UPLOAD_WARM_START, 0)
                ctl.use_graph(graph)
                U = bench.NUM_CONTROLS[model]
                for _ in range(10):
                    bench.plant_step(model, states, ctl.solve(states, 0.1).reshape(R, T - 1, U), 0.1)
                ts = []
                for _ in range(40):
                    t0 = time.perf_counter()
                    u = ctl.solve(states, 0.1)
                    ts.append(time.perf_counter() - t0)
                    bench.plant_step(model, states, u.reshape(R, T - 1, U), 0.1)
                row["graph_us" if graph else "stream_us"] = round(float(np.median(ts)) * 1e6, 1)
                row["fused"] = int(ctl.get_option(_capi.INFO_FUSED_CONTROLS))
                ctl.close()
            print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
