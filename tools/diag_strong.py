#!/usr/bin/env python
"""Device time of the diff-drive T=100 solve for the per-GPU sample counts of the strong-scaling runs
(K_global = 2^20 over 1/2/4/8 GPUs -> K per GPU = 2^20 .. 2^17), one GPU, no exchange, for the combinations of
MPPI_OPT_FUSE_CONTROLS x MPPI_OPT_NOISE_PREFETCH (x grid-builder lanes): back-to-back enqueues with and without the
CUDA graph.  One JSON line per configuration.  Diagnostic only (not a bench value)."""
import itertools
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from ccv_mppi_path_tracker_b200 import CONTROLLERS, _capi, params, paths
    torch.cuda.set_device(0)
    model = sys.argv[1] if len(sys.argv) > 1 else "diff_drive"
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    exps = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [14, 16, 17, 18, 19, 20]
    path = paths.sin_path(**params.LAUNCH_PATH[model])
    S = params.NUM_STATES[model]
    for e in exps:
        K = 1 << e
        for fuse, pre, lanes in itertools.product((0, 1), (0, 1), (0, 16)):
            if lanes and not (fuse and pre):
                continue
            ov = {"roll_off": False} if model == "full_body" else {}
            ctl = CONTROLLERS[model](launch=True, device=0, horizon=T, num_samples=K, **ov)
            ctl.set_path(path)
            ctl.set_seed(0x5EED0000 + e, 0)
            ctl.set_option(_capi.OPT_FUSE_CONTROLS, fuse)
            ctl.set_option(_capi.OPT_NOISE_PREFETCH, pre)
            ctl.set_option(_capi.OPT_GRID_LANES, lanes)
            stream = torch.cuda.Stream(priority=-1)
            torch.cuda.set_stream(stream)
            ctl.set_stream(stream.cuda_stream)
            ctl.upload(np.zeros(S), 0.1, with_nominal=True)
            row = {"K": K, "fuse": fuse, "prefetch": pre, "lanes": lanes}
            for graph in (False, True):
                ctl.use_graph(graph)
                for _ in range(5):
                    ctl.enqueue()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(20):
                    ctl.enqueue()
                e1.record(stream)
                torch.cuda.synchronize()
                row["graph_ms" if graph else "stream_ms"] = round(e0.elapsed_time(e1) / 20, 5)
            ctl.use_graph(False)
            row["launches"] = ctl.launch_count()
            if not lanes:
                row["kernels_us"] = {k: round(v * 1e3, 1) for k, v in ctl.time_kernels(5).items()}
            print(json.dumps(row), flush=True)
            ctl.close()


if __name__ == "__main__":
    main()
