#!/usr/bin/env python
"""bench.py -- rollout-steps/s of the MPPI solve on N B200s (one process per GPU), plus the single-robot latency.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one full MPPI solve (sampling -> predict_States -> calc_Weights -> determine_OptimalSolution,
reference src/diff_drive_mppi.cpp:352-358) over one batch of synthetic input.  Default workload = BASELINE.json
config 4, the one the >= 1e11 rollout-steps/s target is quoted on: diff_drive, K = 2^20 samples per GPU, T = 100.
With N > 1 every rank owns its own 2^20-sample shard of ONE solve (global K = N * 2^20, weak scaling) and the
ranks exchange (c_min, sum w, sum w*u) once per solve.

One JSON line on stdout (rank 0).  `value` = whole-job rollout-steps/s with inputs resident in HBM; `e2e` = the
same metric through mppi_solve() with host buffers (H2D of state/window/warm start and D2H of the controls inside
the timed region); `latency` = host-observed p50 of the K=4096, T=50 steering solve (BASELINE config 2).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (model, K per GPU, T, n_robots per GPU, path course_length)
    "diff_drive_K1M_T100": ("diff_drive", 1 << 20, 100, 1, 10.0),          # BASELINE config 4 (default)
    "steering_K4096_T50": ("steering", 4096, 50, 1, 10.0),                 # config 2
    "full_body_K16384_T100": ("full_body", 16384, 100, 1, 20.0),           # config 3
    "diff_drive_K1000_T15": ("diff_drive", 1000, 15, 1, 10.0),             # config 1
    "batched_1024robots_K1024_T50": ("diff_drive", 1024, 50, 1024, 10.0),  # config 5: 1024 robots per GPU
}
METRIC = "rollout_steps_per_sec"
UNIT = "rollout-steps/s"


def flop_per_step(model, T):
    """Algorithmic FP32 work of one rollout-step with the literal T-point scan."""
    return 6 * T + (130 if model == "full_body" else 50)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), d.get("sm_max_mhz", 1965.0), "measured"
    return 6650.0, 1965.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.01):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = str(e)

    def sample_once(self):
        """One sample from the calling thread (used while the GPU is busy with the enqueued steps)."""
        if not self.ok:
            return
        nv = self.nv
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            try:
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:  # noqa: BLE001
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for bit, name in ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"),
                              (0x4, "sw_power_cap"), (0x80, "hw_power_brake")):
                if r & bit:
                    self.reasons.add(name)
        except Exception:  # noqa: BLE001
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._stop_evt.wait(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": float(self.max_mhz),
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def synthetic_inputs(model, R, course_length, seed=0):
    """launch-file sine path; one robot at (0,0,0), batched robots scattered along per-robot
    phase-shifted paths."""
    from ccv_mppi_path_tracker_b200 import params, paths
    S = params.NUM_STATES[model]
    kw = dict(params.LAUNCH_PATH[model])
    kw["course_length"] = course_length
    base = paths.sin_path(**kw)
    if R == 1:
        return [base], np.zeros((1, S))
    rng = np.random.default_rng(seed)
    path_list, states = [], np.zeros((R, S))
    n_var = 64  # distinct phase-shifted paths, reused round-robin (host memory, not a kernel input size)
    variants = []
    for k in range(n_var):
        kw2 = dict(kw)
        kw2["delta1"] = 2 * np.pi * k / n_var
        variants.append(paths.sin_path(**kw2))
    for r in range(R):
        pth = variants[r % n_var]
        j = r % pth.shape[0]
        jn = min(j + 1, pth.shape[0] - 1)
        tang = np.arctan2(pth[jn, 1] - pth[max(jn - 1, 0), 1], pth[jn, 0] - pth[max(jn - 1, 0), 0])
        states[r, :2] = pth[j] + 0.1 * rng.standard_normal(2)
        states[r, 2] = tang + 0.1 * rng.standard_normal()
        path_list.append(pth)
    return path_list, states


def _host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001
        return os.cpu_count() or 1


def _cpu_case(model, T, course_length):
    from ccv_mppi_path_tracker_b200 import params, paths
    p = params.node_params(model, launch=True, horizon=T, **({"roll_off": False} if model == "full_body" else {}))
    kw = dict(params.LAUNCH_PATH[model])
    kw["course_length"] = course_length
    return p, params.solve_params(model, p), paths.sin_path(**kw), np.zeros(params.NUM_STATES[model])


def reference_binary_rate(model, T, course_length, n_solves, seconds_budget):
    """The UNMODIFIED reference node (oracle/_ref/ref_*_time: its translation unit compiled -O2 against stub ROS
    headers) running its own cycle body -- sampling, predict_States, calc_Weights, determine_OptimalSolution -- on
    ONE thread, as the node does.  Returns (rollout-steps/s, seconds per solve, K of the bounded sample) or None."""
    from oracle import ref_runner
    from ccv_mppi_path_tracker_b200 import params
    if not ref_runner.timing_available(model):
        return None
    p, _, path, state = _cpu_case(model, T, course_length)
    u0 = np.zeros((T - 1, params.NUM_CONTROLS[model]))
    k0 = 512
    t0 = ref_runner.time_solves(model, p, k0, T, state, 0.1, path, u0, 2)[-1]
    rate = k0 * (T - 1) / max(t0, 1e-6)
    k_s = int(max(k0, rate * seconds_budget / n_solves / (T - 1)))
    k_s = max(512, min(1 << 20, (k_s // 512) * 512))
    ts = ref_runner.time_solves(model, p, k_s, T, state, 0.1, path, u0, n_solves)
    return k_s, ts


def cpu_baseline_run(model, T, course_length, seconds_budget=12.0):
    """The CPU beside the GPU number: the unmodified reference node on one thread (kind "reference") when
    oracle/_ref was built, and the FP64 oracle port (OpenMP over samples on all threads, and single-threaded) as
    additional context.  Bounded samples of the same workload."""
    import oracle
    p, sp, path, state = _cpu_case(model, T, course_length)
    cores = _host_cores()
    # calibrate on a small K, then size the sample for ~seconds_budget
    k0 = 2048
    t0 = oracle.time_solves(model, sp, k0, T, state, 0.1, path, 1, literal_copies=False, nthreads=cores)
    rate = k0 * (T - 1) / max(t0, 1e-6)
    k_s = int(min(1 << 20, max(k0, rate * seconds_budget / 2 / (T - 1))))
    k_s = max(1024, (k_s // 1024) * 1024)
    t = oracle.time_solves(model, sp, k_s, T, state, 0.1, path, 2, literal_copies=False, nthreads=cores)
    val = 2 * k_s * (T - 1) / t
    # single thread (the reference is single-threaded), smaller sample
    k_1 = max(1024, (k_s // max(cores, 1) // 1024) * 1024)
    t1 = oracle.time_solves(model, sp, k_1, T, state, 0.1, path, 1, literal_copies=False, nthreads=1)
    port = {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"2 chained FP64 oracle solves of K={k_s} (of the workload's K), T={T}, OpenMP over samples on {cores} threads",
            "single_thread_value": k_1 * (T - 1) / t1, "single_thread_sample": f"1 solve of K={k_1}, T={T}, 1 thread"}
    ref = reference_binary_rate(model, T, course_length, 3, seconds_budget)
    if ref is None:
        return port
    k_r, ts = ref
    return {"value": k_r * (T - 1) / float(np.mean(ts[1:])), "unit": UNIT, "cores": 1, "kind": "reference",
            "sample": f"2 timed cycles (after 1 warm-up) of the unmodified reference node (oracle/_ref, -O2, its own "
                      f"mt19937 sampling included), K={k_r} samples of the workload's K, T={T}, 1 thread (the node is "
                      f"single-threaded)",
            "port_all_threads_value": port["value"], "port_all_threads_cores": cores, "port_all_threads_sample": port["sample"],
            "port_single_thread_value": port["single_thread_value"]}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the solve on the host.  When oracle/_ref holds the
    compiled reference node (built in the container where /root/reference is mounted; the binaries travel to the GPU
    box) that is what runs -- single-threaded, like the node; otherwise the FP64 oracle port on all host threads.
    Each step is one solve of a bounded sample of the workload's K."""
    if rank != 0:
        return
    import oracle
    model, K, T, R, L = WORKLOADS[args.workload]
    cores = _host_cores()
    total = args.steps + args.warmup
    ref = reference_binary_rate(model, T, L, total, 60.0)
    if ref is not None:
        k_s, ts = ref
        t = float(np.sum(ts[args.warmup:]))
        kind, used = "reference", 1
        sample = (f"each step = one cycle of the unmodified reference node (oracle/_ref, -O2) on K={k_s} samples "
                  f"(bounded sample of K={K * R}), T={T}, 1 thread (the node is single-threaded)")
    else:
        p, sp, path, state = _cpu_case(model, T, L)
        k0 = 2048
        t0 = oracle.time_solves(model, sp, k0, T, state, 0.1, path, 1, False, cores)
        rate = k0 * (T - 1) / max(t0, 1e-6)
        k_s = int(min(K * R, max(1024, rate * (90.0 / total) / (T - 1))))
        k_s = max(1024, (k_s // 1024) * 1024)
        for _ in range(args.warmup):
            oracle.time_solves(model, sp, k_s, T, state, 0.1, path, 1, False, cores)
        t = oracle.time_solves(model, sp, k_s, T, state, 0.1, path, args.steps, False, cores)
        kind, used = "port", cores
        sample = f"each step = one FP64 solve of K={k_s} samples (bounded sample of K={K * R}), T={T}, {cores} OpenMP threads"
    val = args.steps * k_s * (T - 1) / t
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": args.workload, "model": model, "K_sample": k_s, "T": T},
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": used, "kind": kind, "sample": sample,
                            "host_cores": cores},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def latency_probe(device, n_solves=1000):
    """BASELINE config 2: steering, K=4096, T=50, one robot; host-observed mppi_solve latency with the CUDA graph."""
    from ccv_mppi_path_tracker_b200 import SteeringDiffDriveMPPI, params, paths
    path = paths.sin_path(**params.LAUNCH_PATH["steering"])
    ctl = SteeringDiffDriveMPPI(launch=True, horizon=50, num_samples=4096, device=device)
    ctl.set_path(path)
    ctl.set_seed(0x5EED0002, 0)
    ctl.use_graph(True)
    state = np.zeros(3)
    for _ in range(20):
        ctl.solve(state, 0.1)
    ts = np.empty(n_solves)
    for k in range(n_solves):
        t0 = time.perf_counter()
        ctl.solve(state, 0.1)
        ts[k] = time.perf_counter() - t0
    km = ctl.time_kernels(20)
    launches = ctl.launch_count()
    ctl.close()
    return {"workload": "steering_K4096_T50", "p50_us": float(np.percentile(ts, 50) * 1e6),
            "p99_us": float(np.percentile(ts, 99) * 1e6), "mean_us": float(ts.mean() * 1e6), "solves": n_solves,
            "device_us": km["total"] * 1e3, "launches_per_solve": launches,
            "what": "host-observed mppi_solve(): host state in -> controls on host, CUDA graph replay"}


def sweep_k(device, model="diff_drive", T=100, n_solves=60):
    """BASELINE.json's metric is quoted "vs K": rollout-steps/s (device-resident, back-to-back enqueues, CUDA events)
    and host-observed solve latency p50 (mppi_solve() with host buffers; CUDA graph for K <= 2^16) for K = 2^10 .. 2^20,
    one robot, launch-file parameters.  One GPU; the sharded runs scale the K = 2^20 row (weak scaling)."""
    import torch
    from ccv_mppi_path_tracker_b200 import CONTROLLERS, params, paths
    kw = dict(params.LAUNCH_PATH[model])
    path = paths.sin_path(**kw)
    S = params.NUM_STATES[model]
    rows = []
    for e in range(10, 21):
        K = 1 << e
        ov = {"roll_off": False} if model == "full_body" else {}
        ctl = CONTROLLERS[model](launch=True, device=device, horizon=T, num_samples=K, **ov)
        ctl.set_path(path)
        ctl.set_seed(0x5EED0000 + e, 0)
        stream = torch.cuda.Stream()
        torch.cuda.set_stream(stream)
        ctl.set_stream(stream.cuda_stream)
        state = np.zeros(S)
        ctl.upload(state, 0.1, with_nominal=True)
        for _ in range(5):
            ctl.enqueue()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(n_solves):
            ctl.enqueue()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n_solves
        ctl.use_graph(K <= (1 << 16))
        for _ in range(5):
            ctl.solve(state, 0.1)
        ts = np.empty(n_solves)
        for k in range(n_solves):
            t0 = time.perf_counter()
            ctl.solve(state, 0.1)
            ts[k] = time.perf_counter() - t0
        rows.append({"K": K, "T": T, "device_ms_per_solve": ms, "rollout_steps_per_sec": K * (T - 1) / (ms * 1e-3),
                     "solve_p50_us": float(np.percentile(ts, 50) * 1e6), "solve_p99_us": float(np.percentile(ts, 99) * 1e6),
                     "launches_per_solve": ctl.launch_count()})
        ctl.close()
    return {"metric": "rollout_steps_per_sec and solve p50 latency vs K", "model": model, "T": T, "n_gpus": 1,
            "solves_per_point": n_solves, "data": "synthetic", "rows": rows}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="diff_drive_K1M_T100", choices=sorted(WORKLOADS))
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scan", default="auto", choices=["auto", "literal", "pruned"])
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="sample-sharded solves: records exchanged by NVLink peer stores inside the finalize kernel, or by ncclAllGather")
    ap.add_argument("--sweep", action="store_true",
                    help="instead of the bench line: rollout-steps/s and solve latency p50 vs K = 2^10..2^20 (one GPU)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.sweep:
        import torch
        torch.cuda.set_device(0)
        print(json.dumps(sweep_k(0)), flush=True)
        return

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from ccv_mppi_path_tracker_b200 import CONTROLLERS, _capi, comm_unique_id

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 MPPI core has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    model, K, T, R, L = WORKLOADS[args.workload]
    U = {"diff_drive": 2, "steering": 3, "full_body": 5}[model]
    shard_robots = R > 1  # config 5 partitions robots (no collective); the others shard samples of one solve
    ov = {"roll_off": False} if model == "full_body" else {}
    ctl = CONTROLLERS[model](launch=True, n_robots=R, device=local, horizon=T, num_samples=K, **ov)
    paths_, states = synthetic_inputs(model, R, L, seed=rank)
    for r in range(R):
        ctl.set_path(paths_[r], robot=r)
    ctl.set_seed(0x5EED0000 + 4, 0)
    ctl.set_scan_mode({"auto": _capi.SCAN_AUTO, "literal": _capi.SCAN_LITERAL, "pruned": _capi.SCAN_PRUNED}[args.scan])
    if world > 1:
        if shard_robots:
            ctl.set_shard(0, K, rank * R)
        else:
            ctl.set_shard(rank * K, world * K, 0)
            exchange = args.exchange
            if exchange == "p2p":
                def all_ok(ok):
                    f = torch.tensor([1 if ok else 0], device="cuda")
                    dist.all_reduce(f, op=dist.ReduceOp.MIN)
                    return int(f.item()) == 1
                mine, err = None, None
                try:
                    mine = ctl.comm_export(world)
                except _capi.MppiError as e:
                    err = e
                if all_ok(mine is not None):
                    t = torch.frombuffer(bytearray(mine), dtype=torch.uint8).cuda()
                    allh = [torch.zeros_like(t) for _ in range(world)]
                    dist.all_gather(allh, t)
                    try:
                        ctl.comm_connect(b"".join(bytes(x.cpu().numpy().tobytes()) for x in allh), rank, world)
                        connected = True
                    except _capi.MppiError as e:
                        connected, err = False, e
                    if not all_ok(connected):
                        raise SystemExit(f"peer exchange connected on some ranks only ({err})")
                else:  # CUDA IPC not available here: every rank falls back to NCCL
                    if rank == 0:
                        print(f"[bench] peer exchange unavailable ({err}); using NCCL", file=sys.stderr)
                    exchange = "nccl"
            if exchange == "nccl":
                idt = torch.zeros(_capi.COMM_ID_BYTES, dtype=torch.uint8, device="cuda")
                if rank == 0:
                    idt.copy_(torch.frombuffer(bytearray(comm_unique_id()), dtype=torch.uint8))
                dist.broadcast(idt, 0)
                ctl.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)
            args.exchange = exchange
    # a non-default torch stream: the handle launches on it, so torch.cuda.Event timing sees the kernels
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctl.set_stream(stream.cuda_stream)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident throughput: inputs uploaded once, K solves enqueued back to back -------------------
    # NVML polling takes a driver lock and nvmlInit is slow: rank 0 only, and set up BEFORE the barrier so that the
    # other ranks' timed regions do not contain rank 0's NVML start-up
    sampler = ClockSampler(local) if (rank == 0 and not os.environ.get("MPPI_BENCH_NO_CLOCKS")) else None
    if sampler:
        sampler.sample_once()
        sampler.samples.clear()
    ctl.upload(states, 0.1, with_nominal=True)
    for _ in range(args.warmup):
        ctl.enqueue()
    if sampler:
        sampler.start()
    sync_all()
    if sampler:
        sampler.samples.clear()  # keep only what is sampled inside the timed region
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        ctl.enqueue()
    e1.record(stream)
    if sampler:
        sampler.sample_once()  # the steps are queued and running: at least one sample under load
    sync_all()
    ms_total = e0.elapsed_time(e1)
    if os.environ.get("MPPI_BENCH_DEBUG"):
        print(f"[bench] rank {rank}: {ms_total / args.steps:.4f} ms/step device-resident", file=sys.stderr, flush=True)
    launches = ctl.launch_count() * args.steps
    clocks = sampler.stop() if sampler else None
    if world > 1:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    steps_per_solve = K * (T - 1) * R * world
    value = steps_per_solve * args.steps / (ms_total * 1e-3)

    # ---- end to end through mppi_solve(): host buffers in, host buffers out, every step ---------------------
    # (unsharded handles replay the solve -- H2D copy, kernels, D2H copy -- as one CUDA graph: mppi_use_graph, the
    # same public switch the latency path uses; MPPI_BENCH_NO_GRAPH=1 times the plain stream launches instead)
    e2e_graph = world == 1 and not os.environ.get("MPPI_BENCH_NO_GRAPH")
    if e2e_graph:
        ctl.use_graph(True)
    for _ in range(min(args.warmup, 3)):
        ctl.solve(states, 0.1)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctl.solve(states, 0.1)
    sync_all()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    h2d, d2h = ctl.io_bytes()
    e2e = {"value": steps_per_solve * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_s / args.steps, "cuda_graph": bool(e2e_graph)}
    if e2e_graph:
        ctl.use_graph(False)

    # ---- per-kernel device time (CUDA events between the launches, on the launching stream) -----------------
    km = ctl.time_kernels(max(3, min(args.steps, 10)))
    stats = ctl.stats(0)
    ctl.close()

    if rank == 0:
        hbm_peak, sm_max_mhz, peak_src = peaks()
        clocks = clocks or {"sm_mhz": None, "sm_max_mhz": sm_max_mhz, "reasons": ["sampling disabled"]}
        clk = clocks.get("sm_mhz") or sm_max_mhz
        fp32_peak_max = 148 * 128 * 2 * sm_max_mhz * 1e6 / 1e12
        fp32_peak_obs = 148 * 128 * 2 * clk * 1e6 / 1e12
        local_steps = K * (T - 1) * R
        fl = flop_per_step(model, T)
        t_k2 = km["rollout_cost"] * 1e-3
        achieved = local_steps * fl / t_k2 / 1e12
        # DRAM traffic of one K2 launch from the committed ncu --set full capture of this workload
        # (profiles/r01_ncu_full.txt: dram__bytes_read.sum + dram__bytes_write.sum); other workloads: not captured
        traffic = 838.8e6 if args.workload == "diff_drive_K1M_T100" else None
        k2_bytes = 4 * U * local_steps  # algorithmic: the normals, read once (4*U B per rollout-step)
        roofline = {"bound": "hbm", "kernel": "rollout_cost", "achieved": k2_bytes / t_k2 / 1e9, "peak": hbm_peak,
                    "unit": "GB/s", "frac": k2_bytes / t_k2 / 1e9 / hbm_peak, "traffic": traffic,
                    "traffic_note": "dram bytes per launch (ncu --set full, profiles/r01_ncu_full.txt); algorithmic "
                                    "bytes = 4*U per rollout-step",
                    "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_src})",
                    "note": "the dominant kernel is FP32/ALU-issue bound, not HBM bound (DESIGN.md section 4): `fp32_view` "
                            "is SURVEY.md 8d's roofline for it, `executed` what it really issues",
                    # algorithmic flop of the LITERAL T-point scan (6 flop/pair) + dynamics against the FP32
                    # peak; the exact pruned scan skips ~95 % of the pairs, so this fraction exceeds 1
                    "fp32_view": {"bound": "fp32", "achieved": achieved, "peak": fp32_peak_max, "unit": "TFLOP/s",
                                  "frac": achieved / fp32_peak_max, "frac_at_observed_clock": achieved / fp32_peak_obs,
                                  "peak_source": f"148 SM x 128 lanes x 2 x {sm_max_mhz:.0f} MHz (clocks.max.sm, {peak_src})",
                                  "algorithmic_flop_per_rollout_step": fl},
                    # from the committed ncu --set full capture of this workload at steady state
                    "executed": ({"warp_instructions_per_warp_step": 81.9, "issue_slot_utilisation": 0.736,
                                  "fp32_pipe_cycles_active": 0.501, "source": "profiles/r01_ncu_full.txt"}
                                 if args.workload == "diff_drive_K1M_T100" else None),
                    "kernel_ms": km}
        nbytes = 4 * U * local_steps
        roof_noise = {"bound": "hbm", "kernel": "noise", "achieved": nbytes / (km["noise"] * 1e-3) / 1e9,
                      "peak": hbm_peak, "unit": "GB/s", "frac": nbytes / (km["noise"] * 1e-3) / 1e9 / hbm_peak,
                      "algorithmic_bytes_per_rollout_step": 4 * U, "peak_source": peak_src}
        if km["weighted_controls"] > 0:
            roof_k4 = {"bound": "hbm", "kernel": "weighted_controls",
                       "achieved": nbytes / (km["weighted_controls"] * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                       "frac": nbytes / (km["weighted_controls"] * 1e-3) / 1e9 / hbm_peak}
        else:
            roof_k4 = {"kernel": "weighted_controls", "fused": "per-CTA records inside rollout_cost (K2); kernel_ms.weights "
                                                               "is the rescale of those records"}
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": {"workload": args.workload, "model": model, "K_per_gpu": K, "K_global": K * (1 if shard_robots else world),
                          "T": T, "U": U, "robots_per_gpu": R, "sharding": "robots" if shard_robots else "samples",
                          "collective": "none" if (shard_robots or world == 1) else (
                              "records (c_min, sum w, sum w^2, sum w*u) stored into every peer's buffer over NVLink by the finalize kernel, merge waits on flags"
                              if args.exchange == "p2p" else "one ncclAllGather of (c_min, sum w, sum w^2, sum w*u) per solve"),
                          "l2": f"noise tensor {4 * U * local_steps / 1e6:.0f} MB per solve vs 126 MB L2 (inputs larger than L2, no flush)",
                          "scan": args.scan, "ess": stats["ess"]},
               "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
               "roofline_noise": roof_noise, "roofline_weighted_controls": roof_k4}
        if not args.no_latency:
            out["latency"] = latency_probe(local)
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline_run(model, T, L)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
