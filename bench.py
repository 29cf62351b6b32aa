#!/usr/bin/env python
"""bench.py -- rollout-steps/s of the MPPI solve on N B200s (one process per GPU), plus the single-robot latency.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME] [--scaling strong|weak]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one full MPPI solve (sampling -> predict_States -> calc_Weights -> determine_OptimalSolution,
reference src/diff_drive_mppi.cpp:352-358) over one batch of synthetic input.  Default workload = BASELINE.json
config 4, the one the >= 1e11 rollout-steps/s target is quoted on: diff_drive, K = 2^20 samples, T = 100.
With N > 1 the K = 2^20 samples of ONE solve are split over the ranks (strong scaling, BASELINE config 4: rank g owns
samples [g K/N, (g+1) K/N)) and the ranks exchange (c_min, sum w, sum w*u) once per solve; the weak figure (2^20
samples per rank) rides along in `weak`.  Config 5 (many robots) partitions robots over the ranks, no collective.

How a step is made stationary: the synthetic robot first tracks the path in closed loop for a few cycles (state
advanced by the first control of each solve, like the reference's plant), so the un-shifted warm start has converged
to steady tracking; the timed device-resident steps then repeat the solve of THAT state and warm start with fresh
noise (MPPI_OPT_FEEDBACK_WARM_START = 0): every timed step does the same work, `value` does not depend on --steps.
`e2e` keeps the closed loop: every step takes the host state in, returns the controls, and the host advances the plant.

One JSON line on stdout (rank 0).  `value` = whole-job rollout-steps/s with inputs resident in HBM; `e2e` = the
same metric through mppi_solve() with host buffers (H2D of header/pose/state [+ windows, warm start] and D2H of the
controls inside the timed region); `latency` = host-observed p50 of the K=4096, T=50 steering solve (BASELINE config
2); `workloads` = the other BASELINE configurations, measured the same way in a few seconds each.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (model, K, T, n_robots per GPU)
    "diff_drive_K1M_T100": ("diff_drive", 1 << 20, 100, 1),          # BASELINE config 4 (default)
    "steering_K4096_T50": ("steering", 4096, 50, 1),                 # config 2
    "full_body_K16384_T100": ("full_body", 16384, 100, 1),           # config 3
    "diff_drive_K1000_T15": ("diff_drive", 1000, 15, 1),             # config 1
    "batched_1024robots_K1024_T50": ("diff_drive", 1024, 50, 1024),  # config 5: 1024 robots per GPU
}
METRIC = "rollout_steps_per_sec"
UNIT = "rollout-steps/s"
NUM_CONTROLS = {"diff_drive": 2, "steering": 3, "full_body": 5}
CLOSED_LOOP_WARMUP = 25  # cycles of closed-loop tracking before anything is timed


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


def ncu_counters(workload):
    """DRAM traffic and issue statistics of the dominant kernel from the committed ncu --set full captures
    (profiles/ncu_counters.json, keyed by workload; written by tools/ncu_summary.py).  None when not captured."""
    p = os.path.join(ROOT, "profiles", "ncu_counters.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        return json.load(f).get(workload)


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed regions run."""

    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"),
               (0x4, "sw_power_cap"), (0x80, "hw_power_brake"))

    def __init__(self, index, period=0.001):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self._active = threading.Event()
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
        if not self.ok:
            return
        nv = self.nv
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            try:
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:  # noqa: BLE001
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for bit, name in self.REASONS:
                if r & bit:
                    self.reasons.add(name)
        except Exception:  # noqa: BLE001
            pass

    def run(self):
        while not self._stop_evt.is_set():
            if self._active.is_set():
                self.sample_once()
                self._stop_evt.wait(self.period)
            else:
                self._stop_evt.wait(0.0005)

    def resume(self):
        self._active.set()

    def pause(self):
        self._active.clear()

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_min_mhz": float(np.min(self.samples)),
                "sm_max_mhz": float(self.max_mhz), "reasons": sorted(self.reasons), "samples": len(self.samples),
                "sampled": "NVML, only while a timed region (device-resident steps, end-to-end steps) is running"}


# ---- synthetic inputs and the plant -------------------------------------------------------------------------------

def bench_path(model, min_length, delta1=0.0):
    """The launch file's sine path (reference_path_creator parameters), long enough for the closed-loop cycles."""
    from ccv_mppi_path_tracker_b200 import params, paths
    kw = dict(params.LAUNCH_PATH[model])
    kw["course_length"] = max(kw["course_length"], float(min_length))
    kw["delta1"] = delta1
    return paths.sin_path(**kw)


def synthetic_inputs(model, R, n_cycles, seed=0):
    """launch-file sine path; one robot at its start, batched robots scattered along
    phase-shifted copies of it.  Paths are long enough for n_cycles of tracking at v_max."""
    from ccv_mppi_path_tracker_b200 import params
    S = params.NUM_STATES[model]
    need = 0.2 * n_cycles + 25.0  # v_max * dt per cycle + the T = 100 window
    if R == 1:
        return [bench_path(model, need)], np.zeros((1, S))
    rng = np.random.default_rng(seed)
    n_var = 64  # distinct phase-shifted paths, reused round-robin (host memory, not a kernel input size)
    variants = [bench_path(model, need + 10.0, 2 * np.pi * k / n_var) for k in range(n_var)]
    path_list, states = [], np.zeros((R, S))
    for r in range(R):
        pth = variants[r % n_var]
        j = r % 100
        tang = np.arctan2(pth[j + 1, 1] - pth[j, 1], pth[j + 1, 0] - pth[j, 0])
        states[r, :2] = pth[j] + 0.1 * rng.standard_normal(2)
        states[r, 2] = tang + 0.1 * rng.standard_normal()
        path_list.append(pth)
    return path_list, states


def plant_step(model, states, u, dt):
    """The kinematic plant, i.e. the controllers' own predict_NextState (diff_drive_mppi.cpp:104-109,
    steering_diff_drive_mppi.cpp:120-125, full_body_mppi.cpp:445-452) applied to the first control of the horizon --
    what mppi_harness.cpp does between cycles.  states [R][S], u [R][T-1][U]."""
    u0 = u[:, 0, :]
    heading = states[:, 2] if model == "diff_drive" else states[:, 2] + u0[:, 2]
    states[:, 0] += u0[:, 0] * np.cos(heading) * dt
    states[:, 1] += u0[:, 0] * np.sin(heading) * dt
    states[:, 2] += u0[:, 1] * dt
    if model == "full_body":
        states[:, 3] += u0[:, 3] * dt
        states[:, 4] += u0[:, 4] * dt
    return states


def _host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001
        return os.cpu_count() or 1


# ---- CPU side ------------------------------------------------------------------------------------------------------

def _cpu_case(model, T):
    from ccv_mppi_path_tracker_b200 import params
    p = params.node_params(model, launch=True, horizon=T, **({"roll_off": False} if model == "full_body" else {}))
    return p, params.solve_params(model, p), bench_path(model, 0.0), np.zeros(params.NUM_STATES[model])


def reference_binary_rate(model, T, n_solves, seconds_budget):
    """The UNMODIFIED reference node (oracle/_ref/ref_*_time: its translation unit compiled -O2 against stub ROS
    headers) running its own cycle body -- sampling, predict_States, calc_Weights, determine_OptimalSolution -- on
    ONE thread, as the node does.  Returns (K of the bounded sample, seconds per solve) or None."""
    from oracle import ref_runner
    from ccv_mppi_path_tracker_b200 import params
    if not ref_runner.timing_available(model):
        return None
    p, _, path, state = _cpu_case(model, T)
    u0 = np.zeros((T - 1, params.NUM_CONTROLS[model]))
    k0 = 512
    t0 = ref_runner.time_solves(model, p, k0, T, state, 0.1, path, u0, 2)[-1]
    rate = k0 * (T - 1) / max(t0, 1e-6)
    k_s = int(max(k0, rate * seconds_budget / n_solves / (T - 1)))
    k_s = max(512, min(1 << 20, (k_s // 512) * 512))
    ts = ref_runner.time_solves(model, p, k_s, T, state, 0.1, path, u0, n_solves)
    return k_s, ts


REF_BUILD_NOTE = ("built -O2 (the reference ships CMAKE_BUILD_TYPE=Debug, -g without optimisation: its own build is "
                  "slower), against stub ROS/tf/Eigen headers -- predict_States() still fills its K marker messages "
                  "(stub structs), as the node does every cycle")


def cpu_baseline_run(model, T, seconds_budget=12.0):
    """The CPU beside the GPU number: the unmodified reference node on one thread (kind "reference") when
    oracle/_ref was built, and the FP64 oracle port (OpenMP over samples on all threads, and single-threaded) as
    additional context.  Bounded samples of the same workload."""
    import oracle
    p, sp, path, state = _cpu_case(model, T)
    cores = _host_cores()
    k0 = 2048
    t0 = oracle.time_solves(model, sp, k0, T, state, 0.1, path, 1, literal_copies=False, nthreads=cores)
    rate = k0 * (T - 1) / max(t0, 1e-6)
    k_s = int(min(1 << 20, max(k0, rate * seconds_budget / 2 / (T - 1))))
    k_s = max(1024, (k_s // 1024) * 1024)
    t = oracle.time_solves(model, sp, k_s, T, state, 0.1, path, 2, literal_copies=False, nthreads=cores)
    val = 2 * k_s * (T - 1) / t
    k_1 = max(1024, (k_s // max(cores, 1) // 1024) * 1024)
    t1 = oracle.time_solves(model, sp, k_1, T, state, 0.1, path, 1, literal_copies=False, nthreads=1)
    port = {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"2 chained FP64 oracle solves of K={k_s} (of the workload's K), T={T}, OpenMP over samples on {cores} threads",
            "single_thread_value": k_1 * (T - 1) / t1, "single_thread_sample": f"1 solve of K={k_1}, T={T}, 1 thread"}
    ref = reference_binary_rate(model, T, 3, seconds_budget)
    if ref is None:
        return port
    k_r, ts = ref
    return {"value": k_r * (T - 1) / float(np.mean(ts[1:])), "unit": UNIT, "cores": 1, "kind": "reference",
            "sample": f"2 timed cycles (after 1 warm-up) of the unmodified reference node (oracle/_ref; {REF_BUILD_NOTE}; "
                      f"its own mt19937 sampling included), K={k_r} samples -- a bounded sample of the workload's K, the "
                      f"rate per rollout-step does not depend on K -- T={T}, 1 thread (the node is single-threaded)",
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
    model, K, T, R = WORKLOADS[args.workload]
    cores = _host_cores()
    total = args.steps + args.warmup
    ref = reference_binary_rate(model, T, total, 60.0)
    if ref is not None:
        k_s, ts = ref
        t = float(np.sum(ts[args.warmup:]))
        kind, used = "reference", 1
        sample = (f"each step = one cycle of the unmodified reference node (oracle/_ref; {REF_BUILD_NOTE}) on K={k_s} "
                  f"samples -- a bounded sample of the workload's K={K * R}; the per-rollout-step rate does not depend on "
                  f"K -- T={T}, 1 thread (the node is single-threaded)")
    else:
        p, sp, path, state = _cpu_case(model, T)
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
           "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
           "scaling": "none (one host thread, whatever --gpus says)", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": args.workload, "model": model, "K_sample": k_s, "K_workload": K * R, "T": T,
                      "same_config_note": "a bounded K-sample of the same workload: the metric is a rate per rollout-step"},
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": used, "kind": kind, "sample": sample,
                            "host_cores": cores},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


# ---- GPU side ------------------------------------------------------------------------------------------------------

class Dist:
    """Thin view of torch.distributed for the three things the bench needs."""

    def __init__(self, world):
        self.world = world
        if world > 1:
            import torch.distributed as dist
            self.dist = dist

    def barrier(self):
        import torch
        torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            torch.cuda.synchronize()

    def max(self, v):
        if self.world == 1:
            return v
        import torch
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def all_ok(self, ok):
        if self.world == 1:
            return ok
        import torch
        f = torch.tensor([1 if ok else 0], device="cuda")
        self.dist.all_reduce(f, op=self.dist.ReduceOp.MIN)
        return int(f.item()) == 1

    def gather_bytes(self, b):
        """all-gather of equal-sized byte strings -> list over ranks"""
        import torch
        t = torch.frombuffer(bytearray(b), dtype=torch.uint8).cuda()
        out = [torch.zeros_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [bytes(x.cpu().numpy().tobytes()) for x in out]


def connect_exchange(ctl, D, rank, exchange):
    """Sample-sharded handles: NVLink peer exchange (CUDA IPC) or NCCL.  Returns the transport in use."""
    import torch
    from ccv_mppi_path_tracker_b200 import _capi, comm_unique_id
    world = D.world
    if exchange == "p2p":
        mine, err = None, None
        try:
            mine = ctl.comm_export(world)
        except _capi.MppiError as e:
            err = e
        if D.all_ok(mine is not None):
            try:
                ctl.comm_connect(b"".join(D.gather_bytes(mine)), rank, world)
                connected = True
            except _capi.MppiError as e:
                connected, err = False, e
            if not D.all_ok(connected):
                raise SystemExit(f"peer exchange connected on some ranks only ({err})")
            return "p2p"
        if rank == 0:  # CUDA IPC not available here: every rank falls back to NCCL
            print(f"[bench] peer exchange unavailable ({err}); using NCCL", file=sys.stderr)
    idt = torch.zeros(_capi.COMM_ID_BYTES, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(comm_unique_id()), dtype=torch.uint8))
    D.dist.broadcast(idt, 0)
    ctl.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)
    return "nccl"


def measure(name, K_local, shard, args, D, rank, local, steps, warmup, sampler=None, exchange="p2p", scan="auto",
            kernel_iters=5, e2e_graph=None):
    """One workload on this rank's GPU: closed-loop warm-up, device-resident steps, end-to-end steps, per-kernel times.
    shard: "none" | "samples" (K_local samples of one K_local * world solve) | "robots" (robots partitioned)."""
    import torch
    from ccv_mppi_path_tracker_b200 import CONTROLLERS, _capi
    model, _, T, R = WORKLOADS[name]
    U = NUM_CONTROLS[model]
    world = D.world
    if e2e_graph is None:  # CUDA-graph replay of the synchronous solve where launch overhead matters (small solves)
        e2e_graph = not os.environ.get("MPPI_BENCH_NO_GRAPH") and (
            bool(os.environ.get("MPPI_BENCH_GRAPH")) or K_local * R * (T - 1) <= (1 << 22))
    ov = {"roll_off": False} if model == "full_body" else {}
    ctl = CONTROLLERS[model](launch=True, n_robots=R, device=local, horizon=T, num_samples=K_local, **ov)
    n_cycles = CLOSED_LOOP_WARMUP + 2 * (steps + warmup) + 40
    paths_, states = synthetic_inputs(model, R, n_cycles, seed=rank if shard == "robots" else 0)
    for r in range(R):
        ctl.set_path(paths_[r], robot=r)
    ctl.set_seed(0x5EED0000 + 4, 0)
    ctl.set_scan_mode({"auto": _capi.SCAN_AUTO, "literal": _capi.SCAN_LITERAL, "pruned": _capi.SCAN_PRUNED}[scan])
    transport = "none"
    if world > 1 and shard == "robots":
        ctl.set_shard(0, K_local, rank * R)
    elif world > 1 and shard == "samples":
        ctl.set_shard(rank * K_local, world * K_local, 0)
        transport = connect_exchange(ctl, D, rank, exchange)
    # a high-priority torch stream: the handle launches on it (torch.cuda.Event timing sees the kernels), its own side
    # stream (candidate grid, noise prefetch) has the lowest priority
    stream = torch.cuda.Stream(priority=-1)
    torch.cuda.set_stream(stream)
    ctl.set_stream(stream.cuda_stream)
    dt = 0.1
    D.barrier()  # every rank connected before the first sharded solve

    # ---- closed-loop warm-up: the warm start converges to steady tracking ----------------------------------------
    for _ in range(CLOSED_LOOP_WARMUP):
        u = ctl.solve(states, dt)
        plant_step(model, states, u.reshape(R, T - 1, U), dt)

    # ---- device-resident throughput: this state and warm start resident, K identical solves back to back --------
    ctl.set_option(_capi.OPT_FEEDBACK_WARM_START, 0)
    ctl.upload(states, dt, with_nominal=True)
    for _ in range(warmup):
        ctl.enqueue()
    D.barrier()
    if sampler:
        sampler.resume()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        ctl.enqueue()
    e1.record(stream)
    D.barrier()
    if sampler:
        sampler.pause()
    ms_total = D.max(e0.elapsed_time(e1))
    launches = ctl.launch_count() * steps
    units = K_local * (T - 1) * R * world
    res = {"value": units * steps / (ms_total * 1e-3), "ms_per_step": ms_total / steps, "gpu_launches": launches,
           "launches_per_step": ctl.launch_count(), "transport": transport}

    # ---- end to end through mppi_solve(): host state in, controls out, closed loop ---------------------------------
    # The closed loop is recorded first (untimed): the pose of every cycle as the plant -- the controllers' own
    # predict_NextState on the first control -- produces it.  The timed loop then replays exactly those poses, one
    # mppi_solve() per step with host buffers in and out, the warm start carried from solve to solve as the reference
    # does; the Python plant step itself (numpy, not part of the product) stays outside the timed region.
    ctl.set_option(_capi.OPT_FEEDBACK_WARM_START, 1)
    if R >= 8:  # a fleet host keeps the warm start on the device (it never edits optimal_solution between cycles)
        ctl.set_option(_capi.OPT_UPLOAD_WARM_START, 0)
    graph = transport != "nccl" and e2e_graph
    ctl.use_graph(graph)
    n_pre = max(min(warmup, 3), 2)
    u_start, s = ctl.optimal_solution.copy(), states.copy()
    traj = []
    for _ in range(n_pre + steps):
        traj.append(s.copy())
        plant_step(model, s, ctl.solve(s, dt).reshape(R, T - 1, U), dt)
    ctl.optimal_solution[...] = u_start
    ctl.upload(traj[0], dt, with_nominal=True)  # the device-resident warm start back to the start of the recording
    ctl.synchronize()
    for k in range(n_pre):
        ctl.solve(traj[k], dt)
    D.barrier()
    if sampler:
        sampler.resume()
    t0 = time.perf_counter()
    for k in range(steps):
        ctl.solve(traj[n_pre + k], dt)
    D.barrier()
    e2e_s = D.max(time.perf_counter() - t0)
    if sampler:
        sampler.pause()
    states = traj[-1]
    h2d, d2h = ctl.io_bytes()
    res["e2e"] = {"value": units * steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                  "ms_per_step": 1e3 * e2e_s / steps, "cuda_graph": bool(graph),
                  "what": "closed loop replayed: mppi_solve(host pose of cycle k) -> host controls, every step; H2D + D2H inside"}
    ctl.use_graph(False)

    # ---- per-kernel device time (CUDA events between the launches, kernels one after the other) ------------------
    ctl.set_option(_capi.OPT_FEEDBACK_WARM_START, 0)
    ctl.upload(states, dt, with_nominal=True)
    res["kernel_ms"] = ctl.time_kernels(kernel_iters)
    res["ess"] = ctl.stats(0)["ess"]
    res["fused_controls"] = bool(ctl.get_option(_capi.INFO_FUSED_CONTROLS))
    res["paths"] = paths_
    return ctl, res, states


def roofline_objects(name, res, K_local, hbm_peak, sm_max_mhz, peak_src, clocks):
    model, _, T, R = WORKLOADS[name]
    U = NUM_CONTROLS[model]
    km = res["kernel_ms"]
    local_steps = K_local * (T - 1) * R
    t_k2 = km["rollout_cost"] * 1e-3
    k2_bytes = 4 * U * local_steps  # algorithmic: the normals, read once (4*U B per rollout-step)
    fl = flop_per_step(model, T)
    fp32_peak = 148 * 128 * 2 * sm_max_mhz * 1e6 / 1e12
    ncu = ncu_counters(name)
    roof = {"bound": "hbm", "kernel": "rollout_cost", "achieved": k2_bytes / t_k2 / 1e9, "peak": hbm_peak,
            "unit": "GB/s", "frac": k2_bytes / t_k2 / 1e9 / hbm_peak,
            "traffic": ncu.get("dram_bytes_per_launch") if ncu else None,
            "traffic_note": (f"dram bytes per launch from {ncu['source']}" if ncu else "no committed ncu capture of this workload")
                            + "; algorithmic bytes = 4*U per rollout-step (+ the re-read of the CTA's tile when the "
                              "weighted controls are reduced inside K2)",
            "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_src})",
            "note": "the dominant kernel is FP32/ALU-issue bound, not HBM bound (DESIGN.md section 4): `fp32_view` is "
                    "SURVEY.md 8d's roofline for it, `executed` what it really issues",
            "fp32_view": {"bound": "fp32", "achieved": local_steps * fl / t_k2 / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
                          "frac": local_steps * fl / t_k2 / 1e12 / fp32_peak,
                          "peak_source": f"148 SM x 128 lanes x 2 x {sm_max_mhz:.0f} MHz (clocks.max.sm, {peak_src})",
                          "algorithmic_flop_per_rollout_step": fl,
                          "note": "algorithmic flop of the LITERAL T-point scan; the exact pruned scan skips most pairs, "
                                  "so this fraction can exceed 1"},
            "executed": ncu.get("executed") if ncu else None,
            "kernel_ms": km, "fused_controls": res["fused_controls"]}
    nbytes = 4 * U * local_steps
    roof_noise = {"bound": "hbm", "kernel": "noise", "achieved": nbytes / (km["noise"] * 1e-3) / 1e9,
                  "peak": hbm_peak, "unit": "GB/s", "frac": nbytes / (km["noise"] * 1e-3) / 1e9 / hbm_peak,
                  "algorithmic_bytes_per_rollout_step": 4 * U, "peak_source": peak_src,
                  "note": "timed alone; in a solve it runs on the side stream under the previous/current K2 (noise prefetch)"}
    if not res["fused_controls"]:
        roof_k4 = {"bound": "hbm", "kernel": "weighted_controls",
                   "achieved": nbytes / (km["weighted_controls"] * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                   "frac": nbytes / (km["weighted_controls"] * 1e-3) / 1e9 / hbm_peak}
    else:
        roof_k4 = {"kernel": "weighted_controls", "fused": "per-CTA records inside rollout_cost (K2); kernel_ms.finalize is "
                                                           "the one-kernel tail (rescale + finalize + merge)"}
    return roof, roof_noise, roof_k4


def latency_probe(device, n_solves=1000):
    """BASELINE config 2: steering, K=4096, T=50, one robot; host-observed mppi_solve latency with the CUDA graph,
    closed loop (the host advances the plant between solves)."""
    from ccv_mppi_path_tracker_b200 import SteeringDiffDriveMPPI
    paths_, states = synthetic_inputs("steering", 1, n_solves + 60)
    ctl = SteeringDiffDriveMPPI(launch=True, horizon=50, num_samples=4096, device=device)
    ctl.set_path(paths_[0])
    ctl.set_seed(0x5EED0002, 0)
    ctl.use_graph(True)
    for _ in range(40):
        plant_step("steering", states, ctl.solve(states, 0.1).reshape(1, 49, 3), 0.1)
    ts = np.empty(n_solves)
    for k in range(n_solves):
        t0 = time.perf_counter()
        u = ctl.solve(states, 0.1)
        ts[k] = time.perf_counter() - t0
        plant_step("steering", states, u.reshape(1, 49, 3), 0.1)
    launches = ctl.launch_count()
    ctl.use_graph(False)
    ctl.upload(states, 0.1, with_nominal=True)
    km = ctl.time_kernels(20)
    ctl.close()
    out = {"workload": "steering_K4096_T50", "p50_us": float(np.percentile(ts, 50) * 1e6),
           "p99_us": float(np.percentile(ts, 99) * 1e6), "mean_us": float(ts.mean() * 1e6), "solves": n_solves,
           "device_us_kernels_serialised": km["total"] * 1e3, "launches_per_solve": launches,
           "what": "host-observed mppi_solve() from Python (ctypes): host state in -> controls on host, CUDA graph replay, closed loop"}
    # the same solve from the C++ host classes (csrc/host/controllers.hpp): the north star's host language
    exe = os.path.join(ROOT, "ccv_mppi_path_tracker_b200", "mppi_harness")
    if os.path.exists(exe):
        try:
            r = subprocess.run([exe, "--model", "sd", "--launch", "--K", "4096", "--T", "50", "--graph", "--cycles", "600",
                                "--sin", "200", "1.0", "0.25", "--quiet"], capture_output=True, text=True, timeout=120)
            j = json.loads(r.stdout.strip().splitlines()[-1])
            out["cpp_harness"] = {"p50_us": j["solve_p50_us"], "p99_us": j["solve_p99_us"], "rmse_m": j["rmse_m"],
                                  "cycles": j["cycles"], "what": "mppi_harness (C++ SteeringDiffDriveMPPI::solve), closed loop"}
        except Exception as e:  # noqa: BLE001
            out["cpp_harness"] = {"error": str(e)[:200]}
    return out


def exchange_check(ctl, name, K_local, D, rank, local, states, paths_):
    """Sample-sharded runs: one more solve, then (a) every rank's record all-gathered and merged on the host with
    mppi_merge_partials against the controls the device merge produced, (b) the controls of all ranks compared bit for
    bit, (c) the same solve over the other transport (NCCL all-gather) compared bit for bit."""
    from ccv_mppi_path_tracker_b200 import CONTROLLERS, _capi, merge_partials, params
    model, _, T, R = WORKLOADS[name]
    U = NUM_CONTROLS[model]
    world = D.world
    ctl.use_graph(False)
    ctl.set_option(_capi.OPT_FEEDBACK_WARM_START, 1)
    ctl.set_seed(0x5EED0000 + 99, 1000)
    ctl.optimal_solution[...] = 0.0
    u_dev = ctl.solve(states, 0.1).copy().reshape(-1)
    rec = ctl.record(0)
    recs = np.stack([np.frombuffer(b, dtype=np.float32) for b in D.gather_bytes(rec.tobytes())])
    lam = params.node_params(model, launch=True)["lambda_"]
    u_host, _ = merge_partials(recs, lam)
    sp = params.solve_params(model, ctl.p)
    rng_ = np.tile(np.array(sp["u_max"][:U]) - np.array(sp["u_min"][:U]), T - 1)
    du = float(np.max(np.abs(u_dev.astype(np.float32) - u_host) / rng_))
    all_u = D.gather_bytes(u_dev.astype(np.float64).tobytes())
    identical = all(b == all_u[0] for b in all_u)
    # the other transport on a second handle: same seed, counter, shard, inputs
    p2p_equals_nccl = None
    try:
        other = CONTROLLERS[model](launch=True, n_robots=R, device=local, horizon=T, num_samples=K_local)
        other.set_path(paths_[0])
        other.set_seed(0x5EED0000 + 99, 1000)
        other.set_shard(rank * K_local, world * K_local, 0)
        used = connect_exchange(other, D, rank, "nccl")
        D.barrier()
        u_other = other.solve(states, 0.1).copy().reshape(-1)
        D.barrier()
        other.close()
        p2p_equals_nccl = bool(D.all_ok(np.array_equal(u_other, u_dev))) if used == "nccl" else None
    except Exception as e:  # noqa: BLE001
        if rank == 0:
            print(f"[bench] exchange_check: second transport failed: {e}", file=sys.stderr)
    return {"max_du_over_range": du, "ranks_bit_identical": bool(identical), "p2p_equals_nccl": p2p_equals_nccl,
            "what": "host merge (mppi_merge_partials) of the all-gathered per-rank records vs the device merge; limit 2e-5"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="diff_drive_K1M_T100", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N > 1, sample-sharded workloads: split the workload's K over the ranks (strong, BASELINE config 4) "
                         "or give every rank the whole K (weak)")
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-workloads", action="store_true", help="skip the other BASELINE configurations")
    ap.add_argument("--scan", default="auto", choices=["auto", "literal", "pruned"])
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="sample-sharded solves: records exchanged by NVLink peer stores inside the tail kernel, or by ncclAllGather")
    ap.add_argument("--sweep", action="store_true",
                    help="instead of the bench line: rollout-steps/s and solve latency p50 vs K = 2^10..2^20 (one GPU)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    args.steps = max(args.steps, 1)
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

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 MPPI core has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    D = Dist(world)

    model, K, T, R = WORKLOADS[args.workload]
    U = NUM_CONTROLS[model]
    shard = "none" if world == 1 else ("robots" if R > 1 else "samples")
    strong = shard == "samples" and args.scaling == "strong"
    K_local = K // world if strong else K
    if strong and (K % (4 * world)) != 0:
        raise SystemExit("strong scaling needs K divisible by 4 * n_gpus")

    # NVML polling takes a driver lock and nvmlInit is slow: rank 0 only, set up before anything is timed
    sampler = ClockSampler(local) if (rank == 0 and not os.environ.get("MPPI_BENCH_NO_CLOCKS")) else None
    if sampler:
        sampler.sample_once()
        sampler.samples.clear()
        sampler.start()

    ctl, res, states = measure(args.workload, K_local, shard, args, D, rank, local, args.steps, args.warmup, sampler,
                               exchange=args.exchange, scan=args.scan, kernel_iters=max(3, min(args.steps, 10)))
    xcheck = None
    if shard == "samples":
        xcheck = exchange_check(ctl, args.workload, K_local, D, rank, local, states, res["paths"])
    ctl.close()
    res.pop("paths", None)
    clocks = sampler.stop() if sampler else None

    weak = None
    if strong:  # the weak figure beside it: every rank solves the whole K (K_global = N * K)
        c2, r2, _ = measure(args.workload, K, "samples", args, D, rank, local, max(10, args.steps // 2), 3, None,
                            exchange=args.exchange, scan=args.scan, kernel_iters=3)
        c2.close()
        weak = {"value": r2["value"], "unit": UNIT, "ms_per_step": r2["ms_per_step"], "K_per_gpu": K, "K_global": K * world,
                "e2e_value": r2["e2e"]["value"], "scaling": "weak"}

    # ---- the other BASELINE configurations, a few seconds each -------------------------------------------------
    workloads = {}
    if not args.no_workloads:
        if world == 1:
            names = [n for n in WORKLOADS if n != args.workload]
        else:  # multi-GPU: the robot-partitioned configuration (config 5: 1024 robots per GPU, no collective)
            names = [n for n in WORKLOADS if WORKLOADS[n][3] > 1 and n != args.workload]
        hbm_peak, sm_max_mhz, peak_src = peaks()
        for n in names:
            m_, K_, T_, R_ = WORKLOADS[n]
            sh = "none" if world == 1 else "robots"
            c3, r3, _ = measure(n, K_, sh, args, D, rank, local, 20, 3, None, scan="auto", kernel_iters=5)
            c3.close()
            if rank == 0:
                roof, _, _ = roofline_objects(n, r3, K_, hbm_peak, sm_max_mhz, peak_src, None)
                key = n if world == 1 or R_ == 1 else f"batched_{R_ * world}robots_K{K_}_T{T_}"
                workloads[key] = {"value": r3["value"], "unit": UNIT, "ms_per_step": r3["ms_per_step"],
                                  "e2e": r3["e2e"], "e2e_over_value": r3["e2e"]["value"] / r3["value"],
                                  "kernel_ms": r3["kernel_ms"], "launches_per_step": r3["launches_per_step"],
                                  "roofline": {k: roof[k] for k in ("bound", "kernel", "achieved", "peak", "unit", "frac", "traffic",
                                                                    "executed", "fused_controls")},
                                  "config": {"model": m_, "K": K_, "T": T_, "robots_per_gpu": R_, "n_gpus": world,
                                             "steps": 20, "warmup": 3, "sharding": sh}}

    if rank == 0:
        hbm_peak, sm_max_mhz, peak_src = peaks()
        clocks = clocks or {"sm_mhz": None, "sm_max_mhz": sm_max_mhz, "reasons": ["sampling disabled"], "samples": 0}
        roof, roof_noise, roof_k4 = roofline_objects(args.workload, res, K_local, hbm_peak, sm_max_mhz, peak_src, clocks)
        # sample-sharded workloads: as asked (strong = the workload's K split over the ranks); robots partitioned: weak
        scaling = args.scaling if R == 1 else "weak"
        coll = "none"
        if shard == "samples":
            coll = ("records (c_min, sum w, sum w^2, sum w*u) stored into every peer's buffer over NVLink by the tail kernel, "
                    "which then waits on the peers' flags and merges" if res["transport"] == "p2p"
                    else "one ncclAllGather of (c_min, sum w, sum w^2, sum w*u) per solve")
        out = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
               "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": {"workload": args.workload, "model": model, "K_per_gpu": K_local,
                          "K_global": K_local * (world if shard == "samples" else 1), "T": T, "U": U, "robots_per_gpu": R,
                          "sharding": shard, "collective": coll,
                          "l2": f"noise tensor {4 * U * K_local * (T - 1) * R / 1e6:.0f} MB per solve and GPU vs 126 MB L2"
                                + (" (inputs larger than L2, no flush)" if 4 * U * K_local * (T - 1) * R > 126e6 else
                                   " -- fits: between two timed steps the generator rewrites the OTHER buffer of the "
                                   "double-buffered tensor (2x this size), so no step re-reads lines it left in L2"),
                          "scan": args.scan, "ess": res["ess"],
                          "stationary": f"{CLOSED_LOOP_WARMUP} closed-loop cycles, then the same state + warm start every timed step (fresh noise)"},
               "e2e": res["e2e"], "gpu_launches": res["gpu_launches"], "launches_per_step": res["launches_per_step"],
               "clocks": clocks, "roofline": roof, "roofline_noise": roof_noise, "roofline_weighted_controls": roof_k4}
        if weak:
            out["weak"] = weak
        if xcheck:
            out["exchange_check"] = xcheck
        if workloads:
            out["workloads"] = workloads
        if not args.no_latency and world == 1:
            out["latency"] = latency_probe(local)
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline_run(model, T)
        print(json.dumps(out), flush=True)
    bad = bool(xcheck) and (xcheck["max_du_over_range"] > 2e-5 or not xcheck["ranks_bit_identical"]
                            or xcheck["p2p_equals_nccl"] is False)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if bad:
        raise SystemExit(f"exchange check failed: {xcheck}")


def sweep_k(device, model="diff_drive", T=100, n_solves=60):
    """BASELINE.json's metric is quoted "vs K": rollout-steps/s (device-resident, back-to-back enqueues, CUDA events)
    and host-observed solve latency p50 (mppi_solve() with host buffers; CUDA graph up to K = 2^15) for K = 2^10 .. 2^20, one robot,
    launch-file parameters, stationary steps as in the bench line.  One GPU."""
    import torch
    from ccv_mppi_path_tracker_b200 import CONTROLLERS, _capi
    rows = []
    U = NUM_CONTROLS[model]
    for e in range(10, 21):
        K = 1 << e
        ov = {"roll_off": False} if model == "full_body" else {}
        ctl = CONTROLLERS[model](launch=True, device=device, horizon=T, num_samples=K, **ov)
        paths_, states = synthetic_inputs(model, 1, CLOSED_LOOP_WARMUP + 2 * n_solves + 40)
        ctl.set_path(paths_[0])
        ctl.set_seed(0x5EED0000 + e, 0)
        stream = torch.cuda.Stream(priority=-1)
        torch.cuda.set_stream(stream)
        ctl.set_stream(stream.cuda_stream)
        for _ in range(CLOSED_LOOP_WARMUP):
            plant_step(model, states, ctl.solve(states, 0.1).reshape(1, T - 1, U), 0.1)
        ctl.set_option(_capi.OPT_FEEDBACK_WARM_START, 0)
        ctl.upload(states, 0.1, with_nominal=True)
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
        ctl.set_option(_capi.OPT_FEEDBACK_WARM_START, 1)
        ctl.use_graph(K * (T - 1) <= (1 << 22))  # graph replay where launch overhead matters, as in the bench line
        for _ in range(5):
            plant_step(model, states, ctl.solve(states, 0.1).reshape(1, T - 1, U), 0.1)
        ts = np.empty(n_solves)
        for k in range(n_solves):
            t0 = time.perf_counter()
            u = ctl.solve(states, 0.1)
            ts[k] = time.perf_counter() - t0
            plant_step(model, states, u.reshape(1, T - 1, U), 0.1)
        rows.append({"K": K, "T": T, "device_ms_per_solve": ms, "rollout_steps_per_sec": K * (T - 1) / (ms * 1e-3),
                     "solve_p50_us": float(np.percentile(ts, 50) * 1e6), "solve_p99_us": float(np.percentile(ts, 99) * 1e6),
                     "launches_per_solve": ctl.launch_count()})
        ctl.close()
    return {"metric": "rollout_steps_per_sec and solve p50 latency vs K", "model": model, "T": T, "n_gpus": 1,
            "solves_per_point": n_solves, "data": "synthetic", "rows": rows}


if __name__ == "__main__":
    main()
