#!/usr/bin/env python
"""Generate the golden vectors of tests/golden/*.npz from the UNMODIFIED reference nodes (oracle/_ref, built by
`make ref` from /root/reference -- available in the build container only).  Each file holds the inputs of one MPPI
cycle (parameters, state, path, warm start, the standard-normal noise tensor) and what the reference's own
predict_States / calc_Weights / calc_Cost / determine_OptimalSolution produced for them.

    python tests/golden/make_golden.py

Reference UB handled as documented in oracle/ref_shim/ref_driver.cpp: the out-of-bounds v_[T-1] read is 0.0.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from common import make_case  # noqa: E402
from oracle import ref_runner  # noqa: E402
from ccv_mppi_path_tracker_b200 import paths  # noqa: E402

CASES = [
    # name, model, K, T, seed, extra overrides, path kind
    ("dd_launch_K96_T15", "diff_drive", 96, 15, 101, {}, "launch"),
    ("dd_tailclamp_K64_T40", "diff_drive", 64, 40, 102, {}, "launch_near_end"),
    ("dd_datacsv_K64_T15", "diff_drive", 64, 15, 103, {}, "one_point"),
    ("sd_launch_K96_T15", "steering", 96, 15, 104, {}, "launch"),
    ("sd_K48_T50", "steering", 48, 50, 105, {}, "launch"),
    ("fb_allterms_K96_T15", "full_body", 96, 15, 106, {"roll_off": False}, "launch"),
    ("fb_rolloff_K64_T15", "full_body", 64, 15, 107, {"roll_off": True}, "launch"),
    ("fb_steeroff_K64_T20", "full_body", 64, 20, 108, {"roll_off": False, "steer_off": True}, "launch"),
    ("dd_defaults_K64_T15", "diff_drive", 64, 15, 109, {"launch": False}, "launch"),
]


def main():
    if not ref_runner.available():
        raise SystemExit("oracle/_ref is missing: run `make ref` where /root/reference is mounted")
    for name, model, K, T, seed, ov, kind in CASES:
        ov = dict(ov)
        launch = ov.pop("launch", True)
        case = make_case(model, K, T, seed=seed, launch=launch, **ov)
        path, state = case["path"], case["state"].copy()
        if kind == "one_point":
            path = np.array([[-5.45606, -6.61448]])  # /root/reference/data/data.csv:1
            state[:2] = [-5.2, -6.4]
        elif kind == "launch_near_end":
            state[:2] = path[-12] + [0.05, -0.08]
        rng = np.random.default_rng(seed)
        u0 = case["u0"] + 0.1 * rng.standard_normal(case["u0"].shape)
        r = ref_runner.run(model, case["p"], K, T, state, case["dt"], path, case["eps"], u0)
        keys = sorted(case["p"])
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"), model=model, K=K, T=T, dt=case["dt"], state=state, path=path,
            eps=case["eps"], u0=u0, param_names=np.array(keys), param_values=np.array([float(case["p"][k]) for k in keys]),
            **{"ref_" + k: v for k, v in r.items()})
        print(name, "cost range", r["cost"].min(), r["cost"].max(), "u_new[0]", r["u_new"][0])


def cmd_golden():
    """publish_CmdVel + publish_CmdPos of the three nodes (DD:248-263, SD:266-296, FB:238-275) for first controls that
    cover both turning directions, w = 0 (R = inf), v = w = 0 (R = NaN), the roll clamp, steer_off and roll_off."""
    rng = np.random.default_rng(55)
    n = 40
    c = np.zeros((n, 8))
    c[:, 0] = rng.uniform(-1.0, 2.0, n)                  # v0
    c[:, 1] = rng.uniform(-1.0, 1.0, n)                  # w0
    c[:, 2] = rng.uniform(-0.52, 0.52, n)                # steer0 / direction0
    c[:, 3] = rng.uniform(-0.6, 0.6, n)                  # roll_v0
    c[:, 4] = rng.uniform(-0.6, 0.6, n)                  # roll of the current state (beyond +-30 deg in places)
    c[:, 5] = rng.uniform(0.08, 0.12, n)                 # dt
    c[0, 1] = 0.0                                        # w = 0: R = inf
    c[1, 0], c[1, 1] = 0.0, 0.0                          # v = w = 0: R = NaN
    c[2, 1], c[2, 2] = 0.0, 0.0                          # w = 0 and steer = 0: inf * 0
    c[3, 1] = -0.0
    c[4, 3], c[4, 4] = 3.0, 0.5                          # roll command far beyond roll_max
    c[5, 3], c[5, 4] = -3.0, -0.5
    c[6:12, 6] = 1.0                                     # steer_off
    c[10:16, 7] = 1.0                                    # roll_off
    out = {}
    for model, tag in (("diff_drive", "dd"), ("steering", "sd"), ("full_body", "fb")):
        out["ref_" + tag] = ref_runner.run_cmd(model, c)
    np.savez_compressed(os.path.join(HERE, "cmd_golden.npz.dat"), cases=c, **out)
    os.replace(os.path.join(HERE, "cmd_golden.npz.dat.npz"), os.path.join(HERE, "cmd_golden.dat"))
    print("cmd_golden", {k: v[0] for k, v in out.items()})


def estimator_golden():
    """Eight cycles of the full-body node's ZMP monitors (calc_true_ZMP + get_CurrentState, FB:528-596)."""
    rng = np.random.default_rng(77)
    n = 8
    cyc = np.zeros((n, 30))
    cyc[:, 0] = rng.uniform(0.08, 0.12, n)           # dt
    cyc[:, 1:3] = rng.normal(0, 0.08, (n, 2))        # imu roll, pitch
    cyc[:, 3:5] = rng.normal(0, 0.6, (n, 2))         # accel x, y (base frame)
    cyc[:, 5:8] = rng.normal(0, 0.3, (n, 3))         # angular velocity
    forces = rng.normal(0, 5.0, (n, 6, 3))
    forces[:, :, 2] = rng.uniform(-20, 150, (n, 6))  # some contacts unloaded (f_z <= 0 is skipped)
    forces[3, :, 2] = -1.0                           # a cycle with no contact at all: denom ~ 0 -> estimate kept
    cyc[:, 8:26] = forces.reshape(n, 18)
    cyc[:, 26:29] = rng.normal(0, 1.0, (n, 3))
    out = ref_runner.run_estimator(cyc)
    np.savez_compressed(os.path.join(HERE, "fb_estimator.npz.dat"), cycles=cyc, ref_out=out)
    os.replace(os.path.join(HERE, "fb_estimator.npz.dat.npz"), os.path.join(HERE, "fb_estimator.dat"))
    print("fb_estimator", out[-1])


if __name__ == "__main__":
    estimator_golden()
    cmd_golden()
    main()
