"""Run one MPPI cycle of the UNMODIFIED reference node (oracle/_ref/ref_{dd,sd,fb}, built by oracle/ref_shim).
TEST INFRASTRUCTURE ONLY: used by tests/test_oracle_vs_ref.py and tests/golden/make_golden.py in the build
container, where /root/reference is mounted.  The GPU box only sees the golden vectors made from it."""
import os
import struct
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
BIN = {"diff_drive": "ref_dd", "steering": "ref_sd", "full_body": "ref_fb"}
NUM_CONTROLS = {"diff_drive": 2, "steering": 3, "full_body": 5}
NUM_STATES = {"diff_drive": 3, "steering": 3, "full_body": 5}


def available():
    return all(os.path.exists(os.path.join(_HERE, "_ref", b)) for b in BIN.values())


def _write_input(fin, model, node_params, K, T, state, dt, path_xy, eps, u_nominal):
    U, S = NUM_CONTROLS[model], NUM_STATES[model]
    p = {("lambda" if k == "lambda_" else k): float(v) for k, v in node_params.items()}
    p["horizon"] = float(T)
    p["num_samples"] = float(K)
    path_xy = np.ascontiguousarray(path_xy, dtype=np.float64).reshape(-1, 2)
    st = np.zeros(5)
    st[:S] = np.asarray(state, dtype=np.float64).reshape(S)
    eps = np.ascontiguousarray(eps, dtype=np.float32).reshape(T - 1, K, U)
    u0 = np.ascontiguousarray(u_nominal, dtype=np.float64).reshape(T - 1, U)
    with open(fin, "wb") as f:
        f.write(struct.pack("<4i", K, T, path_xy.shape[0], len(p)))
        for k, v in p.items():
            f.write(k.encode()[:31].ljust(32, b"\0"))
            f.write(struct.pack("<d", v))
        f.write(st.tobytes())
        f.write(struct.pack("<d", float(dt)))
        f.write(path_xy.tobytes())
        f.write(u0.tobytes())
        f.write(eps.tobytes())


def timing_available(model=None):
    names = [BIN[model]] if model else list(BIN.values())
    return all(os.path.exists(os.path.join(_HERE, "_ref", b + "_time")) for b in names)


def time_solves(model, node_params, K, T, state, dt, path_xy, u_nominal, n_solves):
    """Seconds of each of n_solves consecutive cycles of the UNMODIFIED reference node (its own sampling(),
    predict_States(), calc_Weights(), determine_OptimalSolution(); -O2 build, one thread like the node)."""
    U = NUM_CONTROLS[model]
    with tempfile.TemporaryDirectory() as td:
        fin = os.path.join(td, "in.bin")
        _write_input(fin, model, node_params, K, T, state, dt, path_xy, np.zeros((T - 1, K, U), np.float32), u_nominal)
        r = subprocess.run([os.path.join(_HERE, "_ref", BIN[model] + "_time"), "--time", fin, str(int(n_solves))],
                           check=True, stderr=subprocess.DEVNULL, stdout=subprocess.PIPE, text=True)
    return [float(x) for x in r.stdout.split()]


def run(model, node_params, K, T, state, dt, path_xy, eps, u_nominal):
    """node_params: the node's ROS parameters by their reference names (ccv_mppi_path_tracker_b200.params.node_params
    with `lambda_` -> `lambda`); horizon / num_samples are overridden by T / K."""
    U, S = NUM_CONTROLS[model], NUM_STATES[model]
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.bin"), os.path.join(td, "out.bin")
        _write_input(fin, model, node_params, K, T, state, dt, path_xy, eps, u_nominal)
        subprocess.run([os.path.join(_HERE, "_ref", BIN[model]), fin, fout], check=True, stderr=subprocess.DEVNULL)
        raw = open(fout, "rb").read()
    off = 0

    def take(n, dtype=np.float64):
        nonlocal off
        a = np.frombuffer(raw, dtype=dtype, count=n, offset=off)
        off += a.nbytes
        return a.copy()

    out = {"yaw_used": float(take(1)[0]), "window": take(3 * T).reshape(T, 3), "cost": take(K), "weights": take(K),
           "u_new": take((T - 1) * U).reshape(T - 1, U), "states": take(K * T * S).reshape(K, T, S)}
    if model == "full_body":
        out["zmp"] = take(K * max(T - 2, 0) * 2).reshape(K, max(T - 2, 0), 2)
    out["current_index"] = int(take(1, np.int32)[0])
    return out


def run_estimator(cycles):
    """cycles: [n][30] doubles {dt, imu_roll, imu_pitch, accel_x, accel_y, omega[3], forces[6][3], x, y, yaw, 0};
    returns [n][5] {zmp_x, zmp_y, true_ZMP xyz} from the reference's calc_true_ZMP() + get_CurrentState()."""
    cycles = np.ascontiguousarray(cycles, dtype=np.float64).reshape(-1, 30)
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.bin"), os.path.join(td, "out.bin")
        with open(fin, "wb") as f:
            f.write(struct.pack("<i", cycles.shape[0]))
            f.write(cycles.tobytes())
        subprocess.run([os.path.join(_HERE, "_ref", BIN["full_body"]), "--estimator", fin, fout], check=True,
                       stderr=subprocess.DEVNULL)
        return np.fromfile(fout, dtype=np.float64).reshape(-1, 5)


def run_cmd(model, cases):
    """cases: [n][8] doubles {v0, w0, steer0 / direction0, roll_v0, roll_state, dt, steer_off, roll_off}; returns [n][7]
    {cmd_vel.linear.x, cmd_vel.angular.z, steer_l, steer_r, fore, rear, roll} from the reference's own
    publish_CmdVel() + publish_CmdPos() (default constructor parameters)."""
    cases = np.ascontiguousarray(cases, dtype=np.float64).reshape(-1, 8)
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.bin"), os.path.join(td, "out.bin")
        with open(fin, "wb") as f:
            f.write(struct.pack("<i", cases.shape[0]))
            f.write(cases.tobytes())
        subprocess.run([os.path.join(_HERE, "_ref", BIN[model]), "--cmd", fin, fout], check=True, stderr=subprocess.DEVNULL)
        return np.fromfile(fout, dtype=np.float64).reshape(-1, 7)
