"""Host-side mirror of the reference controller classes over the C ABI (Python flavour, used by tests and bench).

Same class names, parameter names/defaults and per-cycle method names as the reference nodes
(include/ccv_mppi_path_tracker/{diff_drive,steering_diff_drive,full_body}_mppi.h); ROS I/O is replaced by plain
attributes: `set_path()` stands for pathCallback, `current_state` for get_Transform / get_CurrentState, and
`solve()` for one pass of run()'s `sampling(); predict_States(); calc_Weights(); determine_OptimalSolution();`.
The C++ equivalents used by the harness live in csrc/host/controllers.hpp.  All compute happens in
libmppi_b200.so on the GPU; nothing here falls back to the CPU.
"""
import ctypes as C

import numpy as np

from . import _capi
from .params import MODEL_ID, NUM_CONTROLS, NUM_STATES, node_params, solve_params


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class _MPPIBase:
    MODEL = None

    def __init__(self, launch=True, n_robots=1, device=0, **overrides):
        self.lib = _capi.load()
        self.p = node_params(self.MODEL, launch=launch, **overrides)
        self.model = self.MODEL
        self.horizon_ = int(self.p["horizon"])
        self.num_samples_ = int(self.p["num_samples"])
        self.dt_ = float(self.p["dt"])
        self.U = NUM_CONTROLS[self.MODEL]
        self.S = NUM_STATES[self.MODEL]
        self.n_robots = int(n_robots)
        self._h = C.c_void_p()
        cp = self._cparams()
        rc = self.lib.mppi_create(C.byref(self._h), MODEL_ID[self.MODEL], C.byref(cp), self.num_samples_,
                                  self.horizon_, self.n_robots, device)
        if rc != 0:
            raise _capi.MppiError(rc, self.lib.mppi_last_error(None).decode())
        # optimal_solution (DDh:100): [n_robots][T-1][U], zero-initialised like RobotStates::init
        self.optimal_solution = np.zeros((self.n_robots, self.horizon_ - 1, self.U), dtype=np.float64)
        self.current_state = np.zeros((self.n_robots, self.S), dtype=np.float64)

    # -- plumbing ---------------------------------------------------------------------------------------
    def _cparams(self):
        sp = solve_params(self.MODEL, self.p)
        cp = _capi.MppiParams()
        for k, v in sp.items():
            if k in ("u_min", "u_max"):
                setattr(cp, k, (C.c_double * 5)(*v))
            else:
                setattr(cp, k, v)
        return cp

    def _check(self, rc):
        if rc != 0:
            raise _capi.MppiError(rc, self.lib.mppi_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.mppi_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- inputs -----------------------------------------------------------------------------------------
    def set_param(self, **kw):
        self.p.update(kw)
        cp = self._cparams()
        self._check(self.lib.mppi_set_params(self._h, C.byref(cp)))

    def set_path(self, xy, robot=0):
        """pathCallback: the full reference path (N x 2)."""
        xy = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2)
        self._check(self.lib.mppi_set_path(self._h, robot, _dptr(xy), xy.shape[0]))

    def set_window(self, window, robot=0):
        w = np.ascontiguousarray(window, dtype=np.float64).reshape(self.horizon_, 3)
        self._check(self.lib.mppi_set_window(self._h, robot, _dptr(w)))

    def set_seed(self, seed, counter=0):
        self._check(self.lib.mppi_set_seed(self._h, seed, counter))

    def set_shard(self, sample_offset, num_samples_global, robot_offset=0):
        self._check(self.lib.mppi_set_shard(self._h, sample_offset, num_samples_global, robot_offset))

    def set_noise(self, eps):
        """eps: [n_robots][T-1][K][U] float32 standard normals, or None for the internal Philox stream."""
        if eps is None:
            self._check(self.lib.mppi_set_noise(self._h, None))
            return
        e = np.ascontiguousarray(eps, dtype=np.float32).reshape(self.n_robots, self.horizon_ - 1, self.num_samples_, self.U)
        self._check(self.lib.mppi_set_noise(self._h, _fptr(e)))

    def set_debug(self, flags):
        self._check(self.lib.mppi_set_debug(self._h, flags))

    def set_scan_mode(self, mode):
        self._check(self.lib.mppi_set_scan_mode(self._h, mode))

    def set_window_builder(self, mode):
        self._check(self.lib.mppi_set_window_builder(self._h, mode))

    def set_option(self, option, value):
        """mppi_set_option: tuning / behaviour switches (_capi.OPT_*)."""
        self._check(self.lib.mppi_set_option(self._h, int(option), float(value)))

    def get_option(self, option):
        v = C.c_double(0.0)
        self._check(self.lib.mppi_get_option(self._h, int(option), C.byref(v)))
        return v.value

    def set_stream(self, cuda_stream_ptr):
        self._check(self.lib.mppi_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def use_graph(self, enable=True):
        self._check(self.lib.mppi_use_graph(self._h, int(enable)))

    def comm_init(self, unique_id: bytes, rank, n_ranks):
        buf = C.create_string_buffer(unique_id, _capi.COMM_ID_BYTES)
        self._check(self.lib.mppi_comm_init(self._h, buf, rank, n_ranks))

    def comm_export(self, n_ranks) -> bytes:
        """NVLink peer exchange, step 1: this rank's exchange-buffer IPC handle (all-gather these across ranks)."""
        buf = C.create_string_buffer(_capi.IPC_HANDLE_BYTES)
        self._check(self.lib.mppi_comm_export(self._h, int(n_ranks), buf))
        return buf.raw

    def comm_connect(self, handles: bytes, rank, n_ranks):
        """NVLink peer exchange, step 2: map every peer's buffer (handles = all ranks' handles concatenated)."""
        buf = C.create_string_buffer(handles, _capi.IPC_HANDLE_BYTES * n_ranks)
        self._check(self.lib.mppi_comm_connect(self._h, buf, rank, n_ranks))

    # -- the cycle --------------------------------------------------------------------------------------
    def solve(self, state=None, dt=None):
        """sampling(); predict_States(); calc_Weights(); determine_OptimalSolution();  (DD:352-358)
        Returns optimal_solution ([T-1][U] for one robot, else [R][T-1][U])."""
        if state is not None:
            self.current_state[...] = np.asarray(state, dtype=np.float64).reshape(self.n_robots, self.S)
        if dt is not None:
            self.dt_ = float(dt)
        self._check(self.lib.mppi_solve(self._h, _dptr(self.current_state), self.dt_, _dptr(self.optimal_solution)))
        return self.optimal_solution[0] if self.n_robots == 1 else self.optimal_solution

    def upload(self, state=None, dt=None, with_nominal=True):
        if state is not None:
            self.current_state[...] = np.asarray(state, dtype=np.float64).reshape(self.n_robots, self.S)
        if dt is not None:
            self.dt_ = float(dt)
        nom = _dptr(self.optimal_solution) if with_nominal else None
        self._check(self.lib.mppi_upload(self._h, _dptr(self.current_state), self.dt_, nom))

    def enqueue(self):
        self._check(self.lib.mppi_enqueue(self._h))

    def download(self):
        self._check(self.lib.mppi_download(self._h, _dptr(self.optimal_solution)))
        return self.optimal_solution[0] if self.n_robots == 1 else self.optimal_solution

    def synchronize(self):
        self._check(self.lib.mppi_synchronize(self._h))

    # -- outputs beyond the controls ----------------------------------------------------------------------
    def costs(self, robot=0):
        out = np.empty(self.num_samples_, dtype=np.float32)
        self._check(self.lib.mppi_get_costs(self._h, robot, _fptr(out)))
        return out

    def weights(self, robot=0):
        out = np.empty(self.num_samples_, dtype=np.float32)
        self._check(self.lib.mppi_get_weights(self._h, robot, _fptr(out)))
        return out

    def nearest(self, robot=0):
        out = np.empty((self.num_samples_, self.horizon_), dtype=np.int32)
        self._check(self.lib.mppi_get_nearest(self._h, robot, out.ctypes.data_as(C.POINTER(C.c_int32))))
        return out

    def states(self, robot=0):
        """Candidate trajectories [K][T][S] (world frame) of the last solve; needs set_debug(DEBUG_STATES)."""
        out = np.empty((self.num_samples_, self.horizon_, self.S), dtype=np.float64)
        self._check(self.lib.mppi_get_states(self._h, robot, _dptr(out)))
        return out

    def optimal_path(self, robot=0):
        """publish_OptimalPath (DD:295-312): optimal_solution re-rolled from the current pose with predict_NextState,
        in double on the host like the reference; returns [T][3] (x, y, yaw)."""
        T = self.horizon_
        st = np.zeros((T, 3))
        st[0] = self.current_state[robot, :3]
        u = self.optimal_solution[robot]
        for t in range(T - 1):
            heading = st[t, 2] if self.MODEL == "diff_drive" else st[t, 2] + u[t, 2]
            st[t + 1, 0] = st[t, 0] + u[t, 0] * np.cos(heading) * self.dt_
            st[t + 1, 1] = st[t, 1] + u[t, 0] * np.sin(heading) * self.dt_
            st[t + 1, 2] = st[t, 2] + u[t, 1] * self.dt_
        return st

    def noise(self, robot=0):
        out = np.empty((self.horizon_ - 1, self.num_samples_, self.U), dtype=np.float32)
        self._check(self.lib.mppi_get_noise(self._h, robot, _fptr(out)))
        return out

    def window(self, robot=0):
        out = np.empty((self.horizon_, 3), dtype=np.float64)
        idx = C.c_int(0)
        self._check(self.lib.mppi_get_window(self._h, robot, _dptr(out), C.byref(idx)))
        return out, idx.value

    def stats(self, robot=0):
        out = np.empty(3, dtype=np.float64)
        self._check(self.lib.mppi_get_stats(self._h, robot, _dptr(out)))
        return dict(c_min=out[0], sum_w=out[1], ess=out[2])

    def record(self, robot=0):
        """This rank's partial {c_min, sum w, sum w^2, 0, N[(T-1)*U]} of the last solve (what the collective exchanges)."""
        out = np.empty(4 + (self.horizon_ - 1) * self.U, dtype=np.float32)
        self._check(self.lib.mppi_get_record(self._h, robot, _fptr(out)))
        return out

    def time_kernels(self, n_iters=5):
        """Average device ms per kernel: noise, rollout_cost, weights, weighted_controls, finalize, merge, total."""
        ms = np.zeros(8, dtype=np.float32)
        self._check(self.lib.mppi_time_kernels(self._h, int(n_iters), _fptr(ms)))
        names = ("noise", "rollout_cost", "weights", "weighted_controls", "finalize", "merge", "total", "candidate_grid")
        return dict(zip(names, (float(v) for v in ms)))

    def io_bytes(self):
        """(host->device, device->host) bytes moved by one solve()."""
        a, b = C.c_size_t(0), C.c_size_t(0)
        self._check(self.lib.mppi_get_io_bytes(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def launch_count(self):
        return self.lib.mppi_last_launch_count(self._h)

    # first control of the horizon = what publish_CmdVel sends (DD:250-251)
    def cmd_vel(self, robot=0):
        return float(self.optimal_solution[robot, 0, 0]), float(self.optimal_solution[robot, 0, 1])

    TREAD = 0.501  # tread_ (diff_drive_mppi.h:97)

    def _steer_in_out(self, v, w, delta):
        """Steering angles of the inner / outer wheel (steering_diff_drive_mppi.cpp:273-296, full_body_mppi.cpp:
        246-262): R = |v / w| -- inf for w = 0, NaN for v = w = 0, exactly as the reference computes it."""
        with np.errstate(divide="ignore", invalid="ignore"):
            R = np.abs(np.float64(v) / np.float64(w))
            s_in = np.arctan2(R * np.sin(delta), R * np.cos(delta) - self.TREAD / 2.0)
            s_out = np.arctan2(R * np.sin(delta), R * np.cos(delta) + self.TREAD / 2.0)
        return (float(s_in), float(s_out)) if w > 0.0 else (float(s_out), float(s_in))

    def cmd_pos(self, robot=0):
        """What publish_CmdPos sends (ccv_dynamixel_msgs/CmdPoseByRadian): dict(steer_l, steer_r, fore, rear, roll).
        DiffDriveMPPI: diff_drive_mppi.cpp:255-263."""
        po = float(self.p["pitch_offset"])
        return dict(steer_l=0.0, steer_r=0.0, fore=po, rear=po, roll=0.0)


class DiffDriveMPPI(_MPPIBase):
    """class DiffDriveMPPI (diff_drive_mppi.h:52): unicycle, controls (v, w)."""
    MODEL = "diff_drive"


class SteeringDiffDriveMPPI(_MPPIBase):
    """class SteeringDiffDriveMPPI (steering_diff_drive_mppi.h:56): controls (v, w, steer)."""
    MODEL = "steering"

    def cmd_pos(self, robot=0):
        """steering_diff_drive_mppi.cpp:273-296"""
        u = self.optimal_solution[robot, 0]
        l, r = self._steer_in_out(u[0], u[1], u[2])
        po = float(self.p["pitch_offset"])
        return dict(steer_l=l, steer_r=r, fore=po, rear=po, roll=0.0)


class FullBodyMPPI(_MPPIBase):
    """class FullBodyMPPI (full_body_mppi.h:68): controls (v, w, direction, roll_v, pitch_v), ZMP cost."""
    MODEL = "full_body"

    # the node's ZMP monitors (topics zmp_y / true_zmp, full_body_mppi.cpp:628-633); host-side, not inputs of the solve
    MASS, BODY_H, BODY_D, BODY_W, ALPHA = 60.0, 0.8075, 0.208, 0.208, 0.3  # full_body_mppi.h:213-218
    CONTACTS = np.array([[0.0, 0.225, 0.075], [0.0, -0.225, 0.075], [0.245, 0.167, -0.003], [0.245, -0.167, -0.004],
                         [-0.245, -0.167, -0.004], [-0.245, 0.167, -0.003]])  # full_body_mppi.cpp:57-63

    def cmd_pos(self, robot=0):
        """full_body_mppi.cpp:246-275: steer_off zeroes the steering, the roll command is the current roll advanced by
        roll_v[0] * dt, clamped to [roll_min, roll_max], and zeroed by roll_off."""
        u = self.optimal_solution[robot, 0]
        l, r = (0.0, 0.0) if self.p.get("steer_off", False) else self._steer_in_out(u[0], u[1], u[2])
        roll = float(self.current_state[robot, 3] + u[3] * self.dt_)
        if roll > self.p["roll_max"]:
            roll = float(self.p["roll_max"])
        elif roll < self.p["roll_min"]:
            roll = float(self.p["roll_min"])
        if self.p.get("roll_off", False):
            roll = 0.0
        po = float(self.p["pitch_offset"])
        return dict(steer_l=l, steer_r=r, fore=po, rear=po, roll=roll)

    def reset_monitors(self):
        self.zmp_x_, self.zmp_y_ = 0.0, 0.0
        self.true_ZMP = np.zeros(3)
        self.last_HG = np.zeros(3)

    def computeZMPfromModel(self, CoM, accel, HGdot):
        """full_body_mppi.cpp:597-603"""
        g = np.array([0.0, 0.0, -9.8])
        z = np.array([0.0, 0.0, 1.0])
        M_O = np.cross(CoM, self.MASS * g) - np.cross(CoM, self.MASS * np.asarray(accel)) - np.asarray(HGdot)
        return np.cross(z, M_O) / (self.MASS * np.dot(g - np.asarray(accel), z))

    def update_model_zmp(self, imu_roll, imu_pitch, accel_x, accel_y, omega, dt=None):
        """The ZMP part of get_CurrentState (full_body_mppi.cpp:551-566): low-passed model ZMP."""
        if not hasattr(self, "last_HG"):
            self.reset_monitors()
        dt = self.dt_ if dt is None else dt
        m, b = self.MASS, self.BODY_H / 2
        I = np.array([(m * (self.BODY_W ** 2 + self.BODY_H ** 2)) / 12 + m * b * b,
                      (m * (self.BODY_H ** 2 + self.BODY_D ** 2)) / 12 + m * b * b,
                      (m * (self.BODY_D ** 2 + self.BODY_W ** 2)) / 12])
        CoM = np.array([b * np.sin(imu_pitch), -b * np.sin(imu_roll), b * np.cos(imu_pitch) * np.cos(imu_roll)])
        HG = I * np.asarray(omega, dtype=np.float64)
        HGdot = (HG - self.last_HG) / dt
        self.last_HG = HG
        Z = self.computeZMPfromModel(CoM, [accel_x, accel_y, 0.0], HGdot)
        self.zmp_x_ = self.ALPHA * Z[0] + (1 - self.ALPHA) * self.zmp_x_
        self.zmp_y_ = self.ALPHA * Z[1] + (1 - self.ALPHA) * self.zmp_y_
        return self.zmp_x_, self.zmp_y_

    def calc_true_ZMP(self, forces):
        """full_body_mppi.cpp:568-596: forces [6][3] (wheels l, r, casters fl, fr, bl, br) in the base frame."""
        if not hasattr(self, "true_ZMP"):
            self.reset_monitors()
        f = np.asarray(forces, dtype=np.float64).reshape(6, 3)
        sumF, sumM = np.zeros(3), np.zeros(3)
        for r, fi in zip(self.CONTACTS, f):
            if fi[2] > 0.0:
                sumF += fi
                sumM += np.cross(r, fi)
        n = np.array([0.0, 0.0, 1.0])
        denom = np.dot(sumF, n)
        if abs(denom) < 1e-6:
            return self.true_ZMP
        self.true_ZMP = self.ALPHA * (np.cross(n, sumM) / denom) + (1 - self.ALPHA) * self.true_ZMP
        return self.true_ZMP


CONTROLLERS = {"diff_drive": DiffDriveMPPI, "steering": SteeringDiffDriveMPPI, "full_body": FullBodyMPPI}


def comm_unique_id():
    lib = _capi.load()
    buf = C.create_string_buffer(_capi.COMM_ID_BYTES)
    rc = lib.mppi_comm_get_unique_id(buf)
    if rc != 0:
        raise _capi.MppiError(rc, lib.mppi_last_error(None).decode())
    return buf.raw


def merge_partials(partials, lambda_):
    """Host merge of per-rank records [G][4+n] -> (u[n], stats) -- same arithmetic as the device merge kernel."""
    lib = _capi.load()
    p = np.ascontiguousarray(partials, dtype=np.float32)
    G, rec = p.shape
    n = rec - 4
    u = np.empty(n, dtype=np.float32)
    st = np.empty(3, dtype=np.float64)
    rc = lib.mppi_merge_partials(_fptr(p), G, n, float(lambda_), _fptr(u), _dptr(st))
    if rc != 0:
        raise _capi.MppiError(rc, "mppi_merge_partials")
    return u, dict(c_min=st[0], sum_w=st[1], ess=st[2])


def calc_ref_path(path_xy, px, py, v_ref, dt, resolution, horizon):
    lib = _capi.load()
    xy = np.ascontiguousarray(path_xy, dtype=np.float64).reshape(-1, 2)
    win = np.empty((horizon, 3), dtype=np.float64)
    idx = C.c_int(0)
    rc = lib.mppi_calc_ref_path(_dptr(xy), xy.shape[0], px, py, v_ref, dt, resolution, horizon, _dptr(win), C.byref(idx))
    if rc != 0:
        raise _capi.MppiError(rc, "mppi_calc_ref_path")
    return win, idx.value


def philox4x32_10(counter, key):
    lib = _capi.load()
    c = (C.c_uint32 * 4)(*counter)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib.mppi_philox4x32_10(c, k, o)
    return [int(v) for v in o]
