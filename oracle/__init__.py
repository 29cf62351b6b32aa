"""Python access to the CPU oracle.  TEST INFRASTRUCTURE ONLY.

  liboracle.so  mppi_oracle.c   FP64 literal restatement of the reference solve
  libtwin.so    mppi_twin.cpp   FP32 twin built from the product's mppi_math.h (bit-exact GPU contract)
  _ref/         unmodified reference TUs against stub headers (only when /root/reference was present at build)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
Nothing in ccv_mppi_path_tracker_b200/ does.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_P = C.POINTER


class OracleParams(C.Structure):
    _fields_ = [
        ("control_noise", C.c_double), ("lambda_", C.c_double), ("v_ref", C.c_double), ("resolution", C.c_double),
        ("u_min", C.c_double * 5), ("u_max", C.c_double * 5),
        ("path_weight", C.c_double), ("v_weight", C.c_double), ("zmp_weight", C.c_double),
        ("roll_v_weight", C.c_double), ("back_weight", C.c_double), ("yaw_weight", C.c_double),
        ("steer_off", C.c_int32), ("reserved", C.c_int32),
    ]


class OracleOutputs(C.Structure):
    _fields_ = [("controls", _P(C.c_double)), ("states", _P(C.c_double)), ("zmp", _P(C.c_double)),
                ("cost", _P(C.c_double)), ("nearest", _P(C.c_int)), ("weights", _P(C.c_double)),
                ("window", _P(C.c_double)), ("stats", _P(C.c_double)), ("current_index", _P(C.c_int))]


_oracle = None
_twin = None


def _load_oracle():
    global _oracle
    if _oracle is None:
        lib = C.CDLL(os.path.join(_HERE, "liboracle.so"))
        lib.oracle_solve.restype = C.c_int
        lib.oracle_solve.argtypes = [C.c_int, _P(OracleParams), C.c_int, C.c_int, _P(C.c_double), C.c_double,
                                     _P(C.c_double), C.c_int, _P(C.c_float), _P(C.c_double), C.c_int, C.c_int,
                                     _P(OracleOutputs)]
        lib.oracle_time_solves.restype = C.c_double
        lib.oracle_time_solves.argtypes = [C.c_int, _P(OracleParams), C.c_int, C.c_int, _P(C.c_double), C.c_double,
                                           _P(C.c_double), C.c_int, C.c_int, C.c_int, C.c_int]
        lib.oracle_calc_ref_path.restype = C.c_int
        lib.oracle_calc_ref_path.argtypes = [C.c_double, C.c_double, _P(C.c_double), C.c_int, C.c_double, C.c_double,
                                             C.c_double, C.c_int, _P(C.c_double), _P(C.c_double), _P(C.c_double)]
        lib.oracle_current_index.restype = C.c_int
        lib.oracle_current_index.argtypes = [C.c_double, C.c_double, _P(C.c_double), C.c_int]
        lib.oracle_min_distance.restype = C.c_double
        lib.oracle_min_distance.argtypes = [C.c_double, C.c_double, _P(C.c_double), _P(C.c_double), C.c_int, _P(C.c_int)]
        lib.oracle_zmp_from_model.restype = None
        lib.oracle_zmp_from_model.argtypes = [_P(C.c_double)] * 4
        lib.oracle_make_sin_path.restype = C.c_int
        lib.oracle_make_sin_path.argtypes = [C.c_double] * 13 + [_P(C.c_double), C.c_int]
        _oracle = lib
    return _oracle


def _load_twin():
    global _twin
    if _twin is None:
        lib = C.CDLL(os.path.join(_HERE, "libtwin.so"))
        lib.twin_rollout_cost.restype = C.c_int
        lib.twin_rollout_cost.argtypes = [C.c_int, _P(OracleParams), C.c_int, C.c_int, _P(C.c_double), C.c_double,
                                          _P(C.c_double), _P(C.c_float), _P(C.c_double), _P(C.c_float), _P(C.c_int),
                                          _P(C.c_float), _P(C.c_float), _P(C.c_float), _P(C.c_float)]
        lib.twin_atan2.restype = None
        lib.twin_atan2.argtypes = [_P(C.c_float), _P(C.c_float), C.c_int, _P(C.c_float)]
        lib.twin_sincos.restype = None
        lib.twin_sincos.argtypes = [_P(C.c_float), C.c_int, _P(C.c_float), _P(C.c_float)]
        lib.twin_sincos_increment.restype = None
        lib.twin_sincos_increment.argtypes = [_P(C.c_float), C.c_int, _P(C.c_float), _P(C.c_float)]
        lib.twin_heading.restype = None
        lib.twin_heading.argtypes = [C.c_float, _P(C.c_float), C.c_int, _P(C.c_float), _P(C.c_float), _P(C.c_float)]
        lib.twin_clamp.restype = None
        lib.twin_clamp.argtypes = [_P(C.c_float), C.c_int, C.c_float, C.c_float, _P(C.c_float)]
        _twin = lib
    return _twin


def _d(a):
    return a.ctypes.data_as(_P(C.c_double))


def _f(a):
    return a.ctypes.data_as(_P(C.c_float))


def make_params(sp):
    """sp: dict in mppi_params field order (ccv_mppi_path_tracker_b200.params.solve_params)."""
    p = OracleParams()
    for k, v in sp.items():
        if k in ("u_min", "u_max"):
            setattr(p, k, (C.c_double * 5)(*v))
        else:
            setattr(p, k, v)
    return p


MODEL_ID = {"diff_drive": 0, "steering": 1, "full_body": 2}
NUM_CONTROLS = {"diff_drive": 2, "steering": 3, "full_body": 5}
NUM_STATES = {"diff_drive": 3, "steering": 3, "full_body": 5}


def solve(model, sp, K, T, state, dt, path_xy, eps, u_nominal, shifted=True, nthreads=1, want=("cost", "nearest", "weights", "window", "stats")):
    """One FP64 oracle solve.  eps [T-1][K][U] float32; returns dict with u_new [T-1][U] plus requested taps."""
    lib = _load_oracle()
    U, S = NUM_CONTROLS[model], NUM_STATES[model]
    state = np.ascontiguousarray(state, dtype=np.float64).reshape(S)
    path_xy = np.ascontiguousarray(path_xy, dtype=np.float64).reshape(-1, 2)
    eps = np.ascontiguousarray(eps, dtype=np.float32).reshape(T - 1, K, U)
    u = np.array(u_nominal, dtype=np.float64).reshape(T - 1, U).copy()
    bufs = {}
    out = OracleOutputs()
    shapes = dict(controls=((K, T - 1, U), np.float64), states=((K, T, S), np.float64),
                  zmp=((K, max(T - 2, 0), 2), np.float64), cost=((K,), np.float64), nearest=((K, T), np.int32),
                  weights=((K,), np.float64), window=((T, 3), np.float64), stats=((3,), np.float64),
                  current_index=((1,), np.int32))
    for name in want:
        shp, dt_ = shapes[name]
        bufs[name] = np.zeros(shp, dtype=dt_)
        ptr_t = _P(C.c_int) if dt_ == np.int32 else _P(C.c_double)
        setattr(out, name, bufs[name].ctypes.data_as(ptr_t))
    p = make_params(sp)
    rc = lib.oracle_solve(MODEL_ID[model], C.byref(p), K, T, _d(state), float(dt), _d(path_xy), path_xy.shape[0],
                          _f(eps), _d(u), int(bool(shifted)), int(nthreads), C.byref(out))
    if rc != 0:
        raise RuntimeError(f"oracle_solve failed: {rc}")
    bufs["u_new"] = u
    return bufs


def time_solves(model, sp, K, T, state, dt, path_xy, n_solves, literal_copies=False, nthreads=1):
    lib = _load_oracle()
    S = NUM_STATES[model]
    state = np.ascontiguousarray(state, dtype=np.float64).reshape(S)
    path_xy = np.ascontiguousarray(path_xy, dtype=np.float64).reshape(-1, 2)
    p = make_params(sp)
    return lib.oracle_time_solves(MODEL_ID[model], C.byref(p), K, T, _d(state), float(dt), _d(path_xy),
                                  path_xy.shape[0], int(n_solves), int(bool(literal_copies)), int(nthreads))


def calc_ref_path(path_xy, px, py, v_ref, dt, resolution, T):
    lib = _load_oracle()
    path_xy = np.ascontiguousarray(path_xy, dtype=np.float64).reshape(-1, 2)
    xr, yr, yawr = (np.zeros(T) for _ in range(3))
    cur = lib.oracle_calc_ref_path(px, py, _d(path_xy), path_xy.shape[0], v_ref, dt, resolution, T, _d(xr), _d(yr), _d(yawr))
    return np.stack([xr, yr, yawr], axis=1), cur


def min_distance(x, y, x_ref, y_ref):
    lib = _load_oracle()
    xr = np.ascontiguousarray(x_ref, dtype=np.float64)
    yr = np.ascontiguousarray(y_ref, dtype=np.float64)
    arg = C.c_int(-1)
    d = lib.oracle_min_distance(x, y, _d(xr), _d(yr), xr.shape[0], C.byref(arg))
    return d, arg.value


def zmp_from_model(com, accel, hgdot):
    lib = _load_oracle()
    a, b, c = (np.ascontiguousarray(v, dtype=np.float64) for v in (com, accel, hgdot))
    out = np.zeros(3)
    lib.oracle_zmp_from_model(_d(a), _d(b), _d(c), _d(out))
    return out


def make_sin_path(course_length=10.0, resolution=0.1, A1=0.0, omega1=0.0, delta1=1.57, A2=0.0, omega2=0.0,
                  delta2=1.57, A3=0.0, omega3=0.0, delta3=1.57, init_x=0.0, init_y=0.0):
    lib = _load_oracle()
    cap = int(course_length / resolution) + 16
    xy = np.zeros((cap, 2))
    n = lib.oracle_make_sin_path(course_length, resolution, A1, omega1, delta1, A2, omega2, delta2, A3, omega3, delta3,
                                 init_x, init_y, _d(xy), cap)
    return xy[:n].copy()


def twin_rollout_cost(model, sp, K, T, state, dt, window, eps, u_nominal, want=("nearest",)):
    """FP32 twin of the rollout+cost kernel.  window [T][3] float64 (absolute), eps [T-1][K][U] float32."""
    lib = _load_twin()
    U, S = NUM_CONTROLS[model], NUM_STATES[model]
    state = np.ascontiguousarray(state, dtype=np.float64).reshape(S)
    window = np.ascontiguousarray(window, dtype=np.float64).reshape(T, 3)
    eps = np.ascontiguousarray(eps, dtype=np.float32).reshape(T - 1, K, U)
    u = np.ascontiguousarray(u_nominal, dtype=np.float64).reshape(T - 1, U)
    out = {"cost": np.zeros(K, dtype=np.float32)}
    shapes = dict(nearest=((K, T), np.int32), d2=((K, T), np.float32), states=((K, T, 5), np.float32),
                  zmp=((K, T, 2), np.float32), controls=((K, T - 1, U), np.float32))
    ptrs = {}
    for name, (shp, dt_) in shapes.items():
        if name in want:
            out[name] = np.zeros(shp, dtype=dt_)
            ptrs[name] = out[name].ctypes.data_as(_P(C.c_int) if dt_ == np.int32 else _P(C.c_float))
        else:
            ptrs[name] = None
    p = make_params(sp)
    rc = lib.twin_rollout_cost(MODEL_ID[model], C.byref(p), K, T, _d(state), float(dt), _d(window), _f(eps), _d(u),
                               _f(out["cost"]), ptrs["nearest"], ptrs["d2"], ptrs["states"], ptrs["zmp"], ptrs["controls"])
    if rc != 0:
        raise RuntimeError(f"twin_rollout_cost failed: {rc}")
    return out


def twin_sincos(a):
    lib = _load_twin()
    a = np.ascontiguousarray(a, dtype=np.float32)
    s = np.zeros_like(a)
    c = np.zeros_like(a)
    lib.twin_sincos(_f(a), a.size, _f(s), _f(c))
    return s, c


def twin_atan2(y, x):
    lib = _load_twin()
    y = np.ascontiguousarray(y, dtype=np.float32)
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.zeros_like(y)
    lib.twin_atan2(_f(y), _f(x), y.size, _f(out))
    return out


def twin_sincos_increment(a):
    lib = _load_twin()
    a = np.ascontiguousarray(a, dtype=np.float32)
    s = np.zeros_like(a)
    c = np.zeros_like(a)
    lib.twin_sincos_increment(_f(a), a.size, _f(s), _f(c))
    return s, c


def twin_heading(angle0, increments):
    """(cos, sin) carried by the contract's rotation recurrence, and the plainly accumulated FP32 angle."""
    lib = _load_twin()
    inc = np.ascontiguousarray(increments, dtype=np.float32)
    c, s, a = (np.zeros_like(inc) for _ in range(3))
    lib.twin_heading(float(angle0), _f(inc), inc.size, _f(c), _f(s), _f(a))
    return c, s, a


def twin_clamp(v, lo, hi):
    lib = _load_twin()
    v = np.ascontiguousarray(v, dtype=np.float32)
    out = np.zeros_like(v)
    lib.twin_clamp(_f(v), v.size, float(lo), float(hi), _f(out))
    return out
