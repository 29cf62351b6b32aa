"""ctypes binding of the C ABI in include/mppi_b200.h (libmppi_b200.so, built in-tree by `make lib`).

There is no Python or CPU fallback: if the shared library is missing the import of this module raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MPPI_B200_LIB") or os.path.join(_HERE, "libmppi_b200.so")  # env: tuning builds only

MPPI_OK = 0
MPPI_ERR_INVALID, MPPI_ERR_CUDA, MPPI_ERR_STATE, MPPI_ERR_NCCL, MPPI_ERR_ALLOC = -1, -2, -3, -4, -5
MODEL_DIFF_DRIVE, MODEL_STEERING, MODEL_FULL_BODY = 0, 1, 2
DEBUG_NONE, DEBUG_NEAREST, DEBUG_STATES = 0, 1, 2
SCAN_AUTO, SCAN_LITERAL, SCAN_PRUNED = 0, 1, 2
WINDOW_AUTO, WINDOW_HOST, WINDOW_DEVICE = 0, 1, 2
(OPT_GRID_MAX_CELLS, OPT_GRID_H_MIN, OPT_GRID_MARGIN, OPT_GRID_LANES, OPT_REDUCE_GROUPS, OPT_FUSE_CONTROLS,
 OPT_NOISE_PREFETCH, OPT_EXCHANGE_TIMEOUT_MS, OPT_FEEDBACK_WARM_START, OPT_UPLOAD_WARM_START) = range(1, 11)
INFO_FUSED_CONTROLS = 100
COMM_ID_BYTES = 128
IPC_HANDLE_BYTES = 64


class MppiParams(C.Structure):
    """mppi_params of include/mppi_b200.h (same field order as oracle_params)."""
    _fields_ = [
        ("control_noise", C.c_double), ("lambda_", C.c_double), ("v_ref", C.c_double), ("resolution", C.c_double),
        ("u_min", C.c_double * 5), ("u_max", C.c_double * 5),
        ("path_weight", C.c_double), ("v_weight", C.c_double), ("zmp_weight", C.c_double),
        ("roll_v_weight", C.c_double), ("back_weight", C.c_double), ("yaw_weight", C.c_double),
        ("steer_off", C.c_int32), ("reserved", C.c_int32),
    ]


# every symbol include/mppi_b200.h declares: name -> (restype, argtypes)
_P = C.POINTER
SYMBOLS = {
    "mppi_create": (C.c_int, [_P(C.c_void_p), C.c_int, _P(MppiParams), C.c_int, C.c_int, C.c_int, C.c_int]),
    "mppi_destroy": (C.c_int, [C.c_void_p]),
    "mppi_last_error": (C.c_char_p, [C.c_void_p]),
    "mppi_abi_version": (C.c_int, []),
    "mppi_set_params": (C.c_int, [C.c_void_p, _P(MppiParams)]),
    "mppi_set_debug": (C.c_int, [C.c_void_p, C.c_int]),
    "mppi_set_scan_mode": (C.c_int, [C.c_void_p, C.c_int]),
    "mppi_set_window_builder": (C.c_int, [C.c_void_p, C.c_int]),
    "mppi_set_option": (C.c_int, [C.c_void_p, C.c_int, C.c_double]),
    "mppi_get_option": (C.c_int, [C.c_void_p, C.c_int, _P(C.c_double)]),
    "mppi_set_path": (C.c_int, [C.c_void_p, C.c_int, _P(C.c_double), C.c_int]),
    "mppi_set_window": (C.c_int, [C.c_void_p, C.c_int, _P(C.c_double)]),
    "mppi_set_seed": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64]),
    "mppi_set_shard": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int]),
    "mppi_set_noise": (C.c_int, [C.c_void_p, _P(C.c_float)]),
    "mppi_solve": (C.c_int, [C.c_void_p, _P(C.c_double), C.c_double, _P(C.c_double)]),
    "mppi_upload": (C.c_int, [C.c_void_p, _P(C.c_double), C.c_double, _P(C.c_double)]),
    "mppi_enqueue": (C.c_int, [C.c_void_p]),
    "mppi_download": (C.c_int, [C.c_void_p, _P(C.c_double)]),
    "mppi_synchronize": (C.c_int, [C.c_void_p]),
    "mppi_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mppi_use_graph": (C.c_int, [C.c_void_p, C.c_int]),
    "mppi_get_costs": (C.c_int, [C.c_void_p, C.c_int, _P(C.c_float)]),
    "mppi_get_weights": (C.c_int, [C.c_void_p, C.c_int, _P(C.c_float)]),
    "mppi_get_nearest": (C.c_int, [C.c_void_p, C.c_int, _P(C.c_int32)]),
    "mppi_get_states": (C.c_int, [C.c_void_p, C.c_int, _P(C.c_double)]),
    "mppi_get_noise": (C.c_int, [C.c_void_p, C.c_int, _P(C.c_float)]),
    "mppi_get_window": (C.c_int, [C.c_void_p, C.c_int, _P(C.c_double), _P(C.c_int)]),
    "mppi_get_stats": (C.c_int, [C.c_void_p, C.c_int, _P(C.c_double)]),
    "mppi_get_record": (C.c_int, [C.c_void_p, C.c_int, _P(C.c_float)]),
    "mppi_get_info": (C.c_int, [C.c_void_p, _P(C.c_int), _P(C.c_int), _P(C.c_int), _P(C.c_int), _P(C.c_int)]),
    "mppi_time_kernels": (C.c_int, [C.c_void_p, C.c_int, _P(C.c_float)]),
    "mppi_get_io_bytes": (C.c_int, [C.c_void_p, _P(C.c_size_t), _P(C.c_size_t)]),
    "mppi_last_launch_count": (C.c_int, [C.c_void_p]),
    "mppi_comm_get_unique_id": (C.c_int, [C.c_void_p]),
    "mppi_comm_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "mppi_comm_export": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "mppi_comm_connect": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "mppi_merge_partials": (C.c_int, [_P(C.c_float), C.c_int, C.c_int, C.c_double, _P(C.c_float), _P(C.c_double)]),
    "mppi_calc_ref_path": (C.c_int, [_P(C.c_double), C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                     C.c_double, C.c_int, _P(C.c_double), _P(C.c_int)]),
    "mppi_philox4x32_10": (None, [_P(C.c_uint32), _P(C.c_uint32), _P(C.c_uint32)]),
}

_lib = None


def load():
    """Load libmppi_b200.so; raises (loudly) when it has not been built -- there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `make lib` (or __graft_entry__.build()); "
                              "the B200 MPPI core has no Python/CPU fallback")
        lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the ABI lost a symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class MppiError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"mppi error {code}: {msg}")
        self.code = code
