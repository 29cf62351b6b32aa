"""CPU tests (no GPU): the C-ABI library loads and exports every symbol include/mppi_b200.h declares, fails loudly
without a device, and the host-side pieces (window construction, path sources, parameter surface, Philox block,
partial merge) agree with the oracle / known answers."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle
from ccv_mppi_path_tracker_b200 import (CONTROLLERS, _capi, calc_ref_path, merge_partials, params, paths,
                                        philox4x32_10)
from common import make_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "mppi_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mppi_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    lib = _capi.load()
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mppi_b200.h but not exported by libmppi_b200.so"
        assert n in _capi.SYMBOLS, f"{n} has no ctypes prototype in _capi.SYMBOLS"
    assert set(_capi.SYMBOLS) == set(names)
    assert lib.mppi_abi_version() == 1


def test_params_struct_layout_matches_header():
    # 4 doubles + 2 x 5 doubles + 6 doubles + 2 int32
    assert C.sizeof(_capi.MppiParams) == 8 * (4 + 10 + 6) + 8
    assert C.sizeof(oracle.OracleParams) == C.sizeof(_capi.MppiParams)


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(_capi.MppiError) as e:
        CONTROLLERS["diff_drive"](horizon=15, num_samples=64)
    assert e.value.code == _capi.MPPI_ERR_CUDA
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_create_rejects_bad_arguments():
    lib = _capi.load()
    h = C.c_void_p()
    cp = _capi.MppiParams()
    cp.lambda_ = 1.0
    cp.resolution = 0.1
    assert lib.mppi_create(C.byref(h), 7, C.byref(cp), 64, 15, 1, 0) == _capi.MPPI_ERR_INVALID       # unknown model
    assert lib.mppi_create(C.byref(h), 0, C.byref(cp), 0, 15, 1, 0) == _capi.MPPI_ERR_INVALID        # K < 1
    assert lib.mppi_create(C.byref(h), 0, C.byref(cp), 64, 1, 1, 0) == _capi.MPPI_ERR_INVALID        # T < 2
    cp.lambda_ = 0.0
    assert lib.mppi_create(C.byref(h), 0, C.byref(cp), 64, 15, 1, 0) == _capi.MPPI_ERR_INVALID       # lambda <= 0
    assert b"lambda" in lib.mppi_last_error(None)
    cp.lambda_ = 1.0
    cp.u_min[1], cp.u_max[1] = 2.0, -2.0                                                              # w_min > w_max
    assert lib.mppi_create(C.byref(h), 0, C.byref(cp), 64, 15, 1, 0) == _capi.MPPI_ERR_INVALID
    assert b"u_min" in lib.mppi_last_error(None)
    assert lib.mppi_destroy(None) == _capi.MPPI_OK


def test_philox4x32_10_known_answers():
    """Random123 kat_vectors for philox4x32 with 10 rounds."""
    assert philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


@pytest.mark.parametrize("T", [1, 15, 50, 100])
def test_library_window_builder_equals_oracle(T):
    rng = np.random.default_rng(T)
    for model in ("diff_drive", "full_body"):
        path = paths.sin_path(**params.LAUNCH_PATH[model])
        for _ in range(20):
            j = rng.integers(0, path.shape[0])
            px, py = path[j] + rng.normal(0, 0.3, 2)
            v_ref, dt = rng.uniform(0.2, 2.5), rng.uniform(0.05, 0.2)
            w_lib, c_lib = calc_ref_path(path, px, py, v_ref, dt, 0.1, T)
            w_or, c_or = oracle.calc_ref_path(path, px, py, v_ref, dt, 0.1, T)
            assert c_lib == c_or and np.array_equal(w_lib, w_or)
    one = np.array([[-5.45606, -6.61448]])
    assert np.array_equal(calc_ref_path(one, 0.0, 0.0, 1.2, 0.1, 0.1, T)[0], oracle.calc_ref_path(one, 0.0, 0.0, 1.2, 0.1, 0.1, T)[0])


def test_path_sources_and_csv_roundtrip(tmp_path):
    p = paths.sin_path(**params.LAUNCH_PATH["diff_drive"])
    assert p.shape == (101, 2) and p[0, 0] == 0.0 and abs(p[-1, 0] - 10.0) < 1e-9
    assert np.array_equal(p, oracle.make_sin_path(**params.LAUNCH_PATH["diff_drive"]))
    assert np.allclose(p[:, 1], np.cos(2 * np.pi * 0.25 * p[:, 0]) - 1.0)
    assert paths.sin_path(**params.LAUNCH_PATH["full_body"]).shape == (200, 2)
    d = paths.dkan_path()
    assert np.allclose(d[0], [0, 0]) and abs(d[:, 0].max() - 17.7) < 1e-9 and abs(d[:, 1].max() - 8.0) < 0.11
    f = tmp_path / "data.csv"
    f.write_text("-5.45606,-6.61448,\n")  # the reference's data/data.csv:1
    one = paths.load_csv(str(f))
    assert one.shape == (1, 2) and one[0, 0] == -5.45606
    paths.save_csv(str(f), p)
    assert np.allclose(paths.load_csv(str(f)), p, atol=1e-5)


def test_parameter_surface_defaults_and_launch_quirks():
    dd = params.node_params("diff_drive", launch=False)
    assert (dd["horizon"], dd["num_samples"], dd["control_noise"], dd["lambda_"]) == (15, 1000, 0.5, 1.0)
    assert (dd["v_max"], dd["v_min"], dd["w_max"], dd["v_ref"], dd["path_weight"]) == (1.2, -1.2, 2.0, 0.8, 1.0)
    ddl = params.node_params("diff_drive", launch=True)
    assert (ddl["v_max"], ddl["v_ref"], ddl["path_weight"]) == (2.0, 1.2, 10.0)
    assert ddl["control_weight"] == 1.0  # the launch file's v_weight is never read (DD:34)
    sd = params.node_params("steering", launch=True)
    assert sd["num_samples"] == 1000 and sd["w_max"] == 1.0 and abs(sd["steer_max"] - np.pi / 6) < 1e-15
    fb = params.node_params("full_body", launch=True)
    sp = params.solve_params("full_body", fb)
    assert fb["roll_off"] and sp["zmp_weight"] == 0.0 and sp["roll_v_weight"] == 0.0  # FB:43-46
    sp2 = params.solve_params("full_body", params.node_params("full_body", launch=True, roll_off=False))
    assert (sp2["zmp_weight"], sp2["roll_v_weight"], sp2["yaw_weight"], sp2["u_min"][0]) == (10.0, 0.5, 2.0, -3.0)


def test_merge_partials_is_a_log_sum_exp_merge():
    """Sharded softmax: per-rank {c_min, sum w, sum w^2, -, N[]} merged on the host == the unsharded weighted mean."""
    rng = np.random.default_rng(0)
    K, n, G, lam = 4096, 28, 4, 0.7
    cost = rng.uniform(40, 90, K)
    cost[1234] = 12.5  # one rank holds a far better sample: the other ranks' partials are scaled down to ~0
    u = rng.normal(0, 1, (K, n))
    w = np.exp(-(cost - cost.min()) / lam)
    u_ref = (w[:, None] * u).sum(0) / w.sum()
    recs = []
    for g in range(G):
        sl = slice(g * K // G, (g + 1) * K // G)
        m = cost[sl].min()
        wg = np.exp(-(cost[sl] - m) / lam)
        recs.append(np.concatenate([[m, wg.sum(), (wg ** 2).sum(), 0.0], (wg[:, None] * u[sl]).sum(0)]))
    um, st = merge_partials(np.array(recs, dtype=np.float32), lam)
    assert np.abs(um - u_ref).max() < 1e-5
    assert abs(st["c_min"] - cost.min()) < 1e-5 and abs(st["ess"] - w.sum() ** 2 / (w ** 2).sum()) < 1e-3 * st["ess"]
    # one rank: identity
    u1, _ = merge_partials(np.array(recs[:1], dtype=np.float32), lam)
    assert np.allclose(u1, recs[0][4:] / recs[0][1], rtol=1e-6)


def test_make_case_is_deterministic():
    a, b = make_case("steering", 64, 10, seed=3), make_case("steering", 64, 10, seed=3)
    assert np.array_equal(a["eps"], b["eps"]) and a["sp"] == b["sp"]


def test_full_body_zmp_monitors_match_the_reference(tmp_path):
    """f4: the full-body node's ZMP monitors (calc_true_ZMP + the ZMP part of get_CurrentState, FB:528-596) in the
    Python and the C++ host classes against eight golden cycles of the UNMODIFIED reference node."""
    import struct
    import subprocess
    from ccv_mppi_path_tracker_b200.controllers import FullBodyMPPI
    from common import GOLDEN_DIR
    g = np.load(os.path.join(GOLDEN_DIR, "fb_estimator.dat"), allow_pickle=False)
    cyc, ref = g["cycles"], g["ref_out"]
    assert np.array_equal(ref[3, 2:], ref[2, 2:])  # the no-contact cycle keeps the previous force-sensor estimate
    fb = FullBodyMPPI.__new__(FullBodyMPPI)  # monitors only: no device handle
    fb.dt_ = 0.1
    fb.reset_monitors()
    got = []
    for v in cyc:
        tz = fb.calc_true_ZMP(v[8:26].reshape(6, 3)).copy()
        zx, zy = fb.update_model_zmp(v[1], v[2], v[3], v[4], v[5:8], dt=v[0])
        got.append([zx, zy, *tz])
    assert np.allclose(np.array(got), ref, rtol=1e-12, atol=1e-15)
    exe = os.path.join(ROOT, "tests", "host", "fb_monitor_check")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", ROOT, "hosttest"], check=True)
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    fin.write_bytes(struct.pack("<i", cyc.shape[0]) + np.ascontiguousarray(cyc).tobytes())
    subprocess.run([exe, str(fin), str(fout)], check=True)
    cpp = np.fromfile(fout, dtype=np.float64).reshape(-1, 5)
    assert np.allclose(cpp, ref, rtol=1e-12, atol=1e-15)


def test_cmd_vel_and_cmd_pos_match_the_reference(tmp_path):
    """f2: what publish_CmdVel / publish_CmdPos of the three UNMODIFIED reference nodes publish (golden cases incl.
    w = 0 -> R = inf, v = w = 0 -> R = NaN, the w > 0 left/right swap, the roll clamp, steer_off and roll_off;
    diff_drive_mppi.cpp:248-263, steering_diff_drive_mppi.cpp:266-296, full_body_mppi.cpp:238-275) against cmd_vel() /
    cmd_pos() of the C++ host classes (bit for bit) and of the Python mirror."""
    import struct
    import subprocess
    from ccv_mppi_path_tracker_b200 import controllers
    from common import GOLDEN_DIR
    g = np.load(os.path.join(GOLDEN_DIR, "cmd_golden.dat"), allow_pickle=False)
    cases = g["cases"]
    # the golden covers the special cases it claims
    assert cases[0, 1] == 0.0 and np.isnan(g["ref_sd"][1, 2]) and (cases[:, 6] == 1).any() and (cases[:, 7] == 1).any()
    assert (g["ref_fb"][:, 6] == np.deg2rad(30.0)).any() and (g["ref_fb"][:, 6] == -np.deg2rad(30.0)).any()
    exe = os.path.join(ROOT, "tests", "host", "cmd_check")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", ROOT, "hosttest"], check=True)
    fin = tmp_path / "in.bin"
    fin.write_bytes(struct.pack("<i", cases.shape[0]) + np.ascontiguousarray(cases).tobytes())
    for tag, model in (("dd", "diff_drive"), ("sd", "steering"), ("fb", "full_body")):
        ref = g["ref_" + tag]
        fout = tmp_path / f"out_{tag}.bin"
        subprocess.run([exe, tag, str(fin), str(fout)], check=True)
        cpp = np.fromfile(fout, dtype=np.float64).reshape(-1, 7)
        assert np.array_equal(cpp, ref, equal_nan=True), tag
        # Python mirror: no device handle needed for the command post-processing
        cls = controllers.CONTROLLERS[model]
        ctl = cls.__new__(cls)
        ctl.p = params.node_params(model, launch=False)
        U, S = params.NUM_CONTROLS[model], params.NUM_STATES[model]
        ctl.optimal_solution = np.zeros((1, 2, U))
        ctl.current_state = np.zeros((1, S))
        got = []
        for v in cases:
            ctl.optimal_solution[0, 0, :2] = v[:2]
            if U > 2:
                ctl.optimal_solution[0, 0, 2] = v[2]
            ctl.dt_ = v[5]
            if model == "full_body":
                ctl.optimal_solution[0, 0, 3] = v[3]
                ctl.current_state[0, 3] = v[4]
                ctl.p["steer_off"], ctl.p["roll_off"] = bool(v[6]), bool(v[7])
            cv, cp = ctl.cmd_vel(), ctl.cmd_pos()
            got.append([cv[0], cv[1], cp["steer_l"], cp["steer_r"], cp["fore"], cp["rear"], cp["roll"]])
        assert np.allclose(np.array(got), ref, rtol=1e-14, atol=1e-16, equal_nan=True), tag


def test_circle_path_generator():
    """reference_path_creator.cpp:57-68: the circle course (its step formula `resolution / 2 * pi * R` kept as is)."""
    p = paths.circle_path(R=2.0, resolution=0.1, init_x=1.0, init_y=-0.5)
    step = 0.1 / 2 * np.pi * 2.0
    n = int(np.floor(200 * np.pi / step)) + 1
    assert p.shape == (n, 2) or p.shape == (n + 1, 2)  # accumulated `s += step` may land a hair below the bound
    s = np.arange(p.shape[0]) * step
    assert np.allclose(p[:, 0], 1.0 + 2.0 * np.cos(s), atol=1e-9) and np.allclose(p[:, 1], -0.5 + 2.0 * np.sin(s) + 2.0, atol=1e-9)
    assert np.allclose(np.hypot(p[:, 0] - 1.0, p[:, 1] - 1.5), 2.0)
    assert np.array_equal(p[0], [3.0, 1.5])


def test_shipped_library_contains_the_blackwell_instructions():
    """The built libmppi_b200.so (sm_100a SASS, read with cuobjdump here on the CPU) really contains what DESIGN.md
    claims for the production rollout kernel: TMA tensor loads + mbarrier waits, packed FP32, three-input min,
    programmatic-dependent-launch instructions -- and no sm_90-only or library kernels."""
    import shutil
    import subprocess
    from ccv_mppi_path_tracker_b200 import _capi
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([exe, "-sass", _capi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in sass or "SM100" in sass.upper() or "sm_100" in sass
    # split per function, pick the production K2 instantiation (diff drive, no tap)
    parts = sass.split("Function : ")
    k2 = [p for p in parts if p.startswith("_ZN4mppi23rollout_cost_tma_kernelILi0ELb0E")]
    assert len(k2) == 1
    body = k2[0]
    for mnemonic in ("UTMALDG", "SYNCS", "ELECT", "FFMA2", "FMUL2", "FADD2", "FMNMX3", "VIMNMX", "ACQBULK", "PREEXIT"):
        assert mnemonic in body, mnemonic
    assert "WGMMA" not in sass and "HGMMA" not in sass  # nothing sm_90-only, no tensor-core library code
    names = [p.split("\n", 1)[0] for p in parts[1:]]
    assert all(n.startswith("_ZN4mppi") for n in names), [n for n in names if not n.startswith("_ZN4mppi")]
