"""CPU tests of the oracle (no GPU): pinned against the golden vectors of the UNMODIFIED reference nodes
(tests/golden, made by tests/golden/make_golden.py from oracle/_ref), against the reference build itself when it
is present (build container only), and the FP32 twin against the FP64 oracle."""
import numpy as np
import pytest

import oracle
from oracle import ref_runner
from ccv_mppi_path_tracker_b200 import params, paths
from common import golden_names, load_golden, make_case, oob_cost_offset


def _check_against_reference(case, r):
    model, K, T = case["model"], case["K"], case["T"]
    st = np.array(case["state"], dtype=np.float64)
    st[2] = r["yaw_used"]  # the yaw the node read back from its quaternion
    o = oracle.solve(model, case["sp"], K, T, st, case["dt"], case["path"], case["eps"], case["u0"], shifted=False,
                     want=("cost", "weights", "window", "states", "stats", "current_index", "zmp", "controls"))
    assert abs(r["yaw_used"] - case["state"][2]) < 1e-15
    # a8/a9: window and current index, bit for bit
    assert int(r["current_index"]) == int(o["current_index"][0])
    assert np.array_equal(r["window"], o["window"])
    # a5/a6: predicted states (and ZMP) bit for bit
    assert np.array_equal(r["states"], o["states"])
    if model == "full_body":
        assert np.array_equal(r["zmp"], o["zmp"])
    # a11: cost -- FB bit for bit; DD/SD up to the constant of the out-of-bounds term (decision D1)
    if model == "full_body":
        assert np.array_equal(r["cost"], o["cost"])
    else:
        assert np.allclose(r["cost"] - oob_cost_offset(case), o["cost"], rtol=1e-13, atol=1e-12)
    # a12/a13: normalised literal weights and the new control sequence
    # (the literal exp(-c/lambda) of a far-from-best sample is sub-normal and keeps only a few bits)
    big = o["weights"] > 1e-100
    assert np.allclose(r["weights"][big], o["weights"][big], rtol=1e-9, atol=0)
    assert np.allclose(r["weights"][~big], o["weights"][~big], rtol=1e-2, atol=1e-300)
    assert np.allclose(r["u_new"], o["u_new"], rtol=0, atol=1e-12)
    # D3: the shifted weights are the same numbers after normalisation
    o2 = oracle.solve(model, case["sp"], K, T, st, case["dt"], case["path"], case["eps"], case["u0"], shifted=True)
    # (exp(-c) underflows below c ~ 744: relative weights under e^-(744 - c_min) vanish in the literal form exp(-c/lambda) and survive in the shifted one)
    assert np.allclose(o2["weights"], o["weights"], rtol=1e-9, atol=1e-30)
    assert np.allclose(o2["u_new"], o["u_new"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_golden_reference_cycle(name):
    case = load_golden(name)
    _check_against_reference(case, case["ref"])


@pytest.mark.skipif(not ref_runner.available(), reason="oracle/_ref needs /root/reference (build container only)")
@pytest.mark.parametrize("model,K,T,seed", [("diff_drive", 257, 15, 1), ("diff_drive", 64, 100, 2), ("steering", 200, 20, 3),
                                            ("full_body", 150, 15, 4), ("full_body", 64, 60, 5)])
def test_oracle_matches_live_reference_build(model, K, T, seed):
    case = make_case(model, K, T, seed=seed)
    rng = np.random.default_rng(seed)
    case["u0"] = case["u0"] + 0.2 * rng.standard_normal(case["u0"].shape)
    r = ref_runner.run(model, case["p"], K, T, case["state"], case["dt"], case["path"], case["eps"], case["u0"])
    _check_against_reference(case, r)


def test_golden_files_cover_the_three_models_and_the_quirks():
    names = golden_names()
    assert len(names) >= 9
    models = {load_golden(n)["model"] for n in names}
    assert models == {"diff_drive", "steering", "full_body"}


def test_calc_ref_path_quirks():
    """a9: truncated double index, tail clamp to the last pose, atan2(0,0) = 0 on duplicates, yaw_ref[T-1] = 0."""
    path = paths.sin_path(**params.LAUNCH_PATH["diff_drive"])
    assert path.shape == (101, 2)  # accumulated s += 0.1 loop, reference_path_creator.cpp:38
    win, cur = oracle.calc_ref_path(path, 0.0, 0.0, 1.2, 0.1, 0.1, 100)
    assert cur == 0
    idx = [int(0 + i * (1.2 * 0.1 / 0.1)) for i in range(100)]
    exp = path[np.minimum(idx, 100)]
    assert np.array_equal(win[:, :2], exp)
    assert win[-1, 2] == 0.0 and np.all(win[85:, 2] == 0.0)  # clamped tail: identical points -> atan2(0, 0)
    # farther than 100 m from everything: min_distance never drops below its 100.0 start -> index 0
    assert oracle.calc_ref_path(path, 500.0, 0.0, 1.2, 0.1, 0.1, 15)[1] == 0
    # the 1-point data/data.csv path
    one = np.array([[-5.45606, -6.61448]])
    w1, c1 = oracle.calc_ref_path(one, -5.0, -6.0, 1.2, 0.1, 0.1, 15)
    assert c1 == 0 and np.all(w1[:, 0] == one[0, 0]) and np.all(w1[:, 2] == 0.0)


def test_min_distance_cap_and_first_minimum():
    xr = np.array([0.0, 1.0, 1.0, 2.0])
    yr = np.zeros(4)
    d, j = oracle.min_distance(1.0, 0.5, xr, yr)
    assert d == 0.5 and j == 1  # first of the two equal minima
    d, j = oracle.min_distance(1000.0, 0.0, xr, yr)
    assert d == 100.0 and j == -1


def test_zmp_model_matches_closed_form():
    com, acc, hg = np.array([0.01, -0.02, 0.4]), np.array([0.3, -0.2, 0.0]), np.array([0.5, -0.4, 0.1])
    z = oracle.zmp_from_model(com, acc, hg)
    gz, m = -9.8, 60.0
    assert np.isclose(z[0], com[0] + com[2] * acc[0] / gz + hg[1] / (m * gz), rtol=1e-12)
    assert np.isclose(z[1], com[1] + com[2] * acc[1] / gz - hg[0] / (m * gz), rtol=1e-12)


@pytest.mark.parametrize("model,K,T", [("diff_drive", 1000, 15), ("steering", 512, 50), ("full_body", 512, 100),
                                       ("diff_drive", 512, 100)])
def test_fp32_twin_against_fp64_oracle(model, K, T):
    """The FP32 contract of the kernels (mppi_math.h, compiled for the host): nearest indices equal to FP64 except at
    near-ties, costs within 1e-5 relative."""
    case = make_case(model, K, T, seed=7)
    o = oracle.solve(model, case["sp"], K, T, case["state"], case["dt"], case["path"], case["eps"], case["u0"],
                     want=("cost", "nearest", "window", "states"))
    tw = oracle.twin_rollout_cost(model, case["sp"], K, T, case["state"], case["dt"], o["window"], case["eps"],
                                  case["u0"], want=("nearest", "d2", "states"))
    Tc = T - 2 if model == "full_body" else T
    mism = tw["nearest"][:, :Tc] != o["nearest"][:, :Tc]
    if mism.any():
        # every mismatch must be a near-tie: the two candidates' distances differ by < 1e-5 relative
        win = o["window"][:, :2]
        ii, tt = np.nonzero(mism)
        p = o["states"][ii, tt, :2]
        da = np.linalg.norm(p - win[tw["nearest"][ii, tt]], axis=1)
        db = np.linalg.norm(p - win[o["nearest"][ii, tt]], axis=1)
        assert np.all(np.abs(da - db) <= 1e-5 * np.maximum(da, db) + 1e-7)
    assert mism.mean() < 1e-3
    assert np.all(np.abs(tw["cost"] - o["cost"]) <= 1e-5 * np.abs(o["cost"]) + 1e-5)
    # robot-centred FP32 states track the FP64 states to ~1e-5 m over the horizon
    S = 5 if model == "full_body" else 3
    st64 = o["states"].copy()
    st64[:, :, :2] -= case["state"][:2]
    assert np.abs(tw["states"][:, :, :S] - st64).max() < 5e-5


def test_twin_sincos_accuracy():
    a = np.linspace(-40, 40, 200001).astype(np.float32)
    s, c = oracle.twin_sincos(a)
    assert np.abs(s - np.sin(a.astype(np.float64))).max() < 3e-7
    assert np.abs(c - np.cos(a.astype(np.float64))).max() < 3e-7


def test_twin_atan2_accuracy():
    rng = np.random.default_rng(0)
    y = np.concatenate([rng.normal(0, 1, 100000), [0.0, 0.0, 1.0, -1.0, 0.0, 1e-30, 3.0]]).astype(np.float32)
    x = np.concatenate([rng.normal(0, 1, 100000), [0.0, 1.0, 0.0, 0.0, -1.0, 1e-30, -3.0]]).astype(np.float32)
    a = oracle.twin_atan2(y, x)
    ref = np.arctan2(y.astype(np.float64), x.astype(np.float64))
    assert np.abs(a - ref).max() < 4e-7
    assert a[100000] == 0.0  # atan2(0, 0) = 0: duplicate window points (DD:175-178)


def test_twin_sincos_increment_accuracy():
    """The per-step angle increments (w*dt, roll_v*dt, pitch_v*dt): short polynomials inside +-0.35 rad, the general
    sincos beyond; both below one FP32 ulp of error, no jump at the switch-over."""
    a = np.concatenate([np.linspace(-0.35, 0.35, 400001), np.linspace(-3.0, 3.0, 100001),
                        np.nextafter(np.float32(0.35), np.float32([0.0, 1.0])), [0.0, -0.0, 1e-20]]).astype(np.float32)
    s, c = oracle.twin_sincos_increment(a)
    a64 = a.astype(np.float64)
    assert np.abs(s - np.sin(a64)).max() < 1.2e-7
    assert np.abs(c - np.cos(a64)).max() < 1.2e-7
    small = np.abs(a) <= 0.35
    assert np.abs(s[small] - np.sin(a64[small])).max() < 4e-8   # |sin| < 0.35: half an ulp is 1.5e-8
    assert s[-3] == 0.0 and c[-3] == 1.0


def test_heading_recurrence_tracks_the_angle():
    """(cos yaw, sin yaw) carried by rotations (mppi_math.h rotate_by) against FP64 cos/sin of the exactly summed
    angle: after 400 steps the pair is as accurate as -- for large accumulated yaw more accurate than -- the FP32
    angle accumulation it replaces."""
    rng = np.random.default_rng(3)
    for mean in (0.0, 0.2, -0.19):
        inc = (mean + 0.05 * rng.standard_normal(400)).astype(np.float32)
        c, s, a32 = oracle.twin_heading(0.3, inc)
        exact = np.float64(np.float32(0.3)) + np.cumsum(inc.astype(np.float64))
        err_pair = np.hypot(c - np.cos(exact), s - np.sin(exact))
        err_angle = np.abs(a32.astype(np.float64) - exact)
        assert err_pair.max() < 8e-6
        assert np.abs(np.hypot(c.astype(np.float64), s.astype(np.float64)) - 1.0).max() < 4e-6
        if mean != 0.0:
            assert err_pair[-1] <= err_angle[-1] + 2e-6


def test_twin_clamp_matches_the_reference_clamp():
    """clamp (DD:62-67) as min(max(v, lo), hi) with NaN pass-through: same value as the reference's two tests."""
    v = np.concatenate([np.linspace(-5, 5, 1001), [np.nan, np.inf, -np.inf, 0.0, -0.0]]).astype(np.float32)
    for lo, hi in ((-1.2, 2.0), (0.0, 0.0), (-0.5, -0.25)):
        out = oracle.twin_clamp(v, lo, hi)
        lo32, hi32 = np.float32(lo), np.float32(hi)
        ref = np.where(v < lo32, lo32, np.where(v > hi32, hi32, v)).astype(np.float32)
        assert np.array_equal(np.isnan(out), np.isnan(ref))
        ok = ~np.isnan(ref)
        assert np.array_equal(out[ok], ref[ok])   # equal as values (a zero may differ in sign only)


@pytest.mark.skipif(not ref_runner.timing_available(), reason="oracle/_ref timing binaries not built (no /root/reference)")
def test_reference_timing_binary_runs():
    """bench.py --impl reference: the unmodified node's own cycle body, timed."""
    from ccv_mppi_path_tracker_b200 import params, paths
    p = params.node_params("diff_drive", launch=True, horizon=15, num_samples=200)
    ts = ref_runner.time_solves("diff_drive", p, 200, 15, np.zeros(3), 0.1,
                                paths.sin_path(**params.LAUNCH_PATH["diff_drive"]), np.zeros((14, 2)), 3)
    assert len(ts) == 3 and all(t > 0 for t in ts)
