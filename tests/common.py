"""Shared test scaffolding: the five BASELINE.json configurations at parity-test sizes, synthetic inputs."""
import numpy as np

from ccv_mppi_path_tracker_b200 import params, paths


def make_case(model, K, T, seed=0, launch=True, state=None, **overrides):
    if model == "full_body" and "roll_off" not in overrides:
        overrides["roll_off"] = False  # exercise every cost term
    p = params.node_params(model, launch=launch, horizon=T, num_samples=K, **overrides)
    sp = params.solve_params(model, p)
    path = paths.sin_path(**params.LAUNCH_PATH[model])
    U, S = params.NUM_CONTROLS[model], params.NUM_STATES[model]
    rng = np.random.default_rng(seed)
    eps = rng.standard_normal((T - 1, K, U)).astype(np.float32)
    if state is None:
        state = np.zeros(S)
        state[:3] = [0.25, -0.15, 0.2]
        if S == 5:
            state[3:] = [0.03, -0.02]
    u0 = np.zeros((T - 1, U))
    u0[:, 0] = 0.8 * p["v_ref"]
    return dict(model=model, K=K, T=T, U=U, S=S, p=p, sp=sp, path=path, eps=eps, state=np.asarray(state, float),
                u0=u0, dt=0.1, overrides=overrides)


def near_tie_mask(d2_sorted_two, rel=4 * 2.0 ** -23):
    """True where the two smallest squared distances are within `rel` relative: index mismatches allowed there."""
    a, b = d2_sorted_two[..., 0], d2_sorted_two[..., 1]
    return np.abs(b - a) <= rel * np.maximum(np.abs(a), np.abs(b)) + 1e-30


GOLDEN_DIR = __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "golden")


def load_golden(name):
    """One committed golden cycle of the UNMODIFIED reference (tests/golden/make_golden.py)."""
    import os
    g = np.load(os.path.join(GOLDEN_DIR, name), allow_pickle=False)
    model = str(g["model"])
    p = {str(k): float(v) for k, v in zip(g["param_names"], g["param_values"])}
    for k in ("roll_off", "steer_off", "use_gazebo_pose"):
        if k in p:
            p[k] = bool(p[k])
    K, T = int(g["K"]), int(g["T"])
    p["horizon"], p["num_samples"] = T, K
    return dict(model=model, K=K, T=T, p=p, sp=params.solve_params(model, p), dt=float(g["dt"]), state=g["state"],
                path=g["path"], eps=g["eps"], u0=g["u0"], ref={k[4:]: g[k] for k in g.files if k.startswith("ref_")},
                U=params.NUM_CONTROLS[model], S=params.NUM_STATES[model])


def golden_names():
    import os
    return sorted(f for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz"))


def oob_cost_offset(case):
    """The reference's calc_Cost reads v_[T-1] one past the end (diff_drive_mppi.cpp:204); in the golden build that
    read is 0.0, which adds v_weight * (0 - v_ref)^2 to EVERY sample of DD/SD (cancels in the weights). D1."""
    return 0.0 if case["model"] == "full_body" else case["sp"]["v_weight"] * case["sp"]["v_ref"] ** 2
