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
