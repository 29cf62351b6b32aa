"""Parameter surface of the three reference nodes: constructor defaults and launch-file overrides.

Sources (relative to /root/reference):
  DiffDriveMPPI          src/diff_drive_mppi.cpp:17-34,           launch/diff_drive_mppi.launch:6-9
  SteeringDiffDriveMPPI  src/steering_diff_drive_mppi.cpp:18-36,  launch/steering_diff_drive_mppi.launch:7-11
  FullBodyMPPI           src/full_body_mppi.cpp:8-46,             launch/full_body_mppi.launch:7-18
Quirks kept on purpose: DD/SD read the velocity weight from "control_weight", so the launch
files' "v_weight" is ignored; "exploration_noise" is read and never used; "dt" is overwritten every cycle.
"""
import math

DEG = math.pi / 180.0

DEFAULTS = {
    "diff_drive": dict(dt=0.1, horizon=15, num_samples=1000, control_noise=0.5, lambda_=1.0, v_max=1.2, w_max=2.0,
                       v_min=-1.2, w_min=-2.0, pitch_offset=3.0 * DEG, v_ref=0.8, resolution=0.1,
                       exploration_noise=0.5, path_weight=1.0, control_weight=1.0),
    "steering": dict(dt=0.1, horizon=15, num_samples=10000, control_noise=0.5, lambda_=1.0, v_max=1.2, w_max=1.0,
                     steer_max=30.0 * DEG, v_min=-1.2, w_min=-1.0, steer_min=-30.0 * DEG, pitch_offset=3.0 * DEG,
                     v_ref=0.8, resolution=0.1, exploration_noise=0.1, path_weight=1.0, control_weight=1.0),
    "full_body": dict(dt=0.1, horizon=15, num_samples=10000, control_noise=0.5, lambda_=1.0, v_max=1.2, w_max=1.0,
                      steer_max=30.0 * DEG, roll_max=30.0 * DEG, pitch_max=15.0 * DEG, roll_v_max=30.0 * DEG,
                      pitch_v_max=15.0 * DEG, v_min=-3.0, w_min=-1.0, steer_min=-30.0 * DEG, roll_min=-30.0 * DEG,
                      pitch_min=-15.0 * DEG, roll_v_min=-30.0 * DEG, pitch_v_min=-15.0 * DEG, pitch_offset=0.0,
                      v_ref=1.2, resolution=0.1, exploration_noise=0.1, path_weight=1.0, v_weight=1.0,
                      zmp_weight=1.0, roll_v_weight=1.0, back_weight=1.0, yaw_weight=1.0, roll_off=False,
                      steer_off=False, use_gazebo_pose=True),
}

LAUNCH = {
    "diff_drive": dict(path_weight=10.0, v_weight=1.0, v_ref=1.2, v_max=2.0),
    "steering": dict(v_ref=1.2, v_max=2.0, path_weight=10.0, v_weight=1.0, num_samples=1000),
    "full_body": dict(v_ref=2.0, v_max=2.0, path_weight=10.0, v_weight=1.0, zmp_weight=10.0, roll_v_weight=0.5,
                      back_weight=1.0, yaw_weight=2.0, roll_off=True, steer_off=False, use_gazebo_pose=False),
}

# reference_path_creator node parameters in the same launch files
LAUNCH_PATH = {
    "diff_drive": dict(course_length=10.0, A1=1.0, omega1=0.25, delta1=0.0, delta2=0.0, delta3=0.0),
    "steering": dict(course_length=10.0, A1=1.0, omega1=0.25, delta1=0.0, delta2=0.0, delta3=0.0),
    "full_body": dict(course_length=20.0, A1=1.5, omega1=0.127, delta1=0.0, delta2=0.0, delta3=0.0),
}

MODEL_ID = {"diff_drive": 0, "steering": 1, "full_body": 2}
NUM_CONTROLS = {"diff_drive": 2, "steering": 3, "full_body": 5}
NUM_STATES = {"diff_drive": 3, "steering": 3, "full_body": 5}


def node_params(model, launch=True, **overrides):
    """Parameters as the node would hold them after construction (defaults <- launch file <- overrides)."""
    p = dict(DEFAULTS[model])
    if launch:
        for k, v in LAUNCH[model].items():
            # DD/SD never read "v_weight" (they read "control_weight"): the launch value is dropped
            if k == "v_weight" and model != "full_body":
                continue
            p[k] = v
    p.update(overrides)
    return p


def solve_params(model, p):
    """Flatten node parameters into the mppi_params / oracle_params field order."""
    if model == "diff_drive":
        u_min = [p["v_min"], p["w_min"], 0.0, 0.0, 0.0]
        u_max = [p["v_max"], p["w_max"], 0.0, 0.0, 0.0]
    elif model == "steering":
        u_min = [p["v_min"], p["w_min"], p["steer_min"], 0.0, 0.0]
        u_max = [p["v_max"], p["w_max"], p["steer_max"], 0.0, 0.0]
    else:
        u_min = [p["v_min"], p["w_min"], p["steer_min"], p["roll_v_min"], p["pitch_v_min"]]
        u_max = [p["v_max"], p["w_max"], p["steer_max"], p["roll_v_max"], p["pitch_v_max"]]
    fb = model == "full_body"
    roll_off = bool(p.get("roll_off", False))
    return dict(
        control_noise=p["control_noise"], lambda_=p["lambda_"], v_ref=p["v_ref"], resolution=p["resolution"],
        u_min=u_min, u_max=u_max, path_weight=p["path_weight"],
        v_weight=p["v_weight"] if fb else p["control_weight"],
        # FB:43-46: roll_off zeroes both weights
        zmp_weight=(0.0 if roll_off else p["zmp_weight"]) if fb else 0.0,
        roll_v_weight=(0.0 if roll_off else p["roll_v_weight"]) if fb else 0.0,
        back_weight=p["back_weight"] if fb else 0.0, yaw_weight=p["yaw_weight"] if fb else 0.0,
        steer_off=int(bool(p.get("steer_off", False))),
    )
