"""Named workloads of BASELINE.json (synthetic inputs) -- shared by bench.py and the tests.
Pure numpy; nothing here touches oracle/ or the GPU."""
import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ROBOT_JSON = os.path.join(os.path.dirname(_HERE), "tests", "golden", "a1_robot_consts.json")

# initial foot positions of test/mpc_test.cpp:97-101
EE_NOMINAL = np.array([[0.1526, 0.12523, 0.011089], [0.1526, -0.12523, 0.011089],
                       [-0.208321844, 0.1363286, 0.01444], [-0.208321844, -0.1363286, 0.01444]])

CONFIGS = {
    # apps/a1_configuration.yaml (num_nodes:79, integrator_dt:80, friction_coef:104, force_bound:147, swing_height:148,
    # ee_box_size:151, foot_offset:158, force_cost:159, Q_srbd_diag:179, srb_init:67-70)
    "a1_configuration": dict(num_nodes=20, integrator_dt=0.05, friction_coef=0.5, force_bound=150.0, swing_height=0.075,
                             foot_offset=0.015, ee_box_size=(0.15, 0.15), force_cost=0.0,
                             Q=[340, 340, 4000, 0.100, 0.100, 10, 3000, 3000, 3000, 1, 1, 1],
                             srb_init=[0, 0, 0.3, 0, 0, 0, 0, 0, 0, 1.0, 0, 0, 0],
                             srb_target=[0, 0, 0.3, 0, 0, 0, 0, 0, 0, 1.0, 0, 0, 0]),
    # apps/a1_gait_opt_config.yaml (it has no srb_target entry: the target is the initial state)
    "a1_gait_opt_config": dict(num_nodes=50, integrator_dt=0.02, friction_coef=0.6, force_bound=200.0, swing_height=0.1,
                               foot_offset=0.001, ee_box_size=(0.15, 0.15), force_cost=0.0,
                               Q=[55, 40, 500, 0.1, 0.1, 0.1, 5000, 5000, 5000, 0.1, 0.1, 0.1],
                               srb_init=[0, 0, 0.34, 0, 0, 0, 0, 0, 0, 1.0, 0, 0, 0],
                               srb_target=[0, 0, 0.34, 0, 0, 0, 0, 0, 0, 1.0, 0, 0, 0]),
    # apps/a1_config_distr_rejection.yaml
    "a1_config_distr_rejection": dict(num_nodes=50, integrator_dt=0.02, friction_coef=0.6, force_bound=200.0, swing_height=0.075,
                                      foot_offset=0.015, ee_box_size=(0.15, 0.15), force_cost=0.001,
                                      Q=[140, 140, 12000, 0.015, 0.015, 10, 3000, 3000, 3000, 1, 1, 1],
                                      srb_init=[0, 0, 0.3, 2.5, 0, 0, 0, 0, 0, 1.0, 0, 0, 0],
                                      srb_target=[0, 0, 0.3, 0, 0, 0, 0, 0, 0, 1.0, 0, 0, 0]),
}


def robot():
    with open(ROBOT_JSON) as f:
        return json.load(f)


def mpc_kwargs(cfg):
    return {k: cfg[k] for k in ("friction_coef", "force_bound", "swing_height", "foot_offset", "ee_box_size", "force_cost")}


def quat_exp(v):
    """exp map of a rotation vector to an xyzw quaternion (plain closed form; inputs are far from the Taylor branch)."""
    v = np.asarray(v, float)
    t = np.linalg.norm(v)
    if t < 1e-12:
        return np.array([0.5 * v[0], 0.5 * v[1], 0.5 * v[2], 1.0])
    return np.concatenate([np.sin(t / 2) * v / t, [np.cos(t / 2)]])


def target_tangent(cfg):
    """srb_target in the tangent space (identity reference quaternion -> rotation part 0 for the shipped targets)."""
    s = np.asarray(cfg["srb_target"], float)
    assert np.allclose(s[6:10], [0, 0, 0, 1])
    return np.concatenate([s[:6], np.zeros(3), s[10:]])


def batched_trot_inputs(cfg, batch, seed=0):
    """Config #2: state_b = srb_init + U(-1,1)*[.05,.05,.02 m; 1,1,.5 kg m/s; axis-angle .1 rad; .05 x3],
    foot starts = nominal + U(-.02,.02) m in xy, t0 = 0."""
    rng = np.random.default_rng(seed)
    init = np.asarray(cfg["srb_init"], float)
    states = np.tile(init, (batch, 1))
    states[:, :3] += rng.uniform(-1, 1, (batch, 3)) * [0.05, 0.05, 0.02]
    states[:, 3:6] += rng.uniform(-1, 1, (batch, 3)) * [1.0, 1.0, 0.5]
    aa = rng.uniform(-1, 1, (batch, 3)) * 0.1
    for b in range(batch):
        states[b, 6:10] = quat_exp(aa[b])
    states[:, 10:] += rng.uniform(-1, 1, (batch, 3)) * 0.05
    ee = np.tile(EE_NOMINAL, (batch, 1, 1))
    ee[:, :, :2] += rng.uniform(-0.02, 0.02, (batch, 4, 2))
    return states, np.zeros(batch), ee


def disturbance_sweep_inputs(cfg, batch, seed=0):
    """Config #5: scenario_b starts from srb_init with the linear momentum 2.5 (cos phi, sin phi) U(.5, 1.5) kg m/s,
    phi ~ U(0, 2 pi), nominal feet, t0 = 0."""
    rng = np.random.default_rng(seed)
    init = np.asarray(cfg["srb_init"], float)
    states = np.tile(init, (batch, 1))
    phi = rng.uniform(0, 2 * np.pi, batch)
    mag = 2.5 * rng.uniform(0.5, 1.5, batch)
    states[:, 3] = mag * np.cos(phi)
    states[:, 4] = mag * np.sin(phi)
    ee = np.tile(EE_NOMINAL, (batch, 1, 1))
    return states, np.zeros(batch), ee
