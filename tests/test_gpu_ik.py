"""GPU parity of the joint-space targets: bgg_ik_batch / bgg_targets_from_traj_batch (csrc/bgg_ik.cu) against the
oracle's restatement of SingleRigidBodyModel::InverseKinematics and MPCController::GetTargetsFromTraj.

Tolerance: the loop is a contraction (each iteration removes a tenth of the error), so rounding differences between the two
implementations do not grow: 1e-9 on q after ~90 iterations per foot; iteration counts equal."""
import numpy as np
import pytest

import common
from common import wl
import pyoracle as po
from test_oracle_ik import NOMINAL_JOINTS, random_problems

pytestmark = pytest.mark.gpu


def test_ik_batch_matches_the_oracle():
    kin = po.kin_flat(wl.robot())
    n = 96
    st, ee, guess = random_problems(n, 11)
    ee[5, 0] = [2.0, 2.0, 0.0]    # first foot unreachable: the reference throws
    ee[6, 2] = [-2.0, 2.0, 0.0]   # a later foot unreachable: the reference's success flag is already set, it carries on
    gpu = common.make_gpu("a1_configuration", 1)
    gpu.SetKinematics(wl.robot())
    out = gpu.InverseKinematics(st, ee, guess)
    for b in range(n):
        rc, q, it = po.ik(kin, st[b], ee[b], guess[b])
        assert out["status"][b] == rc, b
        if rc == 0 and b != 6:   # 1000 non-converging iterations towards an unreachable target are chaotic: only the outcome is compared
            assert np.array_equal(out["iters"][b], it), (b, out["iters"][b], it)
            assert np.abs(out["q"][b] - q).max() < 1e-9, b
    assert out["status"][5] == 1 and out["status"][6] == 0 and out["iters"][6][2] == 1000


def test_targets_from_traj_match_the_oracle_along_a_solved_horizon():
    cfg_name = "a1_configuration"
    cfg = wl.CONFIGS[cfg_name]
    B = 6
    states, t0, ee0 = wl.batched_trot_inputs(cfg, B, seed=3)
    gpu = common.make_gpu(cfg_name, B, states)
    gpu.SetKinematics(wl.robot())
    kin, rob, dt = po.kin_flat(wl.robot()), wl.robot(), cfg["integrator_dt"]
    oracles = [common.make_oracle(cfg_name, states[b]) for b in range(B)]
    for _ in range(3):
        for b, o in enumerate(oracles):
            o.solve(states[b], 0.0, ee0[b], real_time=True)
    for b, o in enumerate(oracles):   # same trajectories on both sides: what is compared is the targets computation
        common.mirror_oracle_to_gpu(o, gpu, b)
    q0 = np.concatenate([states[:, :3], states[:, 6:10], np.tile(NOMINAL_JOINTS, (B, 1))], axis=1)
    times = np.array([0.0, 0.013, 0.05, 0.12, 0.31, 0.049999])
    import mpc_controller
    ctl = mpc_controller.MPCController(gpu, gait_opt_freq=5)   # the C++ class (host/mpc_controller_b200.cpp) keeps q_des_ between calls
    ctl.SetInitialConfig(q0)
    for step in range(3):
        out = ctl.GetTargetsFromTraj(times)
        if step == 0:   # the plain C-ABI call gives the same numbers
            direct = gpu.GetTargetsFromTraj(times, q0)
            assert np.array_equal(direct["q_des"], out["q_des"]) and np.array_equal(direct["v_des"], out["v_des"])
        for b, o in enumerate(oracles):
            rc, q, v, f = po.targets_from_traj(o, kin, rob, times[b], dt, q0[b])
            assert out["status"][b] == rc == 0, (b, out["status"][b], rc)
            assert np.abs(out["q_des"][b] - q).max() < 1e-9
            assert np.abs(out["v_des"][b] - v).max() < 1e-9 / dt * 10   # a difference of two IK solutions over dt
            assert np.abs(out["force_des"][b] - f).max() < 1e-9
        q0 = out["q_des"]
        times = times + 0.004
    late = gpu.GetTargetsFromTraj(10.0, q0)
    assert np.all(late["status"] == 3)
