// TEST INFRASTRUCTURE ONLY -- CPU oracle, never linked into the product.
//
// Inverse kinematics of the reference's controller:
//   * SingleRigidBodyModel::InverseKinematics      mpc/models/single_rigid_body_model.cpp:314-425
//   * SingleRigidBodyModel::ComputeJacobianForIK   :430-441
//   * SingleRigidBodyModel::GetEndEffectorLocations :443-455
//   * MPCController::GetTargetsFromTraj            controllers/mpc_controller.cpp:414-511
// The reference does the rigid-body arithmetic with pinocchio (forwardKinematics, updateFramePlacements,
// computeFrameJacobian / computeJointJacobian in the LOCAL frame, log6, Jlog6, integrate, neutral), which is absent from this
// image: the functions below restate pinocchio's published formulas (spatial/explog.hpp, multibody/liegroup/
// special-euclidean.hpp, algorithm/frames.hxx) -- "parity unpinned" against pinocchio itself; pinned by properties in
// tests/test_oracle_ik.py (Jlog6 and the Jacobians against finite differences, exp6 / log6 round trips, forward kinematics of
// the converged configuration hitting the targets) and, for the iteration built on top of them, by the reference's own
// InverseKinematics compiled over oracle/ref_shim/pinocchio (which forwards to these same functions).
#pragma once

namespace oracle {
namespace kin {

// One leg of the A1: hip (axis x), thigh (axis y), calf (axis y) revolute joints and the foot frame, each placed in the frame of
// the joint before it (the hip in the floating base; fixed joints merged the way pinocchio's URDF parser does).
struct LegChain {
    double t[4][3];      // translation of hip / thigh / calf joint and of the foot frame
    double R[4][9];      // rotation of the same placements, row-major
    double axis[3][3];   // joint axes in the joint frame
};
struct RobotKin {
    LegChain leg[4];     // FL, FR, RL, RR: pinocchio's joint order for a1.urdf (q = [p, quat xyzw, 4 x (hip, thigh, calf)])
};
struct Se3 {
    double R[9];
    double p[3];
};
constexpr int kNq = 19, kNv = 18;

void QuatToMatrix(const double q_xyzw[4], double R[9]);   // Eigen::Quaternion::toRotationMatrix (no normalisation)
void MatrixToQuat(const double R[9], double q_xyzw[4]);   // Eigen's rotation matrix -> quaternion, = pinocchio assignQuaternion
void Exp6(const double nu[6], Se3& M);                    // nu = (v, w)
void Log3(const double R[9], double w[3], double& theta);
void Jlog3(double theta, const double w[3], double J[9]);
void Log6(const Se3& M, double out[6]);
void Jlog6(const Se3& M, double J[36]);
Se3 Inverse(const Se3& a);
Se3 Mul(const Se3& a, const Se3& b);
Se3 ActInv(const Se3& a, const Se3& b);                   // a^-1 b

// joints[0] = floating base, joints[1 + 3 ee + j] = joint j of leg ee after its rotation; feet[ee] = foot frame.  World frame.
void ForwardKinematics(const RobotKin& rk, const double q[kNq], Se3 joints[13], Se3 feet[4]);
// computeFrameJacobian(model, data, q, foot frame, J): 6 x 18, LOCAL (expressed in the foot frame), rows (v, w)
void FootJacobianLocal(const RobotKin& rk, const Se3 joints[13], const Se3 feet[4], int ee, double J[6 * kNv]);
// pinocchio::integrate for a free-flyer followed by 12 revolute joints
void Integrate(const double q[kNq], const double v[kNv], double out[kNq]);

// SingleRigidBodyModel::InverseKinematics.  state: SRB manifold state [p, linear momentum, quat xyzw, angular momentum].
// Returns 0, or 1 where the reference throws "IK did not converge."; iters[ee] = iterations the loop of that foot ran.
int InverseKinematics(const RobotKin& rk, const double state[13], const double ee_des[4][3], const double joint_guess[12], double q_out[kNq],
                      int iters[4]);

}  // namespace kin

class Traj;
namespace kin {
// MPCController::GetTargetsFromTraj (controllers/mpc_controller.cpp:414-511): joint-space targets at `time` from the MPC
// trajectory.  q_des is the running IK guess on entry (q_des_) and the configuration target on exit; v_des = [p_dot, omega, finite
// difference of two IK solutions]; force_des = spline forces at `time`.
// Returns 0; 1 "IK did not converge."; 2 "bad interp."; 3 node + 1 beyond the trajectory (the reference's .at() throws).
int GetTargetsFromTraj(const RobotKin& rk, const Traj& traj, double time, double integrator_dt, double mass, const double Ir_inv[9],
                       double q_des[kNq], double v_des[kNv], double force_des[12]);
}  // namespace kin
}  // namespace oracle
