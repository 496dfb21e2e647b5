// TEST INFRASTRUCTURE ONLY.  Stand-in for the pinocchio calls the reference's model sources make, so that
// mpc/models/model.cpp and single_rigid_body_model.cpp compile unmodified from /root/reference (pinocchio is not in this
// image).  What the MPC hot path takes from pinocchio is FUNCTIONAL here:
//   * the robot constants read at construction (computeTotalMass; computeCentroidalMap -> oMi[1].actInv(oYcrb[0]).inertia();
//     oMi[hip joint].translation()): injected through pinocchio::stub::consts() by the test driver -- the same numbers the
//     oracle and the CUDA path get (tests/golden/a1_robot_consts.json);
//   * quaternion::log3 / exp3 / firstOrderNormalize: forwarded to the oracle's restatement of pinocchio's published formulas
//     (oracle/srb_mpc.cpp; unpinned third-party arithmetic) -- one restatement, not two.
//   * what SingleRigidBodyModel::InverseKinematics / GetEndEffectorLocations call (forwardKinematics, updateFramePlacements,
//     computeFrameJacobian / computeJointJacobian in the LOCAL frame, log6, Jlog6, integrate, neutral): forwarded to the oracle's
//     restatement of pinocchio's published formulas (oracle/leg_kinematics.cpp), with the A1 leg chains injected through
//     pinocchio::stub::consts().kin -- again one restatement, so the reference's own IK loop runs over the same arithmetic.
// The whole-body helpers (crba, nonLinearEffects, Jacobian time variations) are declared so the files compile and throw when called.
#pragma once
#include <Eigen/Core>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../leg_kinematics.hpp"

namespace oracle {   // oracle/srb_mpc.cpp
void QuatLog3(const double q[4], double out[3]);
void QuatExp3(const double v[3], double q[4]);
void QuatFirstOrderNormalize(double q[4]);
}  // namespace oracle

namespace pinocchio {

enum ReferenceFrame { WORLD = 0, LOCAL = 1, LOCAL_WORLD_ALIGNED = 2 };
struct JointModelFreeFlyer {};

namespace stub {
struct Consts {
    double mass = 0;
    Eigen::Matrix3d Ir = Eigen::Matrix3d::Identity();
    std::map<std::string, Eigen::Vector3d> joint_translation;   // oMi[joint].translation() after computeCentroidalMap(nom_state)
    oracle::kin::RobotKin kin;                                    // leg chains (tests/golden/a1_robot_consts.json["legs"])
    bool have_kin = false;
};
inline Consts& consts() { static Consts c; return c; }
[[noreturn]] inline void unavailable(const char* what) { throw std::runtime_error(std::string("ref_shim/pinocchio: ") + what + " is a compile-only stand-in"); }
}  // namespace stub

struct Inertia {
    Eigen::Matrix3d I = Eigen::Matrix3d::Zero();
    Eigen::Matrix3d inertia() const { return I; }
};

struct Motion {
    Eigen::Matrix<double, 6, 1> v;
    Eigen::Matrix<double, 6, 1> toVector() const { return v; }
};

struct SE3 {
    Eigen::Matrix3d R = Eigen::Matrix3d::Identity();
    Eigen::Vector3d p = Eigen::Vector3d::Zero();
    SE3() {}
    SE3(const Eigen::MatX& R_, const Eigen::MatX& p_) : R(R_), p(p_) {}
    static SE3 Identity() { return SE3(); }
    const Eigen::Vector3d& translation() const { return p; }
    Eigen::Vector3d& translation() { return p; }
    const Eigen::Matrix3d& rotation() const { return R; }
    Eigen::Matrix3d& rotation() { return R; }
    SE3 inverse() const { SE3 o; o.R = R.transpose(); o.p = -(o.R * p); return o; }
    SE3 actInv(const SE3& m) const { SE3 o; o.R = R.transpose() * m.R; o.p = R.transpose() * (m.p - p); return o; }
    Inertia actInv(const Inertia& y) const { return y; }   // stub: oYcrb[0] already holds the inertia expressed in the base frame
};

struct Frame { std::string name; };

struct Model {
    int nq = 19, nv = 18, njoints = 14;
    std::vector<Frame> frames;
    std::vector<std::string> names;
    int getJointId(const std::string& n) const {
        for (size_t i = 0; i < names.size(); i++) if (names[i] == n) return static_cast<int>(i);
        return static_cast<int>(names.size());
    }
    int getFrameId(const std::string& n) const {
        for (size_t i = 0; i < frames.size(); i++) if (frames[i].name == n) return static_cast<int>(i);
        return static_cast<int>(frames.size());
    }
};

struct Data {
    typedef Eigen::Matrix<double, 6, Eigen::Dynamic> Matrix6x;
    typedef Eigen::Matrix<double, 6, 6> Matrix6;
    std::vector<SE3> oMi, oMf, feet_;
    std::vector<Inertia> oYcrb;
    Eigen::MatrixXd M;
    Data() {}
    explicit Data(const Model& m) : oMi(m.names.size()), oMf(m.frames.size()), oYcrb(m.names.size()), M(m.nv, m.nv) {}
};

namespace urdf {
// the A1 joint list in URDF order (models/a1_description/urdf/a1.urdf) and the four foot frames
inline void buildModel(const std::string&, const JointModelFreeFlyer&, Model& m, bool = false) {
    m.names = {"universe", "root_joint", "FL_hip_joint", "FL_thigh_joint", "FL_calf_joint", "FR_hip_joint", "FR_thigh_joint", "FR_calf_joint",
               "RL_hip_joint", "RL_thigh_joint", "RL_calf_joint", "RR_hip_joint", "RR_thigh_joint", "RR_calf_joint"};
    m.njoints = static_cast<int>(m.names.size());
    for (const char* f : {"universe", "root_joint", "trunk", "FL_foot", "FR_foot", "RL_foot", "RR_foot"}) m.frames.push_back(Frame{f});
}
}  // namespace urdf

inline double computeTotalMass(const Model&) { return stub::consts().mass; }
inline void computeCentroidalMap(const Model& m, Data& d, const Eigen::MatX&) {
    d.oYcrb[0].I = stub::consts().Ir;
    for (size_t i = 0; i < m.names.size(); i++) {
        auto it = stub::consts().joint_translation.find(m.names[i]);
        d.oMi[i].p = (it == stub::consts().joint_translation.end()) ? Eigen::Vector3d::Zero() : it->second;
    }
}

namespace quaternion {
// pinocchio/spatial/explog-quaternion.hpp, math/quaternion.hpp: third-party arithmetic that is absent here; ONE restatement
// serves the oracle and this stand-in (oracle/srb_mpc.cpp: QuatLog3 / QuatExp3 / QuatFirstOrderNormalize, linked in)
inline Eigen::Vector3d log3(const Eigen::Quaterniond& q) {
    const double in[4] = {q.x(), q.y(), q.z(), q.w()};
    double out[3];
    oracle::QuatLog3(in, out);
    return Eigen::Vector3d(out[0], out[1], out[2]);
}
inline void exp3(const Eigen::MatX& v, Eigen::Quaterniond& q) {
    const double in[3] = {v(0), v(1), v(2)};
    double out[4];
    oracle::QuatExp3(in, out);
    q.x() = out[0]; q.y() = out[1]; q.z() = out[2]; q.w() = out[3];
}
inline void firstOrderNormalize(Eigen::Quaterniond& q) {
    double c[4] = {q.x(), q.y(), q.z(), q.w()};
    oracle::QuatFirstOrderNormalize(c);
    q.x() = c[0]; q.y() = c[1]; q.z() = c[2]; q.w() = c[3];
}
}  // namespace quaternion

// ---- inverse kinematics: functional, over oracle/leg_kinematics.cpp.  Joint i of the model is joints[i - 1] there (0 = universe);
// frames 3 .. 6 are the four feet, every other frame sits on the floating base.
namespace stub {
inline void q_in(const Eigen::MatX& q, double out[oracle::kin::kNq]) {
    if (!consts().have_kin) throw std::runtime_error("ref_shim/pinocchio: leg chains were not injected");
    for (int i = 0; i < oracle::kin::kNq; i++) out[i] = q(i);
}
inline SE3 se3(const oracle::kin::Se3& m) {
    SE3 o;
    for (int i = 0; i < 3; i++) { o.p(i) = m.p[i]; for (int j = 0; j < 3; j++) o.R(i, j) = m.R[3 * i + j]; }
    return o;
}
inline oracle::kin::Se3 se3(const SE3& m) {
    oracle::kin::Se3 o;
    for (int i = 0; i < 3; i++) { o.p[i] = m.p(i); for (int j = 0; j < 3; j++) o.R[3 * i + j] = m.R(i, j); }
    return o;
}
}  // namespace stub
inline void forwardKinematics(const Model& m, Data& d, const Eigen::MatX& q) {
    double qa[oracle::kin::kNq];
    stub::q_in(q, qa);
    oracle::kin::Se3 joints[13], feet[4];
    oracle::kin::ForwardKinematics(stub::consts().kin, qa, joints, feet);
    d.oMi[0] = SE3();
    for (int i = 1; i < m.njoints; i++) d.oMi[i] = stub::se3(joints[i - 1]);
    d.feet_.assign(4, SE3());
    for (int e = 0; e < 4; e++) d.feet_[e] = stub::se3(feet[e]);
}
inline void updateFramePlacements(const Model& m, Data& d) {
    for (size_t f = 0; f < m.frames.size(); f++) d.oMf[f] = (f >= 3 && f < 7) ? d.feet_[f - 3] : (f == 0 ? SE3() : d.oMi[1]);
}
inline void framesForwardKinematics(const Model& m, Data& d, const Eigen::MatX& q) { forwardKinematics(m, d, q); updateFramePlacements(m, d); }
inline void computeFrameJacobian(const Model&, Data&, const Eigen::MatX& q, int frame_id, Eigen::MatX& J) {   // LOCAL is pinocchio's default here
    if (frame_id < 3 || frame_id > 6) stub::unavailable("computeFrameJacobian of a frame that is not a foot");
    double qa[oracle::kin::kNq], Ja[6 * oracle::kin::kNv];
    stub::q_in(q, qa);
    oracle::kin::Se3 joints[13], feet[4];
    oracle::kin::ForwardKinematics(stub::consts().kin, qa, joints, feet);
    oracle::kin::FootJacobianLocal(stub::consts().kin, joints, feet, frame_id - 3, Ja);
    for (int i = 0; i < 6; i++) for (int j = 0; j < oracle::kin::kNv; j++) J(i, j) = Ja[oracle::kin::kNv * i + j];
}
inline void computeFrameJacobian(const Model& m, Data& d, const Eigen::MatX& q, int frame_id, ReferenceFrame rf, Eigen::MatX& J) {
    if (rf != LOCAL) stub::unavailable("computeFrameJacobian outside the LOCAL frame");
    computeFrameJacobian(m, d, q, frame_id, J);
}
inline void computeJointJacobian(const Model&, Data&, const Eigen::MatX&, int joint_id, Eigen::MatX& J) {
    if (joint_id != 1) stub::unavailable("computeJointJacobian of a joint that is not the floating base");
    for (int i = 0; i < 6; i++) for (int j = 0; j < oracle::kin::kNv; j++) J(i, j) = (i == j) ? 1.0 : 0.0;   // free flyer, LOCAL: S = I6
}
inline Eigen::VectorXd neutral(const Model& m) {
    Eigen::VectorXd q = Eigen::VectorXd::Zero(m.nq);
    q(6) = 1.0;
    return q;
}
inline Eigen::VectorXd integrate(const Model& m, const Eigen::MatX& q, const Eigen::MatX& v) {
    double qa[oracle::kin::kNq], va[oracle::kin::kNv], out[oracle::kin::kNq];
    stub::q_in(q, qa);
    for (int i = 0; i < oracle::kin::kNv; i++) va[i] = v(i);
    oracle::kin::Integrate(qa, va, out);
    Eigen::VectorXd r = Eigen::VectorXd::Zero(m.nq);
    for (int i = 0; i < oracle::kin::kNq; i++) r(i) = out[i];
    return r;
}
inline Motion log6(const SE3& M) {
    double out[6];
    oracle::kin::Log6(stub::se3(M), out);
    Motion mo;
    mo.v = Eigen::Matrix<double, 6, 1>::Zero();
    for (int i = 0; i < 6; i++) mo.v(i) = out[i];
    return mo;
}
inline void Jlog6(const SE3& M, Eigen::MatX& J) {
    double Ja[36];
    oracle::kin::Jlog6(stub::se3(M), Ja);
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) J(i, j) = Ja[6 * i + j];
}

// ---- whole-body helpers, off every path built here: compile-only
inline void computeJointJacobiansTimeVariation(const Model&, Data&, const Eigen::MatX&, const Eigen::MatX&) { stub::unavailable("computeJointJacobiansTimeVariation"); }
inline void getFrameJacobianTimeVariation(const Model&, Data&, int, ReferenceFrame, Eigen::MatX&) { stub::unavailable("getFrameJacobianTimeVariation"); }
inline void crba(const Model&, Data&, const Eigen::MatX&) { stub::unavailable("crba"); }
inline Eigen::VectorXd nonLinearEffects(const Model&, Data&, const Eigen::MatX&, const Eigen::MatX&) { stub::unavailable("nonLinearEffects"); }

}  // namespace pinocchio
