// TEST INFRASTRUCTURE ONLY.  C entry points over the REFERENCE'S OWN mpc::MPCSingleRigidBody, compiled from /root/reference
// against the stand-in headers of this directory (../Makefile, target _ref/libref_mpc.so).  Same function names and
// argument meaning as the orc_mpc_* functions of ../oracle_capi.cpp, so that tests drive the oracle's restatement and the
// reference's code through one Python wrapper (pyoracle.SrbMpc(which="ref")).
//
// What is the reference's and what is not:
//   reference sources, unmodified : trajectory, splines, sparse_matrix_builder, qp_data, model, single_rigid_body_model, mpc,
//                                   mpc_single_rigid_body, rk_integrator, qp_interface, clarabel_interface, gait_optimizer
//   stand-ins written here        : Eigen (dense / sparse containers, eager evaluation), pinocchio (robot constants injected,
//                                   quaternion maps shared with the oracle, kinematics compile-only), the Clarabel SOLVER (the
//                                   oracle's restatement of its algorithm behind clarabel::DefaultSolver), OsqpEigen (compile-only)
#include <cstring>
#include <limits>
#include <memory>
#include <string>

#include "mpc_single_rigid_body.h"
#include "OsqpEigen/OsqpEigen.h"
#define private public   // read access to GaitOptimizer::dHdth (every header gait_optimizer.h pulls in is already included above)
#include "gait_optimizer.h"
#undef private

#include "../qp_ipm.hpp"

using namespace mpc;

// ------------------------------------------------------------------------------------------------ the solver stand-in
namespace {
oracle::IpmSettings g_ipm;
int g_last_iters = 0;
double g_last_prim = 0, g_last_dual = 0;
oracle::Csc ToCsc(const Eigen::SparseMatrix<double>& m) {
    oracle::Csc c;
    c.rows = m.rows();
    c.cols = m.cols();
    c.colptr.assign(m.outerIndexPtr(), m.outerIndexPtr() + m.cols() + 1);
    c.rowidx.assign(m.innerIndexPtr(), m.innerIndexPtr() + m.nonZeros());
    c.val.assign(m.valuePtr(), m.valuePtr() + m.nonZeros());
    return c;
}
}  // namespace

namespace clarabel {
namespace detail {
void Solve(const Eigen::SparseMatrix<double>& P, const Eigen::VectorXd& q, const Eigen::SparseMatrix<double>& A, const Eigen::VectorXd& b,
           const std::vector<SupportedConeT<double>>& cones, const DefaultSettings<double>& /*settings*/, DefaultSolution<double>& out) {
    // Tolerances: the oracle's (Clarabel's documented defaults), not the 1e-15 gap the reference asks for -- see
    // oracle/qp_ipm.hpp.  Equality rows and the dynamics block (first Zero cone) from the cone list the reference built.
    std::vector<char> is_eq;
    int num_dyn = 0;
    for (size_t i = 0; i < cones.size(); i++) {
        is_eq.insert(is_eq.end(), cones[i].dim, cones[i].zero ? 1 : 0);
        if (i == 0 && cones[i].zero) num_dyn = cones[i].dim;
    }
    const oracle::Vec qv(q.data(), q.data() + q.size()), bv(b.data(), b.data() + b.size());
    const oracle::IpmResult r = oracle::IpmSolveMpcOrder(ToCsc(P), qv, ToCsc(A), bv, is_eq, num_dyn, g_ipm);
    out.x.resize(static_cast<int>(r.x.size()));
    out.z.resize(static_cast<int>(r.y.size()));
    out.s.resize(static_cast<int>(r.s.size()));
    for (size_t i = 0; i < r.x.size(); i++) out.x(static_cast<int>(i)) = r.x[i];
    for (size_t i = 0; i < r.y.size(); i++) out.z(static_cast<int>(i)) = r.y[i];
    for (size_t i = 0; i < r.s.size(); i++) out.s(static_cast<int>(i)) = r.s[i];
    // Rows of A without any stored entry (the touch-down sample of every stance) read 0 + s = b: the restated solver keeps them out
    // of its iteration and reports s = b, exactly 0 for the cone / lower-force rows, where an interior-point method (Clarabel) ends
    // on a tiny POSITIVE slack.  The reference's derivative system (clarabel_interface.cpp:262-602) has D(s) on its diagonal and is
    // singular with an exact zero there; the row decouples whatever the positive value, so 1 is reported (as oracle/gait_oracle.py).
    {
        std::vector<int> cnt(A.rows(), 0);
        for (int k = 0; k < A.nonZeros(); k++) cnt[A.innerIndexPtr()[k]]++;
        for (int i = 0; i < A.rows(); i++)
            if (!is_eq[i] && cnt[i] == 0 && out.s(i) == 0.0) out.s(i) = 1.0;
    }
    out.iterations = r.iters;
    g_last_iters = r.iters;
    g_last_prim = r.prim_res;
    g_last_dual = r.dual_res;
    switch (r.status) {
        case oracle::Solved: out.status = SolverStatus::Solved; break;
        case oracle::SolvedInacc: out.status = SolverStatus::AlmostSolved; break;
        case oracle::MaxIter: out.status = SolverStatus::MaxIterations; break;
        case oracle::PrimalInfeasible: out.status = SolverStatus::PrimalInfeasible; break;
        case oracle::PrimalInfeasibleInacc: out.status = SolverStatus::AlmostPrimalInfeasible; break;
        case oracle::DualInfeasible: out.status = SolverStatus::DualInfeasible; break;
        case oracle::DualInfeasibleInacc: out.status = SolverStatus::AlmostDualInfeasible; break;
        case oracle::Unsolved: out.status = SolverStatus::Unsolved; break;
        default: out.status = SolverStatus::NumericalError; break;
    }
}
}  // namespace detail
}  // namespace clarabel

// ------------------------------------------------------------------------------------------------ C API
namespace {
// The reference prints timing lines to std::cout from inside its derivative code whatever the verbosity
// (clarabel_interface.cpp:592-600); they are muted here (test logs), std::cerr -- "Primal infeasible. ..." -- is left alone.
struct MuteCout { MuteCout() { std::cout.setstate(std::ios_base::failbit); } } g_mute_cout;
thread_local std::string g_err;
struct Probe : public MPCSingleRigidBody {   // read access to protected members of the reference class
    using MPCSingleRigidBody::MPCSingleRigidBody;
    const vector_t& PrevQpSol() const { return prev_qp_sol; }
    const Trajectory& Traj() const { return prev_traj_; }
    double Alpha() const { return alpha_.empty() ? 0.0 : alpha_.back(); }
    double EqViolation() const { return equality_constraint_violations_.empty() ? 0.0 : equality_constraint_violations_.back(); }
    double StepNorm() const { return step_norm_.empty() ? 0.0 : step_norm_.back(); }
    double CostResult() const { return cost_result_.empty() ? 0.0 : cost_result_.back(); }
    double MeritResult() const { return merit_result_.empty() ? 0.0 : merit_result_.back(); }
    double MeritDd() const { return merit_directional_deriv_.empty() ? 0.0 : merit_directional_deriv_.back(); }
    const Eigen::Vector2d& EeBox() const { return info_.ee_box_size; }   // the adapted size (IncreaseEEBox / DecreaseEEBox)
    const vector_t& DualSol() const { return prev_dual_sol_; }
    double InitTime() const { return init_time_; }
    const matrix_t& NodeB() const { return B_; }
};
struct Handle {
    std::unique_ptr<Probe> mpc;
    std::unique_ptr<GaitOptimizer> gait;
    double force_cost = 0;
};
Probe& M(void* h) { return *static_cast<Handle*>(h)->mpc; }
std::vector<vector_3t> EE(const double* ee) {
    std::vector<vector_3t> v(4);
    for (int e = 0; e < 4; e++) v[e] = vector_3t(ee[3 * e], ee[3 * e + 1], ee[3 * e + 2]);
    return v;
}
vector_t V(const double* p, int n) {
    vector_t v(n);
    for (int i = 0; i < n; i++) v(i) = p[i];
    return v;
}
}  // namespace
#define TRY try {
#define CATCH(ret) } catch (const std::exception& e) { g_err = e.what(); return ret; } catch (const std::string& s) { g_err = s; return ret; }

struct OrcMpcInfo {
    int num_nodes;
    double friction_coef, integrator_dt, force_bound, swing_height, foot_offset, ee_box_x, ee_box_y, force_cost;
};
struct OrcRobotConsts {
    double mass, Ir[9], Ir_inv[9], hip_xy[8], gravity[3];
};

extern "C" {
const char* orc_mpc_last_error() { return g_err.c_str(); }

// hip_raw: oMi[hip joint].translation() - oMi[root].translation() for FL, FR, RL, RR (the reference adds its own +-0.1 / +0.025
// offsets on top, single_rigid_body_model.cpp:289-305)
void* orc_mpc_create_ref(const OrcMpcInfo* ci, const OrcRobotConsts* cr, const double* hip_raw) {
    TRY
    auto& c = pinocchio::stub::consts();
    c.mass = cr->mass;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) c.Ir(i, j) = cr->Ir[3 * i + j];
    const char* hips[4] = {"FL_hip_joint", "FR_hip_joint", "RL_hip_joint", "RR_hip_joint"};
    for (int e = 0; e < 4; e++) c.joint_translation[hips[e]] = Eigen::Vector3d(hip_raw[3 * e], hip_raw[3 * e + 1], hip_raw[3 * e + 2]);
    MPCInfo info;
    info.num_nodes = ci->num_nodes;
    info.num_qp_iterations = 1;
    info.num_contacts = 4;
    info.friction_coef = ci->friction_coef;
    info.ee_frames = {"FL_foot", "FR_foot", "RL_foot", "RR_foot"};
    info.discretization_steps = 1;
    info.num_switches = 10;
    info.integrator_dt = ci->integrator_dt;
    info.force_bound = ci->force_bound;
    info.swing_height = ci->swing_height;
    info.foot_offset = ci->foot_offset;
    info.nom_state = vector_t::Zero(19);
    info.ee_box_size = Eigen::Vector2d(ci->ee_box_x, ci->ee_box_y);
    info.real_time_iters = 6000;
    info.verbose = Nothing;
    info.force_cost = ci->force_cost;
    auto* h = new Handle;
    h->force_cost = ci->force_cost;
    h->mpc.reset(new Probe(info, std::string("a1.urdf")));
    return h;
    CATCH(nullptr)
}
void orc_mpc_destroy(void* h) { delete static_cast<Handle*>(h); }

// SingleRigidBodyModel::InverseKinematics (single_rigid_body_model.cpp:314-425) as the reference wrote it, over the pinocchio
// stand-in.  kin_flat: the 228 packed doubles of oracle::kin::RobotKin.  Returns 0, 1 when it throws "IK did not converge.", -1 else.
int orc_ref_ik(const double* kin_flat, const double* state, const double* ee_des, const double* joint_guess, double* q_out) {
    try {
        auto& c = pinocchio::stub::consts();
        std::memcpy(&c.kin, kin_flat, sizeof c.kin);
        c.have_kin = true;
        if (c.mass == 0) c.mass = 1.0;   // the model's constructor reads mass / inertia; any positive value serves the IK
        SingleRigidBodyModel model("a1.urdf", {"FL_foot", "FR_foot", "RL_foot", "RR_foot"}, 1, 0.05, vector_t::Zero(19));
        man_state_t st = man_state_t::Zero(13);
        for (int i = 0; i < 13; i++) st(i) = state[i];
        std::vector<vector_3t> ee(4);
        for (int e = 0; e < 4; e++) ee[e] = vector_3t(ee_des[3 * e], ee_des[3 * e + 1], ee_des[3 * e + 2]);
        vector_t guess = vector_t::Zero(19);
        for (int i = 0; i < 12; i++) guess(7 + i) = joint_guess[i];
        const vector_t ub = vector_t::Zero(12), lb = vector_t::Zero(12);
        const vector_t q = model.InverseKinematics(st, ee, guess, ub, lb);
        for (int i = 0; i < 19; i++) q_out[i] = q(i);
        return 0;
    } catch (const std::runtime_error& e) {
        g_err = e.what();
        return g_err == "IK did not converge." ? 1 : -1;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}
void orc_mpc_set_ipm(void*, double tol_feas, double tol_gap, int max_iter, int refine) {
    if (tol_feas > 0) g_ipm.tol_feas = tol_feas;
    if (tol_gap > 0) g_ipm.tol_gap = tol_gap;
    if (max_iter > 0) g_ipm.max_iter = max_iter;
    if (refine >= 0) g_ipm.refine = refine;
}
// the call sequence of test/mpc_test.cpp:17-39 (CreateMPC)
void orc_mpc_set_costs(void* h, const double* state_des, const double* Q, const double* Phi, const double* Phi_w) {
    matrix_t Qm(12, 12), Pm(12, 12);
    for (int i = 0; i < 12; i++) for (int j = 0; j < 12; j++) { Qm(i, j) = Q[12 * i + j]; Pm(i, j) = Phi[12 * i + j]; }
    M(h).AddQuadraticTrackingCost(V(state_des, 12), Qm);
    M(h).AddForceCost(static_cast<Handle*>(h)->force_cost);
    M(h).SetQuadraticFinalCost(Pm);
    M(h).SetLinearFinalCost(V(Phi_w, 12));
}
void orc_mpc_set_warm_states(void* h, const double* states) {
    const int N = M(h).GetQPData().num_dynamics_constraints / 12 - 1;
    std::vector<vector_t> s;
    for (int i = 0; i <= N; i++) s.push_back(V(states + 13 * i, 13));
    M(h).SetStateTrajectoryWarmStart(s);
}
int orc_mpc_solve(void* h, const double* state, double t0, const double* ee_start, int real_time) {
    TRY
    if (real_time) M(h).GetRealTimeUpdate(V(state, 13), t0, EE(ee_start), false);
    else M(h).Solve(V(state, 13), t0, EE(ee_start));
    return M(h).GetSolveQuality();
    CATCH(-1)
}
int orc_mpc_initial_run(void* h, const double* state, const double* ee_start) {
    TRY M(h).CreateInitialRun(V(state, 13), EE(ee_start));
    return M(h).GetSolveQuality();
    CATCH(-1)
}
void orc_mpc_sizes(void* h, int* out) {
    const QPData& d = M(h).GetQPData();
    out[0] = d.num_decision_vars; out[1] = d.GetTotalNumConstraints(); out[2] = d.sparse_constraint_.nonZeros(); out[3] = d.sparse_cost_.nonZeros();
    out[4] = d.num_dynamics_constraints; out[5] = d.num_force_box_constraints_; out[6] = d.num_cone_constraints_; out[7] = d.num_ee_location_constraints_;
    out[8] = d.num_td_pos_constraints_; out[9] = d.num_start_ee_constraints_; out[10] = M(h).Traj().GetTotalForceSplineVars();
    out[11] = M(h).Traj().GetTotalPosSplineVars(); out[12] = d.num_equality_; out[13] = d.num_inequality_;
}
void orc_mpc_get_A(void* h, int* colptr, int* rowidx, double* val) {
    const auto& A = M(h).GetQPData().sparse_constraint_;
    std::copy(A.outerIndexPtr(), A.outerIndexPtr() + A.cols() + 1, colptr);
    std::copy(A.innerIndexPtr(), A.innerIndexPtr() + A.nonZeros(), rowidx);
    std::copy(A.valuePtr(), A.valuePtr() + A.nonZeros(), val);
}
void orc_mpc_get_P(void* h, int* colptr, int* rowidx, double* val) {
    const auto& P = M(h).GetQPData().sparse_cost_;
    std::copy(P.outerIndexPtr(), P.outerIndexPtr() + P.cols() + 1, colptr);
    std::copy(P.innerIndexPtr(), P.innerIndexPtr() + P.nonZeros(), rowidx);
    std::copy(P.valuePtr(), P.valuePtr() + P.nonZeros(), val);
}
void orc_mpc_get_vectors(void* h, double* q, double* ub, char* is_eq) {
    const QPData& d = M(h).GetQPData();
    for (int i = 0; i < d.cost_linear.size(); i++) q[i] = d.cost_linear(i);
    for (int i = 0; i < d.ub_.size(); i++) ub[i] = d.ub_(i);
    int row = 0;   // cone kinds in constraint-list order, as ClarabelInterface::SetupQP builds them (clarabel_interface.cpp:29-64)
    for (const auto& c : d.constraints_) {
        int n = 0, eq = 0;
        switch (c) {
            case Constraints::Dynamics: n = d.num_dynamics_constraints; eq = 1; break;
            case Constraints::ForceBox: n = d.num_force_box_constraints_; break;
            case Constraints::FrictionCone: n = d.num_cone_constraints_; break;
            case Constraints::EndEffectorLocation: n = d.num_ee_location_constraints_; break;
            case Constraints::TDPosition: n = d.num_td_pos_constraints_; eq = 1; break;
            case Constraints::EndEffectorStart: n = d.num_start_ee_constraints_; eq = 1; break;
            case Constraints::Raibert: n = d.num_raibert_constraints_; eq = 1; break;
            default: break;
        }
        for (int i = 0; i < n; i++) is_eq[row++] = static_cast<char>(eq);
    }
}
void orc_mpc_get_prev_qp_sol(void* h, double* z) {
    const vector_t& v = M(h).PrevQpSol();
    for (int i = 0; i < v.size(); i++) z[i] = v(i);
}
void orc_mpc_get_qp_solution(void* h, double* x, double* dual, double* /*slack*/, double* info) {
    if (dual) { const vector_t& v = M(h).DualSol(); for (int i = 0; i < v.size(); i++) dual[i] = v(i); }
    (void)x;
    info[0] = M(h).GetSolveQuality(); info[1] = g_last_iters; info[2] = g_last_prim; info[3] = g_last_dual;
}
void orc_mpc_get_stats(void* h, double* out) {
    out[0] = M(h).Alpha(); out[1] = M(h).EqViolation(); out[2] = M(h).StepNorm(); out[3] = M(h).CostResult(); out[4] = M(h).MeritResult();
    out[5] = M(h).MeritDd(); out[6] = M(h).GetSolveQuality(); out[7] = g_last_iters; out[8] = M(h).EeBox()(0); out[9] = M(h).EeBox()(1);
}
void orc_mpc_get_states(void* h, double* states) {
    const Trajectory& t = M(h).Traj();
    const int N = M(h).GetQPData().num_dynamics_constraints / 12;
    for (int i = 0; i < N; i++) { const vector_t s = t.GetState(i); for (int k = 0; k < 13; k++) states[13 * i + k] = s(k); }
}
double orc_mpc_init_time(void* h) { return M(h).InitTime(); }
double orc_mpc_cost(void* h) { return M(h).GetCost(); }
void orc_mpc_force_at(void* h, int ee, double t, double* out) { const auto f = M(h).Traj().GetForce(ee, t); for (int c = 0; c < 3; c++) out[c] = f(c); }
void orc_mpc_ee_at(void* h, int ee, double t, double* out) { const auto f = M(h).Traj().GetEndEffectorLocation(ee, t); for (int c = 0; c < 3; c++) out[c] = f(c); }
int orc_mpc_num_contacts(void* h, int ee) { return static_cast<int>(M(h).Traj().GetContactTimes().at(ee).size()); }
void orc_mpc_get_contact_times(void* h, int ee, double* t, int* type) {
    const auto ct = M(h).Traj().GetContactTimes();
    for (size_t i = 0; i < ct.at(ee).size(); i++) { t[i] = ct[ee][i].GetTime(); type[i] = static_cast<int>(ct[ee][i].GetType()); }
}
int orc_mpc_set_contact_times(void* h, int ee, const double* t, int n) {
    TRY
    auto ct = M(h).Traj().GetContactTimes();
    if (static_cast<int>(ct.at(ee).size()) != n) throw std::runtime_error("contact time count mismatch");
    for (int i = 0; i < n; i++) ct[ee][i].SetTime(t[i]);
    M(h).UpdateContactTimes(ct);
    return 0;
    CATCH(-1)
}

// MPC::AdjustForCurrentContacts (mpc.cpp:1195-1203), the reference's own
int orc_mpc_adjust_for_contacts(void* h, double time, const int* in_contact) {
    TRY
    controller::Contact c(4);
    for (int ee = 0; ee < 4; ee++) c.in_contact_.at(ee) = in_contact[ee] != 0;
    M(h).AdjustForCurrentContacts(time, c);
    return 0;
    CATCH(-1)
}

// The derivative chain of MPCController::GaitOpt (controllers/mpc_controller.cpp:518-552), run on the reference's own
// ClarabelInterface::SetupDerivativeCalcs / CalcDerivativeWrtMats / Vecs, MPCSingleRigidBody::ComputeParamPartialsClarabel and
// GaitOptimizer::ComputeCostFcnDerivWrtContactTimes.  out: dH/dtheta, foot-major; returns the number of contact times,
// 0 when the last solve was not `Solved` (the reference's calls return false), -1 on error.
int orc_mpc_gait_gradient(void* hv, double* out, int cap) {
    TRY
    Handle* h = static_cast<Handle*>(hv);
    Probe& mpc = *h->mpc;
    if (!mpc.ComputeDerivativeTerms()) return 0;
    const Trajectory traj = mpc.GetTrajectory();
    if (!h->gait) h->gait.reset(new GaitOptimizer(4, 10, 10, 10, 1, 0.05));   // as MPCController constructs it (mpc_controller.cpp:47)
    GaitOptimizer& g = *h->gait;
    g.SetContactTimes(traj.GetContactTimes());
    g.UpdateSizes(mpc.GetNumDecisionVars(), mpc.GetNumConstraints());
    if (!mpc.GetQPPartials(g.GetQPPartials())) return 0;
    for (int ee = 0; ee < 4; ee++) {
        g.SetNumContactTimes(ee, traj.GetNumContactNodes(ee));
        for (int idx = 0; idx < traj.GetNumContactNodes(ee); idx++)
            mpc.ComputeParamPartialsClarabel(traj, g.GetParameterPartials(ee, idx), ee, idx);
    }
    g.ModifyQPPartials(mpc.GetQPSolution());
    g.ComputeCostFcnDerivWrtContactTimes();
    const vector_t& d = g.dHdth;
    if (d.size() > cap) throw std::runtime_error("capacity too small");
    for (int i = 0; i < d.size(); i++) out[i] = d(i);
    return d.size();
    CATCH(-1)
}

// GaitOptimizer::OptimizeContactTimes (gait_optimizer.cpp:185-364) as the reference wrote it, on the optimiser orc_mpc_gait_gradient
// set up (contact times of the MPC's trajectory).  grad: dH/dtheta to use (n entries).  solution == NULL: the reference builds its LP,
// the stand-in solver records it and the call returns 2 with dims = [rows, cols] and the dense A [rows][cols], lb, ub, q filled;
// solution != NULL: it is handed back as the LP's solution, the reference finishes the step, new_times [n] = its contact times after.
int orc_mpc_gait_lp(void* hv, double time, const double* grad, const double* solution, int* dims, double* A_dense, double* lb, double* ub,
                    double* q, double* new_times) {
    Handle* h = static_cast<Handle*>(hv);
    if (!h->gait) { g_err = "call orc_mpc_gait_gradient first"; return -1; }
    GaitOptimizer& g = *h->gait;
    auto& rec = OsqpEigen::recorded();
    int n = 0;
    for (int ee = 0; ee < 4; ee++) n += static_cast<int>(g.GetContactTimes().at(ee).size());
    g.dHdth = vector_t::Zero(n);
    for (int i = 0; i < n; i++) g.dHdth(i) = grad[i];
    rec.have_x = solution != nullptr;
    if (solution) {
        rec.x = vector_t::Zero(n);
        for (int i = 0; i < n; i++) rec.x(i) = solution[i];
    }
    auto export_lp = [&]() {
        dims[0] = rec.A.rows();
        dims[1] = rec.A.cols();
        for (int i = 0; i < rec.A.rows(); i++)
            for (int j = 0; j < rec.A.cols(); j++) A_dense[static_cast<size_t>(i) * rec.A.cols() + j] = rec.A.coeff(i, j);
        for (int i = 0; i < rec.l.size(); i++) { lb[i] = rec.l(i); ub[i] = rec.u(i); }
        for (int i = 0; i < rec.q.size(); i++) q[i] = rec.q(i);
    };
    try {
        MuteCout mute;
        g.OptimizeContactTimes(time, 0.0);
    } catch (const std::runtime_error& e) {
        g_err = e.what();
        if (!solution && g_err.find("no solution was injected") != std::string::npos) {
            export_lp();
            return 2;
        }
        return -1;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
    export_lp();
    int k = 0;
    for (int ee = 0; ee < 4; ee++)
        for (const auto& t : g.GetContactTimes().at(ee)) new_times[k++] = t.GetTime();
    return 0;
}
}  // extern "C"
