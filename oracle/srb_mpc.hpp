// TEST INFRASTRUCTURE ONLY -- CPU oracle.  Nothing under oracle/ may be imported, linked or executed by the
// product path; see foot_spline.hpp.
//
// Restatement (no Eigen / pinocchio / Clarabel) of the reference's live single-rigid-body RTI MPC:
//   mpc/trajectory.cpp (whole file)                      -> class Traj
//   mpc/models/single_rigid_body_model.cpp:19-42,55-256  -> class SrbModel
//   mpc/mpc.cpp:153-209,352-414,542-564,610-624,730-816,1076-1127,1205-1214 and
//   mpc/mpc_single_rigid_body.cpp:25-475,849-887,929-937 -> class SrbMpc
//   mpc/qp/qp_data.cpp:61-289, utils/sparse_matrix_builder.cpp:11-42 -> struct QpData / TripletBuilder
// pinocchio's quaternion log3 / exp3 / firstOrderNormalize are restated from its published algorithm
// (pinocchio/spatial/explog-quaternion.hpp, pinocchio/math/quaternion.hpp; dependency unpinned and absent here).
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "foot_spline.hpp"

namespace oracle {

using Vec = std::vector<double>;

struct Mat {   // small dense row-major matrix
    int r = 0, c = 0;
    Vec a;
    Mat() {}
    Mat(int r_, int c_) : r(r_), c(c_), a(static_cast<size_t>(r_) * c_, 0.0) {}
    double& operator()(int i, int j) { return a[static_cast<size_t>(i) * c + j]; }
    double operator()(int i, int j) const { return a[static_cast<size_t>(i) * c + j]; }
    void zero() { std::fill(a.begin(), a.end(), 0.0); }
};

struct Csc {   // compressed sparse column, what Eigen::SparseMatrix::setFromTriplets produces
    int rows = 0, cols = 0;
    std::vector<int> colptr, rowidx;
    Vec val;
    int nnz() const { return static_cast<int>(val.size()); }
    void mul(const double* x, double* y) const;      // y = M x
    void mul_t(const double* x, double* y) const;    // y = M^T x
};

// utils/sparse_matrix_builder.cpp:11-42
struct TripletBuilder {
    std::vector<int> ri, ci;
    Vec v;
    void Reserve() { ri.clear(); ci.clear(); v.clear(); }
    void Push(int r, int c, double val) { ri.push_back(r); ci.push_back(c); v.push_back(val); }
    void SetDiagonalMatrix(double val, int r0, int c0, int n) { for (int i = 0; i < n; i++) Push(r0 + i, c0 + i, val); }
    void SetMatrix(const Mat& M, int r0, int c0) {
        for (int i = 0; i < M.r; i++)
            for (int j = 0; j < M.c; j++)
                if (M(i, j) != 0) Push(r0 + i, c0 + j, M(i, j));
    }
    void SetRow(const Vec& row, int r0, int c0, double scale = 1.0) {   // SetMatrix(scale * row^T, r0, c0)
        for (size_t j = 0; j < row.size(); j++) {
            const double x = scale * row[j];
            if (x != 0) Push(r0, c0 + static_cast<int>(j), x);
        }
    }
    Csc Build(int rows, int cols) const;   // duplicates summed, rows sorted within a column (setFromTriplets)
};

enum Constraint { Dynamics, ForceBox, FrictionCone, EndEffectorLocation, TDPosition, EndEffectorStart };

enum SolveQuality { Solved = 0, SolvedInacc = 1, MaxIter = 2, PrimalInfeasible = 3, DualInfeasible = 4,
                    PrimalInfeasibleInacc = 5, DualInfeasibleInacc = 6, Unsolved = 7, Other = 8 };   // qp_interface.h:12-22

// mpc/include/qp/qp_data.h:51-126 (the fields the live Clarabel path uses)
struct QpData {
    std::vector<Constraint> constraints;
    TripletBuilder constraint_mat, cost_mat;
    Csc A, P;
    Vec ub;                 // Clarabel form: A z + s = ub, s in the cone of the block
    Vec dynamics_constants, friction_cone_ub, force_box_lb, force_box_ub, ee_location_lb, ee_location_ub,
        start_ee_constants, td_pos_constants, cost_linear;
    int num_dynamics = 0, num_vars = 0, num_cone = 0, num_force_box = 0, num_ee_location = 0, num_start_ee = 0,
        num_td_pos = 0;
    int num_equality = 0, num_inequality = 0;
    int Total() const;
    void InitQPMats();
    void ConstructSparseMats();
    void ConstructVectors();
    // is_eq[row] for the stacked rows (Zero cone vs Nonnegative cone), clarabel_interface.cpp:29-64
    std::vector<char> RowIsEquality() const;
};

// mpc/trajectory.cpp
class Traj {
public:
    Traj(int len, const std::vector<std::vector<double>>& switching_times, double node_dt, double swing_height,
         double foot_offset);
    void SetState(int idx, const Vec& s) { states_.at(idx) = s; }
    const Vec& GetState(int node) const { return states_.at(node); }
    int NumStates() const { return static_cast<int>(states_.size()); }
    int GetTotalPosSplineVars() const { return pos_vars_; }
    int GetTotalForceSplineVars() const { return force_vars_; }
    void UpdateForceSpline(int ee, int coord, const double* vars, int n);
    void UpdatePositionSpline(int ee, int coord, const double* vars, int n);
    std::pair<int, int> GetPositionSplineIndex(int ee, double time, int coord) const;
    std::pair<int, int> GetForceSplineIndex(int ee, double time, int coord) const;
    void AddPolys(double final_time);
    void RemoveUnusedPolys(double init_time);
    void SetInitTime(double t) { init_time_ = t; }
    std::vector<bool> GetContacts(double time) const;
    std::vector<std::vector<KnotTime>> GetContactTimes() const;
    bool IsForceMutable(int ee, double time) const { return ee_[ee].IsForceMutable(time); }
    Vec GetForceSplineLin(int ee, int coord, double time) const { return ee_[ee].GetPolyVarsLin(Force, coord, time); }
    Vec GetPositionSplineLin(int ee, int coord, double time) const;
    void GetForce(int ee, double time, double out[3]) const;
    void GetEndEffectorLocation(int ee, double time, double out[3]) const;
    double GetTime(int node) const { return init_time_ + node_dt_ * node; }
    int GetTotalVariables() const { return force_vars_ + pos_vars_ + NumStates() * 12; }
    Vec SplinesAsVec() const;
    void GetForcePartialWrtContactTime(int ee, double time, int contact_idx, double out[3]) const;
    void GetPositionPartialWrtContactTime(int ee, double time, int contact_idx, double out[3]) const;
    void UpdateContactTimes(std::vector<std::vector<KnotTime>>& ct);
    double GetNextContactTime(int ee, double time) const { return ee_[ee].GetNextTouchDownTime(time); }
    void SetEEInContact(int ee, double time) { ee_[ee].SetToTouchdown(time); }
    double GetCurrentSwingTime(int ee) const { return ee_[ee].GetSwingTime(init_time_); }
    int NumEE() const { return static_cast<int>(ee_.size()); }
    FootSpline& Foot(int ee) { return ee_[ee]; }
    const FootSpline& Foot(int ee) const { return ee_[ee]; }
    double InitTime() const { return init_time_; }
    double NodeDt() const { return node_dt_; }

private:
    void UpdateSplineVarsCount();
    void SetSwingPosZ();
    std::vector<Vec> states_;
    std::vector<FootSpline> ee_;
    int pos_vars_ = 0, force_vars_ = 0;
    double swing_height_, foot_offset_, init_time_ = 0, node_dt_;
};

struct RobotConsts {   // what the reference pulls out of pinocchio at construction
    double mass;
    double Ir[9], Ir_inv[9];       // row-major 3x3
    double hip_xy[4][2];           // GetCOMToHip(ee).head<2>() incl. the hard-coded offsets
    double gravity[3];
};

// pinocchio quaternion helpers (restated, see header comment)
void QuatLog3(const double q_xyzw[4], double out[3]);
void QuatExp3(const double v[3], double q_xyzw[4]);
void QuatFirstOrderNormalize(double q_xyzw[4]);

// mpc/models/single_rigid_body_model.cpp
class SrbModel {
public:
    explicit SrbModel(const RobotConsts& rc) : rc_(rc) {}
    void GetLinearDynamics(const Vec& state, const Vec& ref_state, const Traj& traj, double dt, double time, Mat& A,
                           Mat& B, Vec& C) const;                                                      // :55-169
    Vec CalcDynamics(const Vec& tan_state, const Traj& traj, double time) const;                       // :222-256
    Vec ManifoldToTangent(const Vec& s) const;                                                         // :188-200
    Vec TangentToManifold(const Vec& s) const;                                                         // :202-220
    void ComputeLinearizationPartialWrtContactTimes(Mat& dA, Mat& dB, Vec& dC, const Vec& state, const Traj& traj,
                                                    double time, int ee, int contact_idx) const;       // :458-555
    const RobotConsts& Consts() const { return rc_; }

private:
    RobotConsts rc_;
};

// mpc/include/mpc.h:39-62 (the fields that act on the live path)
struct MpcInfo {
    int num_nodes = 20;
    double friction_coef = 0.5;
    double integrator_dt = 0.05;
    double force_bound = 150;
    double swing_height = 0.075;
    double foot_offset = 0.015;
    double ee_box_size[2] = {0.15, 0.15};
    double force_cost = 0.0;
    int real_time_iters = 6000;
};

struct QpSolution {
    Vec x, dual, slack;    // Clarabel conventions: A x + s = b ; dual >= 0 on Nonnegative rows
    SolveQuality status = Unsolved;
    int iters = 0;
    double prim_res = 0, dual_res = 0;
    bool no_iterate = false;   // the solver broke down before its first finite residual evaluation: x holds nothing
};

// The solver seam (mpc/include/qp/qp_interface.h:30-65).  The oracle's implementation is the ADMM restatement in
// qp_admm.hpp; tests may also plug a reference solution in.
class QpSolver {
public:
    virtual ~QpSolver() {}
    virtual QpSolution Solve(const QpData& data, const Vec& warm_start, bool real_time) = 0;
};

struct SolveStats {       // MPC::RecordStats, mpc.cpp:804-816
    double alpha = 0, eq_violation = 0, step_norm = 0, cost = 0, merit = 0, merit_dd = 0;
    SolveQuality status = Unsolved;
    int qp_iters = 0;
};

// mpc/include/qp/qp_partials.h:15-36 (the fields the Clarabel path fills): partials of the QP data with respect to
// one contact time, as triplets in the equality-row / inequality-row numbering of ClarabelInterface::
// SetupDerivativeCalcs (equalities: Dynamics | TDPosition | EndEffectorStart; inequalities: ForceBox | FrictionCone |
// EndEffectorLocation).
struct ParamPartials {
    TripletBuilder dA, dG;
    Vec db, dh;
    int num_eq = 0, num_ineq = 0, num_vars = 0;
};

class SrbMpc {
public:
    // gait_partials.cpp
    bool ComputeParamPartialsClarabel(const Traj& traj, ParamPartials& out, int ee, int contact_idx) const;   // mpc_single_rigid_body.cpp:642-792
    void AddForceBoxConstraintPartials(TripletBuilder& b, int contact_idx, int start_idx, int ee) const;       // mpc.cpp:416-531
    void AddFrictionConeConstraintPartials(TripletBuilder& b, int contact_idx, int start_idx, int ee) const;   // mpc.cpp:240-350
    void AddTDPositionConstraintPartial(TripletBuilder& b, Vec& db, int contact_idx, int eq_idx, int ee) const;   // mpc_single_rigid_body.cpp:889-927

    SrbMpc(const MpcInfo& info, const RobotConsts& rc, std::shared_ptr<QpSolver> solver);

    void AddQuadraticTrackingCost(const Vec& state_des, const Mat& Q);   // mpc.cpp:533-540
    void SetQuadraticFinalCost(const Mat& Phi) { Phi_ = Phi; }
    void SetLinearFinalCost(const Vec& w) { Phi_w_ = w; }
    void SetStateTrajectoryWarmStart(const std::vector<Vec>& states);    // mpc.cpp:660-666
    void SetWarmStartTrajectory(const Traj& t);                          // mpc.cpp:110-119
    void UpdateContactTimes(std::vector<std::vector<KnotTime>>& ct) { prev_traj_.UpdateContactTimes(ct); }
    void AdjustForCurrentContacts(double time, const std::vector<bool>& in_contact);   // mpc.cpp:1195-1203

    const Traj& CreateInitialRun(const Vec& state, const std::vector<std::array<double, 3>>& ee_start);   // mpc.cpp:78-90
    const Traj& GetRealTimeUpdate(const Vec& state, double init_time, const std::vector<std::array<double, 3>>& ee_start);
    const Traj& Solve(const Vec& state, double init_time, const std::vector<std::array<double, 3>>& ee_start);

    // Stops after step 6 of SURVEY 3.1 (assembly) -- the parity tap for kernels 1-3.
    void AssembleOnly(const Vec& state, double init_time, const std::vector<std::array<double, 3>>& ee_start);

    const QpData& Data() const { return data_; }
    const Traj& Trajectory() const { return prev_traj_; }
    const Vec& PrevQpSol() const { return prev_qp_sol_; }
    const QpSolution& LastQp() const { return last_qp_; }
    const SolveStats& LastStats() const { return stats_; }
    double GetCost() const { return GetCostValue(prev_qp_sol_); }
    const SrbModel& Model() const { return model_; }
    void SetSolver(std::shared_ptr<QpSolver> s) { solver_ = std::move(s); }
    const MpcInfo& Info() const { return info_; }
    // per-node discretised dynamics of the last assembly (dense, as the reference holds them transiently)
    const std::vector<Mat>& NodeA() const { return node_A_; }
    const std::vector<Mat>& NodeB() const { return node_B_; }
    const std::vector<Vec>& NodeC() const { return node_C_; }

    Vec ConvertTrajToQPVec(const Traj& traj) const;                                  // mpc_single_rigid_body.cpp:343-357
    Traj ConvertQPSolToTrajectory(const Vec& qp_sol) const;                          // :275-321
    double GetCostValue(const Vec& x) const;                                         // mpc.cpp:759-761
    Vec GetEqualityConstraintValues(const Traj& traj) const;                         // mpc.cpp:764-776
    double GetMeritValue(const Vec& x) const;                                        // mpc.cpp:749-753
    double GetMeritGradient(const Vec& x, const Vec& p) const;                       // mpc.cpp:783-788
    double LineSearch(const Vec& direction) const;                                   // mpc.cpp:730-747
    int ForceSplineStartIdx() const { return 12 * (1 + info_.num_nodes); }
    int PosSplineStartIdx() const { return ForceSplineStartIdx() + prev_traj_.GetTotalForceSplineVars(); }
    double GetTime(int node) const { return node * info_.integrator_dt + init_time_; }

private:
    void Prepare(const Vec& state, double init_time, const std::vector<std::array<double, 3>>& ee_start);
    void UpdateQPSizes();
    int NumForceBoxConstraints() const;
    int NumFricConeConstraints() const;
    int NumTDConstraints() const;
    void AddCosts();
    void AddDynamicsConstraints(const Vec& state);
    void AddForceBoxConstraints();
    void AddFrictionConeConstraints();
    void AddEELocationConstraints();
    void AddTDPositionConstraints();
    void AddEEStartConstraints(const std::vector<std::array<double, 3>>& ee_start);

    MpcInfo info_;
    SrbModel model_;
    std::shared_ptr<QpSolver> solver_;
    Traj prev_traj_;
    QpData data_;
    Mat Q_, Phi_;
    Vec w_, Phi_w_;
    Vec prev_qp_sol_;
    double init_time_ = 0, mu_ = 5000, td_fraction_ = 0.75;
    double friction_pyramid_[4][3];
    double ee_bounds_[2];
    int num_inputs_ = 0, constraint_idx_ = 0;
    bool in_real_time_ = false;
    QpSolution last_qp_;
    SolveStats stats_;
    std::vector<Mat> node_A_, node_B_;
    std::vector<Vec> node_C_;
};

}  // namespace oracle
