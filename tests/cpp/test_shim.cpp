// Host-shim test in the shape of the reference's test/mpc_test.cpp:17-135: build an MPCSingleRigidBody from an MPCInfo,
// set the costs the way CreateMPC does (:21-40), run CreateInitialRun and a real-time update, read sizes / QP data /
// solve quality, then the gait optimiser's derivative -> LP -> line search sequence of MPCController::GaitOpt
// (controllers/mpc_controller.cpp:518-573, 322-343).  Prints "key value" lines that tests/test_host_shim.py checks
// against the oracle.  Needs a GPU (the library has no CPU fallback).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>

#include "config_parser.h"
#include "mpc_b200.h"

using namespace mpc;

static bgg_robot ReadRobot(const char* path) {
    // "mass Ir[9] Ir_inv[9] hip_xy[8]" as plain numbers (written by the python side from tests/golden/a1_robot_consts.json)
    bgg_robot rb{};
    FILE* f = std::fopen(path, "r");
    if (!f) throw std::runtime_error("cannot open robot constants file");
    double* dst[4] = {&rb.mass, rb.Ir, rb.Ir_inv, rb.hip_xy};
    const int cnt[4] = {1, 9, 9, 8};
    for (int k = 0; k < 4; ++k)
        for (int i = 0; i < cnt[k]; ++i)
            if (std::fscanf(f, "%lf", dst[k] + i) != 1) throw std::runtime_error("short robot constants file");
    std::fclose(f);
    rb.gravity[2] = -9.81;
    return rb;
}

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    try {
        MPCInfo info;   // apps/a1_configuration.yaml
        info.num_nodes = 20;
        info.integrator_dt = 0.05;
        info.friction_coef = 0.5;
        info.force_bound = 150;
        info.swing_height = 0.075;
        info.foot_offset = 0.015;
        info.ee_box_size = vector_2t(0.15, 0.15);
        info.force_cost = 0;
        MPCSingleRigidBody mpc(info, ReadRobot(argv[1]));

        const double q[12] = {340, 340, 4000, .1, .1, 10, 3000, 3000, 3000, 1, 1, 1};
        matrix_t Q = matrix_t::Zero(12, 12);
        for (int i = 0; i < 12; ++i) Q(i, i) = q[i];
        vector_t init_state(13), des_alg(12);
        const double s0[13] = {0, 0, .3, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0};
        for (int i = 0; i < 13; ++i) init_state(i) = s0[i];
        for (int i = 0; i < 12; ++i) des_alg(i) = (i == 2) ? 0.3 : 0.0;   // ConvertManifoldStateToTangentState(srb_target)
        mpc.SetStateTrajectoryWarmStart(std::vector<vector_t>(info.num_nodes + 1, init_state));
        mpc.AddQuadraticTrackingCost(des_alg, Q);
        mpc.AddForceCost(info.force_cost);
        mpc.SetQuadraticFinalCost(Q);
        vector_t lin(12);
        for (int i = 0; i < 12; ++i) lin(i) = -1 * q[i] * des_alg(i);
        mpc.SetLinearFinalCost(lin);

        const double ee_pos[4][3] = {{0.1526, 0.12523, 0.011089}, {0.1526, -0.12523, 0.011089}, {-0.208321844, 0.1363286, 0.01444},
                                     {-0.208321844, -0.1363286, 0.01444}};   // test/mpc_test.cpp:97-101
        std::vector<vector_3t> ee;
        for (auto& p : ee_pos) ee.emplace_back(p[0], p[1], p[2]);
        mpc.SetDefaultGaitTrajectory(Trot, 3, ee);

        bool threw = false;   // wrong-sized cost matrix -> std::runtime_error (mpc.cpp:122-124)
        try {
            mpc.SetQuadraticFinalCost(matrix_t::Zero(3, 3));
        } catch (const std::runtime_error&) {
            threw = true;
        }
        std::printf("bad_cost_throws %d\n", threw ? 1 : 0);

        Trajectory traj = mpc.CreateInitialRun(init_state, ee);
        std::printf("initial_quality %d\n", static_cast<int>(mpc.GetSolveQuality()));
        traj = mpc.GetRealTimeUpdate(init_state, 0.0, ee, false);
        std::printf("rt_quality %d\n", static_cast<int>(mpc.GetSolveQuality()));
        std::printf("num_decision_vars %d\n", mpc.GetNumDecisionVars());
        std::printf("num_constraints %d\n", mpc.GetNumConstraints());
        std::printf("cost %.12e\n", mpc.GetCost());
        const double cost_after_rt = mpc.GetCost();
        const QPData& data = mpc.GetQPData();
        std::printf("qp_rows %d\nqp_cols %d\nqp_nnz %d\n", data.sparse_constraint_.rows, data.sparse_constraint_.cols, data.sparse_constraint_.nonZeros());
        std::printf("num_equality %d\nnum_inequality %d\n", data.num_equality_, data.num_inequality_);
        std::printf("state1_z %.12e\n", traj.GetState(1)(2));
        std::printf("force_ee1_z_t01 %.12e\n", traj.GetForce(1, 0.1)(2));
        std::printf("ee0_x_t04 %.12e\n", traj.GetEndEffectorLocation(0, 0.4)(0));
        const std::vector<time_v> ct = traj.GetContactTimes();
        std::printf("num_contact_nodes %d %d %d %d\n", traj.GetNumContactNodes(0), traj.GetNumContactNodes(1), traj.GetNumContactNodes(2),
                    traj.GetNumContactNodes(3));

        // a copy solves to the same cost (MPC is a value type; the line search depends on it, gait_optimizer.cpp:696)
        MPCSingleRigidBody copy = mpc;
        copy.GetRealTimeUpdate(init_state, 0.0, ee, false);
        MPCSingleRigidBody again = mpc;
        again.GetRealTimeUpdate(init_state, 0.0, ee, false);
        std::printf("copy_cost_equal %d\n", copy.GetCost() == again.GetCost() ? 1 : 0);

        // MPCController::GaitOpt
        GaitOptimizer gait_opt(4, 10, mpc.GetNumDecisionVars(), mpc.GetNumConstraints(), 1.0, 0.05);
        const bool ok = mpc.ComputeDerivativeTerms();
        std::printf("derivative_terms %d\n", ok ? 1 : 0);
        gait_opt.SetContactTimes(mpc.GetTrajectory().GetContactTimes());
        gait_opt.UpdateSizes(mpc.GetNumDecisionVars(), mpc.GetNumConstraints());
        mpc.GetQPPartials(gait_opt.GetQPPartials());
        const Trajectory prev_traj = mpc.GetTrajectory();
        for (int e = 0; e < 4; ++e) {
            gait_opt.SetNumContactTimes(e, prev_traj.GetNumContactNodes(e));
            for (int idx = 0; idx < prev_traj.GetNumContactNodes(e); ++idx)
                mpc.ComputeParamPartialsClarabel(prev_traj, gait_opt.GetParameterPartials(e, idx), e, idx);
        }
        gait_opt.ModifyQPPartials(mpc.GetQPSolution());
        gait_opt.ComputeCostFcnDerivWrtContactTimes();
        std::printf("gradient");
        for (double g : gait_opt.GetGradient()) std::printf(" %.10e", g);
        std::printf("\n");
        gait_opt.OptimizeContactTimes(0.0, 0.0);
        std::printf("step");
        for (double s : gait_opt.GetStep()) std::printf(" %.10e", s);
        std::printf("\n");
        auto res = gait_opt.LineSearch(mpc, 0.0, ee, init_state);
        std::printf("ls_cost_min %.12e\n", res.second);
        std::printf("ls_first_times");
        for (const auto& t : res.first.at(0)) std::printf(" %.10e", t.GetTime());
        std::printf("\n");
        {   // per-solve statistics log in the reference's layout (mpc.cpp:901-989)
            std::ofstream log(std::string(argv[1]) + ".log");
            mpc.PrintStatLineToFile(log);
            mpc.GetRealTimeUpdate(init_state, 0.05, ee, false);
            mpc.PrintStatLineToFile(log);
        }
        {   // configuration file in the reference's YAML layout -> MPCInfo (test/mpc_test.cpp:43-83)
            if (argc > 2) {
                utils::ConfigParser config(argv[2]);
                const MPCInfo from_file = MPCInfoFromConfig(config);
                std::printf("yaml_num_nodes %d\nyaml_friction %.6f\nyaml_frames %zu\n", from_file.num_nodes, from_file.friction_coef,
                            from_file.ee_frames.size());
            }
        }
        {   // "Clarabel Solver" of test/mpc_test.cpp:857-953: the 3-variable QP through the solver seam (QPData + ClarabelInterface::SetupQP /
            // Solve), here on the CUDA path.  Rows in Clarabel form: the two equalities, then G x <= ub and -G x <= -lb.
            QPData qp;
            qp.num_decision_vars = 3;
            qp.constraints_ = {Dynamics, ForceBox};   // a Zero cone of 2 rows and a Nonnegative cone of 4 rows
            qp.num_dynamics_constraints = 2;
            qp.num_force_box_constraints_ = 4;
            qp.num_equality_ = 2;
            qp.num_inequality_ = 4;
            const double Pd[3] = {3.001, 4.0, 0.5};
            qp.sparse_cost_.rows = qp.sparse_cost_.cols = 3;
            qp.sparse_cost_.outer = {0, 1, 2, 3};
            qp.sparse_cost_.inner = {0, 1, 2};
            qp.sparse_cost_.values = {Pd[0], Pd[1], Pd[2]};
            const double A[6][3] = {{1, 1, 0}, {1.3, 0, 0.2}, {-2, 0, 0.9}, {1, 8, 5}, {2, 0, -0.9}, {-1, -8, -5}};
            const double b[6] = {1, 3, 3.1, 13.3, 2, 5};
            qp.sparse_constraint_.rows = 6;
            qp.sparse_constraint_.cols = 3;
            qp.sparse_constraint_.outer.push_back(0);
            for (int j = 0; j < 3; ++j) {
                for (int i = 0; i < 6; ++i)
                    if (A[i][j] != 0.0) {   // SparseMatrixBuilder drops exact zeros (utils/sparse_matrix_builder.cpp:25)
                        qp.sparse_constraint_.inner.push_back(i);
                        qp.sparse_constraint_.values.push_back(A[i][j]);
                    }
                qp.sparse_constraint_.outer.push_back(static_cast<int>(qp.sparse_constraint_.inner.size()));
            }
            qp.cost_linear = vector_t(3);
            qp.cost_linear(0) = 0.1; qp.cost_linear(1) = 4.6; qp.cost_linear(2) = 2.0;
            qp.ub_ = vector_t(6);
            for (int i = 0; i < 6; ++i) qp.ub_(i) = b[i];
            ClarabelInterface clarabel(qp, false);
            clarabel.SetupQP(qp, vector_t::Zero(3));
            const vector_t xs = clarabel.Solve(qp);
            std::printf("qp3_quality %d\nqp3_x %.12e %.12e %.12e\n", static_cast<int>(clarabel.GetSolveQuality()), xs(0), xs(1), xs(2));
            const vector_t dx = clarabel.Computedx(qp.sparse_cost_, qp.cost_linear, xs);
            const vector_t dual = clarabel.GetDualSolution();
            double stat = 0;   // stationarity dx + A'y = 0 (the dx of :955-958 with the returned multipliers)
            for (int j = 0; j < 3; ++j) {
                double r = dx(j);
                for (int i = 0; i < 6; ++i) r += A[i][j] * dual(i);
                stat = std::max(stat, std::abs(r));
            }
            std::printf("qp3_stationarity %.3e\n", stat);
            // an infeasible QP throws the string the reference's Solve() catches (clarabel_interface.cpp:112-114)
            QPData bad = qp;
            bad.ub_(2) = -50.0;   // -2 x0 + 0.9 x2 <= -50 contradicts the equalities and the other box rows
            bad.ub_(4) = -50.0;
            bool caught = false;
            try {
                clarabel.SetupQP(bad, vector_t::Zero(3));
                clarabel.Solve(bad);
            } catch (const std::string& e) {
                caught = (e == "Primal infeasible.");
            }
            std::printf("qp3_infeasible_throws %d\n", caught ? 1 : 0);
        }
        {   // MPC::AdjustForCurrentContacts (mpc.cpp:1195-1203): a foot that is measured in contact up to 70 ms before its planned
            // touch-down is put in contact now; one that is far from its touch-down is left alone
            MPCSingleRigidBody m2 = mpc;
            const Trajectory before = m2.GetTrajectory();
            const double t_now = before.GetTime(0);
            int early = -1;
            double td = 0;
            for (int e = 0; e < 4; ++e)
                if (!before.GetDesiredContacts(t_now).in_contact_.at(e)) { early = e; td = before.GetNextContactTime(e, t_now); break; }
            std::printf("adjust_swing_foot %d\n", early);
            if (early >= 0) {
                controller::Contact c(4);
                for (int e = 0; e < 4; ++e) c.in_contact_.at(e) = before.GetDesiredContacts(td - 0.05).in_contact_.at(e);
                c.in_contact_.at(early) = true;
                m2.AdjustForCurrentContacts(td - 0.05, c);    // 50 ms early: adjusted
                std::printf("adjust_near %d\n", m2.GetTrajectory().GetDesiredContacts(td - 0.05).in_contact_.at(early) ? 1 : 0);
                std::printf("adjust_near_times %.12e", td - 0.05);   // the time of the call, then the foot's contact times after it
                const std::vector<time_v> adjusted = m2.GetTrajectory().GetContactTimes();
                for (const auto& tv : adjusted.at(early)) std::printf(" %.12e", tv.GetTime());
                std::printf("\nadjust_before_times");
                const std::vector<time_v> unadjusted = before.GetContactTimes();
                for (const auto& tv : unadjusted.at(early)) std::printf(" %.12e", tv.GetTime());
                std::printf("\n");
                MPCSingleRigidBody m3 = mpc;
                m3.AdjustForCurrentContacts(td - 0.15, c);    // 150 ms early: outside the 70 ms window, unchanged
                std::printf("adjust_far %d\n", m3.GetTrajectory().GetDesiredContacts(td - 0.15).in_contact_.at(early) ? 1 : 0);
            }
        }
        {   // mpc::MPCCentroidal adapter (mpc/include/mpc_centroidal.h:15-221): the header's call sequence reaches the live MPC
            MPCCentroidal cen(info, ReadRobot(argv[1]));
            cen.SetStateTrajectoryWarmStart(std::vector<vector_t>(info.num_nodes + 1, init_state));
            cen.AddQuadraticTrackingCost(des_alg, Q);
            cen.SetQuadraticFinalCost(Q);
            cen.SetLinearFinalCost(lin);
            cen.CreateInitialRun(init_state);
            cen.GetRealTimeUpdate(6000, init_state, 0.0);
            std::printf("centroidal_vars %d\ncentroidal_cost_equal %d\n", cen.GetNumDecisionVars(), cen.Live().GetCost() == cost_after_rt ? 1 : 0);
        }
        {   // "Model Partials" of test/mpc_test.cpp:113-236: ComputeParamPartialsClarabel's dA / dG against finite differences of the
            // assembled constraint matrix when one contact time moves (dynamics, force-box and friction rows, margin 1e-4)
            const double h = std::sqrt(1e-16);
            MPCSingleRigidBody base = mpc;
            base.SetExportParamPartials(true);
            base.GetRealTimeUpdate(init_state, 0.0, ee, false);
            const Trajectory traj_u = base.GetTrajectory();          // the trajectory the partials are taken at
            MPCSingleRigidBody at = base;                            // its QP assembled at traj_u
            at.SetWarmStartTrajectory(traj_u);
            at.GetRealTimeUpdate(init_state, 0.0, ee, false);
            const QPData d1 = at.GetQPData();
            const std::vector<time_v> times = traj_u.GetContactTimes();
            double worst_dyn = 0, worst_fb = 0, worst_cone = 0;
            int checked = 0, nnz_a = 0, nnz_g = 0;
            const int nd = d1.num_dynamics_constraints, nfb = d1.num_force_box_constraints_, nc = d1.num_cone_constraints_;
            for (int e = 0; e < 4; ++e)
                for (int idx = 1; idx < static_cast<int>(times.at(e).size()); ++idx) {
                    std::vector<time_v> mod = times;
                    mod.at(e).at(idx).SetTime(mod.at(e).at(idx).GetTime() + h);
                    MPCSingleRigidBody moved = base;
                    moved.SetWarmStartTrajectory(traj_u);
                    moved.UpdateContactTimes(mod);
                    moved.GetRealTimeUpdate(init_state, 0.0, ee, false);
                    const QPData d2 = moved.GetQPData();
                    QPPartials part;
                    if (!base.ComputeParamPartialsClarabel(traj_u, part, e, idx)) throw std::runtime_error("no partials");
                    nnz_a += part.dA.nonZeros();
                    nnz_g += part.dG.nonZeros();
                    for (int c = 0; c < d1.sparse_constraint_.cols; ++c) {
                        for (int r = 0; r < nd; ++r)
                            worst_dyn = std::max(worst_dyn, std::abs(part.dA.coeff(r, c) - (d2.sparse_constraint_.coeff(r, c) - d1.sparse_constraint_.coeff(r, c)) / h));
                        for (int r = 0; r < nfb; ++r)
                            worst_fb = std::max(worst_fb, std::abs(part.dG.coeff(r, c) - (d2.sparse_constraint_.coeff(nd + r, c) - d1.sparse_constraint_.coeff(nd + r, c)) / h));
                        for (int r = 0; r < nc; ++r)
                            worst_cone = std::max(worst_cone, std::abs(part.dG.coeff(nfb + r, c) -
                                                                       (d2.sparse_constraint_.coeff(nd + nfb + r, c) - d1.sparse_constraint_.coeff(nd + nfb + r, c)) / h));
                    }
                    checked++;
                }
            std::printf("partials_checked %d\npartials_nnz %d %d\npartials_fd_dyn %.3e\npartials_fd_fb %.3e\npartials_fd_cone %.3e\n", checked, nnz_a, nnz_g,
                        worst_dyn, worst_fb, worst_cone);
        }
        std::printf("done 1\n");
    } catch (const std::exception& e) {
        std::printf("exception %s\n", e.what());
        return 1;
    }
    return 0;
}
